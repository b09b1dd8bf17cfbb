#!/bin/bash
# round 2, GPU call P: ticket + dependent launch as the first thing a CTA does (shipped) vs after the set-up (lateticket)
# vs trigger before the ticket (experiment, ordering not guaranteed); tiles-per-CTA sweep; flag-mode tests
mkdir -p gpurun_out
{
S="131072 262144 1048576"
for v in "" lateticket trigfirst; do
  if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
  echo "=== ${v:-shipped}"
  timeout 300 python tools/kbench.py --games $S --steps 400 --overlap
done; unset ORX_LIB
echo "=== shipped, tiles per CTA at 2^17 (default 7)"
for t in 2 3 4 5 10 14; do timeout 120 python tools/kbench.py --games 131072 --steps 400 --overlap --tpc $t; done
echo "=== shipped, tiles per CTA at 2^20 (default 32 = the cap)"
for t in 10 14 20 28; do timeout 120 python tools/kbench.py --games 1048576 --steps 400 --overlap --tpc $t; done
} > gpurun_out/p_sweep.log 2>&1; cat gpurun_out/p_sweep.log
unset ORX_LIB
timeout 900 python -m pytest tests/test_gpu_tile_flags.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 3
