#!/bin/bash
# round 2, GPU call M: default run length in throughput mode (half a CTA per SM), chunk size variants
mkdir -p gpurun_out
{
S="131072 262144 524288 1048576"
for v in "" chunk8 chunk2; do
  if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
  echo "=== ${v:-shipped (chunk 4)}"
  timeout 300 python tools/kbench.py --games $S --steps 400 --overlap
  timeout 300 python tools/kbench.py --games 131072 524288 --steps 400 --overlap --batches 1
done; unset ORX_LIB
echo "=== shipped, 16 / 32 tiles per CTA"
for tpc in 16 32; do timeout 300 python tools/kbench.py --games $S --steps 400 --overlap --tpc $tpc; done
} > gpurun_out/m_sweep.log 2>&1; cat gpurun_out/m_sweep.log
timeout 600 python -m pytest tests/test_gpu_tile_flags.py -m gpu -q -x 2>&1 | tail -n 3
