#!/bin/bash
# round 2, GPU call L: what the 64-register cap of a 320-thread CTA costs (idle second helper warp), bench graph sizing check
mkdir -p gpurun_out
{
for v in "" w2; do
  if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
  echo "=== ${v:-shipped}"
  timeout 300 python tools/kbench.py --games 131072 1048576 --steps 400 --overlap
  timeout 300 python tools/kbench.py --games 131072 1048576 4194304 --steps 200
done; unset ORX_LIB; } > gpurun_out/l_w2.log 2>&1; cat gpurun_out/l_w2.log
unset ORX_LIB
timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/l_bench.err
python -c "
import json; d=json.load(open('gpurun_out/l_bench.json')); print(d['value'], d['ms_per_step'], d['replays'], d['roofline']['frac'], d['e2e']['value'])"
