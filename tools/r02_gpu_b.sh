#!/bin/bash
# round 2, GPU call B: flag-mode tests, then what each fence of the hand-over costs (tuning variants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile_flags.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/b_tests.log
tail -n 3 gpurun_out/b_tests.log
S="131072 1048576"
{
for v in "" v1 v2 v3 v4 v5; do
  if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; fi
  echo "=== variant ${v:-shipped}"
  timeout 300 python tools/kbench.py --games $S --steps 400 --tpc 6
  timeout 300 python tools/kbench.py --games 131072 --steps 400 --tpc 6 --batches 1
done
} > gpurun_out/b_variants.log 2>&1
cat gpurun_out/b_variants.log
