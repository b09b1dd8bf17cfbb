#!/bin/bash
# round 2, GPU call N: hand-over chunk size 8 / 16 / 32, then the extended fuzz campaign
mkdir -p gpurun_out
{
S="131072 262144 524288 1048576"
for v in chunk8 chunk16 chunk32; do
  export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so
  echo "=== $v"
  timeout 300 python tools/kbench.py --games $S --steps 400 --overlap
done; unset ORX_LIB
} > gpurun_out/n_sweep.log 2>&1; cat gpurun_out/n_sweep.log
unset ORX_LIB
timeout 400 python tests/fuzz_campaign.py 240 > gpurun_out/n_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -n 40 gpurun_out/n_fuzz.log
