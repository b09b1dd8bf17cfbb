#!/bin/bash
# round 2, GPU call S: R1 entry points (replication log, bots, replay, host buffers) against the oracle; same-box A/B of
# the common-case prefilter and of 5 CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r1.py -m gpu -q -x 2>&1 | tail -n 15
{ for v in "" r1nopre r1mb5 r1nopremb5; do
    if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
    echo "=== R1 variant ${v:-shipped}"; timeout 300 python tools/r1bench.py 65536 2; timeout 300 python tools/r1bench.py 65536 0; timeout 300 python tools/r1bench.py 1048576 2
  done; unset ORX_LIB; } > gpurun_out/s_r1.log 2>&1; grep -v "^  \|Traceback\|\^" gpurun_out/s_r1.log
