#!/bin/bash
# round 2, GPU call V: R1 rows prefetched into L2 ahead of the block's turn (shipped) vs not (r1nopf)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r1.py -m gpu -q -x -k "overlapped" 2>&1 | tail -n 3
{ for v in "" r1nopf ""; do
    if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
    echo "=== R1 variant ${v:-shipped}"; timeout 300 python tools/r1bench.py 65536 2; timeout 300 python tools/r1bench.py 1048576 2
  done; unset ORX_LIB; } > gpurun_out/v_r1.log 2>&1; grep -v "^  \|Traceback\|\^" gpurun_out/v_r1.log
