#!/bin/bash
# round 2, GPU call R: R1 with the common-case prefilter; 5 / 6 CTAs per SM (96 / 80 registers); throughput mode up to 2^22 games
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r1.py -m gpu -q -x 2>&1 | tail -n 3
{ for v in "" r1mb5 r1mb6; do
    if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
    echo "=== R1 variant ${v:-shipped}"; timeout 300 python tools/r1bench.py 65536 2; timeout 300 python tools/r1bench.py 65536 0; timeout 300 python tools/r1bench.py 1048576 2
  done; unset ORX_LIB; } > gpurun_out/r_r1.log 2>&1; cat gpurun_out/r_r1.log
unset ORX_LIB
timeout 900 python -m pytest tests/test_gpu_tile_flags.py -m gpu -q -x 2>&1 | tail -n 3
timeout 300 python tools/kbench.py --games 2097152 4194304 --steps 200 --overlap
