#!/bin/bash
# round 2, GPU call I: opt-in throughput mode -- tests, bench N=1, R1 CTA-size variants, evidence round (ncu + kernel trace)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/i_tests.log
tail -n 5 gpurun_out/i_tests.log
{ for v in "" r1t64 r1t256 r1t128b4; do
    if [ -n "$v" ]; then export ORX_LIB=$PWD/optimax_rogue_b200/liborx_$v.so; else unset ORX_LIB; fi
    echo "=== R1 variant ${v:-shipped}"; timeout 300 python tools/r1bench.py 65536 2; timeout 300 python tools/r1bench.py 65536 0
  done; unset ORX_LIB; } > gpurun_out/i_r1.log 2>&1; cat gpurun_out/i_r1.log
unset ORX_LIB
bash tools/profile_round2.sh
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/i_bench.err
