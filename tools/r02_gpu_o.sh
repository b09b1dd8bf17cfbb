#!/bin/bash
# round 2, GPU call O: one hand-over chunk per CTA (chunk 32) as shipped -- full GPU tests, bench N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/o_tests.log
tail -n 5 gpurun_out/o_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/o_bench.err
cat gpurun_out/o_bench.json
