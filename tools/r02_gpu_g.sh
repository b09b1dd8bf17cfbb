#!/bin/bash
# round 2, GPU call G: all tests, A/B vs r1 tree, batch sweep both modes, e2e parts, full bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/g_tests.log
tail -n 5 gpurun_out/g_tests.log
{
echo "=== r1 tree"; (cd build/r1tree && timeout 300 python tools/kbench.py --games 131072 1048576 4194304 --steps 200)
echo "=== current"; timeout 300 python tools/kbench.py --games 131072 262144 524288 1048576 4194304 --steps 400
echo "=== current, grid-wait mode forced"; timeout 300 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --path-flags 32
echo "=== current, flag mode forced at 2^20"; timeout 300 python tools/kbench.py --games 1048576 --steps 400 --path-flags 64
} > gpurun_out/g_ab.log 2>&1; cat gpurun_out/g_ab.log
timeout 300 python tools/e2eparts2.py 1048576 > gpurun_out/g_e2eparts.log 2>&1
timeout 300 python tools/e2eparts2.py 131072 >> gpurun_out/g_e2eparts.log 2>&1
grep -E "host_stepper|G=" gpurun_out/g_e2eparts.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err; echo "ref rc=$?"
tail -n 3 gpurun_out/g_bench.err
