#!/bin/bash
# round 2, GPU call C: all GPU tests, strict-fence build on the flag tests, tiles-per-CTA sweep, new bench.py
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/c_tests.log
tail -n 15 gpurun_out/c_tests.log
ORX_LIB=$PWD/optimax_rogue_b200/liborx_strict.so timeout 600 python -m pytest tests/test_gpu_tile_flags.py -m gpu -q > gpurun_out/c_tests_strict.log 2>&1; echo "strict rc=$?" | tee -a gpurun_out/c_tests_strict.log
{
for tpc in 2 3 4 5 6 8; do timeout 300 python tools/kbench.py --games 131072 262144 --steps 400 --tpc $tpc; done
for tpc in 6 8 10 12 14 16 20; do timeout 300 python tools/kbench.py --games 524288 --steps 400 --tpc $tpc; done
for tpc in 8 10 12 14 16 19 24 28; do timeout 300 python tools/kbench.py --games 1048576 --steps 400 --tpc $tpc; done
timeout 300 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --path-flags 32
echo "--- same state every step"
for tpc in 1 2 4 6; do timeout 300 python tools/kbench.py --games 131072 1048576 --steps 400 --tpc $tpc --batches 1; done
timeout 300 python tools/kbench.py --games 131072 1048576 --steps 400 --path-flags 32 --batches 1
} > gpurun_out/c_sweep.log 2>&1
cat gpurun_out/c_sweep.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err; echo "ref rc=$?"
tail -c 3000 gpurun_out/c_bench.json; tail -n 5 gpurun_out/c_bench.err; cat gpurun_out/c_bench_ref.json
