#!/bin/bash
# round 2, GPU call D: chunk-level hand-over -- tests, sweep, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/d_tests.log
tail -n 8 gpurun_out/d_tests.log
S="131072 262144 524288 1048576"
{
for tpc in 0 4 8 16 32; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32
echo "--- same state every step"
for tpc in 0 4 8 16; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc --batches 1; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32 --batches 1
} > gpurun_out/d_sweep.log 2>&1
cat gpurun_out/d_sweep.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
tail -c 4000 gpurun_out/d_bench.json; tail -n 5 gpurun_out/d_bench.err
