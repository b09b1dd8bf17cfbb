#!/bin/bash
# round 2, GPU call H: R1 with block hand-over (tests + bench leg), flag-mode policy sweep incl. isolated launches, evidence round
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r1.py tests/test_gpu_tile_flags.py -m gpu -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/h_tests.log
tail -n 4 gpurun_out/h_tests.log
{
S="131072 262144 524288"
echo "=== rotating"; for tpc in 0 4; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32
echo "=== isolated (ordinary kernel between ticks), rotating"
for tpc in 0 4; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc --isolate; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32 --isolate
echo "=== isolated, same state"
for tpc in 0 4; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc --isolate --batches 1; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32 --isolate --batches 1
echo "=== same state, back to back"
for tpc in 0 4; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc --batches 1; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32 --batches 1
echo "=== 2^20..2^22, flag mode forced (path 64) vs grid-wait"
for tpc in 0 12 16; do timeout 300 python tools/kbench.py --games 1048576 2097152 --steps 200 --path-flags 64 --tpc $tpc; done
timeout 300 python tools/kbench.py --games 1048576 2097152 --steps 200 --path-flags 32
for tpc in 0 12; do timeout 300 python tools/kbench.py --games 1048576 --steps 200 --path-flags 64 --tpc $tpc --isolate; done
timeout 300 python tools/kbench.py --games 1048576 --steps 200 --path-flags 32 --isolate
for tpc in 0 12; do timeout 300 python tools/kbench.py --games 1048576 --steps 200 --path-flags 64 --tpc $tpc --batches 1; done
timeout 300 python tools/kbench.py --games 1048576 --steps 200 --path-flags 32 --batches 1
} > gpurun_out/h_sweep.log 2>&1; cat gpurun_out/h_sweep.log
{ for g in 65536 262144; do for f in 0 2; do timeout 300 python tools/r1bench.py $g $f; done; done; } > gpurun_out/h_r1bench.log 2>&1; tail -n 12 gpurun_out/h_r1bench.log
