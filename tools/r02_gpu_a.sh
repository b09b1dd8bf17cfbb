#!/bin/bash
# round 2, GPU call A: parity of the tile-flag mode + batch sweep of both ordering modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/a_tests.log
tail -n 3 gpurun_out/a_tests.log
S="131072 262144 524288 1048576"
{
for tpc in 0 1 2 3 6 8 12; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 36
echo "--- same state every step (L2 resident)"
for tpc in 0 1 2 8; do timeout 300 python tools/kbench.py --games $S --steps 400 --tpc $tpc --batches 1; done
timeout 300 python tools/kbench.py --games $S --steps 400 --path-flags 32 --batches 1
} > gpurun_out/a_sweep.log 2>&1
cat gpurun_out/a_sweep.log
