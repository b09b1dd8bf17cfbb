#!/bin/bash
# round 2, GPU call Q: R1 with the rare paths rolled (one copy of the Philox / free-tile code each), quarter-of-the-SMs
# run length in throughput mode, throughput mode above 2^20 games
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r1.py -m gpu -q -x 2>&1 | tail -n 3
{ timeout 300 python tools/r1bench.py 65536 2; timeout 300 python tools/r1bench.py 65536 0; timeout 300 python tools/r1bench.py 1048576 2; } > gpurun_out/q_r1.log 2>&1; cat gpurun_out/q_r1.log
{
echo "=== shipped (run length: tiles / (SMs/4), at most 32)"
timeout 300 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --overlap
for t in 20 28; do timeout 120 python tools/kbench.py --games 131072 --steps 400 --overlap --tpc $t; done
timeout 120 python tools/kbench.py --games 131072 524288 --steps 400 --overlap --batches 1
echo "=== grid-wait mode"
timeout 300 python tools/kbench.py --games 2097152 4194304 --steps 200
echo "=== throughput mode allowed up to 2^22 games (tuning build)"
ORX_LIB=$PWD/optimax_rogue_b200/liborx_flagmax.so timeout 300 python tools/kbench.py --games 2097152 4194304 --steps 200 --overlap
} > gpurun_out/q_sweep.log 2>&1; cat gpurun_out/q_sweep.log
timeout 900 python -m pytest tests/test_gpu_tile_flags.py -m gpu -q -x 2>&1 | tail -n 3
