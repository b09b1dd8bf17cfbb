#!/bin/bash
# round 2, GPU call F: wait-mode regression A/B (r1 tree vs current vs build without flag code), bits e2e with per-chunk command fetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile_flags.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/f_tests.log
tail -n 4 gpurun_out/f_tests.log
{
echo "=== r1 tree"; (cd build/r1tree && timeout 300 python tools/kbench.py --games 1048576 4194304 --steps 200)
echo "=== current"; timeout 300 python tools/kbench.py --games 1048576 4194304 --steps 200
echo "=== current, static tiles"; timeout 300 python tools/kbench.py --games 1048576 4194304 --steps 200 --path-flags 4
echo "=== without flag-mode code"; ORX_LIB=$PWD/optimax_rogue_b200/liborx_noflag.so timeout 300 python tools/kbench.py --games 1048576 4194304 --steps 200
echo "=== r1 tree again"; (cd build/r1tree && timeout 300 python tools/kbench.py --games 1048576 4194304 --steps 200)
} > gpurun_out/f_ab.log 2>&1; cat gpurun_out/f_ab.log
timeout 300 python tools/e2eparts2.py 1048576 > gpurun_out/f_e2eparts.log 2>&1
timeout 300 python tools/e2eparts2.py 131072 >> gpurun_out/f_e2eparts.log 2>&1
cat gpurun_out/f_e2eparts.log
