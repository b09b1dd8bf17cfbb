#!/bin/bash
# round 2, GPU call E: tests (NPC refactor, compat), NPC/EV bench, e2e breakdown, PCIe microbench, 2^22 wait-mode check
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/e_tests.log
tail -n 6 gpurun_out/e_tests.log
timeout 600 python tools/evbench.py > gpurun_out/e_evbench.log 2>&1; cat gpurun_out/e_evbench.log
timeout 300 python tools/kbench.py --games 1048576 4194304 --steps 200 > gpurun_out/e_kbench_big.log 2>&1; cat gpurun_out/e_kbench_big.log
timeout 300 ./tools/pciebench > gpurun_out/e_pcie.log 2>&1; cat gpurun_out/e_pcie.log
timeout 300 python tools/e2eparts2.py 1048576 > gpurun_out/e_e2eparts.log 2>&1
timeout 300 python tools/e2eparts2.py 131072 >> gpurun_out/e_e2eparts.log 2>&1
timeout 300 python tools/e2eparts2.py 131072 32 >> gpurun_out/e_e2eparts.log 2>&1
cat gpurun_out/e_e2eparts.log
