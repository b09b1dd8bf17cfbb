#!/bin/bash
# round 2, GPU call X: the evidence round again without the batch sweep (call W's files exceeded gpurun's 64 MiB and were not brought back)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/w_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/w_tests.log
SKIP_SWEEP=1 bash tools/profile_round3.sh
du -sh gpurun_out
