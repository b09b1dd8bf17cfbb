#!/bin/bash
# Evidence for profiles/: run on a GPU box (gpurun). Each ncu pass follows a plain run of the same command.
#   bash tools/profile_round.sh r01
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python tools/profile_targets.py > gpurun_out/plain_targets.log 2>&1 || { echo "plain targets failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:^k_step_pipe$ --launch-skip 3 --launch-count 6 -f \
    -o gpurun_out/prof_pipe_$TAG python tools/profile_targets.py > gpurun_out/ncu_pipe.log 2>&1
tail -n 2 gpurun_out/ncu_list.log; tail -n 2 gpurun_out/ncu_pipe.log
