#!/bin/bash
# Evidence for profiles/ (round 2): run on a GPU box (gpurun). Each ncu pass follows a plain run of the same command.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_plain.json 2> gpurun_out/p_bench_plain.err || { echo "plain bench failed"; tail -n 5 gpurun_out/p_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/p_ncu_list.log 2>&1
tail -n 2 gpurun_out/p_ncu_list.log
for spec in "1048576 8 0 pipe2p20" "131072 12 32 pipe2p17" "16777216 4 0 pipe2p24"; do
  set -- $spec
  python tools/profile_targets2.py $1 $2 $3 > gpurun_out/p_targets_$4.log 2>&1 || { echo "plain targets $4 failed"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:^k_step_pipe$ --launch-skip 2 --launch-count 2 -f \
      -o gpurun_out/prof_$4_r02 python tools/profile_targets2.py $1 $2 $3 > gpurun_out/p_ncu_$4.log 2>&1
  tail -n 1 gpurun_out/p_ncu_$4.log
done
python tools/kernel_trace.py 1048576 gpurun_out/kernel_trace_2p20_r02.csv 200 > gpurun_out/kernel_trace_2p20_r02.txt 2>&1; tail -n 4 gpurun_out/kernel_trace_2p20_r02.txt
python tools/kernel_trace.py 131072 gpurun_out/kernel_trace_2p17_r02.csv 400 > gpurun_out/kernel_trace_2p17_r02.txt 2>&1; tail -n 4 gpurun_out/kernel_trace_2p17_r02.txt
