#!/bin/bash
# Evidence for profiles/ (round 2, final build): run on a GPU box (gpurun). Each ncu pass follows a plain run of the same command.
# (gpurun brings back at most 64 MiB: one launch per capture, no source import; SKIP_SWEEP=1 skips the tests and the batch sweep.)
mkdir -p gpurun_out
if [ -z "$SKIP_SWEEP" ]; then
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/w_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/w_tests.log
tail -n 3 gpurun_out/w_tests.log
{
S="131072 262144 524288 1048576 2097152 4194304"
echo "=== rotating states, ticks back to back"
timeout 600 python tools/kbench.py --games $S --steps 400 --overlap
timeout 600 python tools/kbench.py --games $S --steps 400
echo "=== one state, ticks back to back"
timeout 600 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --overlap --batches 1
timeout 600 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --batches 1
echo "=== rotating states, an ordinary kernel between ticks"
timeout 600 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --overlap --isolate
timeout 600 python tools/kbench.py --games 131072 262144 524288 1048576 --steps 400 --isolate
} > gpurun_out/w_sweep.log 2>&1; cat gpurun_out/w_sweep.log
fi
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_plain.json 2> gpurun_out/p_bench_plain.err || { echo "plain bench failed"; tail -n 5 gpurun_out/p_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/p_ncu_list.log 2>&1
tail -n 2 gpurun_out/p_ncu_list.log
for spec in "1048576 8 32 pipe2p20t" "1048576 8 0 pipe2p20" "131072 12 32 pipe2p17" "16777216 4 0 pipe2p24"; do
  set -- $spec
  python tools/profile_targets2.py $1 $2 $3 > gpurun_out/p_targets_$4.log 2>&1 || { echo "plain targets $4 failed"; continue; }
  ncu --set full --clock-control none -k regex:^k_step_pipe$ --launch-skip 2 --launch-count 1 -f \
      -o gpurun_out/prof_$4_r02 python tools/profile_targets2.py $1 $2 $3 > gpurun_out/p_ncu_$4.log 2>&1
  tail -n 1 gpurun_out/p_ncu_$4.log
  # gpurun brings back at most 64 MiB: keep the raw page of every capture, and the report itself of one
  ncu -i gpurun_out/prof_$4_r02.ncu-rep --page raw --csv > gpurun_out/prof_$4_r02.ncu-rep.raw.csv 2>/dev/null
  if [ "$4" != "pipe2p20t" ]; then rm -f gpurun_out/prof_$4_r02.ncu-rep; fi
done
python tools/kernel_trace.py 1048576 gpurun_out/kernel_trace_2p20_r02.csv 200 > gpurun_out/kernel_trace_2p20_r02.txt 2>&1; tail -n 4 gpurun_out/kernel_trace_2p20_r02.txt
python tools/kernel_trace.py 131072 gpurun_out/kernel_trace_2p17_r02.csv 400 > gpurun_out/kernel_trace_2p17_r02.txt 2>&1; tail -n 4 gpurun_out/kernel_trace_2p17_r02.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/w_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/w_bench_ref.json 2> gpurun_out/w_bench_ref.err; echo "reference arm rc=$?"
timeout 300 python tests/fuzz_campaign.py 180 > gpurun_out/w_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -n 3 gpurun_out/w_fuzz.log
du -sh gpurun_out; ls -S -l gpurun_out | head -5
