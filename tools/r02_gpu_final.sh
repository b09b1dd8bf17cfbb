#!/bin/bash
# round 2: last check of the final tree -- the driver's own sequence: GPU tests, smoke(), bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/final_tests.log; tail -n 3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 1
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench (no flags) rc=$?"; tail -n 2 gpurun_out/final_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/final_bench.json')); r = json.load(open('gpurun_out/final_bench_ref.json'))
print('value', d['value'], 'steps', d['steps'], 'warmup', d['warmup'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('reference', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'], 'same config', r['config'] == d['config'], 'ratio e2e', d['e2e']['value'] / r['value'])
PY
