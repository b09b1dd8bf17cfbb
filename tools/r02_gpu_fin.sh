#!/bin/bash
# round 2: smoke() and the bench line of the final tree (N = 1, every extra)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/fin_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/fin_bench.json'))
r = d['roofline']
print(d['value'], d['ms_per_step'] * 1e3, r['frac'], 'large', r['large_batch']['frac'], 'grid-wait', r['grid_wait_mode']['us_per_step'], r['grid_wait_mode']['frac'], 'e2e', d['e2e']['value'], 'r1', d['r1']['us_per_step'])
PY
