#!/bin/bash
# round 2, GPU call T: completion word of the synchronous host tick -- tests, e2e parts A/B (path flag 64 = stream sync), bench N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile_flags.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 5
{ for G in 1048576 131072; do for f in 0 64; do timeout 300 python tools/e2eparts2.py $G $f 2>&1 | grep -E "^G=|host_stepper|sync   :|empty"; done; done; } > gpurun_out/t_e2e.log 2>&1; cat gpurun_out/t_e2e.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/t_bench.err
python -c "
import json; d=json.load(open('gpurun_out/t_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('large_batch'), d['e2e']['value'], d['e2e']['us_per_step'], d['r1']['us_per_step'], d['r1']['hbm_frac'])"
