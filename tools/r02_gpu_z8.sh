#!/bin/bash
# round 2, GPU call Z8 (8 GPUs): the bench under torchrun at N = 8, final build
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > gpurun_out/z_bench_n8.json 2> gpurun_out/z_bench_n8.err; echo "bench n8 rc=$?"
tail -n 2 gpurun_out/z_bench_n8.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/z_bench_n8.json'))
print({k: d[k] for k in ('value', 'scaling', 'ms_per_step', 'replays')}, 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e']['value'], 'weak', d['weak_scaling']['value'], 'weak e2e', d['weak_scaling']['e2e']['value'], d['measurement']['numa'])
PY
