#!/bin/bash
# round 2, GPU call K (8 GPUs): the bench under torchrun at N = 8 and N = 4, reference arm at N = 8
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 --no-extras > gpurun_out/k_bench_n$n.json 2> gpurun_out/k_bench_n$n.err; echo "bench n$n rc=$?"
  tail -n 2 gpurun_out/k_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 20 --warmup 3 > gpurun_out/k_bench_ref_n8.json 2> gpurun_out/k_bench_ref_n8.err; echo "ref n8 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    d = json.load(open(f'gpurun_out/k_bench_n{n}.json'))
    print(n, {k: d[k] for k in ('value', 'scaling', 'ms_per_step', 'replays')}, 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e']['value'], 'weak', d['weak_scaling']['value'], 'weak e2e', d['weak_scaling']['e2e']['value'], d['measurement']['numa'])
r = json.load(open('gpurun_out/k_bench_ref_n8.json')); print(r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
