#!/bin/bash
# round 2, GPU call J (2 GPUs): the bench under torchrun, both arms, and the 2-GPU shard-invariance check
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err; echo "bench n2 rc=$?"
tail -n 3 gpurun_out/j_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/j_bench_ref_n2.json 2> gpurun_out/j_bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/multigpu_check.py > gpurun_out/j_multigpu_check.log 2>&1; echo "multigpu rc=$?"; tail -n 4 gpurun_out/j_multigpu_check.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/j_bench_n2.json'))
print({k: d[k] for k in ('value', 'n_gpus', 'scaling', 'ms_per_step', 'replays')}, d['roofline']['frac'], d['e2e']['value'], d['weak_scaling']['value'], d['weak_scaling']['e2e']['value'], d['measurement']['numa'])
r = json.load(open('gpurun_out/j_bench_ref_n2.json'))
print(r['value'], r['cpu_baseline']['kind'], r['config'] == d['config'])
PY
