#!/bin/bash
# round 2, GPU call Z4 (4 GPUs): the bench under torchrun at N = 4 and N = 2, final build; 2-GPU shard-invariance check
mkdir -p gpurun_out
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 --no-extras > gpurun_out/z_bench_n$n.json 2> gpurun_out/z_bench_n$n.err; echo "bench n$n rc=$?"
  tail -n 1 gpurun_out/z_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/multigpu_check.py > gpurun_out/z_multigpu_check.log 2>&1; echo "multigpu rc=$?"; tail -n 4 gpurun_out/z_multigpu_check.log
python - <<'PY'
import json
for n in (4, 2):
    d = json.load(open(f'gpurun_out/z_bench_n{n}.json'))
    print(n, {k: d[k] for k in ('value', 'scaling', 'ms_per_step', 'replays')}, 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e']['value'], 'weak', d['weak_scaling']['value'], 'weak e2e', d['weak_scaling']['e2e']['value'])
PY
