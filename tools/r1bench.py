import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200.r1 import R1GameState
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0     # _abi.R1_PATH_* (2 = block hand-over, the throughput mode)
nb = max(2, min(18, (300_000_000 // (244 * G)) + 1))
K = 4 * nb
bs = [R1GameState(G, max_ticks=1000, auto_reset=True, seed=3, game_id_base=b * G, path_flags=flags).reset() for b in range(nb)]
mv = torch.randint(1, 7, (4, G, 2), dtype=torch.uint8, device='cuda')
res = [torch.empty((G,), dtype=torch.uint8, device='cuda') for _ in range(nb)]
for b in bs: b.rollout(64)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for k in range(3): bs[k % nb].update(mv[k % 4], out=res[k % nb])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for k in range(K): bs[k % nb].update(mv[k % 4], out=res[k % nb])
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    stats = torch.zeros(8, dtype=torch.int64, device='cuda')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for b in range(4): bs[b % nb].rollout(64, stats)
    e1.record(st); torch.cuda.synchronize()
print(f'G={G} path_flags={flags} nb={nb} R1 step: {best * 1e3:.1f} us/step, {G / best * 1e3:.3e} ticks/s; rollout {4 * G * 64 / e0.elapsed_time(e1) * 1e3:.3e} ticks/s', flush=True)
