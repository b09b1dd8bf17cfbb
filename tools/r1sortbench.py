"""Experiment: what would ruleset R1's tick gain if the 32 games of a warp had the same enemy slots alive? The states are
populated, then PHYSICALLY permuted so that games are sorted by their alive-enemy bit mask (a one-off, on the planes;
the Philox streams follow the lane, so trajectories change but the workload is statistically the same), and the tick is
timed on the sorted against the unsorted states over a few steps (the order decays as enemies die and spawn).
    python tools/r1sortbench.py [games] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200 import _abi
from optimax_rogue_b200.r1 import R1GameState
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nb = 18


def make(sort):
    bs = [R1GameState(G, max_ticks=1000, auto_reset=True, seed=3, game_id_base=b * G, overlap_ticks=True).reset() for b in range(nb)]
    for b in bs:
        b.rollout(150)
        if sort:
            alive = ((b.ent_loc[:, 2:10] >> 16) & 1).to(torch.int64)
            key = (alive << torch.arange(8, device='cuda')).sum(1)
            perm = torch.argsort(key, stable=True)
            for name, _, _ in _abi.R1_PLANES:
                p = getattr(b, name)
                p.copy_(p[perm].clone())
    return bs


mv = torch.randint(1, 7, (4, G, 2), dtype=torch.uint8, device='cuda')
for sort in (False, True, False, True):
    bs = make(sort)
    res = [torch.empty((G,), dtype=torch.uint8, device='cuda') for _ in range(nb)]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for k in range(2):
            bs[k].update(mv[k], out=res[k])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()               # K ticks per state; every replay continues from where the last one ended
        with torch.cuda.graph(g, stream=st):
            for k in range(K * nb):
                bs[k % nb].update(mv[k % 4], out=res[k % nb])
        out = []
        for rnd in range(5):                      # the sorted order decays from replay to replay
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st)
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / (K * nb) * 1e3)
    alive = ((bs[0].ent_loc[:, 2:10] >> 16) & 1).float().sum(1)
    print(f'G={G} sorted={int(sort)}: us/step per round of {K} ticks per state: ' + ' '.join(f'{x:.1f}' for x in out) +
          f'; alive enemies per game now: mean {alive.mean():.2f}, games with none {float((alive == 0).float().mean()):.2f}', flush=True)
