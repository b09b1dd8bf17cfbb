"""Tuning aid: orx_observe, orx_step and orx_step_observe over rotating batches (> L2), CUDA graph + events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device('cuda')
cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True)
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
nb = 9
bs = []
for b in range(nb):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); upd.rollout(gs, 1, 1, 20); bs.append(gs)
obs = [torch.empty((G, 2, 12), dtype=torch.int16, device=dev) for _ in range(nb)]   # 9 x 50 MB: rotating > L2, like the planes
mv = torch.randint(1, 6, (4, G, 2), dtype=torch.uint8, device=dev)
res = torch.empty((G,), dtype=torch.uint8, device=dev)
st = torch.cuda.Stream()
K = 90


def timed(fn):
    with torch.cuda.stream(st):
        for k in range(3): fn(k)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for k in range(K): fn(k)
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K)
    return best * 1e3


t_obs = timed(lambda k: upd.observe(bs[k % nb], 4, out=obs[k % nb]))
t_step = timed(lambda k: upd.update(bs[k % nb], mv[k % 4], out=res))
t_both = timed(lambda k: (upd.update(bs[k % nb], mv[k % 4], out=res), upd.observe(bs[k % nb], 4, out=obs[k % nb])))
t_fused = timed(lambda k: upd.update_observe(bs[k % nb], mv[k % 4], stairs_radius=4, out=res, obs_out=obs[k % nb]))
print(f'G={G}: observe {t_obs:.2f} us ({77 * G / t_obs / 1e3:.0f} GB/s of 77 B/game); step {t_step:.2f} us; '
      f'step then observe {t_both:.2f} us; fused step+observe {t_fused:.2f} us ({109 * G / t_fused / 1e3:.0f} GB/s of 109 B/game)')
