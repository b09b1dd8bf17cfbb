"""Tuning aid: host cost of one eager tick call (small batch, so the GPU is never the limit)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

dev = torch.device('cuda')
for G in (4096, 1 << 20):
    cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
    gs = BatchedGameState(cfg, G, dev)
    reset_games(gs)
    mv = torch.randint(1, 6, (G, 2), dtype=torch.uint8, device=dev)
    res = torch.empty((G,), dtype=torch.uint8, device=dev)
    obs = torch.empty((G, 2, 12), dtype=torch.int16, device=dev)
    step = upd.device_stepper(gs, mv, res)
    step_obs = upd.device_stepper(gs, mv, res, obs=obs, stairs_radius=4)
    step_bot = upd.device_stepper(gs, mv, res, bots=(0, 1))
    K = 3000
    for name, f in (('update(gs, mv, out=res)', lambda: upd.update(gs, mv, out=res)), ('device_stepper', step),
                    ('device_stepper + obs', step_obs), ('device_stepper, p2 = RandomBot', step_bot)):
        for _ in range(50):
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            f()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f'G={G} {name}: {1e6 * (t1 - t0) / K:.2f} us of host time per call, {1e6 * (t2 - t0) / K:.2f} us per tick including the GPU', flush=True)
