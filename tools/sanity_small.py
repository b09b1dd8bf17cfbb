"""Small exercise of every kernel, for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimax_rogue_b200 import SimConfig, _abi
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator, FixedDungeonGenerator
from optimax_rogue_b200.r1 import R1GameState
for n in (1000, 4096 + 77):
    cfg = SimConfig(max_ticks=30, seed=3, auto_reset=True, hp=(2, 2))
    gs = BatchedGameState(cfg, n, 'cuda'); reset_games(gs)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 30, auto_reset=True)
    for t in range(40):
        mv = upd.bot_moves(gs, 2, 1)
        upd.update(gs, mv, want_events=(t % 2 == 0))
    upd.rollout(gs, 1, 1, 50)
    upd.observe(gs, 3)
    hm = torch.full((n, 2), 2, dtype=torch.uint8).pin_memory(); hr = torch.empty((n,), dtype=torch.uint8).pin_memory()
    st = upd.host_stepper(gs, hm, hr)
    for _ in range(5): st()
t = np.full((20, 9), 1, np.uint8); t[[0, -1], :] = 2; t[:, [0, -1]] = 2; t[5, 4] = 2; t[10, 3] = 3
cfg = SimConfig(width=20, height=9, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t, max_ticks=50, seed=1, auto_reset=True, n_npc=2)
gs = BatchedGameState(cfg, 700, 'cuda'); reset_games(gs)
gs.set_npc(0, 0, 0, 3, 3, 2)
upd = BatchedUpdater(FixedDungeonGenerator(t), 2, 50, auto_reset=True)
for _ in range(30):
    upd.update(gs, upd.bot_moves(gs, 1, 2), want_events=True)
cfg = SimConfig(width=20, height=9, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t, max_ticks=50, seed=1, auto_reset=True)
gs = BatchedGameState(cfg, 2000, 'cuda'); reset_games(gs)
for _ in range(30):
    upd.update(gs, upd.bot_moves(gs, 1, 2))
upd.rollout(gs, 2, 2, 40)
r1 = R1GameState(3000, max_ticks=100, auto_reset=True, seed=2).reset()
mv = torch.randint(0, 8, (3000, 2), dtype=torch.uint8, device='cuda')
for _ in range(40): r1.update(mv)
r1.rollout(60)
torch.cuda.synchronize()
print('sanity_small ok')
