"""Profiling target: eager launches of the TMA-pipelined kernel in a fixed order, 2^20 games each, on
rotating batches: 3 warm-up ticks, then 2 ticks, 2 observes, 2 fused tick+observe (tools/profile_round.sh
captures launches 4..9 of k_step_pipe with ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
G = 1 << 20
dev = torch.device('cuda')
cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True)
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
bs = []
for b in range(9):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); upd.rollout(gs, 1, 1, 20 + 7 * b); bs.append(gs)
mv = torch.randint(1, 6, (4, G, 2), dtype=torch.uint8, device=dev)
res = torch.empty((G,), dtype=torch.uint8, device=dev)
obs = [torch.empty((G, 2, 12), dtype=torch.int16, device=dev) for _ in range(4)]
k = 0
for _ in range(3 + 2):
    upd.update(bs[k % 9], mv[k % 4], out=res); k += 1
for _ in range(2):
    upd.observe(bs[k % 9], 4, out=obs[k % 4]); k += 1
for _ in range(2):
    upd.update_observe(bs[k % 9], mv[k % 4], stairs_radius=4, out=res, obs_out=obs[k % 4]); k += 1
torch.cuda.synchronize()
print('profile targets done:', k, 'launches')
