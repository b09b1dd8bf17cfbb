"""Profiling target (round 2): eager launches of the plain tick at a given batch size on rotating batches.
    python tools/profile_targets2.py <games> [<launches>] [<path_flags>]
tools/profile_round2.sh captures k_step_pipe launches of it with ncu --set full (2^20: grid-wait mode, 2^17: flag mode,
2^24: DRAM read + write bytes per game with planes four times the L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 8
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device('cuda')
cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True, path_flags=flags)
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
nb = max(2, min(16, -(-300_000_000 // (32 * G))))
bs = []
for b in range(nb):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); upd.rollout(gs, 1, 1, 20 + 7 * b); bs.append(gs)
mv = torch.randint(1, 6, (2, G, 2), dtype=torch.uint8, device=dev)
res = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb)]
for k in range(launches):
    upd.update(bs[k % nb], mv[k % 2], out=res[k % nb])
torch.cuda.synchronize()
print('profile targets done:', launches, 'launches of', G, 'games')
