"""Profiling target: eager launches of the ruleset-R1 tick (parity unpinned) on rotating batches.
    python tools/profile_targets_r1.py <games> [<launches>] [<path_flags>]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200.r1 import R1GameState
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 8
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nb = 6
bs = [R1GameState(G, max_ticks=1000, auto_reset=True, seed=3, game_id_base=b * G, path_flags=flags).reset() for b in range(nb)]
for b in bs: b.rollout(150)          # populated levels: enemies, items, players on different depths
mv = torch.randint(1, 7, (2, G, 2), dtype=torch.uint8, device='cuda')
res = [torch.empty((G,), dtype=torch.uint8, device='cuda') for _ in range(nb)]
for k in range(launches):
    bs[k % nb].update(mv[k % 2], out=res[k % nb])
torch.cuda.synchronize()
print('profile targets done:', launches, 'R1 launches of', G, 'games')
