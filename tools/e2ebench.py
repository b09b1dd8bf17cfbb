"""e2e micro-benchmark: BatchedUpdater.update with pinned host buffers, sync every step."""
import os, sys, time
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
batches = []
for b in range(9):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); batches.append(gs)
hm = [torch.randint(1, 6, (G, 2), dtype=torch.uint8).pin_memory() for _ in range(4)]
hr = torch.empty((G,), dtype=torch.uint8, pin_memory=True)
for k in range(5):
    upd.update(batches[k % 9], hm[k % 4], out=hr); torch.cuda.synchronize()
K = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for k in range(K):
    upd.update(batches[k % 9], hm[k % 4], out=hr)
    torch.cuda.current_stream().synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'staged={os.environ.get("ORX_HOST_STAGED")} G={G}: {ms / K * 1e3:.1f} us/step (events), {(time.perf_counter() - t0) / K * 1e6:.1f} us/step (wall), {G * K / ms * 1e3:.3e} ticks/s', flush=True)
