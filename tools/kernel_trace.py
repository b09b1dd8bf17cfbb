"""Unserialised per-kernel durations of the bench's value leg (nsys is not in the image: torch.profiler's CUPTI
activity trace does the same job). A K-step CUDA graph over rotating batches is replayed under the profiler; every
kernel instance's start and duration go to a CSV, the summary (median / mean per kernel name, and the step period
= spacing of consecutive tick-kernel starts) to stdout.
    python tools/kernel_trace.py <games> <out.csv> [<steps>]"""
import csv, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
G = int(sys.argv[1]); out = sys.argv[2]; K = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device('cuda')
cfg = SimConfig(max_ticks=1000, seed=0x0A11CE, auto_reset=True, overlap_ticks=True)      # as bench.py's value leg
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
nb = max(2, min(64, -(-300_000_000 // (32 * G))))
bs = []
for b in range(nb):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); upd.rollout(gs, 1, 1, 37 * (b + 1) % 997 + 1); bs.append(gs)
mv = torch.randint(1, 6, (16, G, 2), dtype=torch.uint8, device=dev)
res = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb)]
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for k in range(4): upd.update(bs[k % nb], mv[k % 16], out=res[k % nb])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for k in range(K): upd.update(bs[k % nb], mv[k % 16], out=res[k % nb])
    g.replay(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay(); g.replay()
        torch.cuda.synchronize()
rows = [(e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: r[1])
with open(out, 'w', newline='') as f:
    w = csv.writer(f); w.writerow(['kernel', 'start_us', 'duration_us'])
    t0 = rows[0][1] if rows else 0
    for name, s, d in rows: w.writerow([name[:120], f'{s - t0:.3f}', f'{d:.3f}'])
by = {}
for name, s, d in rows: by.setdefault(name[:60], []).append(d)
for name, ds in by.items():
    print(f'{len(ds):5d} x {name}: median {statistics.median(ds):.2f} us, mean {statistics.mean(ds):.2f} us')
ticks = [s for name, s, d in rows if 'k_step_pipe' in name]
if len(ticks) > 2:
    gaps = [b - a for a, b in zip(ticks, ticks[1:])]
    print(f'games={G}: tick-kernel start-to-start period: median {statistics.median(gaps):.2f} us, mean {statistics.mean(gaps[: K - 1]):.2f} us over {len(gaps)} gaps (kernels of consecutive steps may overlap)')
