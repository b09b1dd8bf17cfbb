"""Tuning aid: where does a tick launch spend its time? Builds a traced variant of liborx
(-DORX_PIPE_TRACE: per-CTA %globaltimer stamps), replays the bench's CUDA graph and prints, for the
last launches, the ramp / body / drain of the persistent TMA kernel relative to its neighbours.

  python tools/pipetrace.py [games]     (GPU box)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optimax_rogue_b200 import build as B          # noqa: E402

TRACED = os.path.join(ROOT, 'optimax_rogue_b200', 'liborx_trace.so')
if not os.path.exists(TRACED) or '--rebuild' in sys.argv:
    extra = [a for a in sys.argv[1:] if a.startswith('-D')]
    cmd = ['nvcc'] + B.NVCC_FLAGS + ['-DORX_PIPE_TRACE'] + extra + ['-o', TRACED] + B.SOURCES
    subprocess.run(cmd, check=True)
if '--build-only' in sys.argv:
    sys.exit(0)
os.environ['ORX_LIB'] = TRACED

import numpy as np                                  # noqa: E402
import torch                                        # noqa: E402
from optimax_rogue_b200 import SimConfig, _lib      # noqa: E402
from optimax_rogue_b200.game.state import BatchedGameState   # noqa: E402
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games   # noqa: E402
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator        # noqa: E402

nums = [a for a in sys.argv[1:] if a.isdigit()]
G = int(nums[0]) if nums else 1 << 20
dev = torch.device('cuda')
cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True)
nb = max(2, min(64, -(-300_000_000 // (32 * G))))
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
batches = []
for b in range(nb):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G)
    reset_games(gs)
    upd.rollout(gs, 1, 1, 17 * (b + 1))
    batches.append(gs)
moves = torch.randint(1, 6, (8, G, 2), dtype=torch.uint8, device=dev)
res = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb)]
K = 64
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for k in range(4):
        upd.update(batches[k % nb], moves[k % 8], out=res[k % nb])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for k in range(K):
            upd.update(batches[k % nb], moves[k % 8], out=res[k % nb])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); g.replay(); e1.record(st)
    torch.cuda.synchronize()
print(f'games {G}: {e0.elapsed_time(e1) / K * 1e3:.2f} us per step (traced build)')
lib = _lib.lib()
buf = np.zeros((16, 512, 8), dtype=np.uint64)
lib.orx_debug_trace.restype = C.c_int
lib.orx_debug_trace.argtypes = [C.c_void_p]
assert lib.orx_debug_trace(buf.ctypes.data) == 0
t = buf.astype(np.int64)
n_cta = int((t[0, :, 0] != 0).sum())
order = np.argsort(t[:, 0, 0])                  # launches by start time of CTA 0
t = t[order][:, :n_cta]
print(f'{n_cta} CTAs; stamps: 0 entry, 1 producer past griddepcontrol.wait, 2 first tile landed, '
      f'5 last tile landed, 6 consumers done, 3 last stores read out, 4 last stores complete')
print('per launch, microseconds relative to the previous launch\'s last store completion (col "prev end" = 0):')
hdr = 'launch | entry min/med/max | wait-done min/med/max | 1st tile min/med/max | last tile landed med/max | stores complete min/med/max | span'
print(hdr)
for i in range(1, 16):
    ref = t[i - 1, :, 4].max()
    def mmm(col, i=i, ref=ref):
        v = (t[i, :, col] - ref) / 1e3
        return f'{v.min():6.2f}/{np.median(v):6.2f}/{v.max():6.2f}'
    v5 = (t[i, :, 5] - ref) / 1e3
    span = (t[i, :, 4].max() - t[i - 1, :, 4].max()) / 1e3
    print(f'{i:3d} | {mmm(0)} | {mmm(1)} | {mmm(2)} | {np.median(v5):6.2f}/{v5.max():6.2f} | {mmm(4)} | {span:6.2f}')
# distribution of CTA finishing times of the last launch, by tiles owned
i = 15
ref = t[i - 1, :, 4].max()
end = (t[i, :, 4] - ref) / 1e3
n_tiles = G // 256
owned = np.array([(n_tiles - b + n_cta - 1) // n_cta for b in range(n_cta)])
for k in np.unique(owned):
    e = end[owned == k]
    print(f'CTAs with {k} tiles: {len(e)}; end min {e.min():.2f} med {np.median(e):.2f} max {e.max():.2f} us')
