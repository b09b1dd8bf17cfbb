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

TRACED = os.environ.get('ORX_TRACE_LIB') or os.path.join(ROOT, 'optimax_rogue_b200', 'liborx_trace.so')
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
buf = np.zeros((16, 512, 24), dtype=np.uint64)
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

# which CTAs finish late? slot 7 = SM id | tiles handled << 16 (per launch)
info = t[:, :, 7]
smid = (info & 0xFFFF)
ntl = (info >> 16)
print('last launch: tiles handled per CTA: min %d med %d max %d' % (ntl[15].min(), np.median(ntl[15]), ntl[15].max()))
ends = np.stack([(t[i, :, 4] - t[i - 1, :, 4].max()) / 1e3 for i in range(1, 16)])      # [15, n_cta]
sm_of = smid[1:]
n_sm = int(sm_of.max()) + 1
per_sm = np.zeros(n_sm); cnt = np.zeros(n_sm)
for i in range(15):
    np.add.at(per_sm, sm_of[i], ends[i]); np.add.at(cnt, sm_of[i], 1)
per_sm /= np.maximum(cnt, 1)
o = np.argsort(per_sm)
print('mean CTA end time by SM over 15 launches (us): fastest', ' '.join(f'{k}:{per_sm[k]:.2f}' for k in o[:12]))
print('                                              slowest', ' '.join(f'{k}:{per_sm[k]:.2f}' for k in o[-12:]))
print('per-SM mean end: min %.2f p25 %.2f med %.2f p75 %.2f max %.2f' % (per_sm.min(), np.percentile(per_sm, 25), np.median(per_sm), np.percentile(per_sm, 75), per_sm.max()))
# launch-to-launch consistency: correlation of per-SM end times between two launches
a = np.zeros(n_sm); b = np.zeros(n_sm)
np.add.at(a, sm_of[13], ends[13]); np.add.at(b, sm_of[14], ends[14])
print('correlation of per-SM end-time sums between two launches: %.2f' % np.corrcoef(a, b)[0, 1])
for i in (13, 14):
    e = ends[i]; k = ntl[i + 1]
    for v in np.unique(k):
        print(f'launch {i + 1}: CTAs that handled {v} tiles: {int((k == v).sum())}, end med {np.median(e[k == v]):.2f} max {e[k == v].max():.2f}')

# per-tile completion times of the last launch (slots 8+k), relative to the previous launch's end
i = 15
ref = t[i - 1, :, 4].max()
k = ntl[i]
print('tile k done (producer saw the consumers finish it), us after previous end: k | min med max | n CTAs')
for j in range(12):
    m = k > j
    if m.sum() == 0: break
    v = (t[i, m, 8 + j] - ref) / 1e3
    print(f'  {j:2d} | {v.min():6.2f} {np.median(v):6.2f} {v.max():6.2f} | {int(m.sum())}')
# interval between consecutive tiles, early vs late CTAs
late = ends[14] > np.percentile(ends[14], 80); early = ends[14] < np.percentile(ends[14], 20)
for name, m in (('late 20%', late), ('early 20%', early)):
    d = []
    for j in range(1, 8):
        mm = m & (k > j)
        d.append(np.median((t[i, mm, 8 + j] - t[i, mm, 8 + j - 1]) / 1e3))
    f0 = np.median((t[i, m, 8] - ref) / 1e3)
    print(f'{name}: first tile done at {f0:.2f}; median interval to next tile, tiles 1..7:', ' '.join(f'{x:.2f}' for x in d), '; tiles handled med', np.median(k[m]))

def rel(col):
    v = (t[i, :, col] - ref) / 1e3
    return f'{v.min():.2f}/{np.median(v):.2f}/{v.max():.2f}'
print('last launch, min/med/max us after previous end: wait passed', rel(1), '| prologue loads issued', rel(23), '| first tile landed (warp 0)', rel(2),
      '| first tile ticked (warp 0)', rel(22), '| producer saw tile 0 done', rel(8))
