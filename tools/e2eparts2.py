"""Where does the e2e step time go with the bit-packed streams? (tuning aid; CUDA events around 100 calls)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from optimax_rogue_b200 import SimConfig, _abi, _lib
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.moves import pack_moves5
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device('cuda')
cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True, path_flags=flags)
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
nb = max(2, min(64, -(-300_000_000 // (32 * G))))
bs = []
for b in range(nb):
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); bs.append(gs)
mv = torch.randint(1, 6, (G, 2), dtype=torch.uint8)
hc = [torch.from_numpy(pack_moves5(mv[:, 0].numpy(), mv[:, 1].numpy())).pin_memory() for _ in range(4)]
hr = [torch.empty((_abi.res2_bytes(G),), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
dc = hc[0].cuda(); dr = torch.empty((_abi.res2_bytes(G),), dtype=torch.uint8, device=dev)
L = _lib.lib()
K = 200
def timeit(fn, sync_each):
    for k in range(5): fn(k); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(K):
        fn(k)
        if sync_each: torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3, (time.perf_counter() - t0) / K * 1e6
def raw(k, cmds, res):
    gs = bs[k % nb]; c, st = upd._cfg(gs)
    rc = L.orx_step_bits(C.byref(c), C.byref(st), cmds.data_ptr(), res.data_ptr(), gs.n, gs.game_id_base, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
print(f'G={G} path_flags={flags} batches={nb}; (events us, wall us) per step')
print('bits: device cmds + device res, no sync:', '%.1f %.1f' % timeit(lambda k: raw(k, dc, dr), False))
print('bits: host cmds   + device res, no sync:', '%.1f %.1f' % timeit(lambda k: raw(k, hc[k % 4], dr), False))
print('bits: device cmds + host res,   no sync:', '%.1f %.1f' % timeit(lambda k: raw(k, dc, hr[k % 2]), False))
print('bits: host cmds   + host res,   no sync:', '%.1f %.1f' % timeit(lambda k: raw(k, hc[k % 4], hr[k % 2]), False))
print('bits: host cmds   + host res,   sync   :', '%.1f %.1f' % timeit(lambda k: raw(k, hc[k % 4], hr[k % 2]), True))
print('bits: device cmds + device res, sync   :', '%.1f %.1f' % timeit(lambda k: raw(k, dc, dr), True))
st = [upd.host_stepper(bs[k % nb], hc[k % 4], hr[k % 2], sync=True, bits=True) for k in range(4 * nb)]
print('host_stepper bits (sync inside)        :', '%.1f %.1f' % timeit(lambda k: st[k % len(st)](), False))
t0 = time.perf_counter()
for k in range(2000): torch.cuda.current_stream().synchronize()
print('empty stream sync: %.2f us' % ((time.perf_counter() - t0) / 2000 * 1e6))
