"""Where does the e2e step time go? (tuning aid)"""
import os, sys, time
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
    gs = BatchedGameState(cfg, G, dev, game_id_base=b * G); reset_games(gs); bs.append(gs)
hm = [torch.randint(1, 6, (G, 2), dtype=torch.uint8).pin_memory() for _ in range(4)]
hr = torch.empty((G,), dtype=torch.uint8, pin_memory=True)
dm = hm[0].cuda(); dr = torch.empty((G,), dtype=torch.uint8, device=dev)
K = 100
def timeit(fn, sync_each):
    for k in range(5): fn(k); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        fn(k)
        if sync_each: torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
print('host moves+result, sync each : %.1f us' % timeit(lambda k: upd.update(bs[k % 9], hm[k % 4], out=hr), True))
print('host moves+result, no sync   : %.1f us' % timeit(lambda k: upd.update(bs[k % 9], hm[k % 4], out=hr), False))
print('device moves+result, sync each: %.1f us' % timeit(lambda k: upd.update(bs[k % 9], dm, out=dr), True))
print('device moves+result, no sync  : %.1f us' % timeit(lambda k: upd.update(bs[k % 9], dm, out=dr), False))
t0 = time.perf_counter()
for k in range(1000): torch.cuda.current_stream().synchronize()
print('empty stream sync: %.2f us' % ((time.perf_counter() - t0) / 1000 * 1e6))
# raw PCIe: pinned copies
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K): dm.copy_(hm[k % 4], non_blocking=True)
e1.record(); torch.cuda.synchronize(); print('H2D 2 MB memcpy: %.1f us' % (e0.elapsed_time(e1) / K * 1e3))
e0.record()
for k in range(K): hr.copy_(dr, non_blocking=True)
e1.record(); torch.cuda.synchronize(); print('D2H 1 MB memcpy: %.1f us' % (e0.elapsed_time(e1) / K * 1e3))
st1 = [upd.host_stepper(bs[k % 9], hm[k % 4], hr) for k in range(36)]
def run(k): st1[k % 36]()
print('host_stepper (sync inside)    : %.1f us' % timeit(run, False))
# nibble-packed commands, each direction on its own
from optimax_rogue_b200.logic.moves import pack_moves
hc = [pack_moves(m[:, 0], m[:, 1]).contiguous().pin_memory() for m in hm]
dc = hc[0].cuda()
print('packed: host cmds + host result, no sync : %.1f us' % timeit(lambda k: upd.update(bs[k % 9], hc[k % 4], out=hr, packed=True), False))
print('packed: host cmds + host result, sync    : %.1f us' % timeit(lambda k: upd.update(bs[k % 9], hc[k % 4], out=hr, packed=True), True))
print('packed: device cmds + device result, no sync: %.1f us' % timeit(lambda k: upd.update(bs[k % 9], dc, out=dr, packed=True), False))
# mixed: call the C ABI directly with one side mapped
import ctypes as C
from optimax_rogue_b200 import _lib
L = _lib.lib()
def raw(k, cmds, res, packed):
    gs = bs[k % 9]; cfg, st = upd._cfg(gs)
    fn = L.orx_step_packed if packed else L.orx_step
    rc = fn(C.byref(cfg), C.byref(st), cmds.data_ptr(), res.data_ptr(), None, gs.n, gs.game_id_base, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
print('packed: host cmds + device result, no sync: %.1f us' % timeit(lambda k: raw(k, hc[k % 4], dr, True), False))
print('packed: device cmds + host result, no sync: %.1f us' % timeit(lambda k: raw(k, dc, hr, True), False))
print('bytes : host cmds + device result, no sync: %.1f us' % timeit(lambda k: raw(k, hm[k % 4], dr, False), False))
print('bytes : device cmds + host result, no sync: %.1f us' % timeit(lambda k: raw(k, dm, hr, False), False))
st2 = [upd.host_stepper(bs[k % 9], hc[k % 4], hr) for k in range(36)]
print('host_stepper packed (sync inside): %.1f us' % timeit(lambda k: st2[k % 36](), False))
st3 = [upd.host_stepper(bs[k % 9], hc[k % 4], hr, sync=False) for k in range(36)]
print('host_stepper packed (async)      : %.1f us' % timeit(lambda k: st3[k % 36](), False))
