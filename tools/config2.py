"""BASELINE.json configs[1]: 4096 concurrent games, one shared fixed wall map, players only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimax_rogue_b200 import SimConfig, _abi
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import FixedDungeonGenerator
rng = np.random.default_rng(0)
t = np.full((60, 10), 1, np.uint8); t[[0, -1], :] = 2; t[:, [0, -1]] = 2
t[1:-1, 1:-1][rng.random((58, 8)) < 0.10] = 2
for G in (4096, 65536, 1 << 20):
    cfg = SimConfig(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t, max_ticks=0, seed=1, auto_reset=True)
    nb = max(2, min(64, 300_000_000 // (32 * G)))
    bs = []
    for b in range(nb):
        gs = BatchedGameState(cfg, G, 'cuda', game_id_base=b * G); reset_games(gs); bs.append(gs)
    upd = BatchedUpdater(FixedDungeonGenerator(t), 1, None, auto_reset=True)
    mv = torch.randint(1, 6, (8, G, 2), dtype=torch.uint8, device='cuda')
    res = torch.empty((G,), dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    K = 4096 if G == 4096 else 400
    with torch.cuda.stream(st):
        for k in range(4): upd.update(bs[k % nb], mv[k % 8], out=res)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for k in range(K): upd.update(bs[k % nb], mv[k % 8], out=res)
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K)
    print(f'fixed map, G={G}, {nb} batches: {best * 1e3:.2f} us/step, {G / best * 1e3:.3e} ticks/s', flush=True)
