"""Tuning aid: orx_step with the event log (want_events=True) and with NPC slots, against the plain tick."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
OVERLAP = len(sys.argv) > 2 and sys.argv[2] == 'overlap'      # throughput mode (SimConfig.overlap_ticks)
dev = torch.device('cuda')


def run(n_npc, events, track=True, steps=60, reps=5):
    cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True, n_npc=n_npc, overlap_ticks=OVERLAP)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
    upd.track_order = track
    nb = 9
    batches = []
    for b in range(nb):
        gs = BatchedGameState(cfg, G, dev, game_id_base=b * G)
        reset_games(gs)
        upd.rollout(gs, 1, 1, 17 * (b + 1))
        if n_npc:      # scatter live NPCs over the first two levels (70% of the slots), as tests/test_gpu_parity.py does
            gen = torch.Generator(device=dev).manual_seed(b)
            live = torch.rand((G, n_npc), device=dev, generator=gen) < 0.7
            gs.npc_depth.copy_(torch.where(live, torch.randint(0, 2, (G, n_npc), device=dev, generator=gen), torch.full((G, n_npc), -1, device=dev)).to(torch.int32))
            gs.npc_pos[:, :, 0] = torch.randint(1, 59, (G, n_npc), device=dev, generator=gen).to(torch.uint8)
            gs.npc_pos[:, :, 1] = torch.randint(1, 9, (G, n_npc), device=dev, generator=gen).to(torch.uint8)
            gs.npc_hp.copy_(torch.randint(1, 7, (G, n_npc), device=dev, generator=gen).to(torch.int16))
        batches.append(gs)
    moves = torch.randint(1, 6, (8, G, 2), dtype=torch.uint8, device=dev)
    res = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb)]
    best = 1e9
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for k in range(4):
            upd.update(batches[k % nb], moves[k % 8], out=res[k % nb], want_events=events)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()          # graph replay: kernel time without the Python call overhead
        with torch.cuda.graph(g, stream=st):
            for k in range(steps):
                upd.update(batches[k % nb], moves[k % 8], out=res[k % nb], want_events=events)
        g.replay()
        torch.cuda.synchronize()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return best * 1e3


for n_npc, ev, track in ((0, False, False), (0, True, False), (0, True, True), (1, False, False), (2, False, False), (4, False, False), (8, False, False), (2, True, False)):
    us = run(n_npc, ev, track)
    print(f'G={G} overlap={int(OVERLAP)} npc={n_npc} events={ev} track_order={track}: {us:.1f} us/step (CUDA graph), {G / us * 1e6:.3e} ticks/s', flush=True)
