"""Kernel micro-benchmark used while tuning: k_step alone over rotating batches (> L2), CUDA graph,
CUDA events. Not the judged benchmark (that is bench.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, nargs='+', default=[1 << 20])
ap.add_argument('--steps', type=int, default=200)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--rollout', type=int, default=0, help='also time orx_rollout with this many ticks per launch')
ap.add_argument('--json', default=None, help='write the sweep (one row per batch size) to this file')
ap.add_argument('--path-flags', type=int, default=0, help='OrxConfig.path_flags (include/orx.h ORX_PATH_*)')
ap.add_argument('--tpc', type=int, default=0, help='tiles per CTA in tile-flag mode (0 = built-in default)')
ap.add_argument('--overlap', action='store_true', help='throughput mode: SimConfig.overlap_ticks (ORX_PATH_TILE_FLAGS)')
ap.add_argument('--isolate', action='store_true', help='a tiny ordinary kernel between consecutive steps (what a policy network in the loop does: no overlap between ticks)')
ap.add_argument('--batches', type=int, default=0, help='rotating batches (0 = enough to exceed the L2; 1 = the same state every step)')
args = ap.parse_args()
dev = torch.device('cuda')
rows = []
for G in args.games:
    cfg = SimConfig(max_ticks=1000, seed=1, auto_reset=True, path_flags=args.path_flags | (args.tpc << 8), overlap_ticks=args.overlap)
    nb = args.batches or max(2, min(64, -(-300_000_000 // (32 * G))))
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1000, auto_reset=True)
    batches = []
    for b in range(nb):
        gs = BatchedGameState(cfg, G, dev, game_id_base=b * G)
        reset_games(gs)
        upd.rollout(gs, 1, 1, 17 * (b + 1))
        batches.append(gs)
    moves = torch.randint(1, 6, (8, G, 2), dtype=torch.uint8, device=dev)
    res = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb)]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for k in range(4):
            upd.update(batches[k % nb], moves[k % 8], out=res[k % nb])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        dummy = torch.zeros((1,), device=dev)
        with torch.cuda.graph(g, stream=st):
            for k in range(args.steps):
                upd.update(batches[k % nb], moves[k % 8], out=res[k % nb])
                if args.isolate:
                    dummy.add_(1.0)
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / args.steps)
    if args.rollout:
        stats = torch.zeros(8, dtype=torch.int64, device=dev)
        with torch.cuda.stream(st):
            upd.rollout(batches[0], 1, 1, args.rollout, stats)
            torch.cuda.synchronize()
            rbest = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for b in range(min(nb, 4)):
                    upd.rollout(batches[b], 1, 1, args.rollout, stats)
                e1.record(st)
                torch.cuda.synchronize()
                rbest = min(rbest, e0.elapsed_time(e1) / min(nb, 4))
        print(f'games={G} rollout T={args.rollout}: {rbest * 1e3:.1f} us/launch, {G * args.rollout / rbest * 1e3:.3e} ticks/s', flush=True)
    print(f'overlap={int(args.overlap)} path_flags={args.path_flags} tpc={args.tpc} isolate={int(args.isolate)} games={G} batches={nb} us/step={best * 1e3:.2f} ticks/s={G / best * 1e3:.3e} GB/s(61B)={61 * G / best / 1e6:.0f}', flush=True)
    rows.append({'overlap_ticks': bool(args.overlap), 'isolated': bool(args.isolate), 'path_flags': args.path_flags, 'tiles_per_cta': args.tpc, 'games_per_launch': G, 'rotating_batches': nb, 'us_per_step': round(best * 1e3, 2),
                 'game_ticks_per_s': float(f'{G / best * 1e3:.4g}'), 'alg_GBps_61B': round(61 * G / best / 1e6),
                 'frac_of_measured_hbm_peak': round(61 * G / best / 1e6 / 6548.2, 3)})
    del batches, moves, res
    torch.cuda.empty_cache()
if args.json:
    import json
    with open(args.json, 'w') as f:
        json.dump({'what': 'k_step_pipe (orx_step, ruleset R0, auto-reset, uniform random commands), one launch per step in a CUDA graph '
                           f'of {args.steps} steps, best of {args.reps} replays, CUDA events; rotating batches keep the working set above the 126 MB L2',
                   'peak_GBps': 6548.2, 'rows': rows}, f, indent=1)
