"""-m gpu: the CUDA path against the LIVE reference in ONE process (no fixture, no C oracle in between).

The unmodified reference modules (``optimax_rogue.logic.updater.Updater`` & co.) are imported from
``/root/reference`` in the build container, or from the verbatim copy ``oracle/make_ref.py`` ships to the GPU
box (``oracle/_ref``, git-ignored); ``oracle/ref_harness`` injects the shared Philox draws into the reference's
``random`` / ``numpy.random`` calls. Each episode is played by the reference, tick by tick, and by the CUDA
kernels (bots on device, one lane per episode); position, depth, health of both players, tick, result, the
staircases of the occupied levels and the ordered ``GameStateUpdate`` records are folded into one digest per
episode on each side and must be equal. Skipped when no reference tree is present."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from optimax_rogue_b200 import SimConfig, _abi
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

import trace_util as tu

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rh.reference_available(), reason='no reference tree (run oracle/make_ref.py where /root/reference exists)')]

SEED = 0x5EED1E55


@pytest.mark.parametrize('bots,despawn,start,episodes,max_ticks', [
    (('staircase', 'random'), 'unreachable', 'together', 512, 96),
    (('random', 'random'), 'unreachable', 'together', 384, 128),
    (('staircase', 'staircase'), 'unused', 'together', 256, 64),
    (('random', 'staircase'), 'unreachable', 'separated', 128, 64),
])
def test_cuda_equals_live_reference(bots, despawn, start, episodes, max_ticks):
    gid0 = 7_000_000
    # --- the reference, one episode at a time (updater.py:76-162 under injected draws)
    want = np.empty(episodes, np.uint64)
    ticks = 0
    for k in range(episodes):
        trace, _ = rh.play_episode(SEED, gid0 + k, bots=bots, despawn=despawn, start=start, p_depths=(0, 6),
                                   max_ticks=max_ticks)
        want[k] = rh.digest(trace)
        ticks += len(trace) - 1
    # --- CUDA, all episodes at once
    cfg = SimConfig(seed=SEED, max_ticks=max_ticks, despawn_strat=1 if despawn == 'unreachable' else 2,
                    start_kind=_abi.START_SEPARATED if start == 'separated' else _abi.START_TOGETHER,
                    start_depth=(0, 6) if start == 'separated' else (0, 0))
    gs = BatchedGameState(cfg, episodes, 'cuda', game_id_base=gid0)
    reset_games(gs)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), cfg.despawn_strat, max_ticks)
    codes = [tu.BOT_CODES[b] for b in bots]
    dg = tu.BatchDigest(episodes)
    active = np.ones(episodes, bool)
    p = gs.planes_cpu()
    dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], p['status'], None, active)
    moves = torch.full((episodes, 2), 5, dtype=torch.uint8, device='cuda')
    gpu_ticks = 0
    for _ in range(max_ticks):
        upd.bot_moves(gs, codes[0], codes[1], out=moves)
        res, ev = upd.update(gs, moves, want_events=True)
        p = gs.planes_cpu()
        r = res.cpu().numpy()
        dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], r, ev.cpu().numpy(), active)
        gpu_ticks += int(active.sum())
        active &= r == 1
        if not active.any():
            break
    assert not active.any()
    assert gpu_ticks == ticks
    bad = np.flatnonzero(dg.h != want)
    assert len(bad) == 0, f'{len(bad)} of {episodes} episodes differ from the live reference, first {bad[:5]}'
