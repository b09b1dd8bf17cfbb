"""-m gpu: the CUDA path against the fixtures generated from the LIVE reference
(oracle/gen_golden.py): 2 x 10k replayed episodes, every tick folded into a digest that covers
(x, y, depth, health) x 2, tick, result, the staircases and the ordered event records."""
import json
import os

import numpy as np
import pytest
import torch

from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

import trace_util as tu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


SUITES = sorted(json.load(open(os.path.join(GOLD, 'digests_meta.json')))['suites'])


@pytest.mark.parametrize('suite', SUITES)
@pytest.mark.parametrize('fused_bots', [False, True])
def test_replayed_episodes_bit_exact(suite, fused_bots):
    """Up to 10,000 episodes per suite replayed on the GPU against digests recorded from the live
    reference (StaircaseBot / RandomBot pairings, both despawn strategies, separated starts)."""
    from test_oracle_golden import suite_config
    meta = json.load(open(os.path.join(GOLD, 'digests_meta.json')))
    want = np.load(os.path.join(GOLD, f'digests_{suite}.npy'))
    kw = meta['suites'][suite]
    n = len(want)
    cfg = suite_config(meta, kw)
    gs = BatchedGameState(cfg, n, 'cuda')
    reset_games(gs)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), cfg.despawn_strat, kw['max_ticks'])
    bots = [tu.BOT_CODES[b] for b in kw['bots']]
    dg = tu.BatchDigest(n)
    active = np.ones(n, bool)
    p = gs.planes_cpu()
    dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], p['status'], None, active)
    moves = torch.full((n, 2), 5, dtype=torch.uint8, device='cuda')
    for _ in range(kw['max_ticks']):
        if fused_bots:
            upd.bot_moves(gs, bots[0], bots[1], out=moves)
        else:   # one launch per player, as two independent Bot objects would do
            upd.bot_moves(gs, bots[0], 0, out=moves)
            upd.bot_moves(gs, 0, bots[1], out=moves)
        res, ev = upd.update(gs, moves, want_events=True)
        p = gs.planes_cpu()
        r = res.cpu().numpy()
        dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], r, ev.cpu().numpy(), active)
        active &= r == 1
        if not active.any():
            break
    assert not active.any()
    bad = np.flatnonzero(dg.h != want)
    assert len(bad) == 0, f'{len(bad)} of {n} episodes differ from the reference, first {bad[:5]}'
    hist = {str(k): int((p['status'] == k).sum()) for k in (2, 3, 4)}
    assert hist == kw['result_hist']


def test_truth_table_through_the_cuda_path():
    """The 400-row two-player resolution table recorded from the live reference (8 placements x 25
    command pairs x 2 initiative orders): positions, health and the ordered events."""
    from optimax_rogue_b200 import _abi
    from optimax_rogue_b200.logic.worldgen import FixedDungeonGenerator
    from optimax_rogue_b200.logic.updates import unpack_events
    rows = json.load(open(os.path.join(GOLD, 'truth_table.json')))
    tiles = np.full((9, 9), 1, np.uint8)
    tiles[[0, -1], :] = 2
    tiles[:, [0, -1]] = 2
    cfg = SimConfig(width=9, height=9, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=tiles, seed=0x0A11CE)
    upd = BatchedUpdater(FixedDungeonGenerator(tiles), 1, None)
    by_gid = {}
    for row in rows:
        by_gid.setdefault(row['gid'], []).append(row)
    assert len(by_gid) == 2
    for gid, group in by_gid.items():
        for row in group:
            gs = BatchedGameState(cfg, 1, 'cuda', game_id_base=gid)
            reset_games(gs)
            dx, dy = row['placement']
            gs.pos.copy_(torch.tensor([[4, 4, 4 + dx, 4 + dy]], dtype=torch.uint8))
            mv = torch.tensor([[row['m1'], row['m2']]], dtype=torch.uint8, device='cuda')
            res, ev = upd.update(gs, mv, want_events=True)
            p = gs.planes_cpu()
            ent = [[int(p['pos'][0, 0]), int(p['pos'][0, 1]), int(p['depth'][0, 0]), int(p['hp'][0, 0])],
                   [int(p['pos'][0, 2]), int(p['pos'][0, 3]), int(p['depth'][0, 1]), int(p['hp'][0, 1])]]
            assert ent == row['ent'], row
            assert int(res[0]) == row['result']
            evs = [list(map(int, e)) for e in unpack_events(ev)[0] if e[0] != 0]
            assert evs == row['events'], (row, evs)
