"""Pins the C oracle to the committed fixtures generated from the live reference
(oracle/gen_golden.py): full traces, the two-player truth table and the 2 x 10k episode digests."""
import json
import os

import numpy as np
import pytest

from oracle import cport
from oracle import ref_harness as rh
from optimax_rogue_b200 import SimConfig, _abi

import trace_util as tu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SEED = 0x0A11CE


def fixed_map(stairs):
    from oracle.gen_golden import fixed_map as fm
    return fm(stairs=stairs)


def cfg_from_case(case):
    kw = {}
    for k in ('width', 'height', 'max_ticks'):
        if k in case:
            kw[k] = case[k]
    if case.get('despawn') == 'unused':
        kw['despawn_strat'] = 2
    if case.get('start') == 'separated':
        kw['start_kind'] = _abi.START_SEPARATED
        kw['start_depth'] = tuple(case['p_depths'])
    for k in ('hp', 'damage', 'armor'):
        if k in case:
            v = case[k]
            kw[k] = tuple(v) if isinstance(v, (list, tuple)) else (v, v)
    if 'fixed' in case:
        kw['dgen_kind'] = _abi.DGEN_FIXED
        kw['fixed_tiles'] = fixed_map(case['fixed'] == 'walls_stairs')
    kw['n_npc'] = len(case.get('npcs', ()))
    return SimConfig(seed=SEED, **kw)


def test_full_traces_match_reference_fixtures():
    cases = json.load(open(os.path.join(GOLD, 'cases.json')))
    data = np.load(os.path.join(GOLD, 'traces.npz'))
    for k, case in enumerate(cases):
        cfg = cfg_from_case(case)
        npcs = [tuple(x) for x in case.get('npcs', ())]
        orc = cport.Oracle(cfg, 1, game_id_base=case['game_id'])
        orc.state.episode[:] = case.get('episode', 0)
        orc.reset()
        s = orc.state
        for j, (nd, nx, ny, nhp) in enumerate(npcs):
            s.npc_depth[0, j], s.npc_hp[0, j] = nd, nhp
            s.npc_pos[0, j] = (nx, ny)
        want = data[f'case{k}_records']
        moves = data[f'case{k}_moves']
        max_ev = 4 + len(npcs)
        rec0 = tu.records_from_planes(s.pos, s.hp, s.depth, s.stairs, s.tick, s.status, None, 0)
        assert rh.record_values(rec0, max_ev) == want[0].tolist(), f'case {k} reset'
        bots = [tu.BOT_CODES[b] for b in case['bots']]
        for t in range(len(moves)):
            mv = orc.bot_moves(bots[0], bots[1])
            assert mv[0].tolist() == moves[t].tolist(), f'case {k} tick {t}: bot moves'
            res, ev = orc.step(mv, want_events=True)
            rec = tu.records_from_planes(s.pos, s.hp, s.depth, s.stairs, s.tick, res, ev, 0)
            assert rh.record_values(rec, max_ev) == want[t + 1].tolist(), f'case {k} tick {t}'


def test_truth_table():
    rows = json.load(open(os.path.join(GOLD, 'truth_table.json')))
    assert len(rows) == 8 * 25 * 2
    tiles = np.full((9, 9), 1, np.uint8)
    tiles[[0, -1], :] = 2
    tiles[:, [0, -1]] = 2
    cfg = SimConfig(width=9, height=9, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=tiles, seed=SEED)
    flags = set()
    for row in rows:
        dx, dy = row['placement']
        tr, _ = tu.oracle_episode(cfg, row['gid'], bots=('script', 'script'), scripts=[[row['m1']], [row['m2']]],
                                  limit_ticks=1, place=((4, 4), (4 + dx, 4 + dy)))
        r = tr[1]
        assert [list(e) for e in r['ent']] == row['ent'], row
        assert [list(e) for e in r['events']] == row['events'], row
        for e in row['events']:
            if e[0] == 2:
                flags.add(e[3])
    assert flags == {1, 2, 3}      # Block, Ambush, Flee; Parry is unreachable


def batch_digests(cfg, bots, n, max_steps, chunk=2000):
    out = np.zeros(n, np.uint64)
    for base in range(0, n, chunk):
        cnt = min(chunk, n - base)
        orc = cport.Oracle(cfg, cnt, game_id_base=base)
        orc.reset()
        s = orc.state
        dg = tu.BatchDigest(cnt)
        active = np.ones(cnt, bool)
        dg.update(s.pos, s.hp, s.depth, s.stairs, s.tick, s.status, None, active)
        for _ in range(max_steps):
            mv = orc.bot_moves(bots[0], bots[1])
            res, ev = orc.step(mv, want_events=True)
            dg.update(s.pos, s.hp, s.depth, s.stairs, s.tick, res, ev, active)
            active &= res == 1
            if not active.any():
                break
        assert not active.any()
        out[base:base + cnt] = dg.h
    return out


SUITES = sorted(json.load(open(os.path.join(GOLD, 'digests_meta.json')))['suites'])


def suite_config(meta, kw):
    extra = {}
    if kw.get('despawn') == 'unused':
        extra['despawn_strat'] = 2
    if kw.get('start') == 'separated':
        extra['start_kind'] = _abi.START_SEPARATED
        extra['start_depth'] = tuple(kw['p_depths'])
    return SimConfig(max_ticks=kw['max_ticks'], seed=meta['seed'], **extra)


@pytest.mark.parametrize('suite', SUITES)
def test_replayed_episode_digests(suite):
    meta = json.load(open(os.path.join(GOLD, 'digests_meta.json')))
    want = np.load(os.path.join(GOLD, f'digests_{suite}.npy'))
    kw = meta['suites'][suite]
    cfg = suite_config(meta, kw)
    bots = [tu.BOT_CODES[b] for b in kw['bots']]
    got = batch_digests(cfg, bots, len(want), kw['max_ticks'])
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, f'{len(bad)} of {len(want)} episodes differ, first {bad[:5]}'
