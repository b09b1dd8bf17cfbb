"""Pins the C oracle to the LIVE reference (only where /root/reference exists, i.e. the build
container): same injected draws, every tick compared field by field, events included."""
import pytest

from oracle import ref_harness as rh
from optimax_rogue_b200 import SimConfig, _abi

import trace_util as tu

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')

SEED = 0xBEEF


def compare(gid, bots, cfg_kw, ref_kw, npcs=(), episode=0):
    tr, mv = rh.play_episode(SEED, gid, episode, bots=bots, npcs=npcs, **ref_kw)
    cfg = SimConfig(seed=SEED, n_npc=len(npcs), **cfg_kw)
    if episode:
        pytest.skip('episode offsets are covered by the golden traces')
    to, mo = tu.oracle_episode(cfg, gid, bots=bots, npcs=npcs)
    a = [tu.strip(r) for r in tr]
    assert len(a) == len(to)
    for i, (x, y) in enumerate(zip(a, to)):
        assert x == y, f'game {gid} record {i}: reference {x} != oracle {y}'
    assert mv == mo


@pytest.mark.parametrize('bots', [('staircase', 'random'), ('random', 'random'), ('staircase', 'staircase')])
@pytest.mark.parametrize('despawn', ['unreachable', 'unused'])
def test_default_room(bots, despawn):
    for gid in range(100, 106):
        compare(gid, bots, dict(max_ticks=200, despawn_strat=1 if despawn == 'unreachable' else 2),
                dict(max_ticks=200, despawn=despawn))


def test_small_rooms_and_stats():
    for gid, (w, h) in enumerate([(4, 4), (5, 5), (4, 9), (7, 4)]):
        compare(gid, ('random', 'random'),
                dict(width=w, height=h, max_ticks=150, hp=(3, 5), damage=(3, 2), armor=(1, 0)),
                dict(width=w, height=h, max_ticks=150, hp=(3, 5), damage=(3, 2), armor=(1, 0)))


def test_separated_start():
    for gid in range(4):
        compare(gid, ('staircase', 'staircase'),
                dict(start_kind=_abi.START_SEPARATED, start_depth=(0, 3), max_ticks=200, despawn_strat=1 + gid % 2),
                dict(start='separated', p_depths=(0, 3), max_ticks=200, despawn=('unreachable', 'unused')[gid % 2]))


def test_fixed_map_plugin_generator():
    from oracle.gen_golden import fixed_map
    for gid, stairs in enumerate([False, True, True]):
        tiles = fixed_map(stairs=stairs, seed=gid)
        compare(gid, ('staircase' if stairs else 'random', 'random'),
                dict(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=tiles, max_ticks=200),
                dict(fixed_tiles=tiles.astype('int32'), max_ticks=200))


def test_npc_slots():
    from oracle import cport
    base = [(0, 3, 3, 2), (0, 5, 2, 1), (1, 2, 2, 3), (0, 2, 4, 1)]
    for gid in range(6):
        # an NPC dropped onto a player's spawn tile would corrupt the reference's pos_lookup
        # (add_entity overwrites the player's key, state.py:78-82): keep the setup legal
        orc = cport.Oracle(SimConfig(seed=SEED, width=8, height=6, n_npc=4), 1, gid)
        orc.reset()
        taken = {(0, int(orc.state.pos[0, 0]), int(orc.state.pos[0, 1])), (0, int(orc.state.pos[0, 2]), int(orc.state.pos[0, 3]))}
        npcs = []
        for (d, x, y, hp) in base:
            while (d, x, y) in taken:
                x = x % 6 + 1
            taken.add((d, x, y))
            npcs.append((d, x, y, hp))
        compare(gid, ('random', 'staircase'), dict(width=8, height=6, max_ticks=150, hp=(50, 50)),
                dict(width=8, height=6, max_ticks=150, hp=50), npcs=npcs)


def test_flat_modifier_bonuses():
    """The Modifier seam (game/modifiers.py:102-108 -> game/attribles.py:21-43 -> updater.py:313): players that carry
    a modifier with flat damage / armor / max-health bonuses, in rooms small enough to fight all the time. The oracle
    takes the same sums through OrxState.flat."""
    cases = [((3, 0, 2), (0, 1, 0)), ((0, 0, 0), (5, 2, -3)), ((-1, 0, 0), (0, -2, 7)), ((127, 0, 0), (0, 127, -128))]
    n_hits = 0
    for gid, flat in enumerate(cases):
        kw = dict(width=5, height=5, max_ticks=300, hp=(40, 35), damage=(2, 3), armor=(1, 1))
        tr, mv = rh.play_episode(SEED, gid, bots=('random', 'random'), flat=flat, **kw)
        to, mo = tu.oracle_episode(SimConfig(seed=SEED, **kw), gid, bots=('random', 'random'), flat=flat)
        assert [tu.strip(r) for r in tr] == to, flat
        assert mv == mo
        hits = [e for r in tr for e in r['events'] if e[0] == rh.EV_COMBAT]
        want = {1: 2 + flat[0][0] - 1 - flat[0][1], 2: 3 + flat[1][0] - 1 - flat[1][1]}
        assert all(e[4] == want[e[1]] for e in hits)
        n_hits += len(hits)
    assert n_hits > 3


from hypothesis import given, settings, strategies as st


@settings(max_examples=40, deadline=None)
@given(gid=st.integers(0, 2**53), seed=st.integers(0, 2**64 - 1), w=st.integers(4, 30), h=st.integers(4, 12),
       hp1=st.integers(1, 12), hp2=st.integers(1, 12), dmg=st.integers(0, 4), arm=st.integers(0, 3),
       despawn=st.sampled_from([1, 2]), separated=st.booleans(), d2=st.integers(1, 5),
       bots=st.sampled_from([('random', 'random'), ('staircase', 'random'), ('random', 'staircase'), ('staircase', 'staircase')]))
def test_fuzzed_configurations_against_live_reference(gid, seed, w, h, hp1, hp2, dmg, arm, despawn, separated, d2, bots):
    """Random room sizes, stats, start generators, despawn strategies, bot pairings, 64-bit seeds and
    53-bit game ids: the C oracle must reproduce the live reference tick by tick, events included."""
    ref_kw = dict(width=w, height=h, max_ticks=90, hp=(hp1, hp2), damage=(dmg, dmg + 1), armor=(arm, 0),
                  despawn='unreachable' if despawn == 1 else 'unused')
    cfg_kw = dict(width=w, height=h, max_ticks=90, hp=(hp1, hp2), damage=(dmg, dmg + 1), armor=(arm, 0),
                  despawn_strat=despawn)
    if separated:
        ref_kw.update(start='separated', p_depths=(0, d2))
        cfg_kw.update(start_kind=_abi.START_SEPARATED, start_depth=(0, d2))
    tr, mv = rh.play_episode(seed, gid, bots=bots, **ref_kw)
    to, mo = tu.oracle_episode(SimConfig(seed=seed, **cfg_kw), gid, bots=bots)
    assert [tu.strip(r) for r in tr] == to
    assert mv == mo
