"""The observation records against the LIVE reference (build container / oracle/_ref): what GameState.view_for
(optimax_rogue/game/state.py:53-58) shows each player -- the entities on the viewer's depth, that level alone, the tick --
must be exactly what the numpy restatements of orx_observe / orx_observe_npc (tests/obs_util.py) encode; the GPU tests
compare the CUDA kernels with the same restatements."""
import numpy as np
import pytest

from oracle import ref_harness as rh
from obs_util import expected_obs, expected_npc_obs

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')


def planes_of(rec, n_npc):
    (x1, y1, d1, h1), (x2, y2, d2, h2) = rec['ent']
    p = {'pos': np.array([[x1, y1, x2, y2]], np.uint8), 'hp': np.array([[h1, h2]], np.int16), 'depth': np.array([[d1, d2]], np.int32),
         'stairs': np.array([[rec['stairs'][0][0], rec['stairs'][0][1], rec['stairs'][1][0], rec['stairs'][1][1]]], np.uint8),
         'tick': np.array([rec['tick']], np.int32),
         'npc_depth': np.full((1, n_npc), -1, np.int32), 'npc_hp': np.zeros((1, n_npc), np.int16), 'npc_pos': np.zeros((1, n_npc, 2), np.uint8)}
    for iden, d, x, y, hp in rec['npcs']:                 # slot k carries iden 3 + k; a culled NPC leaves its slot empty
        p['npc_depth'][0, iden - 3], p['npc_hp'][0, iden - 3], p['npc_pos'][0, iden - 3] = d, hp, (x, y)
    return p


def check_views(trace, n_npc):
    seen_apart = seen_npc = 0
    for rec in trace:
        p = planes_of(rec, n_npc)
        obs = expected_obs(p, -1)[0]
        npc = expected_npc_obs(p)[0] if n_npc else None
        for pl, view in enumerate(rec['views']):
            ents = {iden: (d, x, y, hp) for iden, d, x, y, hp in view['ents']}
            me, other = ents[pl + 1], ents.get(2 - pl)
            assert view['levels'] == [me[0]]                                   # the viewer's level alone
            want = [me[1], me[2], me[0], me[3]]
            want += [1, other[1], other[2], other[3]] if other is not None else [0, -1, -1, 0]
            sx, sy = view['stairs']
            want += [1, sx, sy] if sx != 255 else [0, -1, -1]
            want += [view['tick']]
            assert obs[pl].tolist() == want, (rec['tick'], pl)
            seen_apart += other is None
            for k in range(n_npc):
                e = ents.get(3 + k)
                assert npc[pl, k].tolist() == ([1, e[1], e[2], e[3]] if e is not None else [0, -1, -1, 0]), (rec['tick'], pl, k)
                seen_npc += e is not None
            assert set(ents) <= {1, 2} | {3 + k for k in range(n_npc)}
    return seen_apart, seen_npc


@pytest.mark.parametrize('bots,start', [(('staircase', 'random'), 'together'), (('random', 'random'), 'together'),
                                        (('staircase', 'staircase'), 'separated')])
def test_player_observation_is_view_for(bots, start):
    apart = 0
    for gid in range(4):
        trace, _ = rh.play_episode(0xA11CE, gid, bots=bots, start=start, p_depths=(0, 3), width=12, height=7, max_ticks=120)
        a, _ = check_views(trace, 0)
        apart += a
    if bots[0] == 'staircase':
        assert apart > 0                                                        # the players did end up on different depths


def test_npc_observation_is_view_for():
    from oracle import cport
    from optimax_rogue_b200 import SimConfig
    npc_seen = 0
    for gid in range(4):
        orc = cport.Oracle(SimConfig(seed=0xA11CE, width=8, height=6, n_npc=4), 1, gid)
        orc.reset()
        taken = {(0, int(orc.state.pos[0, 0]), int(orc.state.pos[0, 1])), (0, int(orc.state.pos[0, 2]), int(orc.state.pos[0, 3]))}
        npcs = []
        for (d, x, y, hp) in [(0, 3, 3, 2), (0, 5, 2, 1), (1, 2, 2, 3), (0, 2, 4, 1)]:
            while (d, x, y) in taken:                                           # keep the set-up legal (state.py:78-82)
                x = x % 6 + 1
            taken.add((d, x, y))
            npcs.append((d, x, y, hp))
        trace, _ = rh.play_episode(0xA11CE, gid, bots=('random', 'staircase'), width=8, height=6, max_ticks=150, hp=50, npcs=npcs)
        _, s = check_views(trace, 4)
        npc_seen += s
    assert npc_seen > 0
