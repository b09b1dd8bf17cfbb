"""Invariants of the reference dynamics (suite 5), checked with hypothesis on the C oracle
(CPU) and at BASELINE scale on the CUDA path (-m gpu): players stand on Ground tiles only, never two
entities on one (depth, x, y), health never increases, depth never decreases, tick advances by one."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cport
from optimax_rogue_b200 import SimConfig


def check_planes(prev, cur, cfg, same_episode):
    x1, y1, x2, y2 = (cur['pos'][:, k].astype(int) for k in range(4))
    for x, y in ((x1, y1), (x2, y2)):
        assert ((x >= 1) & (x <= cfg.width - 2) & (y >= 1) & (y <= cfg.height - 2)).all(), 'on a wall'
    sx1, sy1, sx2, sy2 = (cur['stairs'][:, k].astype(int) for k in range(4))
    assert not ((x1 == sx1) & (y1 == sy1)).any() and not ((x2 == sx2) & (y2 == sy2)).any(), 'standing on stairs'
    same_depth = cur['depth'][:, 0] == cur['depth'][:, 1]
    assert not (same_depth & (x1 == x2) & (y1 == y2)).any(), 'two players on one tile'
    if prev is not None:
        s = same_episode
        assert (cur['hp'][s] <= prev['hp'][s]).all(), 'health increased'
        assert (cur['depth'][s] >= prev['depth'][s]).all(), 'depth decreased'
        assert (cur['depth'][s] - prev['depth'][s] <= 1).all(), 'descended two levels in a tick'
        assert (cur['tick'][s] == prev['tick'][s] + 1).all(), 'tick did not advance by one'
        moved = np.abs(cur['pos'][s].astype(int) - prev['pos'][s].astype(int))
        stay_level = cur['depth'][s] == prev['depth'][s]
        assert (moved[:, :2].sum(1)[stay_level[:, 0]] <= 1).all() and (moved[:, 2:].sum(1)[stay_level[:, 1]] <= 1).all(), \
            'moved more than one tile without descending'


def snap(state):
    return {k: getattr(state, k).copy() for k in ('pos', 'hp', 'depth', 'stairs', 'tick', 'episode', 'status')}


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**63 - 1), w=st.integers(4, 40), h=st.integers(4, 14), hp=st.integers(1, 6),
       max_ticks=st.integers(2, 80), base=st.integers(0, 2**50), stream=st.integers(0, 2**31))
def test_oracle_invariants(seed, w, h, hp, max_ticks, base, stream):
    cfg = SimConfig(width=w, height=h, hp=(hp, hp), max_ticks=max_ticks, seed=seed, auto_reset=True)
    orc = cport.Oracle(cfg, 96, game_id_base=base)
    orc.reset()
    rng = np.random.default_rng(stream)
    prev = None
    for _ in range(60):
        cur = snap(orc.state)
        check_planes(prev, cur, cfg, None if prev is None else cur['episode'] == prev['episode'])
        prev = cur
        res, _ = orc.step(rng.integers(0, 7, size=(96, 2), dtype=np.uint8))
        assert ((res >= 1) & (res <= 4)).all()
        finished = res != 1
        assert (orc.state.episode[finished] == prev['episode'][finished] + 1).all(), 'auto-reset bumps the episode'
        assert (orc.state.tick[finished] == 1).all()


@pytest.mark.gpu
def test_cuda_invariants_at_scale():
    import torch
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    n = 1 << 20
    cfg = SimConfig(max_ticks=300, seed=99, auto_reset=True, hp=(4, 4))
    gs = BatchedGameState(cfg, n, 'cuda')
    reset_games(gs)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 300, auto_reset=True)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(5)
    prev = None
    for t in range(120):
        pos, hp, depth, tick, ep, stairs = gs.pos.int(), gs.hp.int(), gs.depth.clone(), gs.tick.clone(), gs.episode.clone(), gs.stairs.int()
        for k in (0, 2):
            assert bool(((pos[:, k] >= 1) & (pos[:, k] <= 58) & (pos[:, k + 1] >= 1) & (pos[:, k + 1] <= 8)).all())
            assert not bool(((pos[:, k] == stairs[:, k]) & (pos[:, k + 1] == stairs[:, k + 1])).any())
        assert not bool(((depth[:, 0] == depth[:, 1]) & (pos[:, 0] == pos[:, 2]) & (pos[:, 1] == pos[:, 3])).any())
        if prev is not None:
            s = ep == prev[4]
            assert bool((hp[s] <= prev[1][s]).all()) and bool((depth[s] >= prev[2][s]).all())
            assert bool((tick[s] == prev[3][s] + 1).all())
        prev = (pos, hp, depth, tick, ep)
        mv = torch.randint(1, 6, (n, 2), dtype=torch.uint8, device='cuda', generator=gen)
        if t % 3 == 0:
            mv = upd.bot_moves(gs, 2, 2)      # drive descents as well
        res, _ = upd.update(gs, mv)
        assert bool(((res >= 1) & (res <= 4)).all())


@pytest.mark.gpu
@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 2**64 - 1), base=st.integers(0, 2**53), w=st.integers(4, 64), h=st.integers(4, 20),
       hp=st.tuples(st.integers(1, 9), st.integers(1, 9)), dmg=st.tuples(st.integers(0, 4), st.integers(0, 4)),
       arm=st.tuples(st.integers(0, 2), st.integers(0, 2)), despawn=st.sampled_from([1, 2]),
       separated=st.booleans(), d2=st.integers(1, 4), max_ticks=st.integers(0, 70), auto_reset=st.booleans(),
       n=st.integers(1, 700), bots=st.tuples(st.sampled_from([1, 2]), st.sampled_from([1, 2])))
def test_cuda_matches_oracle_on_fuzzed_configurations(seed, base, w, h, hp, dmg, arm, despawn, separated, d2,
                                                      max_ticks, auto_reset, n, bots):
    """Random room sizes, stats, generators, despawn strategies, batch sizes (pipelined body + ragged
    tail), 64-bit seeds and 53-bit game ids: CUDA == oracle on every plane, result and event, every tick."""
    from optimax_rogue_b200 import _abi
    import gpu_util as gu
    kw = dict(width=w, height=h, hp=hp, damage=dmg, armor=arm, despawn_strat=despawn, max_ticks=max_ticks,
              auto_reset=auto_reset, seed=seed)
    if separated:
        kw.update(start_kind=_abi.START_SEPARATED, start_depth=(0, d2))
    cfg = SimConfig(**kw)
    gu.run_parity(cfg, n, 40, bots=bots, events=bool(seed & 1), game_id_base=base)
