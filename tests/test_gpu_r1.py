"""-m gpu: ruleset R1 (README-only rules, docs/RULESET_R1.md). PARITY UNPINNED with respect to the
reference -- these tests pin the CUDA kernels to the plain-C restatement of the written spec
(oracle/orx_r1_oracle.c), bit for bit, every plane, every tick."""
import numpy as np
import pytest
import torch

from oracle import cport
from optimax_rogue_b200 import _abi
from optimax_rogue_b200.r1 import R1GameState

pytestmark = pytest.mark.gpu


FLAVOUR = {'flags': 0}


@pytest.fixture(params=['thread_per_game', 'thread_per_game_overlapped', 'half_warp'], autouse=True)
def r1_kernel_flavour(request):
    """Both device formulations of the R1 tick (orx_r1t.cuh: one thread per game, the default, with and without the
    block-by-block ordering of consecutive launches; orx_r1.cu: sixteen lanes per game with warp primitives) must
    match the oracle bit for bit."""
    from optimax_rogue_b200 import _abi
    FLAVOUR['flags'] = {'thread_per_game': 0, 'thread_per_game_overlapped': _abi.R1_PATH_BLOCK_FLAGS,
                        'half_warp': _abi.R1_PATH_HALFWARP}[request.param]
    yield request.param
    FLAVOUR['flags'] = 0


def pair(n, base=0, **cfg):
    gs = R1GameState(n, game_id_base=base, path_flags=FLAVOUR['flags'], **cfg).reset()
    orc = cport.R1Oracle(n, game_id_base=base, **cfg)
    orc.reset()
    return gs, orc


def assert_equal(gs, orc, where):
    p = gs.planes_cpu()
    for name, _, _ in _abi.R1_PLANES:
        a, b = p[name], getattr(orc.state, name)
        if not np.array_equal(a, b):
            bad = np.argwhere(a.reshape(len(a), -1) != b.reshape(len(b), -1))[0]
            raise AssertionError(f'{where}: plane {name} differs at game {bad[0]} col {bad[1]}: '
                                 f'cuda={a[bad[0]]} oracle={b[bad[0]]}')


@pytest.mark.parametrize('w,h,density', [(60, 10, 26), (12, 8, 0), (9, 9, 40), (30, 6, 90)])
def test_r1_reset_parity(w, h, density):
    gs, orc = pair(3001, base=(1 << 33) + 5, width=w, height=h, wall_density=density, seed=11)
    assert_equal(gs, orc, 'reset')


@pytest.mark.parametrize('w,h,density,ticks', [(60, 10, 26, 400), (10, 7, 10, 300), (7, 7, 0, 300), (24, 8, 60, 250)])
def test_r1_step_parity_random_commands(w, h, density, ticks):
    n = 2000
    gs, orc = pair(n, width=w, height=h, wall_density=density, seed=3, max_ticks=150, auto_reset=True)
    rng = np.random.default_rng(w * 1000 + h)
    for t in range(ticks):
        mv = rng.integers(0, 8, size=(n, 2), dtype=np.uint8)        # includes Heal (6) and invalid codes (0, 7)
        ro = orc.step(mv)
        rg = gs.update(torch.from_numpy(mv).cuda())
        assert np.array_equal(rg.cpu().numpy(), ro), f'tick {t}: results differ'
        if t % 10 == 0 or t == ticks - 1:
            assert_equal(gs, orc, f'tick {t}')


def test_r1_frozen_without_auto_reset():
    n = 1500
    gs, orc = pair(n, width=8, height=8, wall_density=0, seed=9, max_ticks=60)
    rng = np.random.default_rng(1)
    for t in range(80):
        mv = rng.integers(1, 7, size=(n, 2), dtype=np.uint8)
        ro = orc.step(mv)
        rg = gs.update(torch.from_numpy(mv).cuda())
        assert np.array_equal(rg.cpu().numpy(), ro)
    assert_equal(gs, orc, 'end')
    assert (gs.status != 1).all()


def test_r1_rollout_matches_oracle_and_rules_fire():
    n = 5000
    gs, orc = pair(n, seed=0x0A11CE, max_ticks=500, auto_reset=True)
    stats = gs.rollout(400)
    ostats = orc.rollout(400)
    assert_equal(gs, orc, 'rollout')
    assert np.array_equal(stats.cpu().numpy().astype(np.uint64), ostats)
    p = gs.planes_cpu()
    alive = (p['ent_loc'] >> 16) & 1
    assert alive[:, 2:10].sum() > 0, 'enemies spawn'
    assert ((p['pl_b'] >> 16) & 255).max() > 0, 'items get picked up'
    assert ((p['pl_b'] >> 8) & 255).max() > 1, 'players level up'
    assert int(stats[1]) + int(stats[2]) > 0, 'games end by death'
    assert int(stats[5]) > 0, 'descents happen'


def test_r1_shard_invariance():
    cfg = dict(seed=21, max_ticks=200, auto_reset=True)
    full = R1GameState(2048, **cfg).reset()
    full.rollout(150)
    part = R1GameState(1024, game_id_base=1024, **cfg).reset()
    part.rollout(150)
    for name, _, _ in _abi.R1_PLANES:
        assert torch.equal(getattr(part, name), getattr(full, name)[1024:]), name


@pytest.mark.parametrize('radius', [-1, 0, 4])
def test_r1_observation_parity(radius):
    gs, orc = pair(3000, width=24, height=9, wall_density=50, seed=17, max_ticks=300, auto_reset=True)
    gs.rollout(120)
    orc.rollout(120)
    assert_equal(gs, orc, 'before observe')
    got = gs.observe(radius).cpu().numpy()
    want = orc.observe(radius)
    assert np.array_equal(got, want)
    assert (got[:, :, 23:47:3] >= 0).any(), 'some enemy is visible'
    if radius == 0:
        assert (got[:, :, 20] == 0).all()      # nobody ever stands on the staircase


def test_r1_unsynchronised_ticks_in_a_graph():
    """Ticks enqueued back to back (a CUDA graph replayed several times, two states interleaved): with the hand-over
    words consecutive launches overlap block by block; every result and both end states equal the oracle's."""
    n, per, replays = 128 * 37 + 50, 12, 5
    a, oa = pair(n, 0, max_ticks=40, auto_reset=True, seed=5)
    b, ob = pair(n, 10 * n, max_ticks=40, auto_reset=True, seed=6)
    g = torch.Generator(device='cuda').manual_seed(2)
    mv = torch.randint(0, 8, (per, n, 2), dtype=torch.uint8, device='cuda', generator=g)
    ra = torch.zeros((per, n), dtype=torch.uint8, device='cuda')
    rb = torch.zeros_like(ra)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        a.update(mv[0], out=ra[0]); b.update(mv[0], out=rb[0])
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for t in range(per):
                a.update(mv[t], out=ra[t])
                b.update(mv[t], out=rb[t])
        for _ in range(replays):
            graph.replay()
    torch.cuda.synchronize()
    m = mv.cpu().numpy()
    for gs, orc, res in ((a, oa, ra), (b, ob, rb)):
        orc.step(m[0])
        want = None
        for _ in range(replays):
            want = np.stack([orc.step(m[t]) for t in range(per)])
        assert np.array_equal(res.cpu().numpy(), want)
        assert_equal(gs, orc, 'after the replayed graph')


def _trim(ev):
    """Records up to each game's terminator; slots behind it are not written by the kernel."""
    ev = ev.copy()
    kind = ev[:, :, 0] & 0xFF
    ended = np.cumsum(kind == 0, axis=1) > 0
    ev[ended] = 0
    return ev


@pytest.mark.parametrize('w,h,density,cap', [(60, 10, 26, _abi.R1_MAX_EVENTS), (12, 8, 10, _abi.R1_MAX_EVENTS), (14, 9, 20, 5)])
def test_r1_replication_log_matches_oracle(r1_kernel_flavour, w, h, density, cap):
    """orx_r1_step_events: every record of every tick, in emission order, equals the oracle's (and the state with it);
    the buffer is poisoned first, so slots behind a game's terminator prove to be left alone."""
    if r1_kernel_flavour == 'half_warp':
        gs = R1GameState(256, path_flags=_abi.R1_PATH_HALFWARP).reset()
        with pytest.raises(RuntimeError):
            gs.update_events(torch.full((256, 2), 5, dtype=torch.uint8, device='cuda'))
        return
    n = 3000
    gs, orc = pair(n, width=w, height=h, wall_density=density, seed=21, max_ticks=120, auto_reset=True)
    rng = np.random.default_rng(w + cap)
    ev = torch.empty((n, cap, 2), dtype=torch.int32, device='cuda')
    kinds = set()
    for t in range(300):
        mv = rng.integers(0, 8, size=(n, 2), dtype=np.uint8)
        if t % 3:
            orc.bot_moves(_abi.BOT_STAIRCASE, _abi.BOT_RANDOM, mv)
        ro, eo = orc.step_events(mv, max_events=cap)
        ev.fill_(0x5A5A5A5A)
        rg, _ = gs.update_events(torch.from_numpy(mv).cuda(), events=ev)
        got = ev.cpu().numpy()
        assert np.array_equal(rg.cpu().numpy(), ro), f'tick {t}: results differ'
        assert np.array_equal(_trim(got), _trim(eo)), f'tick {t}: records differ'
        kind = got[:, :, 0] & 0xFF
        behind = np.cumsum(kind == 0, axis=1) > 1                  # strictly behind the terminator
        assert (got[behind] == 0x5A5A5A5A).all()
        kinds |= set(np.unique(_trim(got)[:, :, 0] & 0xFF).tolist())
        if t % 25 == 0:
            assert_equal(gs, orc, f'tick {t}')
    assert_equal(gs, orc, 'end')
    if cap == _abi.R1_MAX_EVENTS:
        assert kinds >= {_abi.EV_MOVE, _abi.EV_COMBAT, _abi.EV_DEATH, _abi.EV_SPAWN, _abi.EV_HEALTH, _abi.EV_DESCEND}


def test_r1_bot_moves_match_oracle():
    n = 5000
    gs, orc = pair(n, base=1 << 35, seed=8, max_ticks=90, auto_reset=True)
    for t in range(120):
        b1, b2 = [(1, 1), (2, 1), (1, 2), (2, 2), (0, 2), (1, 0)][t % 6]
        mo = orc.bot_moves(b1, b2, np.full((n, 2), 4, np.uint8))
        mg = gs.bot_moves(b1, b2, out=torch.full((n, 2), 4, dtype=torch.uint8, device='cuda'))
        assert np.array_equal(mg.cpu().numpy(), mo), f'tick {t}'
        orc.step(mo)
        gs.update(mg)
    assert_equal(gs, orc, 'end')


@pytest.mark.parametrize('auto_reset', [True, False])
def test_r1_replay_equals_step_loop(auto_reset):
    """orx_r1_replay (T queued ticks, state in registers) == T orx_r1_step calls == the oracle."""
    n, T = 3000, 90
    gs, orc = pair(n, width=10, height=8, wall_density=10, seed=13, max_ticks=50, auto_reset=auto_reset)
    rng = np.random.default_rng(4)
    mv = rng.integers(0, 8, size=(T, n, 2), dtype=np.uint8)
    want = np.stack([orc.step(mv[t]) for t in range(T)])
    got = gs.replay(torch.from_numpy(mv).cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    assert_equal(gs, orc, 'after replay')


def test_r1_host_buffers():
    """orx_r1_step_host_sync: pinned host commands in, pinned host results out, readable when the call returns."""
    n = 4096 + 77
    gs, orc = pair(n, seed=17, max_ticks=70, auto_reset=True)
    rng = np.random.default_rng(6)
    hm = torch.empty((n, 2), dtype=torch.uint8).pin_memory()
    hr = torch.empty((n,), dtype=torch.uint8).pin_memory()
    for t in range(60):
        mv = rng.integers(1, 7, size=(n, 2), dtype=np.uint8)
        hm.copy_(torch.from_numpy(mv))
        hr.fill_(0)
        gs.update_host(hm, hr)                 # no synchronize here: the call returns after the stream has drained
        assert np.array_equal(hr.numpy(), orc.step(mv)), f'tick {t}'
    assert_equal(gs, orc, 'end')
    with pytest.raises(ValueError):
        gs.update_host(torch.empty((n, 2), dtype=torch.uint8), hr)       # not pinned


def test_r1_cuda_reproduces_the_committed_trajectory_digests(r1_kernel_flavour):
    """The same fixture from the CUDA kernels: whole trajectories (every result, the planes every 20 ticks) hash to the
    digests the oracle committed (tests/golden/r1_oracle_digests.json; a regression fixture of the written spec, not a
    reference-derived vector)."""
    import json
    from oracle import gen_golden_r1 as gg
    want = json.load(open(gg.PATH))
    for k, case in enumerate(gg.CASES):
        gs = R1GameState(case['n'], game_id_base=case['base'], path_flags=FLAVOUR['flags'], **case['cfg']).reset()
        got = gg.run_case(k, lambda mv: gs.update(torch.from_numpy(mv).cuda()).cpu().numpy(), gs.planes_cpu)
        assert got == want[case['name']], case['name']
