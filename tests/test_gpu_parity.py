"""-m gpu: the CUDA path (through the C ABI) against the C oracle, bit-exact, on the same seeded
inputs. The oracle itself is pinned to the live reference by tests/golden + test_oracle_*."""
import numpy as np
import pytest
import torch

from optimax_rogue_b200 import SimConfig, _abi
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games

import gpu_util as gu

pytestmark = pytest.mark.gpu

SEED = 0x0A11CE


@pytest.mark.parametrize('w,h', [(60, 10), (4, 4), (5, 7), (255, 255), (16, 4)])
@pytest.mark.parametrize('start', [_abi.START_TOGETHER, _abi.START_SEPARATED])
def test_reset_parity(w, h, start):
    cfg = SimConfig(width=w, height=h, start_kind=start,
                    start_depth=(0, 1000) if start == _abi.START_SEPARATED else (0, 0), seed=SEED)
    gs, upd, orc = gu.make_pair(cfg, 5000, game_id_base=(1 << 40) + 17)
    gu.assert_state_equal(gs, orc, 'reset')
    # masked reset with episode bump touches only the masked lanes
    mask = (np.arange(5000) % 3 == 0).astype(np.uint8)
    reset_games(gs, torch.from_numpy(mask), bump_episode=True)
    orc.reset(mask, bump_episode=True)
    gu.assert_state_equal(gs, orc, 'masked reset')
    assert int(gs.episode.sum()) == int(mask.sum())


@pytest.mark.parametrize('bots', [(1, 1), (2, 1), (2, 2), (1, 2)])
@pytest.mark.parametrize('despawn', [1, 2])
def test_step_parity_bots(bots, despawn):
    cfg = SimConfig(max_ticks=150, seed=SEED, despawn_strat=despawn, auto_reset=True)
    gu.run_parity(cfg, 2048, 320, bots=bots, events=True)


def test_step_parity_frozen_without_auto_reset():
    cfg = SimConfig(max_ticks=40, seed=7, auto_reset=False, hp=(2, 2))
    gs, upd, orc = gu.run_parity(cfg, 1024, 60, bots=(1, 1), events=True)
    assert (gs.status != 1).all()


def test_step_parity_small_rooms_fight():
    # tiny rooms force combat, double deaths and descents every few ticks
    for (w, h) in [(4, 4), (5, 5), (6, 4)]:
        cfg = SimConfig(width=w, height=h, max_ticks=64, seed=99, auto_reset=True, hp=(3, 3))
        gu.run_parity(cfg, 1024, 200, bots=(1, 1), events=True)


@pytest.mark.parametrize('w,h', [(255, 255), (255, 4), (4, 255)])
def test_step_parity_extreme_room_sizes(w, h):
    # coordinates are uint8 and the command table packs W-2 / H-2 into bytes: exercise the limits
    cfg = SimConfig(width=w, height=h, max_ticks=120, seed=123, auto_reset=True, hp=(2, 2))
    gu.run_parity(cfg, 1500, 150, bots=(2, 2), events=True)


def test_step_parity_separated_start_and_stats():
    cfg = SimConfig(start_kind=_abi.START_SEPARATED, start_depth=(3, 5), max_ticks=300, seed=5,
                    auto_reset=True, hp=(4, 7), damage=(3, 2), armor=(1, 0))
    gu.run_parity(cfg, 1024, 400, bots=(2, 1), events=True)


@pytest.mark.parametrize('n', [1, 31, 255, 256, 257, 1000])
def test_step_parity_ragged_batch_sizes(n):
    # < 256 games: simple kernel only; 257 / 1000: TMA-pipelined body + ragged tail
    cfg = SimConfig(max_ticks=30, seed=21, auto_reset=True, hp=(2, 2))
    gu.run_parity(cfg, n, 70, bots=(1, 2), events=False, game_id_base=77)


def test_events_do_not_change_the_state():
    # the event-writing kernel variant and the pipelined hot variant must agree
    cfg = SimConfig(max_ticks=90, seed=8, auto_reset=True)
    a, upd, _ = gu.make_pair(cfg, 3000)
    b = a.clone()
    for t in range(120):
        mv = upd.bot_moves(a, 2, 1)
        ra, _ = upd.update(a, mv, want_events=True)
        rb, _ = upd.update(b, mv, want_events=False)
        assert torch.equal(ra, rb)
    for name in gu.PLANES:
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_checkpoint_resume(tmp_path):
    cfg = SimConfig(max_ticks=64, seed=5, auto_reset=True)
    gs, upd, _ = gu.make_pair(cfg, 2048)
    upd.rollout(gs, 1, 1, 33)
    torch.save(gs.state_dict(), tmp_path / 'ck.pt')
    upd.rollout(gs, 1, 1, 50)
    resumed = BatchedGameState(cfg, 2048, 'cuda')
    resumed.load_state_dict(torch.load(tmp_path / 'ck.pt'))
    upd.rollout(resumed, 1, 1, 50)
    for name in gu.PLANES:
        assert torch.equal(getattr(gs, name), getattr(resumed, name)), name


def test_step_out_of_range_commands_are_stay():
    cfg = SimConfig(max_ticks=0, seed=3)
    rng = np.random.default_rng(0)

    def moves_fn(t, orc):
        return rng.integers(0, 256, size=(orc.n, 2), dtype=np.uint8)
    gu.run_parity(cfg, 2048, 50, moves_fn=moves_fn, events=True)


def fixed_map(w=60, h=10, p=0.10, stairs=False, seed=0):
    rng = np.random.default_rng(seed)
    t = np.full((w, h), 1, np.uint8)
    t[[0, -1], :] = 2
    t[:, [0, -1]] = 2
    inner = rng.random((w - 2, h - 2)) < p
    t[1:-1, 1:-1][inner] = 2
    if stairs:
        g = np.argwhere(t == 1)
        for k in (len(g) // 3, 2 * len(g) // 3):
            t[g[k][0], g[k][1]] = 3
    return t


@pytest.mark.parametrize('stairs', [False, True])
def test_step_parity_fixed_map(stairs):
    cfg = SimConfig(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=fixed_map(stairs=stairs), max_ticks=200,
                    seed=SEED, auto_reset=True, hp=(3, 3))
    gu.run_parity(cfg, 4096, 300, bots=(1, 1) if not stairs else (2, 1), events=True)


def test_step_parity_npcs():
    cfg = SimConfig(width=8, height=6, max_ticks=0, seed=11, n_npc=3, hp=(50, 50))

    def setup(gs, orc):
        rng = np.random.default_rng(1)
        for lane in range(gs.n):
            for k in range(3):
                if rng.random() < 0.8:
                    d, x, y, hp = int(rng.integers(0, 3)), int(rng.integers(1, 7)), int(rng.integers(1, 5)), int(rng.integers(1, 4))
                    orc.state.npc_depth[lane, k] = d
                    orc.state.npc_pos[lane, k] = (x, y)
                    orc.state.npc_hp[lane, k] = hp
        gs.npc_depth.copy_(torch.from_numpy(orc.state.npc_depth))
        gs.npc_pos.copy_(torch.from_numpy(orc.state.npc_pos))
        gs.npc_hp.copy_(torch.from_numpy(orc.state.npc_hp))
    gu.run_parity(cfg, 512, 150, bots=(1, 2), events=True, setup=setup)


@pytest.mark.parametrize('bots', [(1, 1), (2, 1)])
def test_rollout_matches_oracle(bots):
    cfg = SimConfig(max_ticks=100, seed=SEED, auto_reset=True)
    gs, upd, orc = gu.make_pair(cfg, 10000, game_id_base=123456789)
    stats = upd.rollout(gs, bots[0], bots[1], 257)
    ostats = orc.rollout(bots[0], bots[1], 257)
    gu.assert_state_equal(gs, orc, 'rollout')
    assert np.array_equal(stats.cpu().numpy().astype(np.uint64), ostats)
    assert int(stats[0]) == 10000 * 257


def test_rollout_equals_step_loop():
    cfg = SimConfig(max_ticks=64, seed=1, auto_reset=True)
    gs, upd, _ = gu.make_pair(cfg, 4096)
    gs2 = gs.clone()
    upd.rollout(gs, 1, 2, 100)
    for _ in range(100):
        mv = upd.bot_moves(gs2, 1, 2)
        upd.update(gs2, mv)
    for name in gu.PLANES:
        assert torch.equal(getattr(gs, name), getattr(gs2, name)), name


@pytest.mark.parametrize('auto_reset', [True, False])
def test_replay_equals_update_loop(auto_reset):
    cfg = SimConfig(max_ticks=40, seed=12, auto_reset=auto_reset, hp=(2, 2))
    gs, upd, orc = gu.make_pair(cfg, 3000, game_id_base=99)
    twin = gs.clone()
    gen = torch.Generator(device='cuda')
    gen.manual_seed(3)
    moves = torch.randint(0, 7, (70, 3000, 2), dtype=torch.uint8, device='cuda', generator=gen)
    results = upd.replay(gs, moves)
    for t in range(70):
        r, _ = upd.update(twin, moves[t])
        assert torch.equal(r, results[t]), t
        ro, _ = orc.step(moves[t].cpu().numpy())
        assert np.array_equal(ro, results[t].cpu().numpy()), t
    for name in gu.PLANES:
        assert torch.equal(getattr(gs, name), getattr(twin, name)), name
    gu.assert_state_equal(gs, orc, 'replay')


def test_shard_invariance():
    """Game g gives the same trajectory whichever shard (game_id_base) holds it."""
    cfg = SimConfig(max_ticks=80, seed=SEED, auto_reset=True)
    full, upd, _ = gu.make_pair(cfg, 4096, game_id_base=0)
    upd.rollout(full, 1, 1, 200)
    for base, cnt in ((0, 1024), (1024, 3072)):
        part = BatchedGameState(cfg, cnt, 'cuda', game_id_base=base)
        reset_games(part)
        upd.rollout(part, 1, 1, 200)
        for name in gu.PLANES:
            assert torch.equal(getattr(part, name), getattr(full, name)[base:base + cnt]), name


def test_step_host_buffers():
    cfg = SimConfig(max_ticks=50, seed=2, auto_reset=True)
    gs, upd, orc = gu.make_pair(cfg, 3000)
    host_moves = torch.empty((3000, 2), dtype=torch.uint8, pin_memory=True)
    host_res = torch.empty((3000,), dtype=torch.uint8, pin_memory=True)
    for t in range(60):
        mv = orc.bot_moves(1, 1)
        host_moves.copy_(torch.from_numpy(mv))
        res, _ = upd.update(gs, host_moves, out=host_res)
        torch.cuda.synchronize()
        res_o, _ = orc.step(mv)
        assert np.array_equal(res.numpy(), res_o)
    gu.assert_state_equal(gs, orc, 'host path')


def test_host_stepper_matches_oracle():
    cfg = SimConfig(max_ticks=50, seed=6, auto_reset=True)
    gs, upd, orc = gu.make_pair(cfg, 2500)
    host_moves = torch.empty((2500, 2), dtype=torch.uint8, pin_memory=True)
    host_res = torch.empty((2500,), dtype=torch.uint8, pin_memory=True)
    step = upd.host_stepper(gs, host_moves, host_res)
    for t in range(70):
        mv = orc.bot_moves(2, 1)
        host_moves.copy_(torch.from_numpy(mv))
        res = step()                                  # synchronous: results are on the host on return
        res_o, _ = orc.step(mv)
        assert np.array_equal(res.numpy(), res_o)
    gu.assert_state_equal(gs, orc, 'host stepper')
    # pageable buffers take the staged-copy path and must agree as well
    pm, pr = torch.empty((2500, 2), dtype=torch.uint8), torch.empty((2500,), dtype=torch.uint8)
    step2 = upd.host_stepper(gs, pm, pr)
    for t in range(20):
        mv = orc.bot_moves(1, 1)
        pm.copy_(torch.from_numpy(mv))
        assert np.array_equal(step2().numpy(), orc.step(mv)[0])
    gu.assert_state_equal(gs, orc, 'host stepper, pageable')


@pytest.mark.parametrize('n,fixed,npc', [(3000, False, 0), (256, False, 0), (100, False, 0), (1500, True, 0), (700, False, 2)])
def test_packed_commands_match_oracle(n, fixed, npc):
    """orx_step_packed (p1 | p2 << 4 in one byte): device, pinned-host and pageable-host commands all
    play the same tick as the oracle does on the unpacked commands, invalid nibbles included."""
    from optimax_rogue_b200.logic.moves import pack_moves, unpack_moves
    kw = {}
    if fixed:
        tiles = np.ones((12, 9), dtype=np.uint8)
        tiles[0, :] = tiles[-1, :] = 2; tiles[:, 0] = tiles[:, -1] = 2; tiles[5, 3:6] = 2
        kw = dict(width=12, height=9, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=tiles)
    cfg = SimConfig(max_ticks=40, seed=11, auto_reset=True, n_npc=npc, **kw)
    gs, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(n)
    pinned = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    pinned_res = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    step = upd.host_stepper(gs, pinned, pinned_res)
    pageable = torch.empty((n,), dtype=torch.uint8)
    for t in range(60):
        mv = rng.integers(0, 16 if t % 7 == 0 else 6, size=(n, 2), dtype=np.uint8)       # nibbles 0 and 6..15 are Stay
        packed = pack_moves(mv[:, 0], mv[:, 1])
        a, b = unpack_moves(packed)
        assert np.array_equal(a, mv[:, 0]) and np.array_equal(b, mv[:, 1])
        res_o, ev_o = orc.step(mv, want_events=True)
        if t % 3 == 0:
            res, ev = upd.update(gs, torch.from_numpy(packed).cuda(), packed=True, want_events=(t % 6 == 0))
            res = res.cpu().numpy()
            if ev is not None:
                assert np.array_equal(ev.cpu().numpy(), ev_o)
        elif t % 3 == 1:
            pinned.copy_(torch.from_numpy(packed))
            res = step().numpy()
        else:
            pageable.copy_(torch.from_numpy(packed))
            res, _ = upd.update(gs, pageable, packed=True)
            torch.cuda.synchronize()
            res = res.numpy()
        assert np.array_equal(res, res_o), f'tick {t}'
    gu.assert_state_equal(gs, orc, 'packed commands')
    with pytest.raises(ValueError):
        upd.update(gs, torch.zeros((n, 2), dtype=torch.uint8, device='cuda'), packed=True)


def test_host_stepper_async_two_batches_in_flight():
    """sync=False: two independent batches ticked back to back, results read after one event wait each."""
    cfg = SimConfig(max_ticks=30, seed=21, auto_reset=True)
    pairs = [gu.make_pair(cfg, 2048, game_id_base=b * 2048) for b in range(2)]
    bufs = [(torch.empty((2048,), dtype=torch.uint8, pin_memory=True), torch.empty((2048,), dtype=torch.uint8, pin_memory=True))
            for _ in range(2)]
    steps = [pairs[b][1].host_stepper(pairs[b][0], bufs[b][0], bufs[b][1], sync=False) for b in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    from optimax_rogue_b200.logic.moves import pack_moves
    for t in range(40):
        want = []
        for b in range(2):
            mv = pairs[b][2].bot_moves(1, 2)
            bufs[b][0].copy_(torch.from_numpy(pack_moves(mv[:, 0], mv[:, 1])))
            steps[b]()
            evs[b].record()
            want.append(pairs[b][2].step(mv)[0])
        for b in range(2):
            evs[b].synchronize()
            assert np.array_equal(bufs[b][1].numpy(), want[b]), (t, b)
    for b in range(2):
        gu.assert_state_equal(pairs[b][0], pairs[b][2], f'async batch {b}')


def test_concurrent_host_threads_on_separate_streams():
    """The library may be driven from several host threads at once (ctypes drops the GIL during the call):
    four threads tick four independent batches on their own CUDA streams; every batch matches the oracle."""
    import threading
    cfg = SimConfig(max_ticks=40, seed=31, auto_reset=True)
    n, ticks = 4096, 50
    pairs = [gu.make_pair(cfg, n, game_id_base=b * n) for b in range(4)]
    rng = np.random.default_rng(5)
    plans = [rng.integers(1, 6, size=(ticks, n, 2), dtype=np.uint8) for _ in range(4)]
    dev_plans = [torch.from_numpy(p).cuda() for p in plans]
    torch.cuda.synchronize()
    errors = []

    def worker(b):
        try:
            gs, upd, _ = pairs[b]
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for t in range(ticks):
                    upd.update(gs, dev_plans[b][t])
            stream.synchronize()
        except Exception as e:      # surfaced below
            errors.append((b, repr(e)))

    threads = [threading.Thread(target=worker, args=(b,)) for b in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for b in range(4):
        for t in range(ticks):
            pairs[b][2].step(plans[b][t])
        gu.assert_state_equal(pairs[b][0], pairs[b][2], f'thread {b}')


@pytest.mark.parametrize('n', [(1 << 19) + 37, 1 << 16, 1000 * 256, 1500 * 256 + 5, 1776 * 256, 1777 * 256,
                               2300 * 256, 2664 * 256, 2665 * 256, 3000 * 256 + 5])
def test_dynamic_tile_scheduler_equals_static_assignment(n):
    """OrxState.sched: tiles beyond the first kStages (6; 4 earlier in the round) per CTA are claimed from a
    counter. The outcome may not depend on who ticks which tile, and the counter words must be zero again
    after every launch. Sizes cover every regime of tiles vs resident CTAs (444 on a B200) for both depths:
    fewer tiles than CTAs, a partial fixed prefix, claimers that all miss (last fixed round partly filled),
    exactly no dynamic tile (1776 / 2664 tiles), one dynamic tile, many."""
    cfg = SimConfig(max_ticks=60, seed=17, auto_reset=True)
    gs, upd, orc = gu.make_pair(cfg, n)
    static = gs.clone()
    static.sched = torch.zeros((4,), dtype=torch.int32)          # not on the device -> OrxState.sched = NULL
    assert static.c_struct().sched is None and gs.c_struct().sched is not None
    upd_static = type(upd)(upd.dgen, upd.despawn_strat, upd.max_ticks, auto_reset=True)
    rng = np.random.default_rng(1)
    for t in range(24):
        mv = torch.from_numpy(rng.integers(1, 6, size=(n, 2), dtype=np.uint8)).cuda()
        r_dyn, _ = upd.update(gs, mv)
        r_sta, _ = upd_static.update(static, mv)
        assert torch.equal(r_dyn, r_sta), t
        assert int(gs.sched.abs().sum()) == 0, (t, gs.sched[:8].tolist())
        orc.step(mv.cpu().numpy())
    for name in gu.PLANES:
        assert torch.equal(getattr(gs, name), getattr(static, name)), name
    gu.assert_state_equal(gs, orc, 'dynamic tiles')


from obs_util import expected_obs, expected_npc_obs      # numpy restatements, pinned to the live view_for in test_observation_vs_reference.py


@pytest.mark.parametrize('n_npc', [1, 3, 8])
def test_observe_npc_slots_on_the_viewers_depth(n_npc):
    """orx_observe_npc: view_for (state.py:53-58) keeps the entities on the viewer's depth -- per player and NPC slot
    { on_my_depth, x, y, health }, against a numpy restatement, while StaircaseBot takes player 1 down the levels and
    hits wear the NPCs down (dead ones leave their slot, updater.py:137-145)."""
    n = 3000
    cfg = SimConfig(max_ticks=0, seed=12, n_npc=n_npc, width=14, height=6, hp=(200, 200))
    gs, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(n_npc)
    live = rng.random((n, n_npc)) < 0.7
    gs.npc_depth.copy_(torch.from_numpy(np.where(live, rng.integers(0, 3, size=(n, n_npc)), -1).astype(np.int32)))
    gs.npc_pos.copy_(torch.from_numpy(np.stack([rng.integers(1, 13, size=(n, n_npc)), rng.integers(1, 5, size=(n, n_npc))], axis=-1).astype(np.uint8)))
    gs.npc_hp.copy_(torch.from_numpy(rng.integers(1, 5, size=(n, n_npc)).astype(np.int16)))
    seen = 0
    for t in range(30):
        p = gs.planes_cpu()
        want = expected_npc_obs(p)
        got = upd.observe_npc(gs).cpu().numpy()
        assert np.array_equal(got, want), t
        seen += int(want[:, :, :, 0].sum())
        upd.update(gs, upd.bot_moves(gs, 2, 1))
    assert seen > 0 and (gs.planes_cpu()['npc_depth'] < 0).sum() > (~live).sum()        # visible NPCs existed, and some died on the way
    assert upd.observe_npc(gu.make_pair(SimConfig(seed=1), 64)[0]).shape == (64, 2, 0, 4)


@pytest.mark.parametrize('n', [100, 2000, 70000])
def test_observe_all_columns_and_paths(n):
    """n = 100: simple kernel; 2000: TMA pipeline + ragged tail; 70000: pipeline with claimed tiles."""
    cfg = SimConfig(max_ticks=0, seed=4)
    gs, upd, orc = gu.make_pair(cfg, n)
    upd.rollout(gs, 2, 1, 40)
    for radius in (-1, 3):
        obs = upd.observe(gs, stairs_radius=radius).cpu().numpy()
        assert np.array_equal(obs, expected_obs(gs.planes_cpu(), radius)), radius
    assert int(gs.sched[:_abi.SCHED_HEADER_WORDS].abs().sum()) == 0


@pytest.mark.parametrize('n,packed,fixed,npc', [(3000, False, False, 0), (70000, True, False, 0), (200, False, False, 0),
                                                (1500, True, True, 0), (700, False, False, 2)])
def test_update_observe_equals_update_then_observe(n, packed, fixed, npc):
    """orx_step_observe = one tick + the observations of the resulting state in one pass: same results,
    same planes, same observations as the two calls, for both command formats and every kernel path."""
    from optimax_rogue_b200.logic.moves import pack_moves
    kw = dict(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=fixed_map(stairs=True)) if fixed else {}
    cfg = SimConfig(max_ticks=30, seed=23, auto_reset=True, n_npc=npc, **kw)
    gs, upd, orc = gu.make_pair(cfg, n)
    ref = gs.clone()
    rng = np.random.default_rng(n)
    for t in range(45):
        mv = rng.integers(1, 6, size=(n, 2), dtype=np.uint8)
        cmds = torch.from_numpy(pack_moves(mv[:, 0], mv[:, 1]) if packed else mv).cuda()
        res, obs = upd.update_observe(gs, cmds, packed=packed, stairs_radius=4)
        res2, _ = upd.update(ref, cmds, packed=packed)
        obs2 = upd.observe(ref, stairs_radius=4)
        assert torch.equal(res, res2), t
        assert torch.equal(obs, obs2), t
        res_o, _ = orc.step(mv)
        assert np.array_equal(res.cpu().numpy(), res_o), t
    gu.assert_state_equal(gs, orc, 'update_observe')
    assert np.array_equal(obs.cpu().numpy(), expected_obs(gs.planes_cpu(), 4))


def test_observe():
    cfg = SimConfig(max_ticks=0, seed=4)
    gs, upd, orc = gu.make_pair(cfg, 2000)
    upd.rollout(gs, 2, 1, 40)
    obs = upd.observe(gs, stairs_radius=3).cpu().numpy()
    p = gs.planes_cpu()
    for pl in range(2):
        o = 1 - pl
        assert np.array_equal(obs[:, pl, 0], p['pos'][:, 2 * pl])
        assert np.array_equal(obs[:, pl, 1], p['pos'][:, 2 * pl + 1])
        assert np.array_equal(obs[:, pl, 2], p['depth'][:, pl])
        assert np.array_equal(obs[:, pl, 3], p['hp'][:, pl])
        same = p['depth'][:, 0] == p['depth'][:, 1]
        assert np.array_equal(obs[:, pl, 4], same.astype(np.int16))
        assert np.array_equal(obs[:, pl, 5], np.where(same, p['pos'][:, 2 * o].astype(int), -1))
        cheb = np.maximum(np.abs(p['stairs'][:, 2 * pl].astype(int) - p['pos'][:, 2 * pl]),
                          np.abs(p['stairs'][:, 2 * pl + 1].astype(int) - p['pos'][:, 2 * pl + 1]))
        vis = cheb <= 3
        assert np.array_equal(obs[:, pl, 8], vis.astype(np.int16))
        assert np.array_equal(obs[:, pl, 9], np.where(vis, p['stairs'][:, 2 * pl].astype(int), -1))
        assert np.array_equal(obs[:, pl, 11], p['tick'])


def test_bad_arguments_fail_loudly():
    from optimax_rogue_b200 import _lib
    import ctypes as C
    cfg = SimConfig(seed=1)
    gs = BatchedGameState(cfg, 16, 'cuda')
    c = gs.c_config()
    st = gs.c_struct()
    lib = _lib.lib()
    assert lib.orx_step(C.byref(c), C.byref(st), None, None, None, 16, 0, None) == _abi.ERR_BAD_ARG
    c.struct_size = 4
    assert lib.orx_reset(C.byref(c), C.byref(st), None, 0, 16, 0, None) == _abi.ERR_BAD_ARG
    c = gs.c_config()
    c.width = 3
    assert lib.orx_reset(C.byref(c), C.byref(st), None, 0, 16, 0, None) == _abi.ERR_BAD_ARG
    with pytest.raises(RuntimeError):
        _lib.check(_abi.ERR_BAD_ARG, 'x')
    # game ids must fit the 54 bits the Philox counter reserves for them
    assert lib.orx_reset(C.byref(gs.c_config()), C.byref(st), None, 0, 16, 1 << 54, None) == _abi.ERR_BAD_ARG
    # n == 0 is a no-op
    c = gs.c_config()
    assert lib.orx_reset(C.byref(c), C.byref(st), None, 0, 0, 0, None) == 0
    # stats the int16 planes cannot hold are refused at the boundary (not wrapped)
    for field, val in (('hp', 40000), ('hp', 0), ('damage', 40000)):
        c = gs.c_config()
        getattr(c, field)[0] = val
        assert lib.orx_reset(C.byref(c), C.byref(st), None, 0, 16, 0, None) == _abi.ERR_BAD_ARG, (field, val)
    # the bit-packed streams: null / misaligned buffers, NPC slots
    c = gs.c_config()
    buf = torch.zeros((64,), dtype=torch.uint8, device='cuda')
    assert lib.orx_step_bits(C.byref(c), C.byref(st), None, buf.data_ptr(), 16, 0, None) == _abi.ERR_BAD_ARG
    assert lib.orx_step_bits(C.byref(c), C.byref(st), buf.data_ptr() + 1, buf.data_ptr() + 32, 16, 0, None) == _abi.ERR_BAD_ARG
    gn = BatchedGameState(SimConfig(seed=1, n_npc=2), 16, 'cuda')
    assert lib.orx_step_bits(C.byref(gn.c_config()), C.byref(gn.c_struct()), buf.data_ptr(), buf.data_ptr() + 32, 16, 0, None) == _abi.ERR_UNSUPPORTED
    assert lib.orx_sched_words(1 << 20) == _abi.sched_words(1 << 20) and lib.orx_abi_version() == _abi.ABI_VERSION


def test_full_size_batch_properties():
    """BASELINE.json size (2^20 games): the oracle cannot replay a million games in seconds, so the
    checks are size-independent properties -- fused rollout == tick-by-tick loop on every lane, and
    four 1024-game windows of the batch equal the oracle run on just those global game ids."""
    n, ticks = 1 << 20, 48
    cfg = SimConfig(max_ticks=40, seed=SEED, auto_reset=True, hp=(3, 3))
    gs = BatchedGameState(cfg, n, 'cuda')
    reset_games(gs)
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 40, auto_reset=True)
    twin = gs.clone()
    stats = upd.rollout(gs, 1, 2, ticks)
    moves = torch.empty((n, 2), dtype=torch.uint8, device='cuda')
    wins = torch.zeros(5, dtype=torch.int64, device='cuda')
    for _ in range(ticks):
        upd.bot_moves(twin, 1, 2, out=moves)
        res, _ = upd.update(twin, moves)
        wins += torch.bincount(res.to(torch.int64), minlength=5)
    for name in gu.PLANES:
        assert torch.equal(getattr(gs, name), getattr(twin, name)), name
    assert int(stats[0]) == n * ticks
    assert [int(stats[1]), int(stats[2]), int(stats[3])] == [int(wins[2]), int(wins[3]), int(wins[4])]
    p = gs.planes_cpu()
    from oracle import cport
    for start in (0, 123 * 1024, 700 * 1024 + 1, n - 1024):
        orc = cport.Oracle(cfg, 1024, game_id_base=start)
        orc.reset()
        orc.rollout(1, 2, ticks)
        for name in gu.PLANES:
            a = p[name][start:start + 1024]
            if name == 'episode':
                a = a.view(np.uint32)
            assert np.array_equal(a, getattr(orc.state, name)), (start, name)


def test_selfplay_example_runs():
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'examples', 'selfplay_loop.py'), '--games', '4096', '--ticks', '20'],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'game-ticks/s' in out.stdout
    out = subprocess.run([sys.executable, os.path.join(root, 'examples', 'selfplay_loop.py'), '--games', '4096', '--ticks', '20',
                          '--opponent', 'staircase'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'game-ticks/s' in out.stdout, out.stderr[-2000:]


def test_single_game_interop_and_event_decoding():
    """Lane <-> host GameState conversion (reference attribute names, reference wire format) and
    decoding of the device event records into GameStateUpdate objects with running orders."""
    from optimax_rogue_b200.game.state import GameState
    from optimax_rogue_b200.logic import updates
    cfg = SimConfig(max_ticks=0, seed=31)
    gs, upd, orc = gu.make_pair(cfg, 64)
    order = 0
    seen = set()
    for t in range(120):
        mv = upd.bot_moves(gs, 2, 2)
        before = gs.planes_cpu()
        res, ev = upd.update(gs, mv, want_events=True)
        orc.step(mv.cpu().numpy())
        recs = updates.unpack_events(ev)
        after = gs.planes_cpu()
        evs = updates.decode_events(recs[0], first_order=order)
        order += len(evs)
        for e in evs:
            seen.add(type(e).__name__)
            if isinstance(e, updates.EntityPositionUpdate):
                k = e.entity_iden - 1
                assert (e.posx, e.posy, e.depth) == (int(after['pos'][0, 2 * k]), int(after['pos'][0, 2 * k + 1]), int(after['depth'][0, k])) \
                    or len([x for x in evs if isinstance(x, updates.EntityPositionUpdate) and x.entity_iden == e.entity_iden]) > 1
                assert e.old_depth == int(before['depth'][0, k])
            if isinstance(e, updates.DungeonCreatedUpdate):
                assert e.dungeon.width == 60 and e.dungeon.height == 10 and (e.dungeon.tiles == 3).sum() == 1
    assert {'EntityPositionUpdate', 'DungeonCreatedUpdate'} <= seen
    assert int(upd.get_incr_upd_order()[0]) == order                       # Updater.get_incr_upd_order, updater.py:71-74
    art = gs.render(5).splitlines()
    assert len(art) == 11 and all(len(r) == 60 for r in art[1:]) and art[1] == '#' * 60 and 'o' in ''.join(art[1:])
    # lane -> GameState -> bytes -> GameState -> lane
    host = gs.to_game_state(5)
    p = gs.planes_cpu()
    assert (host.player_1.x, host.player_1.y, host.player_1.depth, host.player_1.health) == \
        (int(p['pos'][5, 0]), int(p['pos'][5, 1]), int(p['depth'][5, 0]), int(p['hp'][5, 0]))
    assert host.tick == int(p['tick'][5]) and host.world.dungeons[host.player_2.depth].staircase() == \
        (int(p['stairs'][5, 2]), int(p['stairs'][5, 3]))
    back = GameState.from_prims(host.to_prims())
    other = BatchedGameState(cfg, 8, 'cuda', game_id_base=gs.game_id_base + 5)   # lane 0 of `other` has the same global id
    reset_games(other)
    other.load_game_state(0, back)
    q = other.planes_cpu()
    for name in ('pos', 'hp', 'depth', 'stairs', 'tick'):
        assert np.array_equal(q[name][0], p[name][5]), name
    # and the transplanted lane keeps playing exactly like the original (same global game id, same episode)
    mv = upd.bot_moves(gs, 1, 1)
    upd.update(gs, mv)
    upd2 = BatchedUpdater(upd.dgen, 1, None)
    mv2 = upd2.bot_moves(other, 1, 1)
    assert torch.equal(mv2[0], mv[5])
    upd2.update(other, mv2)
    for name in ('pos', 'hp', 'depth', 'stairs', 'tick'):
        assert torch.equal(getattr(other, name)[0], getattr(gs, name)[5]), name


@pytest.mark.parametrize('despawn', [1, 2])
def test_single_game_updater_is_a_drop_in_for_the_reference_updater(despawn):
    """SingleGameUpdater mutates a host GameState in place like Updater.update (updater.py:76-162):
    after every tick the host object equals the device lane, the world holds exactly the levels the
    reference's despawn strategy would keep, and the update orders run without gaps."""
    from optimax_rogue_b200.logic.compat import SingleGameUpdater
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator, TogetherGameStartGenerator
    from optimax_rogue_b200.logic import updates
    dgen = EmptyDungeonGenerator(20, 8)
    seed_state = TogetherGameStartGenerator(dgen).setup_game(1, seed=77, game_id_base=1234)
    host = seed_state.to_game_state(0)
    upd = SingleGameUpdater(dgen, despawn, max_ticks=400, seed=77, game_id=1234)
    orc = cport_oracle_for(dgen, despawn, 400, 77, 1234)
    rng = np.random.default_rng(despawn)
    expect_order = 0
    for t in range(300):
        p1 = host.player_1
        sx, sy = host.world.dungeons[p1.depth].staircase()                        # StaircaseBot, staircasebot.py:9-20
        dx, dy = sx - p1.x, sy - p1.y
        m1 = (2 if dx > 0 else 4) if abs(dx) > abs(dy) else (3 if dy > 0 else 1)
        m2 = int(rng.integers(1, 6)) if t % 3 else m1
        res, evs = upd.update(host, m1, m2)
        ro, _ = orc.step(np.array([[m1, m2]], np.uint8))
        assert int(res) == int(ro[0])
        assert [e.order for e in evs] == list(range(expect_order, expect_order + len(evs)))
        expect_order += len(evs)
        s = orc.state
        assert (host.player_1.x, host.player_1.y, host.player_1.depth, host.player_1.health) == \
            (int(s.pos[0, 0]), int(s.pos[0, 1]), int(s.depth[0, 0]), int(s.hp[0, 0]))
        assert (host.player_2.x, host.player_2.y, host.player_2.depth, host.player_2.health) == \
            (int(s.pos[0, 2]), int(s.pos[0, 3]), int(s.depth[0, 1]), int(s.hp[0, 1]))
        assert host.tick == int(s.tick[0])
        d1, d2 = host.player_1.depth, host.player_2.depth
        want_levels = set(range(min(d1, d2), max(d1, d2) + 1)) if despawn == 1 else {d1, d2}
        assert set(host.world.dungeons) == want_levels
        assert host.world.dungeons[d1].staircase() == (int(s.stairs[0, 0]), int(s.stairs[0, 1]))
        assert (d1, host.player_1.x, host.player_1.y) in host.pos_lookup
        if int(res) != 1:
            break
    assert host.player_1.depth > 3


def cport_oracle_for(dgen, despawn, max_ticks, seed, gid):
    from oracle import cport
    cfg = SimConfig(width=dgen.width, height=dgen.height, despawn_strat=despawn, max_ticks=max_ticks, seed=seed)
    orc = cport.Oracle(cfg, 1, game_id_base=gid)
    orc.reset()
    return orc


def test_long_horizon_soak():
    """2^20 games x 4096 fused ticks (several episodes per lane, depths in the hundreds for the
    StaircaseBot): windows of the batch must equal the oracle run on the same global game ids."""
    from oracle import cport
    n, ticks = 1 << 20, 4096
    cfg = SimConfig(max_ticks=1500, seed=0xC0FFEE, auto_reset=True)
    gs = BatchedGameState(cfg, n, 'cuda', game_id_base=1 << 33)
    reset_games(gs)
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 1500, auto_reset=True)
    stats = upd.rollout(gs, 2, 1, ticks)
    assert int(stats[0]) == n * ticks
    p = gs.planes_cpu()
    assert p['depth'].max() > 50 and p['episode'].view(np.uint32).min() >= 2
    for start in (0, 555_555, n - 256):
        orc = cport.Oracle(cfg, 256, game_id_base=(1 << 33) + start)
        orc.reset()
        orc.rollout(2, 1, ticks)
        for name in gu.PLANES:
            a = p[name][start:start + 256]
            if name == 'episode':
                a = a.view(np.uint32)
            assert np.array_equal(a, getattr(orc.state, name)), (start, name)


@pytest.mark.gpu
def test_plane_placement_does_not_change_results():
    """The tick kernel moves pos/hp/stairs/tick/episode with one tensor-map copy per tile when they share an
    allocation at a common pitch (BatchedGameState does that) and plane by plane otherwise; both equal."""
    import torch
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    dev = torch.device('cuda')
    n = 256 * 37 + 19                                  # full tiles plus a ragged tail
    cfg = SimConfig(max_ticks=40, seed=11, auto_reset=True)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 40, auto_reset=True)
    a = BatchedGameState(cfg, n, dev)
    reset_games(a)
    assert a.hp.data_ptr() - a.pos.data_ptr() == a.stairs.data_ptr() - a.hp.data_ptr()      # stacked
    b = a.clone()
    pitch = (4 * n + 127) // 128 * 128                                                      # scattered:
    backing = torch.zeros((5 * pitch + 256,), dtype=torch.uint8, device=dev)                # uneven spacing
    for (name, dtype, shape), off in zip(BatchedGameState.WORD_PLANES, (0, pitch + 16, 2 * pitch + 64, 3 * pitch + 16, 4 * pitch + 128)):
        plane = backing[off:off + 4 * n].view(dtype).view((n,) + shape)
        plane.copy_(getattr(a, name))
        setattr(b, name, plane)
    pitch_b = b.hp.data_ptr() - b.pos.data_ptr()
    assert (b.stairs.data_ptr() - b.pos.data_ptr(), b.tick.data_ptr() - b.pos.data_ptr()) != (2 * pitch_b, 3 * pitch_b)
    g = torch.Generator(device='cpu').manual_seed(5)
    obs_a = torch.empty((n, 2, 12), dtype=torch.int16, device=dev)
    obs_b = torch.empty_like(obs_a)
    for t in range(90):
        mv = torch.randint(0, 8, (n, 2), dtype=torch.uint8, generator=g).to(dev)
        if t % 3 == 0:
            ra, _ = upd.update_observe(a, mv, obs_out=obs_a)
            rb, _ = upd.update_observe(b, mv, obs_out=obs_b)
            assert torch.equal(obs_a, obs_b)
        else:
            ra, rb = upd.update(a, mv)[0], upd.update(b, mv)[0]
        assert torch.equal(ra, rb)
    for name in BatchedGameState.PLANES:
        assert torch.equal(getattr(a, name), getattr(b, name)), name


@pytest.mark.gpu
def test_event_log_through_the_tile_pipeline_equals_the_simple_kernel():
    """With no NPC slots the event records of full 256-game tiles are staged in shared memory and
    streamed out by the pipelined tick kernel; path flag ORX_PATH_NO_EVENT_PIPE forces the one-thread-per-game
    kernel. Same records, same results, same states; and the running update order
    (Updater.get_incr_upd_order, updater.py:71-74) equals the number of records emitted."""
    import os
    import torch
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    dev = torch.device('cuda')
    n = 256 * 21 + 77
    cfg = SimConfig(max_ticks=50, seed=23, auto_reset=True, width=12, height=7)     # small room: combat and stairs are frequent
    dg = EmptyDungeonGenerator(12, 7)
    ua, ub = BatchedUpdater(dg, 2, 50, auto_reset=True), BatchedUpdater(dg, 2, 50, auto_reset=True)
    ub.path_flags = _abi.PATH_NO_EVENT_PIPE
    a = BatchedGameState(cfg, n, dev)
    reset_games(a)
    b = a.clone()
    g = torch.Generator(device='cpu').manual_seed(9)
    expect = torch.zeros((n,), dtype=torch.int64, device=dev)
    kinds_seen = set()
    for t in range(120):
        mv = torch.randint(0, 7, (n, 2), dtype=torch.uint8, generator=g).to(dev)
        ra, ea = ua.update(a, mv, want_events=True)
        rb, eb = ub.update(b, mv, want_events=True)
        assert torch.equal(ra, rb) and torch.equal(ea, eb), t
        k = ea[:, :, 0] & 0xFF
        expect += (k != 0).sum(dim=1)
        kinds_seen.update(int(x) for x in torch.unique(k).cpu())
    assert {1, 2, 3, 5} <= kinds_seen                   # move, combat, dungeon created, descend (4 = NPC death: no NPCs here)
    for name in BatchedGameState.PLANES:
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(ua.get_incr_upd_order(), expect) and torch.equal(ub.get_incr_upd_order(), expect)


@pytest.mark.gpu
@pytest.mark.parametrize('bots', [(0, 1), (2, 0), (1, 2), (2, 2), (0, 0)])
@pytest.mark.parametrize('fixed', [False, True])
def test_scripted_players_inside_the_tick(bots, fixed):
    """orx_step_bots == orx_bot_moves for the scripted players, merged with the caller's commands for the
    others, then orx_step: same results, planes, event records and observations; pipelined tiles and the ragged
    tail."""
    import torch
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    dev = torch.device('cuda')
    n = 256 * 9 + 31
    if fixed:
        cfg = SimConfig(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=fixed_map(stairs=True), max_ticks=70, seed=41, auto_reset=True)
    else:
        cfg = SimConfig(max_ticks=70, seed=41, auto_reset=True, width=14, height=8)
    a, upd, _ = gu.make_pair(cfg, n)
    b = a.clone()
    upd_b = BatchedUpdater(upd.dgen, upd.despawn_strat, upd.max_ticks, auto_reset=True)
    g = torch.Generator(device='cpu').manual_seed(3)
    for t in range(80):
        mv = torch.randint(1, 6, (n, 2), dtype=torch.uint8, generator=g).to(dev)
        ref_mv = mv.clone()
        scripted = upd_b.bot_moves(b, bots[0], bots[1])
        for p in range(2):
            if bots[p] != 0:
                ref_mv[:, p] = scripted[:, p]
        mode = t % 3
        if mode == 0:
            ra, _, _ = upd.update_with_bots(a, mv, bots[0], bots[1])
            rb, _ = upd_b.update(b, ref_mv)
        elif mode == 1:
            ra, _, oa = upd.update_with_bots(a, mv, bots[0], bots[1], observe=True, stairs_radius=3)
            rb, ob = upd_b.update_observe(b, ref_mv, stairs_radius=3)
            assert torch.equal(oa, ob), t
        else:
            ra, ea, _ = upd.update_with_bots(a, mv, bots[0], bots[1], want_events=True)
            rb, eb = upd_b.update(b, ref_mv, want_events=True)
            assert torch.equal(ea, eb), t
        assert torch.equal(ra, rb), t
    for name in BatchedGameState.PLANES:
        assert torch.equal(getattr(a, name), getattr(b, name)), name


@pytest.mark.gpu
def test_device_stepper_equals_update():
    """BatchedUpdater.device_stepper: the bound, pre-marshalled call equals update / update_observe /
    update_with_bots, with the commands read from the bound buffer at each call."""
    import torch
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.moves import pack_moves
    from optimax_rogue_b200.logic.updater import BatchedUpdater
    n = 256 * 5 + 9
    cfg = SimConfig(max_ticks=50, seed=77, auto_reset=True, width=16, height=9)
    a, upd, _ = gu.make_pair(cfg, n)
    b = a.clone()
    upd_b = BatchedUpdater(upd.dgen, upd.despawn_strat, upd.max_ticks, auto_reset=True)
    mv = torch.empty((n, 2), dtype=torch.uint8, device='cuda')
    cmd = torch.empty((n,), dtype=torch.uint8, device='cuda')
    res = torch.empty((n,), dtype=torch.uint8, device='cuda')
    obs = torch.empty((n, 2, 12), dtype=torch.int16, device='cuda')
    steppers = [upd.device_stepper(a, mv, res), upd.device_stepper(a, cmd, res, packed=True),
                upd.device_stepper(a, mv, res, obs=obs, stairs_radius=2), upd.device_stepper(a, mv, res, bots=(2, 0))]
    g = torch.Generator(device='cpu').manual_seed(1)
    for t in range(60):
        m = torch.randint(1, 6, (n, 2), dtype=torch.uint8, generator=g).cuda()
        mv.copy_(m)
        cmd.copy_(pack_moves(m[:, 0], m[:, 1]))
        k = t % 4
        ra = steppers[k]()
        if k < 2:
            rb, _ = upd_b.update(b, m)
        elif k == 2:
            rb, ob = upd_b.update_observe(b, m, stairs_radius=2)
            assert torch.equal(obs, ob), t
        else:
            rb, _, _ = upd_b.update_with_bots(b, m, 2, 0)
        assert torch.equal(ra, rb), t
    for name in BatchedGameState.PLANES:
        assert torch.equal(getattr(a, name), getattr(b, name)), name


@pytest.mark.gpu
@pytest.mark.parametrize('n_npc', [1, 3, 8])
def test_npc_slots_through_the_tile_pipeline(n_npc):
    """With NPC slots the slot planes travel through the tile pipeline as three more slices per tile;
    path flag ORX_PATH_NO_NPC_PIPE forces the one-thread-per-game kernel. Both equal the oracle: planes, NPC planes,
    results; players bump into NPCs, kill them, and descend next to them."""
    import os
    import torch
    n = 256 * 6 + 40
    cfg = SimConfig(max_ticks=0, seed=19, width=9, height=6, damage=(3, 2), n_npc=n_npc)
    a, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(4)
    live = rng.random((n, n_npc)) < 0.7               # scatter NPCs over the first two levels
    orc.state.npc_depth[:] = np.where(live, rng.integers(0, 2, size=(n, n_npc)), -1)
    orc.state.npc_pos[:, :, 0] = rng.integers(1, 8, size=(n, n_npc))
    orc.state.npc_pos[:, :, 1] = rng.integers(1, 5, size=(n, n_npc))
    orc.state.npc_hp[:] = rng.integers(1, 7, size=(n, n_npc))
    a.npc_depth.copy_(torch.from_numpy(orc.state.npc_depth))
    a.npc_pos.copy_(torch.from_numpy(orc.state.npc_pos))
    a.npc_hp.copy_(torch.from_numpy(orc.state.npc_hp))
    culled0 = int((a.npc_depth < 0).sum())
    b = a.clone()
    upd_b = type(upd)(upd.dgen, upd.despawn_strat, upd.max_ticks, auto_reset=False)
    upd_b.path_flags = _abi.PATH_NO_NPC_PIPE
    for t in range(70):
        mv = orc.bot_moves(1, 2) if t % 2 else rng.integers(1, 6, size=(n, 2), dtype=np.uint8)
        m = torch.from_numpy(mv).cuda()
        ra, _ = upd.update(a, m)
        rb, _ = upd_b.update(b, m)
        ro, _ = orc.step(mv, want_events=False)
        assert torch.equal(ra, rb) and np.array_equal(ra.cpu().numpy(), ro), t
        gu.assert_state_equal(a, orc, f'tick {t} (pipeline)')
    gu.assert_state_equal(b, orc, 'simple kernel')
    assert int((a.npc_depth < 0).sum()) > culled0      # some NPCs were killed and culled (updater.py:137-145)


@pytest.mark.parametrize('n', [256 * 9 + 31, 200])
def test_flat_modifier_bonuses_match_oracle(n):
    """The Modifier seam (OrxState.flat; game/modifiers.py:102-108, game/attribles.py:21-43, updater.py:313): random
    per-game flat damage / armor bonuses for both players, small rooms so that fights are frequent; tile pipeline
    and simple kernel (ragged tail / small batch), with the event log (og_damage is part of the record)."""
    cfg = SimConfig(width=5, height=5, max_ticks=60, seed=77, auto_reset=True, hp=(30, 25), damage=(2, 3), armor=(1, 1))
    gs, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(n)
    flat = rng.integers(-4, 9, size=(n, 2, 3)).astype(np.int8)
    flat[::7] = 0
    flat[5, 0] = (127, -128, 0)
    gs.enable_flat_bonuses().copy_(torch.from_numpy(flat))
    orc.state.enable_flat_bonuses()[:] = flat
    hits = 0
    for t in range(90):
        mv = rng.integers(1, 6, size=(n, 2), dtype=np.uint8)
        ro, eo = orc.step(mv, want_events=True)
        want_ev = t % 2 == 0
        rg, eg = upd.update(gs, torch.from_numpy(mv).cuda(), want_events=want_ev)
        assert np.array_equal(rg.cpu().numpy(), ro), t
        if want_ev:
            assert np.array_equal(eg.cpu().numpy(), eo), t
            hits += int(((eo[:, :, 0] & 0xFF) == _abi.EV_COMBAT).sum())
        gu.assert_state_equal(gs, orc, f'tick {t}')
    assert hits > n // 4
    c = gs.clone()
    assert torch.equal(c.flat, gs.flat) and c.flat.data_ptr() != gs.flat.data_ptr()
