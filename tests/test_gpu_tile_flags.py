"""-m gpu: chunk-by-chunk ordering of consecutive tick launches (ORX_PATH_TILE_FLAGS = SimConfig.overlap_ticks,
OrxState.sched with orx_sched_words(n) words, csrc/orx_pipe.cuh "flag mode"). The other parity tests read the state back after every tick, which
serialises the stream; here many ticks are enqueued back to back, eagerly and as a CUDA graph, on one state
and on interleaved states, so that consecutive launches really overlap -- and the outcome must still be the
oracle's, tick for tick (the tick is Updater.update, optimax_rogue/logic/updater.py:76-162)."""
import numpy as np
import pytest
import torch

from optimax_rogue_b200 import SimConfig, _abi

import gpu_util as gu

pytestmark = pytest.mark.gpu


def _moves(rng, ticks, n):
    return rng.integers(0, 7, size=(ticks, n, 2), dtype=np.uint8)      # includes invalid codes 0 and 6


def _oracle_run(orc, mv):
    res = np.empty(mv.shape[:2], np.uint8)
    for t in range(mv.shape[0]):
        res[t], _ = orc.step(mv[t], want_events=False)
    return res


@pytest.mark.parametrize('n', [256 * 40 + 13, 256 * 3, 131072])
@pytest.mark.parametrize('tpc', [0, 1, 4, 7, 12])
def test_back_to_back_ticks_on_one_state(n, tpc):
    """T ticks enqueued without any synchronisation between them: tick k+1 may start on a tile as soon as
    tick k has written it. Every result and the final planes equal the oracle's."""
    ticks = 48
    cfg = SimConfig(max_ticks=29, seed=77, auto_reset=True, width=11, height=6, overlap_ticks=True,
                    path_flags=tpc << _abi.PATH_TILES_PER_CTA_SHIFT)
    gs, upd, orc = gu.make_pair(cfg, n)
    assert gs.sched.numel() == _abi.sched_words(n)
    rng = np.random.default_rng(n + tpc)
    mv = _moves(rng, ticks, n)
    pad = -(-n // 16) * 16            # every tick's command / result rows 16-byte aligned: all launches take the tile pipeline
    dmv = torch.zeros((ticks, pad, 2), dtype=torch.uint8, device='cuda')
    dmv[:, :n] = torch.from_numpy(mv).cuda()
    res = torch.zeros((ticks, pad), dtype=torch.uint8, device='cuda')
    torch.cuda.synchronize()
    for t in range(ticks):
        upd.update(gs, dmv[t, :n], out=res[t, :n])
    torch.cuda.synchronize()
    want = _oracle_run(orc, mv)
    assert np.array_equal(res[:, :n].cpu().numpy(), want)
    gu.assert_state_equal(gs, orc, 'after back-to-back ticks')
    # the hand-over words are balanced again: tickets handed out == passes completed == ticks for every chunk in use
    # (one chunk per CTA as shipped), zero beyond
    w = gs.sched.cpu().numpy()[_abi.SCHED_HEADER_WORDS:]
    used = int((w != 0).sum())
    assert used >= 2 and used % 2 == 0 and (w[:used] == ticks).all() and (w[used:] == 0).all()
    if tpc:
        assert used == 2 * -(-(n // _abi.TILE) // tpc)


@pytest.mark.parametrize('n,ticks', [(1 << 18, 600), (1 << 20, 160), (1 << 21, 80), (1 << 22, 40)])     # up to the largest batch ticked in this mode (16,384 tiles)
def test_long_unsynchronised_runs(n, ticks):
    """Hundreds of ticks in flight behind each other (a CUDA graph replayed without a pause), the state planes of
    the bigger batch far larger than what is in flight: the end state equals the oracle's after the same commands."""
    cfg = SimConfig(max_ticks=97, seed=4242, auto_reset=True, overlap_ticks=True)
    gs, upd, orc = gu.make_pair(cfg, n)
    per = 40
    g = torch.Generator(device='cuda').manual_seed(n)
    dmv = torch.randint(1, 6, (per, n, 2), dtype=torch.uint8, device='cuda', generator=g)
    res = torch.zeros((per, n), dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        upd.update(gs, dmv[0], out=res[0])
        upd.update(gs, dmv[1], out=res[1])
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for t in range(per):
                upd.update(gs, dmv[t], out=res[t])
        for _ in range(ticks // per):
            graph.replay()
    torch.cuda.synchronize()
    mv = dmv.cpu().numpy()
    _oracle_run(orc, mv[:2])
    want = None
    for _ in range(ticks // per):
        want = _oracle_run(orc, mv)
    assert np.array_equal(res.cpu().numpy(), want)
    gu.assert_state_equal(gs, orc, 'long run')


def test_graph_of_interleaved_states_replayed():
    """Three states ticked round-robin inside one CUDA graph (launches on different states do not wait for
    each other), the graph replayed several times back to back; each state equals its oracle."""
    n, per_replay, replays = 256 * 24, 15, 4
    cfgs = [SimConfig(max_ticks=40, seed=5 + k, auto_reset=True, width=9, height=7, overlap_ticks=True) for k in range(3)]
    trios = [gu.make_pair(c, n, game_id_base=k * n) for k, c in enumerate(cfgs)]
    rng = np.random.default_rng(3)
    mv = _moves(rng, per_replay, n)
    dmv = torch.from_numpy(mv).cuda()
    res = [torch.zeros((per_replay, n), dtype=torch.uint8, device='cuda') for _ in trios]     # one result buffer per state
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for t in range(2):                      # warm-up outside the graph (these ticks count)
            for k, (gs, upd, _) in enumerate(trios):
                upd.update(gs, dmv[t], out=res[k][t])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for t in range(per_replay):
                for k, (gs, upd, _) in enumerate(trios):
                    upd.update(gs, dmv[t], out=res[k][t])
        for _ in range(replays):
            g.replay()
    torch.cuda.synchronize()
    for k, (gs, upd, orc) in enumerate(trios):
        _oracle_run(orc, mv[:2])
        want = None
        for _ in range(replays):
            want = _oracle_run(orc, mv)
        assert np.array_equal(res[k].cpu().numpy(), want), k
        gu.assert_state_equal(gs, orc, f'state {k}')


@pytest.mark.parametrize('n', [256 * 17, 256 * 17 + 5])
def test_other_kernels_between_flagged_ticks(n):
    """Observation passes (same pipeline, no tick), the fused step+observe, a fused rollout and a masked
    reset between unsynchronised ticks of one state: all ordered correctly against the tile hand-over."""
    cfg = SimConfig(max_ticks=33, seed=12, auto_reset=True, overlap_ticks=True)
    gs, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(8)
    mv = _moves(rng, 30, n)
    dmv = torch.from_numpy(mv).cuda()
    res = torch.zeros((30, n), dtype=torch.uint8, device='cuda')
    obs_a = torch.zeros((n, 2, _abi.OBS_LEN), dtype=torch.int16, device='cuda')
    obs_b = torch.zeros_like(obs_a)
    mask = torch.from_numpy((np.arange(n) % 5 == 0).astype(np.uint8)).cuda()
    torch.cuda.synchronize()
    for t in range(10):
        upd.update(gs, dmv[t], out=res[t])
    upd.observe(gs, 3, out=obs_a)                      # reads what tick 9 wrote, before tick 10 overwrites it
    for t in range(10, 20):
        upd.update(gs, dmv[t], out=res[t])
    upd.update_observe(gs, dmv[20], stairs_radius=3, out=res[20], obs_out=obs_b)
    upd.reset(gs, mask, bump_episode=True)
    upd.rollout(gs, 1, 2, 3)                           # fused ticks (RandomBot vs StaircaseBot), an ordinary launch
    for t in range(21, 30):
        upd.update(gs, dmv[t], out=res[t])
    torch.cuda.synchronize()
    want = np.empty((30, n), np.uint8)
    for t in range(10):
        want[t], _ = orc.step(mv[t], want_events=False)
    ref_a = gu.make_pair(cfg, n)                       # observation reference: the CUDA observe of an oracle-equal state
    for t in range(10, 21):
        want[t], _ = orc.step(mv[t], want_events=False)
    orc.reset(mask.cpu().numpy(), bump_episode=True)
    orc.rollout(1, 2, 3, np.zeros(8, np.uint64))
    for t in range(21, 30):
        want[t], _ = orc.step(mv[t], want_events=False)
    assert np.array_equal(res.cpu().numpy(), want)
    gu.assert_state_equal(gs, orc, 'mixed sequence')
    # observations: replay the same prefix synchronously on a second state and observe it there
    gs2, upd2, _ = ref_a
    for t in range(10):
        upd2.update(gs2, dmv[t])
        torch.cuda.synchronize()
    assert torch.equal(upd2.observe(gs2, 3), obs_a)
    for t in range(10, 21):
        upd2.update(gs2, dmv[t])
        torch.cuda.synchronize()
    assert torch.equal(upd2.observe(gs2, 3), obs_b)


def test_flag_mode_equals_grid_wait_mode():
    """The same command stream through both ordering modes (path flag ORX_PATH_TILE_FLAGS) and with the
    static / dynamic tile hand-out of grid-wait mode: identical planes and results."""
    n, ticks = 256 * 31 + 200, 40
    rng = np.random.default_rng(21)
    mv = torch.from_numpy(_moves(rng, ticks, n)).cuda()
    outs = []
    for flags in (_abi.PATH_TILE_FLAGS, 0, _abi.PATH_STATIC_TILES, _abi.PATH_NO_TENSOR_MAP, _abi.PATH_TILE_FLAGS | _abi.PATH_NO_TENSOR_MAP):
        cfg = SimConfig(max_ticks=25, seed=31, auto_reset=True, path_flags=flags)
        gs, upd, _ = gu.make_pair(cfg, n)
        res = torch.zeros((ticks, n), dtype=torch.uint8, device='cuda')
        for t in range(ticks):
            upd.update(gs, mv[t], out=res[t])
        torch.cuda.synchronize()
        outs.append((res.cpu(), gs.planes_cpu()))
    for res, planes in outs[1:]:
        assert torch.equal(res, outs[0][0])
        for name, p in planes.items():
            assert np.array_equal(p, outs[0][1][name]), name


def test_short_scratch_falls_back_to_grid_wait():
    """A caller that asks for the throughput mode but only provides the 4 header words (the round-1 layout) gets
    grid-wait mode and correct ticks."""
    n = 256 * 9
    cfg = SimConfig(max_ticks=25, seed=2, auto_reset=True, overlap_ticks=True)
    gs, upd, orc = gu.make_pair(cfg, n)
    gs.sched = torch.zeros((_abi.SCHED_HEADER_WORDS,), dtype=torch.int32, device='cuda')
    rng = np.random.default_rng(1)
    mv = _moves(rng, 20, n)
    dmv = torch.from_numpy(mv).cuda()
    res = torch.zeros((20, n), dtype=torch.uint8, device='cuda')
    for t in range(20):
        upd.update(gs, dmv[t], out=res[t])
    torch.cuda.synchronize()
    assert np.array_equal(res.cpu().numpy(), _oracle_run(orc, mv))
    gu.assert_state_equal(gs, orc, 'header-only scratch')


# ---------------------------------------------------------------------------------------------------------
# bit-packed streams (orx_step_bits / orx_step_host_bits): 5 bits of command pair in, 2 bits of result out
@pytest.mark.parametrize('n', [256 * 37, 256 * 37 + 91, 77, 131072 + 3])
@pytest.mark.parametrize('flags', [_abi.PATH_TILE_FLAGS, 0])
def test_bit_packed_streams_match_oracle(n, flags):
    """Device-resident cmd5 / res2 streams, unsynchronised ticks; invalid codes are packed as Stay. Planes and
    the unpacked results equal the oracle's (which is fed the plain codes)."""
    from optimax_rogue_b200.logic.moves import pack_moves5, unpack_results2
    ticks = 24
    cfg = SimConfig(max_ticks=19, seed=314, auto_reset=True, width=8, height=7, path_flags=flags)
    gs, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(n)
    mv = _moves(rng, ticks, n)
    nb_in, nb_out = _abi.cmd5_bytes(n), _abi.res2_bytes(n)
    pitch_in, pitch_out = -(-nb_in // 16) * 16, -(-nb_out // 16) * 16
    cmd = torch.zeros((ticks, pitch_in), dtype=torch.uint8)
    for t in range(ticks):
        cmd[t, :nb_in] = torch.from_numpy(pack_moves5(mv[t, :, 0], mv[t, :, 1]))
    cmd = cmd.cuda()
    res = torch.full((ticks, pitch_out), 0xEE, dtype=torch.uint8, device='cuda')
    torch.cuda.synchronize()
    for t in range(ticks):
        upd.update_bits(gs, cmd[t, :nb_in], out=res[t, :nb_out])
    torch.cuda.synchronize()
    want = _oracle_run(orc, mv)
    got = res.cpu().numpy()
    for t in range(ticks):
        assert np.array_equal(unpack_results2(got[t, :nb_out], n), want[t]), t
        if (2 * n) % 8:                                   # padding bits of the last byte are zero
            assert got[t, nb_out - 1] >> ((2 * n) % 8) == 0
    gu.assert_state_equal(gs, orc, 'bit-packed streams')


@pytest.mark.parametrize('n', [256 * 50, 256 * 9 + 17])
@pytest.mark.parametrize('staged', [False, True])
def test_bit_packed_host_buffers(n, staged):
    """The host-buffer tick with bit-packed streams: pinned buffers are read / written by the kernel over PCIe
    (one bulk copy per CTA each way), ORX_PATH_HOST_STAGED forces the staged-copy path pageable buffers take."""
    from optimax_rogue_b200.logic.moves import pack_moves5, unpack_results2
    cfg = SimConfig(max_ticks=23, seed=99, auto_reset=True, path_flags=_abi.PATH_HOST_STAGED if staged else 0)
    gs, upd, orc = gu.make_pair(cfg, n)
    rng = np.random.default_rng(5)
    cmd_host = torch.zeros((_abi.cmd5_bytes(n),), dtype=torch.uint8).pin_memory()
    res_host = torch.zeros((_abi.res2_bytes(n),), dtype=torch.uint8).pin_memory()
    step = upd.host_stepper(gs, cmd_host, res_host, sync=True, bits=True)
    for t in range(30):
        mv = rng.integers(1, 6, size=(n, 2), dtype=np.uint8)
        cmd_host.copy_(torch.from_numpy(pack_moves5(mv[:, 0], mv[:, 1])))
        step()                                            # synchronous: res_host is valid on return
        want, _ = orc.step(mv, want_events=False)
        assert np.array_equal(unpack_results2(res_host.numpy(), n), want), t
    gu.assert_state_equal(gs, orc, 'host bit-packed')
