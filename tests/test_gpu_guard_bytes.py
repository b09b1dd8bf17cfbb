"""-m gpu: no kernel writes outside the buffers it was given. compute-sanitizer is closed on the GPU pool, so this is the
repo's own bounds check: every plane of the state, the scratch words and every command / result / event / observation
buffer lives inside a larger allocation filled with a guard pattern (4 KB before and after, and the padding between the
five word planes that travel as one tensor-map box), every kernel family is run over it -- full tiles and ragged
tails, both orderings of consecutive launches, byte / nibble / bit-packed commands, event log, observations, NPC slots,
fixed maps, bots, rollout, replay, reset, ruleset R1 -- and afterwards every guard byte must still hold the pattern,
while the results equal those of an ordinary, unguarded state. (The tick is Updater.update,
optimax_rogue/logic/updater.py:76-162; the TMA / bulk-copy pipeline is csrc/orx_pipe.cuh.)"""
import ctypes as C

import numpy as np
import pytest
import torch

from optimax_rogue_b200 import SimConfig, _abi, _lib
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.moves import pack_moves, pack_moves5
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator, FixedDungeonGenerator
from optimax_rogue_b200.r1 import R1GameState

pytestmark = pytest.mark.gpu

GUARD, PAT = 4096, 0xA5


class Guards:
    """Allocations filled with PAT; payload regions are handed out as views, everything else must stay PAT."""

    def __init__(self):
        self.bufs = []

    def regions(self, total, regions):
        buf = torch.full((GUARD + total + GUARD,), PAT, dtype=torch.uint8, device='cuda')
        self.bufs.append((buf, [(GUARD + off, nb) for off, nb in regions]))
        views = []
        for off, nb in regions:
            v = buf[GUARD + off:GUARD + off + nb]
            v.zero_()
            views.append(v)
        return views

    def like(self, t):
        """A guarded tensor with the dtype / shape / contents of ``t``."""
        nb = t.numel() * t.element_size()
        v = self.regions((nb + 127) // 128 * 128, [(0, nb)])[0].view(t.dtype).view(t.shape)
        v.copy_(t)
        return v

    def empty(self, shape, dtype):
        return self.like(torch.zeros(shape, dtype=dtype, device='cuda'))

    def check(self, where):
        torch.cuda.synchronize()
        for k, (buf, regions) in enumerate(self.bufs):
            mask = torch.ones_like(buf, dtype=torch.bool)
            for off, nb in regions:
                mask[off:off + nb] = False
            bad = torch.nonzero(mask & (buf != PAT))
            assert bad.numel() == 0, f'{where}: guard bytes of buffer {k} overwritten at byte offsets {bad[:8].flatten().tolist()} (payload regions {regions})'


def rehome(gs, g):
    """Moves every plane of a BatchedGameState into guarded memory: the five word planes at a common pitch with 128
    guard bytes between consecutive planes (still one tensor-map box per tile), the others one allocation each."""
    n = gs.n
    pitch = (4 * n + 127) // 128 * 128 + 128
    views = g.regions(5 * pitch, [(k * pitch, 4 * n) for k in range(5)])
    for v, (name, dtype, shape) in zip(views, gs.WORD_PLANES):
        new = v.view(dtype).view((n,) + shape)
        new.copy_(getattr(gs, name))
        setattr(gs, name, new)
    for name in ('depth', 'status', 'npc_pos', 'npc_hp', 'npc_depth', 'sched'):
        setattr(gs, name, g.like(getattr(gs, name)))
    if gs.flat is not None:
        gs.flat = g.like(gs.flat)
    return gs


def fixed_map():
    t = np.full((20, 9), 1, np.uint8)
    t[[0, -1], :] = 2
    t[:, [0, -1]] = 2
    t[5, 4] = 2
    t[10, 3] = 3
    return t


CASES = [
    dict(n=256 * 9, overlap=False), dict(n=256 * 9, overlap=True), dict(n=256 * 5 + 77, overlap=False), dict(n=256 * 40 + 13, overlap=True),
    dict(n=199, overlap=False), dict(n=256 * 6 + 1, overlap=False, npc=1), dict(n=256 * 6 + 31, overlap=True, npc=3), dict(n=256 * 4, overlap=False, npc=8),
    dict(n=256 * 7 + 5, overlap=False, fixed=True), dict(n=256 * 7, overlap=True, fixed=True, npc=2), dict(n=256 * 8 + 16, overlap=True, flat=True),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join(f'{k}{v}' for k, v in c.items()))
def test_r0_kernels_stay_inside_their_buffers(case):
    n, npc = case['n'], case.get('npc', 0)
    kw = dict(max_ticks=23, seed=31, auto_reset=True, overlap_ticks=case['overlap'], n_npc=npc)
    if case.get('fixed'):
        t = fixed_map()
        kw.update(width=20, height=9, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t)
        dgen = FixedDungeonGenerator(t)
    else:
        kw.update(width=13, height=7)
        dgen = EmptyDungeonGenerator(13, 7)
    cfg = SimConfig(**kw)
    g = Guards()
    gs, twin = BatchedGameState(cfg, n, 'cuda'), BatchedGameState(cfg, n, 'cuda')
    if case.get('flat'):
        for s in (gs, twin):
            s.enable_flat_bonuses().copy_(torch.randint(-1, 3, (n, 2, 3), dtype=torch.int8, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5)))
    rehome(gs, g)
    upd, upd2 = (BatchedUpdater(dgen, 1, 23, auto_reset=True) for _ in range(2))
    upd.track_order = upd2.track_order = False
    for s in (gs, twin):
        reset_games(s)
        for k in range(min(npc, 2)):
            s.set_npc(slice(None), k, 0, 2 + k, 2, 3)
    g.check('after reset')
    rng = np.random.default_rng(n)
    lib = _lib.lib()
    E = _abi.MAX_EVENTS_BASE + npc
    res, res2 = g.empty((n,), torch.uint8), torch.zeros((n,), dtype=torch.uint8, device='cuda')
    ev = g.empty((n, E, 2), torch.int32)
    obs = g.empty((n, 2, _abi.OBS_LEN), torch.int16)
    mvg, nib = g.empty((n, 2), torch.uint8), g.empty((n,), torch.uint8)
    bits_in, bits_out = g.empty((_abi.cmd5_bytes(n),), torch.uint8), g.empty((_abi.res2_bytes(n),), torch.uint8)
    for t in range(36):
        mv = rng.integers(0, 8, size=(n, 2), dtype=np.uint8)
        mvg.copy_(torch.from_numpy(mv))
        kind = t % 9
        if kind == 0:                                  # byte commands
            upd.update(gs, mvg, out=res); upd2.update(twin, mvg, out=res2)
        elif kind == 1:                                # + event log (the caller's guarded buffer, through the C ABI directly)
            c, st = upd._cfg(gs)
            rc = lib.orx_step(C.byref(c), C.byref(st), mvg.data_ptr(), res.data_ptr(), ev.data_ptr(), n, gs.game_id_base,
                              torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            _, ev2 = upd2.update(twin, mvg, want_events=True, out=res2)
            assert torch.equal(ev, ev2), f'tick {t}: events'
        elif kind == 2:                                # nibble-packed commands
            nib.copy_(torch.from_numpy(pack_moves(mv[:, 0], mv[:, 1])))
            upd.update(gs, nib, out=res, packed=True); upd2.update(twin, nib, out=res2, packed=True)
        elif kind == 3 and npc == 0:                   # bit-packed streams
            bits_in.copy_(torch.from_numpy(pack_moves5(mv[:, 0], mv[:, 1])))
            upd.update_bits(gs, bits_in, out=bits_out)
            assert torch.equal(bits_out, upd2.update_bits(twin, bits_in)), f'tick {t}: bit-packed results'
            continue
        elif kind == 4 and npc == 0:                   # tick + observations in one pass
            upd.update_observe(gs, mvg, stairs_radius=3, out=res, obs_out=obs)
            _, o2 = upd2.update_observe(twin, mvg, stairs_radius=3, out=res2)
            assert torch.equal(obs, o2), f'tick {t}: observations'
        elif kind == 5 and npc == 0:                   # scripted players inside the tick
            upd.update_with_bots(gs, mvg, _abi.BOT_STAIRCASE, _abi.BOT_RANDOM, out=res)
            upd2.update_with_bots(twin, mvg, _abi.BOT_STAIRCASE, _abi.BOT_RANDOM, out=res2)
        elif kind == 6:                                # bot commands, observations alone, masked reset
            upd.bot_moves(gs, 2, 1, out=mvg)
            upd.observe(gs, 2, out=obs)
            mask = torch.from_numpy((rng.integers(0, 4, size=n) == 0).astype(np.uint8)).cuda()
            reset_games(gs, mask, True); reset_games(twin, mask, True)
            continue
        elif kind == 7:                                # fused rollout, then a replay of queued commands
            upd.rollout(gs, 1, 2, 5); upd2.rollout(twin, 1, 2, 5)
            q = g.like(torch.from_numpy(rng.integers(0, 8, size=(3, n, 2), dtype=np.uint8)).cuda())
            r = g.empty((3, n), torch.uint8)
            upd.replay(gs, q, out=r)
            assert torch.equal(r, upd2.replay(twin, q)), f'tick {t}: replay'
            continue
        else:
            upd.update(gs, mvg, out=res); upd2.update(twin, mvg, out=res2)
        assert torch.equal(res, res2), f'tick {t} (kind {kind}): results'
    g.check('after the ticks')
    a, b = gs.planes_cpu(), twin.planes_cpu()
    for name in a:
        assert np.array_equal(a[name], b[name]), name
    used = gs.sched.cpu().numpy()
    assert used[0] == 0                               # the tile counter is back at zero


@pytest.mark.parametrize('n,flags', [(128 * 9, 0), (128 * 9 + 5, _abi.R1_PATH_BLOCK_FLAGS), (77, 0), (16 * 40 + 3, _abi.R1_PATH_HALFWARP)])
def test_r1_kernels_stay_inside_their_buffers(n, flags):
    kw = dict(width=12, height=8, wall_density=15, seed=9, max_ticks=40, auto_reset=True, path_flags=flags)
    gs, twin = R1GameState(n, **kw), R1GameState(n, **kw)
    g = Guards()
    for name, _, _ in _abi.R1_PLANES:
        new = g.like(getattr(gs, name))
        setattr(gs, name, new)
        setattr(gs._st, name, new.data_ptr())
    gs.sched = g.like(gs.sched)
    gs._st.sched = gs.sched.data_ptr()
    gs.reset(); twin.reset()
    g.check('after reset')
    rng = np.random.default_rng(n)
    mvg, res = g.empty((n, 2), torch.uint8), g.empty((n,), torch.uint8)
    ev = g.empty((n, 7, 2), torch.int32)
    obs = g.empty((n, 2, _abi.R1_OBS_LEN), torch.int16)
    for t in range(40):
        mvg.copy_(torch.from_numpy(rng.integers(0, 8, size=(n, 2), dtype=np.uint8)))
        kind = t % 5
        if kind == 1 and not flags & _abi.R1_PATH_HALFWARP:
            gs.update_events(mvg, out=res, events=ev)
            r2, e2 = twin.update_events(mvg, max_events=7)
            live = ((ev[:, :, 0] & 0xFF) == 0).int().cumsum(1) == 0      # records ahead of each game's terminator
            assert torch.equal(live, ((e2[:, :, 0] & 0xFF) == 0).int().cumsum(1) == 0), f'tick {t}: event counts'
            assert torch.equal(ev[live], e2[live]), f'tick {t}: events'
        elif kind == 2:
            gs.bot_moves(_abi.BOT_RANDOM, _abi.BOT_STAIRCASE, out=mvg)
            gs.update(mvg, out=res); r2 = twin.update(mvg)
        elif kind == 3:
            gs.observe(3, out=obs)
            gs.rollout(4); twin.rollout(4)
            q = g.like(torch.from_numpy(rng.integers(1, 7, size=(3, n, 2), dtype=np.uint8)).cuda())
            r = g.empty((3, n), torch.uint8)
            gs.replay(q, out=r)
            assert torch.equal(r, twin.replay(q)), f'tick {t}: replay'
            continue
        else:
            gs.update(mvg, out=res); r2 = twin.update(mvg)
        assert torch.equal(res, r2), f'tick {t} (kind {kind}): results'
    g.check('after the ticks')
    a, b = gs.planes_cpu(), twin.planes_cpu()
    for name in a:
        assert np.array_equal(a[name], b[name]), name


def test_the_guard_check_itself_detects_an_overrun():
    """Negative control: one byte written just past a payload, or just ahead of it, fails the check."""
    for where in (-1, 0):
        g = Guards()
        v = g.empty((100,), torch.uint8)
        g.check('clean')
        base = v.untyped_storage()
        raw = torch.empty(0, dtype=torch.uint8, device='cuda').set_(base)
        off = v.storage_offset()
        raw[off + (100 if where == 0 else -1)] = 0
        with pytest.raises(AssertionError):
            g.check('dirty')
