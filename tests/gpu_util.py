"""Helpers for the -m gpu parity tests: drive the CUDA path (through the C ABI, via the
BatchedUpdater host mirror) and the C oracle with the same seeded inputs."""
import numpy as np
import torch

from oracle import cport
from optimax_rogue_b200 import _abi
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import (EmptyDungeonGenerator, FixedDungeonGenerator)

PLANES = ('pos', 'hp', 'depth', 'stairs', 'tick', 'episode', 'status')
NPC_PLANES = ('npc_pos', 'npc_hp', 'npc_depth')


def make_pair(cfg, n, game_id_base=0, device='cuda'):
    """(BatchedGameState, BatchedUpdater, Oracle) for one SimConfig, all freshly reset."""
    gs = BatchedGameState(cfg, n, device, game_id_base)
    reset_games(gs)
    dgen = (FixedDungeonGenerator(cfg.fixed_tiles) if cfg.dgen_kind == _abi.DGEN_FIXED
            else EmptyDungeonGenerator(cfg.width, cfg.height))
    upd = BatchedUpdater(dgen, cfg.despawn_strat, cfg.max_ticks or None, auto_reset=cfg.auto_reset)
    orc = cport.Oracle(cfg, n, game_id_base)
    orc.reset()
    return gs, upd, orc


def assert_state_equal(gs, orc, where=''):
    p = gs.planes_cpu()
    names = PLANES + (NPC_PLANES if gs.cfg.n_npc else ())
    for name in names:
        a = p[name]
        b = getattr(orc.state, name)
        if name == 'episode':
            a = a.view(np.uint32)
        if not np.array_equal(a, b):
            bad = np.argwhere(a.reshape(a.shape[0], -1) != b.reshape(b.shape[0], -1))[0][0]
            raise AssertionError(f'{where}: plane {name} differs at lane {bad}: cuda={a[bad]} oracle={b[bad]}')


def run_parity(cfg, n, ticks, bots=(1, 1), events=True, game_id_base=0, moves_fn=None, check_every=1,
               setup=None):
    gs, upd, orc = make_pair(cfg, n, game_id_base)
    if setup is not None:
        setup(gs, orc)
    assert_state_equal(gs, orc, 'after reset')
    for t in range(ticks):
        if moves_fn is not None:
            mv = moves_fn(t, orc)
        else:
            mv = orc.bot_moves(bots[0], bots[1])
            gmv = upd.bot_moves(gs, bots[0], bots[1])
            assert np.array_equal(gmv.cpu().numpy(), mv), f'tick {t}: bot moves differ'
        res_o, ev_o = orc.step(mv, want_events=events)
        res_g, ev_g = upd.update(gs, torch.from_numpy(mv).to(gs.device), want_events=events)
        if t % check_every == 0 or t == ticks - 1:
            assert np.array_equal(res_g.cpu().numpy(), res_o), f'tick {t}: results differ'
            if events:
                assert np.array_equal(ev_g.cpu().numpy(), ev_o), f'tick {t}: events differ'
            assert_state_equal(gs, orc, f'tick {t}')
    return gs, upd, orc


def run_burst(cfg, n, ticks, game_id_base=0, setup=None, bits=False, seed=0):
    """``ticks`` ticks enqueued back to back WITHOUT synchronisation (random command bytes, invalid codes included),
    results into a [ticks, n] buffer, then every result and the final planes against the oracle. With
    ``cfg.overlap_ticks`` consecutive launches overlap chunk by chunk; ``bits``: the bit-packed streams."""
    from optimax_rogue_b200.logic.moves import pack_moves5, unpack_results2
    gs, upd, orc = make_pair(cfg, n, game_id_base)
    if setup is not None:
        setup(gs, orc)
    rng = np.random.default_rng(seed)
    mv = rng.integers(0, 8, size=(ticks, n, 2), dtype=np.uint8)
    if bits:
        nb_in, nb_out = _abi.cmd5_bytes(n), _abi.res2_bytes(n)
        pin, pout = -(-nb_in // 16) * 16, -(-nb_out // 16) * 16
        cmd = torch.zeros((ticks, pin), dtype=torch.uint8)
        for t in range(ticks):
            cmd[t, :nb_in] = torch.from_numpy(pack_moves5(mv[t, :, 0], mv[t, :, 1]))
        cmd = cmd.to(gs.device)
        res = torch.zeros((ticks, pout), dtype=torch.uint8, device=gs.device)
        torch.cuda.synchronize()
        for t in range(ticks):
            upd.update_bits(gs, cmd[t, :nb_in], out=res[t, :nb_out])
    else:
        pad = -(-n // 16) * 16
        dmv = torch.zeros((ticks, pad, 2), dtype=torch.uint8, device=gs.device)
        dmv[:, :n] = torch.from_numpy(mv).to(gs.device)
        res = torch.zeros((ticks, pad), dtype=torch.uint8, device=gs.device)
        torch.cuda.synchronize()
        for t in range(ticks):
            upd.update(gs, dmv[t, :n], out=res[t, :n])
    torch.cuda.synchronize()
    got = res.cpu().numpy()
    for t in range(ticks):
        want, _ = orc.step(mv[t], want_events=False)
        have = unpack_results2(got[t, :_abi.res2_bytes(n)], n) if bits else got[t, :n]
        assert np.array_equal(have, want), f'burst tick {t}: results differ'
    assert_state_equal(gs, orc, 'after the burst')
    return gs, upd, orc
