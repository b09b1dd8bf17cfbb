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
