"""One-off parity campaign on a GPU box: many random configurations, CUDA (through the C ABI) against the pinned
plain-C oracle on every plane, result and event record, every tick. Larger and longer than the hypothesis test in
tests/test_properties.py (batches of up to a few hundred thousand games, so the multi-tile, dynamically scheduled,
tensor-map and event-pipeline paths are all exercised). The oracle is test infrastructure; this is a test tool.

  python tests/fuzz_campaign.py [seconds] [seed]      (lives under tests/: it calls the oracle, which only tests may do)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                                         # noqa: E402
import torch                                               # noqa: E402
from optimax_rogue_b200 import SimConfig, _abi             # noqa: E402
import gpu_util as gu                                      # noqa: E402

from oracle import cport                                   # noqa: E402
from optimax_rogue_b200.r1 import R1GameState              # noqa: E402


def r1_case(rng):
    """One random ruleset-R1 configuration (parity unpinned: CUDA against oracle/orx_r1_oracle.c, the written spec):
    every plane after the run, every result, and -- in the event-log variant -- every record of every tick."""
    w, h = int(rng.integers(5, 65)), int(rng.integers(5, 21))
    kw = dict(width=w, height=h, wall_density=int(rng.integers(0, 70)), seed=int(rng.integers(0, 2**63)),
              max_ticks=int(rng.integers(0, 120)), auto_reset=bool(rng.integers(0, 2)))
    flavour = int(rng.integers(0, 3))
    flags = (0, _abi.R1_PATH_BLOCK_FLAGS, _abi.R1_PATH_HALFWARP)[flavour]
    n = int(rng.integers(1, 500)) if rng.integers(0, 3) == 0 else int(rng.integers(500, 30000))
    base = int(rng.integers(0, 2**53))
    ticks = int(rng.integers(20, 100))
    gs = R1GameState(n, game_id_base=base, path_flags=flags, **kw).reset()
    orc = cport.R1Oracle(n, game_id_base=base, **kw)
    orc.reset()
    mode = int(rng.integers(0, 4)) if flavour != 2 else int(rng.integers(0, 2))
    mrng = np.random.default_rng(int(rng.integers(0, 2**31)))
    if mode == 3:                                               # queued commands: one replay launch against T oracle steps
        mv = mrng.integers(0, 8, size=(ticks, n, 2), dtype=np.uint8)
        want = np.stack([orc.step(mv[t]) for t in range(ticks)])
        assert np.array_equal(gs.replay(torch.from_numpy(mv).cuda()).cpu().numpy(), want), 'R1 replay'
        what = 'replay'
    elif mode == 2:                                             # the tick with its replication log
        ev = torch.empty((n, _abi.R1_MAX_EVENTS, 2), dtype=torch.int32, device='cuda')
        for t in range(ticks):
            mv = mrng.integers(0, 8, size=(n, 2), dtype=np.uint8)
            ro, eo = orc.step_events(mv)
            ev.fill_(0)
            rg, _ = gs.update_events(torch.from_numpy(mv).cuda(), events=ev)
            assert np.array_equal(rg.cpu().numpy(), ro), f'R1 tick {t}: results'
            got = ev.cpu().numpy()
            live = np.cumsum((got[:, :, 0] & 0xFF) == 0, axis=1) == 0
            assert np.array_equal(live, np.cumsum((eo[:, :, 0] & 0xFF) == 0, axis=1) == 0) and np.array_equal(got[live], eo[live]), f'R1 tick {t}: records'
        what = 'event log'
    elif mode == 1:                                             # scripted players, a burst of unsynchronised ticks
        b1, b2 = int(mrng.integers(1, 3)), int(mrng.integers(1, 3))
        res = torch.empty((ticks, n), dtype=torch.uint8, device='cuda')
        for t in range(ticks):
            gs.update(gs.bot_moves(b1, b2), out=res[t])
        want = np.stack([orc.step(orc.bot_moves(b1, b2)) for _ in range(ticks)])
        assert np.array_equal(res.cpu().numpy(), want), 'R1 bots'
        what = 'bots, burst'
    else:
        for t in range(ticks):
            mv = mrng.integers(0, 8, size=(n, 2), dtype=np.uint8)
            assert np.array_equal(gs.update(torch.from_numpy(mv).cuda()).cpu().numpy(), orc.step(mv)), f'R1 tick {t}'
        what = 'random bytes'
    p = gs.planes_cpu()
    for name, _, _ in _abi.R1_PLANES:
        assert np.array_equal(p[name], getattr(orc.state, name)), f'R1 plane {name}'
    return ('ruleset R1 (parity unpinned)', what, ('thread per game', 'thread per game, throughput mode', 'half warp')[flavour]), n, ticks


budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
t0 = time.time()
runs = games = ticks_total = 0
kinds = {}
while time.time() - t0 < budget:
    if rng.integers(0, 5) == 0:                                 # every fifth configuration: ruleset R1
        key, n, n_ticks = r1_case(rng)
        kinds[key] = kinds.get(key, 0) + 1
        runs += 1
        games += n
        ticks_total += n * n_ticks
        continue
    w, h = int(rng.integers(4, 65)), int(rng.integers(4, 21))
    kw = dict(width=w, height=h, hp=(int(rng.integers(1, 10)), int(rng.integers(1, 10))),
              damage=(int(rng.integers(0, 5)), int(rng.integers(0, 5))), armor=(int(rng.integers(0, 3)), int(rng.integers(0, 3))),
              despawn_strat=int(rng.integers(1, 3)), max_ticks=int(rng.integers(0, 90)), auto_reset=bool(rng.integers(0, 2)),
              seed=int(rng.integers(0, 2**63)))
    if rng.integers(0, 3) == 0:
        kw.update(start_kind=_abi.START_SEPARATED, start_depth=(0, int(rng.integers(1, 5))))
    fixed = rng.integers(0, 4) == 0
    if fixed:
        t = np.full((w, h), 1, np.uint8)
        t[[0, -1], :] = 2
        t[:, [0, -1]] = 2
        inner = rng.random((w - 2, h - 2)) < 0.12
        t[1:-1, 1:-1][inner] = 2
        g = np.argwhere(t == 1)
        if len(g) < 4:
            continue
        if rng.integers(0, 2):
            t[tuple(g[len(g) // 2])] = 3
        kw.update(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t)
    n_npc = int(rng.integers(1, 9)) if rng.integers(0, 4) == 0 else 0
    if n_npc and (w - 2) * (h - 2) - 1 >= 2 + n_npc:
        kw.update(n_npc=n_npc)
    else:
        n_npc = 0
    kw.update(overlap_ticks=bool(rng.integers(0, 2)))          # throughput mode: consecutive launches ordered chunk by chunk
    cfg = SimConfig(**kw)
    size_class = int(rng.integers(0, 4))
    n = int(rng.integers(1, 700)) if size_class == 0 else int(rng.integers(700, 40000)) if size_class < 3 else int(rng.integers(40000, 300000))
    n_ticks = int(rng.integers(10, 60)) if n > 40000 else int(rng.integers(20, 120))
    bots = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
    events = bool(rng.integers(0, 2))
    base = int(rng.integers(0, 2**53))
    setup = None
    if n_npc:
        srng = np.random.default_rng(int(rng.integers(0, 2**31)))

        def setup(gs, orc, srng=srng, n_npc=n_npc, w=w, h=h):      # static NPCs on the first levels (updater.py:116-128)
            live = srng.random((orc.n, n_npc)) < 0.6
            orc.state.npc_depth[:] = np.where(live, srng.integers(0, 3, size=(orc.n, n_npc)), -1)
            orc.state.npc_pos[:, :, 0] = srng.integers(1, w - 1, size=(orc.n, n_npc))
            orc.state.npc_pos[:, :, 1] = srng.integers(1, h - 1, size=(orc.n, n_npc))
            orc.state.npc_hp[:] = srng.integers(1, 6, size=(orc.n, n_npc))
            gs.npc_depth.copy_(torch.from_numpy(orc.state.npc_depth))
            gs.npc_pos.copy_(torch.from_numpy(orc.state.npc_pos))
            gs.npc_hp.copy_(torch.from_numpy(orc.state.npc_hp))
    flat = rng.integers(0, 3) == 0                              # the Modifier seam: random flat bonuses per game and player
    if flat:
        frng = np.random.default_rng(int(rng.integers(0, 2**31)))
        inner = setup

        def setup(gs, orc, inner=inner, frng=frng):
            if inner is not None:
                inner(gs, orc)
            f = frng.integers(-3, 7, size=(orc.n, 2, 3)).astype(np.int8)
            orc.state.enable_flat_bonuses()[:] = f
            gs.enable_flat_bonuses().copy_(torch.from_numpy(f))
    mode = int(rng.integers(0, 4))
    if mode == 3:                                               # a burst of unsynchronised ticks (overlapping launches when in throughput mode)
        bits = n_npc == 0 and bool(rng.integers(0, 2))
        gu.run_burst(cfg, n, min(n_ticks, 48), game_id_base=base, setup=setup, bits=bits, seed=int(rng.integers(0, 2**31)))
        kind, events = 'burst, bit-packed streams' if bits else 'burst', False
    elif mode == 2:
        gu.run_parity(cfg, n, n_ticks, bots=bots, events=events, game_id_base=base, setup=setup)
        kind = 'bots'
    else:
        mrng = np.random.default_rng(int(rng.integers(0, 2**31)))
        gu.run_parity(cfg, n, n_ticks, events=events, game_id_base=base, setup=setup,
                      moves_fn=lambda t, orc: mrng.integers(0, 8, size=(orc.n, 2), dtype=np.uint8))
        kind = 'random bytes'
    key = (kind, 'fixed map' if fixed else 'empty rooms', 'events' if events else 'no events', 'NPC slots' if n_npc else 'players only',
           'throughput mode' if cfg.overlap_ticks else 'grid-wait mode', 'flat bonuses' if flat else 'no modifiers')
    kinds[key] = kinds.get(key, 0) + 1
    runs += 1
    games += n
    ticks_total += n * n_ticks
print(f'fuzz campaign: {runs} configurations, {games} games, {ticks_total} game-ticks compared bit for bit against the oracle '
      f'in {time.time() - t0:.0f} s on {torch.cuda.get_device_name(0)}: all equal')
for k in sorted(kinds):
    print('  ', ' / '.join(k), kinds[k])
