"""-m gpu: the CUDA path against the LIVE reference in ONE process (no fixture, no C oracle in between).

The unmodified reference modules (``optimax_rogue.logic.updater.Updater`` & co.) are imported from
``/root/reference`` in the build container, or from the verbatim copy ``oracle/make_ref.py`` ships to the GPU
box (``oracle/_ref``, git-ignored); ``oracle/ref_harness`` injects the shared Philox draws into the reference's
``random`` / ``numpy.random`` calls. Each episode is played by the reference, tick by tick, and by the CUDA
kernels (bots on device, one lane per episode); position, depth, health of both players, tick, result, the
staircases of the occupied levels and the ordered ``GameStateUpdate`` records are folded into one digest per
episode on each side and must be equal. Skipped when no reference tree is present."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from optimax_rogue_b200 import SimConfig, _abi
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

import trace_util as tu

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rh.reference_available(), reason='no reference tree (run oracle/make_ref.py where /root/reference exists)')]

SEED = 0x5EED1E55


@pytest.mark.parametrize('bots,despawn,start,episodes,max_ticks', [
    (('staircase', 'random'), 'unreachable', 'together', 512, 96),
    (('random', 'random'), 'unreachable', 'together', 384, 128),
    (('staircase', 'staircase'), 'unused', 'together', 256, 64),
    (('random', 'staircase'), 'unreachable', 'separated', 128, 64),
])
def test_cuda_equals_live_reference(bots, despawn, start, episodes, max_ticks):
    gid0 = 7_000_000
    # --- the reference, one episode at a time (updater.py:76-162 under injected draws)
    want = np.empty(episodes, np.uint64)
    ticks = 0
    for k in range(episodes):
        trace, _ = rh.play_episode(SEED, gid0 + k, bots=bots, despawn=despawn, start=start, p_depths=(0, 6),
                                   max_ticks=max_ticks)
        want[k] = rh.digest(trace)
        ticks += len(trace) - 1
    # --- CUDA, all episodes at once
    cfg = SimConfig(seed=SEED, max_ticks=max_ticks, despawn_strat=1 if despawn == 'unreachable' else 2,
                    start_kind=_abi.START_SEPARATED if start == 'separated' else _abi.START_TOGETHER,
                    start_depth=(0, 6) if start == 'separated' else (0, 0))
    gs = BatchedGameState(cfg, episodes, 'cuda', game_id_base=gid0)
    reset_games(gs)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), cfg.despawn_strat, max_ticks)
    codes = [tu.BOT_CODES[b] for b in bots]
    dg = tu.BatchDigest(episodes)
    active = np.ones(episodes, bool)
    p = gs.planes_cpu()
    dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], p['status'], None, active)
    moves = torch.full((episodes, 2), 5, dtype=torch.uint8, device='cuda')
    gpu_ticks = 0
    for _ in range(max_ticks):
        upd.bot_moves(gs, codes[0], codes[1], out=moves)
        res, ev = upd.update(gs, moves, want_events=True)
        p = gs.planes_cpu()
        r = res.cpu().numpy()
        dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], r, ev.cpu().numpy(), active)
        gpu_ticks += int(active.sum())
        active &= r == 1
        if not active.any():
            break
    assert not active.any()
    assert gpu_ticks == ticks
    bad = np.flatnonzero(dg.h != want)
    assert len(bad) == 0, f'{len(bad)} of {episodes} episodes differ from the live reference, first {bad[:5]}'


def test_single_game_adapter_on_cuda_equals_reference_updater():
    """``SingleGameUpdater`` (the drop-in for the reference ``Updater`` on the reference's own GameState objects)
    with a real CUDA lane behind it, NPC entities included, next to the reference ``Updater`` under the same
    injected draws: equal results, equal GameStates (GameState.__eq__, game/state.py:134-153) and equal update
    lists, tick by tick."""
    from optimax_rogue_b200.logic.compat import SingleGameUpdater
    ref = rh.load_reference()
    seed, gid, W, H = 0xFEED, 4242, 7, 6
    inj = rh.Injector(seed)
    inj.game_id = gid
    with inj:
        rdgen = inj.wrap_dgen(ref.worldgen.EmptyDungeonGenerator(W, H))
        inj.site, inj.q = ('reset',), 0
        gs_ref = ref.worldgen.TogetherGameStartGenerator(rdgen).setup_game()
        inj.site = None
        taken = {(e.x, e.y) for e in gs_ref.entities} | {gs_ref.world.dungeons[0].staircase()}
        free = [(x, y) for x in range(1, W - 1) for y in range(1, H - 1) if (x, y) not in taken]
        for iden, (x, y), hp in zip((9, 4, 17), free[::3], (1, 2, 3)):
            gs_ref.add_entity(ref.entities.Entity(iden, 0, x, y, hp, hp, 0, 0, [], dict()))
        gs_ours = ref.state.GameState.from_prims(gs_ref.to_prims())
        upd_ref = inj.wrap_updater(ref.updater.Updater(rdgen, ref.updater.DungeonDespawningStrategy(1), 400), gs_ref)
        adapter = SingleGameUpdater(EmptyDungeonGenerator(W, H), 1, 400, seed=seed, game_id=gid, device='cuda',
                                    updates_module=ref.updates, world_module=ref.world, result_enum=ref.updater.UpdateResult)
        b = [ref.randombot.RandomBot(1), ref.randombot.RandomBot(2)]
        deaths = 0
        for t in range(400):
            inj.tick, inj.shuffle_calls = gs_ref.tick, 0
            gs_ref.on_tick()
            inj.choice_slot = 0
            m1 = b[0].move(gs_ref)
            inj.choice_slot = 1
            m2 = b[1].move(gs_ref)
            res_ref, ev_ref = upd_ref.update(gs_ref, m1, m2)
            res_ours, ev_ours = adapter.update(gs_ours, m1, m2)
            assert res_ours == res_ref
            assert gs_ours == gs_ref, f'tick {t}: GameStates differ'
            assert [(type(e), e.order) for e in ev_ours] == [(type(e), e.order) for e in ev_ref]
            for eo, er in zip(ev_ours, ev_ref):
                if isinstance(er, ref.updates.EntityCombatUpdate):
                    assert (eo.attacker_iden, eo.defender_iden, eo.og_damage) == (er.attacker_iden, er.defender_iden, er.og_damage)
                elif isinstance(er, ref.updates.EntityDeathUpdate):
                    assert eo.entity_iden == er.entity_iden
                    deaths += 1
            if res_ref != ref.updater.UpdateResult.InProgress:
                break
        assert deaths > 0


def test_cuda_equals_live_reference_with_flat_modifiers():
    """Entities that carry a Modifier with flat bonuses (the smallest concrete subclass, oracle/ref_harness.make_flat_modifier):
    the reference folds them into damage.value / armor.value in Entity.on_tick; the CUDA lane takes the same sums through
    OrxState.flat. One lane per episode, every episode with its own bonuses."""
    episodes, max_ticks, gid0 = 96, 150, 9_000_000
    rng = np.random.default_rng(3)
    flat = rng.integers(-2, 6, size=(episodes, 2, 3)).astype(np.int8)
    kw = dict(width=5, height=5, max_ticks=max_ticks, hp=(25, 30), damage=(2, 3), armor=(1, 1))
    want = np.empty(episodes, np.uint64)
    for k in range(episodes):
        trace, _ = rh.play_episode(SEED, gid0 + k, bots=('random', 'random'), flat=[tuple(int(v) for v in flat[k, p]) for p in range(2)], **kw)
        want[k] = rh.digest(trace)
    cfg = SimConfig(seed=SEED, **kw)
    gs = BatchedGameState(cfg, episodes, 'cuda', game_id_base=gid0)
    reset_games(gs)
    gs.enable_flat_bonuses().copy_(torch.from_numpy(flat))
    upd = BatchedUpdater(EmptyDungeonGenerator(5, 5), 1, max_ticks)
    dg = tu.BatchDigest(episodes)
    active = np.ones(episodes, bool)
    p = gs.planes_cpu()
    dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], p['status'], None, active)
    moves = torch.full((episodes, 2), 5, dtype=torch.uint8, device='cuda')
    for _ in range(max_ticks):
        upd.bot_moves(gs, 1, 1, out=moves)
        res, ev = upd.update(gs, moves, want_events=True)
        p = gs.planes_cpu()
        r = res.cpu().numpy()
        dg.update(p['pos'], p['hp'], p['depth'], p['stairs'], p['tick'], r, ev.cpu().numpy(), active)
        active &= r == 1
        if not active.any():
            break
    bad = np.flatnonzero(dg.h != want)
    assert len(bad) == 0, f'{len(bad)} of {episodes} episodes differ from the live reference, first {bad[:5]}'
