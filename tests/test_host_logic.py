"""Host-side mirror of the reference interface: configuration errors, enums, wire formats, event
decoding. No GPU."""
import numpy as np
import pytest

from oracle import ref_harness as rh
from optimax_rogue_b200 import SimConfig, _abi
from optimax_rogue_b200.game.state import GameState, empty_room_tiles
from optimax_rogue_b200.game.entities import Entity
from optimax_rogue_b200.game.world import Dungeon, Tile, World
from optimax_rogue_b200.logic import updates
from optimax_rogue_b200.logic.moves import Move
from optimax_rogue_b200.logic.updater import BatchedUpdater, DungeonDespawningStrategy, UpdateResult
from optimax_rogue_b200.logic.worldgen import (EmptyDungeonGenerator, FixedDungeonGenerator,
                                               SeparatedGameStartGenerator, TogetherGameStartGenerator)


def test_enum_codes_are_the_references():
    assert [m.value for m in Move] == [1, 2, 3, 4, 5] and Move.Stay == 5          # logic/moves.py:6-12
    assert [r.value for r in UpdateResult] == [1, 2, 3, 4]                          # updater.py:16-21
    assert [t.value for t in Tile] == [1, 2, 3]                                     # world.py:10-17
    assert DungeonDespawningStrategy.Unreachable == 1 and DungeonDespawningStrategy.Unused == 2


@pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')
def test_enum_codes_against_live_reference():
    ref = rh.load_reference()
    assert {m.name: m.value for m in Move} == {m.name: m.value for m in ref.moves.Move}
    assert {m.name: m.value for m in UpdateResult} == {m.name: m.value for m in ref.updater.UpdateResult}
    assert {m.name: m.value for m in Tile} == {m.name: m.value for m in ref.world.Tile}
    assert {m.name: m.value for m in DungeonDespawningStrategy} == \
        {m.name: m.value for m in ref.updater.DungeonDespawningStrategy}
    import optimax_rogue.game.modifiers as mods
    assert (mods.CombatFlag.Block, mods.CombatFlag.Ambush, mods.CombatFlag.Flee, mods.CombatFlag.Parry) == (1, 2, 3, 4)


def test_config_validation_errors():
    with pytest.raises(ValueError):
        SimConfig(width=3).validate()
    with pytest.raises(ValueError):
        SimConfig(height=300).validate()
    with pytest.raises(ValueError):
        SimConfig(n_npc=9).validate()
    with pytest.raises(ValueError, match='Unknown despawn strat'):
        SimConfig(despawn_strat=3).validate()
    with pytest.raises(ValueError, match='cannot use SeparatedGameStartGenerator'):
        SeparatedGameStartGenerator(EmptyDungeonGenerator(60, 10), 4, 4)          # worldgen.py:112-114
    with pytest.raises(ValueError):
        SimConfig(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=np.ones((5, 5), np.uint8)).validate()
    with pytest.raises(ValueError, match='Unknown despawn strat'):
        BatchedUpdater(EmptyDungeonGenerator(60, 10), 7)


def test_generators_carry_reference_defaults():
    gen = TogetherGameStartGenerator()
    cfg = gen.sim_config(seed=5)
    assert (cfg.width, cfg.height) == (60, 10)                                      # worldgen.py:67
    assert cfg.hp == (10, 10) and cfg.damage == (2, 2) and cfg.armor == (1, 1)      # worldgen.py:85-86
    sep = SeparatedGameStartGenerator()
    assert sep.sim_config().start_depth == (0, 1000)                                # worldgen.py:110-111
    fx = FixedDungeonGenerator(np.full((7, 5), 1, np.uint8))
    assert (fx.width, fx.height) == (7, 5)


def test_fixed_tables_rank_order_is_x_major():
    t = np.full((5, 4), 2, np.uint8)
    t[1, 1] = t[1, 2] = t[3, 1] = 1
    t[2, 2] = 3
    cfg = SimConfig(width=5, height=4, dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t)
    flat, ground, stairs = cfg.fixed_tables()
    assert ground.tolist() == [1 * 4 + 1, 1 * 4 + 2, 3 * 4 + 1]                     # world.py:60 flat = x*H + y
    assert stairs == (2, 2)
    assert flat[2 * 4 + 2] == 3


def test_dungeon_and_world_wire_format_roundtrip():
    tiles = empty_room_tiles(6, 5, (2, 3))
    d = Dungeon(tiles)
    raw = d.to_prims()
    assert raw[:8] == (6).to_bytes(4, 'big') + (5).to_bytes(4, 'big')               # world.py:76-77
    assert raw[8:] == tiles.astype('uint8').tobytes()
    assert Dungeon.from_prims(raw) == d
    assert d.staircase() == (2, 3) and d.is_blocked(0, 0) and d.is_blocked(-1, 2) and not d.is_blocked(2, 3)
    w = World({0: d, 7: Dungeon(empty_room_tiles(6, 5, (1, 1)))})
    assert World.from_prims(w.to_prims()) == w


@pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')
def test_wire_format_is_byte_identical_to_reference():
    ref = rh.load_reference()
    tiles = empty_room_tiles(60, 10, (17, 4))
    ours = World({0: Dungeon(tiles), 3: Dungeon(empty_room_tiles(60, 10, (2, 2)))})
    theirs = ref.world.World({0: ref.world.Dungeon(tiles.copy()), 3: ref.world.Dungeon(empty_room_tiles(60, 10, (2, 2)))})
    assert ours.to_prims() == theirs.to_prims()
    back = ref.world.World.from_prims(ours.to_prims())
    assert back == theirs


def test_game_state_view_for_filters_by_depth():
    d0, d1 = Dungeon(empty_room_tiles(6, 5, (2, 2))), Dungeon(empty_room_tiles(6, 5, (3, 3)))
    e1, e2 = Entity(1, 0, 1, 1, 10, 10, 2, 1), Entity(2, 1, 2, 2, 10, 10, 2, 1)
    gs = GameState(True, 5, 1, 2, World({0: d0, 1: d1}), [e1, e2])
    v = gs.view_for(e1)
    assert list(v.world.dungeons) == [0] and [e.iden for e in v.entities] == [1] and not v.is_authoritative
    assert gs.view_for(e2, reduce_tick=True).tick == 4                              # state.py:56


def test_event_decoding():
    def rec(kind, iden, a, b, depth):
        return [kind | (iden << 8) | (a << 16) | (b << 24), depth]
    raw = np.array([rec(3, 0, 4, 5, 2), rec(5, 1, 7, 3, 2), rec(2, 2, 1, 3, 1), rec(1, 2, 9, 8, 0), [0, 0]], np.int64).astype(np.int32)
    un = updates.unpack_events(raw)
    evs = updates.decode_events(un, first_order=10, width=12, height=9)
    assert [type(e).__name__ for e in evs] == ['DungeonCreatedUpdate', 'EntityPositionUpdate', 'EntityCombatUpdate', 'EntityPositionUpdate']
    assert [e.order for e in evs] == [10, 11, 12, 13]
    assert evs[0].depth == 2 and evs[0].dungeon.staircase() == (4, 5) and evs[0].dungeon.width == 12
    assert (evs[1].entity_iden, evs[1].depth, evs[1].old_depth, evs[1].posx, evs[1].posy) == (1, 2, 1, 7, 3)
    assert evs[1].depth_changed and not evs[3].depth_changed
    assert (evs[2].attacker_iden, evs[2].defender_iden, evs[2].og_damage, evs[2].tags) == (2, 1, 1, {3})


def _sample_game_state():
    d0, d3 = Dungeon(empty_room_tiles(60, 10, (17, 4))), Dungeon(empty_room_tiles(60, 10, (2, 2)))
    ents = [Entity(1, 0, 5, 6, 9, 10, 2, 1), Entity(2, 3, 40, 3, -1, 10, 2, 1), Entity(3, 0, 7, 7, 2, 2, 0, 0)]
    return GameState(True, 77, 1, 2, World({0: d0, 3: d3}), ents)


def test_game_state_snapshot_roundtrip():
    gs = _sample_game_state()
    raw = gs.to_prims()
    assert raw[0] == 1 and raw[1:5] == (77).to_bytes(4, 'big')                      # state.py:96-98
    back = GameState.from_prims(raw)
    assert back.tick == 77 and back.world == gs.world and back.entities == gs.entities
    assert back.to_prims() == raw


@pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')
def test_game_state_snapshot_is_byte_identical_to_reference():
    ref = rh.load_reference()
    ours = _sample_game_state()
    rd = {d: ref.world.Dungeon(dung.tiles.copy()) for d, dung in ours.world.dungeons.items()}
    rents = [ref.entities.Entity(e.iden, e.depth, e.x, e.y, e.health, e.base_max_health, e.base_damage,
                                 e.base_armor, [], dict()) for e in ours.entities]
    theirs = ref.state.GameState(True, 77, 1, 2, ref.world.World(rd), rents)
    assert ours.to_prims() == theirs.to_prims()
    # and the reference can load what we write
    back = ref.state.GameState.from_prims(ours.to_prims())
    assert back == theirs


def test_pack_moves_roundtrip():
    """Nibble-packed command format of orx_step_packed: p1 | p2 << 4, on numpy and torch alike."""
    import torch
    from optimax_rogue_b200.logic.moves import pack_moves, unpack_moves
    rng = np.random.default_rng(0)
    mv = rng.integers(0, 16, size=(1000, 2), dtype=np.uint8)
    packed = pack_moves(mv[:, 0], mv[:, 1])
    assert packed.dtype == np.uint8 and np.array_equal(packed, mv[:, 0] + 16 * mv[:, 1])
    a, b = unpack_moves(packed)
    assert np.array_equal(a, mv[:, 0]) and np.array_equal(b, mv[:, 1])
    t = torch.from_numpy(mv)
    tp = pack_moves(t[:, 0], t[:, 1])
    assert tp.dtype == torch.uint8 and np.array_equal(tp.numpy(), packed)


def test_word_planes_share_one_allocation_at_a_common_pitch():
    """include/orx.h, layout hint of OrxState: pos, hp, stairs, tick, episode are carved out of one
    allocation, in that order, at a pitch that is a multiple of 128 bytes and >= 4 n, each one an ordinary
    contiguous tensor; clone() keeps the layout, state_dict() round-trips through torch.save."""
    import io
    import torch
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    for n in (1, 255, 1000, 4096):
        gs = BatchedGameState(SimConfig(), n, 'cpu')
        for other in (gs, gs.clone()):
            base = other.pos.data_ptr()
            pitch = other.hp.data_ptr() - base
            assert pitch % 128 == 0 and pitch >= 4 * n
            assert [getattr(other, k).data_ptr() - base for k in ('pos', 'hp', 'stairs', 'tick', 'episode')] == [k * pitch for k in range(5)]
            assert all(getattr(other, k).is_contiguous() for k in ('pos', 'hp', 'stairs', 'tick', 'episode'))
            assert (other.pos.shape, other.hp.shape, other.stairs.shape, other.tick.shape, other.episode.shape) == \
                ((n, 4), (n, 2), (n, 4), (n,), (n,))
            assert (other.pos.dtype, other.hp.dtype, other.tick.dtype) == (torch.uint8, torch.int16, torch.int32)
        gs.pos[:] = 7
        gs.hp[:, 1] = -3
        gs.tick[:] = 11
        assert int(gs.stairs.sum()) == 0 and int(gs.episode.sum()) == 0 and int(gs.hp[:, 0].sum()) == 0      # no overlap
        c = gs.clone()
        c.tick[:] = 5
        assert int(gs.tick[0]) == 11                                                                       # clone owns its planes
        buf = io.BytesIO()
        torch.save(gs.state_dict(), buf)
        buf.seek(0)
        back = BatchedGameState(SimConfig(), n, 'cpu')
        back.load_state_dict(torch.load(buf))
        for name in BatchedGameState.PLANES:
            assert torch.equal(getattr(back, name), getattr(gs, name)), name


def test_tools_and_examples_compile():
    """The tuning / evidence scripts under tools/ and examples/ only run on a GPU box; keep them importable."""
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, 'tools', '*.py')) + glob.glob(os.path.join(root, 'examples', '*.py')) + \
        [os.path.join(root, 'bench.py'), os.path.join(root, '__graft_entry__.py')]
    assert len(files) >= 10
    for f in files:
        with open(f) as fh:
            compile(fh.read(), f, 'exec')


def test_state_dict_is_a_snapshot_and_checks_what_it_is_loaded_into():
    """state_dict() copies EVERY plane (mutating the live state afterwards must not leak into the snapshot),
    carries the batch size + configuration, and load_state_dict() refuses a mismatch instead of broadcasting."""
    import pytest
    import torch
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    cfg = SimConfig(n_npc=2, seed=5)
    gs = BatchedGameState(cfg, 300, 'cpu')
    gs.depth[:] = 3; gs.status[:] = 1; gs.npc_hp[:] = 9; gs.tick[:] = 17
    sd = gs.state_dict()
    keep = gs.clone()
    gs.depth[:] = 8; gs.status[:] = 4; gs.npc_hp[:] = -1; gs.npc_depth[:] = 5; gs.npc_pos[:] = 2; gs.tick[:] = 99      # "tick on"
    gs.load_state_dict(sd)
    for name in BatchedGameState.PLANES:
        assert torch.equal(getattr(gs, name), getattr(keep, name)), name
    assert sd['fingerprint']['n'] == 300 and sd['fingerprint']['seed'] == 5
    with pytest.raises(ValueError, match='n:'):
        BatchedGameState(cfg, 1, 'cpu').load_state_dict(sd)            # would have broadcast before
    with pytest.raises(ValueError, match='seed'):
        BatchedGameState(SimConfig(n_npc=2, seed=6), 300, 'cpu').load_state_dict(sd)
    legacy = {k: v for k, v in sd.items() if k != 'fingerprint'}
    legacy['tick'] = legacy['tick'][:1]
    with pytest.raises(ValueError, match='plane tick'):
        BatchedGameState(cfg, 300, 'cpu').load_state_dict(legacy)


def test_updater_remarshals_when_the_state_or_its_config_changes():
    """BatchedUpdater keeps the C structs of the state it last ticked. The cache must notice a different state
    object (even one that reuses the id of a freed one), a re-bound plane / scratch buffer and a mutated config."""
    import gc
    import torch
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 50)
    gs = BatchedGameState(SimConfig(seed=1), 64, 'cpu')
    cfg, st = upd._cfg(gs)
    assert upd._cfg(gs)[1] is st                                   # hit
    assert cfg.seed == 1 and st.pos == gs.pos.data_ptr()
    gs.cfg.seed = 2                                                # config mutated after the first call
    assert upd._cfg(gs)[0].seed == 2
    gs.sched = torch.zeros((4,), dtype=torch.int32)                # scratch re-bound
    assert upd._cfg(gs)[1].sched_words == 4
    flat = gs.enable_flat_bonuses()                                # bonus plane allocated later
    assert upd._cfg(gs)[1].flat == flat.data_ptr()
    seen = set()
    for k in range(20):                                            # short-lived states: ids get reused, pointers must not go stale
        tmp = BatchedGameState(SimConfig(seed=100 + k), 64, 'cpu')
        c, s = upd._cfg(tmp)
        assert c.seed == 100 + k and s.pos == tmp.pos.data_ptr() and s.tick == tmp.tick.data_ptr()
        seen.add(id(tmp))
        del tmp
        gc.collect()
    upd.max_ticks = 70
    assert upd._cfg(gs)[0].max_ticks == 70
