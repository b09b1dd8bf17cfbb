"""SingleGameUpdater against the LIVE reference objects (build container only): the adapter drives a
real optimax_rogue GameState / World / update classes, the reference Updater runs beside it under the
same injected draws, and after every tick the two GameStates must be equal (GameState.__eq__,
game/state.py:134-153) and the update lists must agree field by field. No GPU here, so the C oracle
stands in for the CUDA lane behind the adapter (tests may do that); the CUDA lane itself is checked
against the same oracle in tests/test_gpu_parity.py."""
import pytest
import torch

from oracle import cport
from oracle import ref_harness as rh
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import empty_room_tiles
from optimax_rogue_b200.logic.compat import SingleGameUpdater
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')

SEED, GID = 0xFEED, 4242


class OracleLane:
    """Stands in for the 1-game BatchedGameState + BatchedUpdater pair."""

    def __init__(self, cfg, game_state):
        self.cfg, self.device = cfg, torch.device('cpu')
        self.orc = cport.Oracle(cfg, 1, game_id_base=GID)
        self.orc.reset()
        s = self.orc.state
        for k, ent in enumerate((game_state.player_1, game_state.player_2)):
            s.pos[0, 2 * k], s.pos[0, 2 * k + 1] = ent.x, ent.y
            s.depth[0, k], s.hp[0, k] = ent.depth, ent.health
            sx, sy = game_state.world.dungeons[ent.depth].staircase()
            s.stairs[0, 2 * k], s.stairs[0, 2 * k + 1] = sx, sy
        s.tick[0] = game_state.tick
        npcs = [e for e in game_state.entities if e.iden not in (game_state.player_1_iden, game_state.player_2_iden)]
        for k, e in enumerate(npcs):
            s.npc_depth[0, k], s.npc_hp[0, k] = e.depth, e.health
            s.npc_pos[0, k] = (e.x, e.y)

        self.flat = None

    def enable_flat_bonuses(self):
        if self.flat is None:
            self.flat = torch.from_numpy(self.orc.state.enable_flat_bonuses())
        return self.flat

    def level_tiles(self, stairs):
        return empty_room_tiles(self.cfg.width, self.cfg.height, stairs)

    def update(self, lane, moves, want_events=True):
        res, ev = self.orc.step(moves.numpy(), want_events=True)
        return res, ev


@pytest.mark.parametrize('despawn', ['unreachable', 'unused'])
@pytest.mark.parametrize('bots', [('staircase', 'random'), ('staircase', 'staircase'), ('random', 'random')])
def test_adapter_equals_reference_updater_on_reference_objects(despawn, bots):
    ref = rh.load_reference()
    strat = 1 if despawn == 'unreachable' else 2
    inj = rh.Injector(SEED)
    inj.game_id = GID
    with inj:
        # the reference side: its own generator + Updater, injected draws
        rdgen = inj.wrap_dgen(ref.worldgen.EmptyDungeonGenerator(20, 8))
        inj.site, inj.q = ('reset',), 0
        gs_ref = ref.worldgen.TogetherGameStartGenerator(rdgen).setup_game()
        inj.site = None
        gs_ours = ref.state.GameState.from_prims(gs_ref.to_prims())          # an independent copy
        upd_ref = inj.wrap_updater(ref.updater.Updater(rdgen, ref.updater.DungeonDespawningStrategy(strat), 300), gs_ref)
        # our side: the adapter over reference classes, oracle lane behind it
        cfg = SimConfig(width=20, height=8, despawn_strat=strat, max_ticks=300, seed=SEED)
        adapter = SingleGameUpdater(EmptyDungeonGenerator(20, 8), strat, 300, seed=SEED, game_id=GID, device='cpu',
                                    updates_module=ref.updates, world_module=ref.world, result_enum=ref.updater.UpdateResult)
        lane = OracleLane(cfg, gs_ours)
        adapter._lane, adapter._moves = lane, torch.empty((1, 2), dtype=torch.uint8)
        adapter._batched = lane
        b = [ref.staircasebot.StaircaseBot(k + 1) if kind == 'staircase' else ref.randombot.RandomBot(k + 1)
             for k, kind in enumerate(bots)]
        for t in range(250):
            inj.tick, inj.shuffle_calls = gs_ref.tick, 0
            gs_ref.on_tick()
            inj.choice_slot = 0
            m1 = b[0].move(gs_ref)
            inj.choice_slot = 1
            m2 = b[1].move(gs_ref)
            res_ref, ev_ref = upd_ref.update(gs_ref, m1, m2)
            res_ours, ev_ours = adapter.update(gs_ours, m1, m2)
            assert res_ours == res_ref and isinstance(res_ours, ref.updater.UpdateResult)
            assert gs_ours == gs_ref, f'tick {t}: GameStates differ'
            assert set(gs_ours.world.dungeons) == set(gs_ref.world.dungeons)
            assert gs_ours.pos_lookup.keys() == gs_ref.pos_lookup.keys()
            assert [type(e) for e in ev_ours] == [type(e) for e in ev_ref]
            for eo, er in zip(ev_ours, ev_ref):
                assert eo.order == er.order
                if isinstance(er, ref.updates.EntityPositionUpdate):
                    assert (eo.entity_iden, eo.depth, eo.old_depth, eo.posx, eo.posy) == \
                        (er.entity_iden, er.depth, er.old_depth, er.posx, er.posy)
                elif isinstance(er, ref.updates.EntityCombatUpdate):
                    assert (eo.attacker_iden, eo.defender_iden, eo.og_damage, {int(x) for x in eo.tags}) == \
                        (er.attacker_iden, er.defender_iden, er.og_damage, {int(x) for x in er.tags})
                elif isinstance(er, ref.updates.DungeonCreatedUpdate):
                    assert eo.depth == er.depth and eo.dungeon == er.dungeon
            assert adapter.current_update_order == upd_ref.current_update_order
            if res_ref != ref.updater.UpdateResult.InProgress:
                break


def test_adapter_carries_npc_entities_like_the_reference():
    """Entities besides the two players (updater.py:116-145) with arbitrary idens: they block tiles, take hits
    (EntityCombatUpdate with their own iden), die (EntityDeathUpdate, removed from the GameState in reverse
    entity order) -- tick by tick equal to the reference Updater on a small room where bumping into them is frequent."""
    ref = rh.load_reference()
    inj = rh.Injector(SEED)
    inj.game_id = GID
    W, H = 7, 6
    with inj:
        rdgen = inj.wrap_dgen(ref.worldgen.EmptyDungeonGenerator(W, H))
        inj.site, inj.q = ('reset',), 0
        gs_ref = ref.worldgen.TogetherGameStartGenerator(rdgen).setup_game()
        inj.site = None
        taken = {(e.x, e.y) for e in gs_ref.entities} | {gs_ref.world.dungeons[0].staircase()}
        free = [(x, y) for x in range(1, W - 1) for y in range(1, H - 1) if (x, y) not in taken]
        for iden, (x, y), hp in zip((9, 4, 17), free[::3], (1, 2, 3)):                 # idens neither contiguous nor ordered
            gs_ref.add_entity(ref.entities.Entity(iden, 0, x, y, hp, hp, 0, 0, [], dict()))
        gs_ours = ref.state.GameState.from_prims(gs_ref.to_prims())
        upd_ref = inj.wrap_updater(ref.updater.Updater(rdgen, ref.updater.DungeonDespawningStrategy(1), 400), gs_ref)
        cfg = SimConfig(width=W, height=H, max_ticks=400, seed=SEED, n_npc=3)
        adapter = SingleGameUpdater(EmptyDungeonGenerator(W, H), 1, 400, seed=SEED, game_id=GID, device='cpu',
                                    updates_module=ref.updates, world_module=ref.world, result_enum=ref.updater.UpdateResult)
        lane = OracleLane(cfg, gs_ours)
        adapter._lane, adapter._moves, adapter._batched = lane, torch.empty((1, 2), dtype=torch.uint8), lane
        adapter._idens = [gs_ours.player_1_iden, gs_ours.player_2_iden, 9, 4, 17]
        adapter._expected = adapter._alive(gs_ours)
        b = [ref.randombot.RandomBot(1), ref.randombot.RandomBot(2)]
        deaths = hits = 0
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
            assert [e.iden for e in gs_ours.entities] == [e.iden for e in gs_ref.entities]
            assert [type(e) for e in ev_ours] == [type(e) for e in ev_ref]
            for eo, er in zip(ev_ours, ev_ref):
                assert eo.order == er.order
                if isinstance(er, ref.updates.EntityCombatUpdate):
                    assert (eo.attacker_iden, eo.defender_iden, eo.og_damage) == (er.attacker_iden, er.defender_iden, er.og_damage)
                    hits += er.defender_iden > 2
                elif isinstance(er, ref.updates.EntityDeathUpdate):
                    assert eo.entity_iden == er.entity_iden
                    deaths += 1
            if res_ref != ref.updater.UpdateResult.InProgress:
                break
        assert hits > 0 and deaths > 0


def test_adapter_honours_flat_modifiers_on_reference_entities():
    """Players that carry a Modifier (game/modifiers.py:92-108): the adapter reads the flat bonuses off the reference
    entities every tick and the hits equal the reference Updater's (updater.py:313 through the attribles)."""
    ref = rh.load_reference()
    inj = rh.Injector(SEED)
    inj.game_id = GID
    W, H = 5, 5
    with inj:
        rdgen = inj.wrap_dgen(ref.worldgen.EmptyDungeonGenerator(W, H))
        inj.site, inj.q = ('reset',), 0
        gs_ref = ref.worldgen.TogetherGameStartGenerator(rdgen).setup_game()
        inj.site = None
        for ent in (gs_ref.player_1, gs_ref.player_2):
            ent.health = ent.base_max_health = 40
        gs_ours = ref.state.GameState.from_prims(gs_ref.to_prims())
        for gs in (gs_ref, gs_ours):
            gs.player_1.modifiers.append(rh.make_flat_modifier(ref, gs.player_1, flat_damage=3, flat_max_health=2))
            gs.player_2.modifiers.append(rh.make_flat_modifier(ref, gs.player_2, flat_armor=-1))
            gs.player_2.modifiers.append(rh.make_flat_modifier(ref, gs.player_2, flat_damage=1))
        upd_ref = inj.wrap_updater(ref.updater.Updater(rdgen, ref.updater.DungeonDespawningStrategy(1), 300), gs_ref)
        cfg = SimConfig(width=W, height=H, max_ticks=300, seed=SEED, hp=(40, 40))
        adapter = SingleGameUpdater(EmptyDungeonGenerator(W, H), 1, 300, seed=SEED, game_id=GID, device='cpu',
                                    updates_module=ref.updates, world_module=ref.world, result_enum=ref.updater.UpdateResult)
        lane = OracleLane(cfg, gs_ours)
        adapter._lane, adapter._moves, adapter._batched = lane, torch.empty((1, 2), dtype=torch.uint8), lane
        b = [ref.randombot.RandomBot(1), ref.randombot.RandomBot(2)]
        dmg = set()
        for t in range(300):
            inj.tick, inj.shuffle_calls = gs_ref.tick, 0
            gs_ref.on_tick()
            gs_ours.on_tick()
            inj.choice_slot = 0
            m1 = b[0].move(gs_ref)
            inj.choice_slot = 1
            m2 = b[1].move(gs_ref)
            res_ref, ev_ref = upd_ref.update(gs_ref, m1, m2)
            res_ours, ev_ours = adapter.update(gs_ours, m1, m2)
            assert res_ours == res_ref
            assert [(e.x, e.y, e.depth, e.health) for e in gs_ours.entities] == [(e.x, e.y, e.depth, e.health) for e in gs_ref.entities], t
            for eo, er in zip(ev_ours, ev_ref):
                if isinstance(er, ref.updates.EntityCombatUpdate):
                    assert (eo.attacker_iden, eo.defender_iden, eo.og_damage) == (er.attacker_iden, er.defender_iden, er.og_damage)
                    dmg.add((er.attacker_iden, er.og_damage))
            if res_ref != ref.updater.UpdateResult.InProgress:
                break
        assert dmg == {(1, 2 + 3 - 1), (2, 2 + 1 - 1 + 1)}
