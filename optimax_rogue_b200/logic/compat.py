"""Drop-in for the reference's single-game ``Updater`` (optimax_rogue/logic/updater.py:52-162).

``SingleGameUpdater`` has the reference's constructor and the reference's
``update(game_state, player1_move, player2_move) -> (UpdateResult, [GameStateUpdate])`` contract,
including the in-place mutation of the ``GameState`` it is given and ``get_incr_upd_order()``
(used by ``networking/server.py:194,203``), so the unmodified reference ``Server`` can be built with it
(``Server.__init__(game_state, updater, ...)``, ``networking/server.py:75-94``). Behind it one lane of the
batched CUDA path does the work; the host objects are duck-typed, so both the reference's classes
(pass ``updates_module=optimax_rogue.logic.updates``, ``world_module=optimax_rogue.game.world``) and
this package's mirrors work.

Differences from the reference, by construction: randomness comes from the Philox schedule
(seed, game_id), not from the process-global MT19937 generators, and a level is a function of
(seed, game_id, depth) -- so levels that already exist in ``game_state.world`` are honoured only for
the depths the players currently stand on.

Entities other than the two players (NPCs, ``updater.py:116-145``) ride in the lane's NPC slots, in the order they
have in ``game_state.entities`` (which is the order the reference culls the dead in): up to ``ORX_MAX_NPC`` of them,
with any idens. They never move (``decide_npc_move`` returns Stay), can be hit and die.
"""
import numpy as np
import torch

from .. import _abi
from ..config import SimConfig
from ..game import world as _world
from ..game.state import BatchedGameState
from . import updates as _updates
from .updater import BatchedUpdater, DungeonDespawningStrategy, UpdateResult


class SingleGameUpdater:
    def __init__(self, dgen, despawn_strat, max_ticks=None, *, seed=0, game_id=0, device='cuda',
                 updates_module=None, world_module=None, result_enum=None):
        self.dgen = dgen
        self.despawn_strat = DungeonDespawningStrategy(int(despawn_strat))
        self.max_ticks = max_ticks
        self.current_update_order = 0
        self._seed, self._game_id, self._device = seed, game_id, device
        self._updates = updates_module or _updates
        self._world = world_module or _world
        self._result_enum = result_enum or UpdateResult
        self._batched = BatchedUpdater(dgen, self.despawn_strat, max_ticks)
        self._lane = None
        self._moves = None
        self._idens, self._expected = [1, 2], ()
        self._flat_rows = None

    def get_incr_upd_order(self):
        """updater.py:71-74"""
        self.current_update_order += 1
        return self.current_update_order - 1

    # ------------------------------------------------------------------------------------------
    def _bind(self, game_state):
        p1, p2 = game_state.iden_lookup[game_state.player_1_iden], game_state.iden_lookup[game_state.player_2_iden]
        npcs = [e for e in game_state.entities if e.iden not in (game_state.player_1_iden, game_state.player_2_iden)]
        if len(npcs) > _abi.MAX_NPC:
            raise ValueError(f'at most {_abi.MAX_NPC} entities besides the two players (got {len(npcs)})')
        self._idens = [game_state.player_1_iden, game_state.player_2_iden] + [e.iden for e in npcs]   # lane iden k + 1 -> the caller's
        separated = p1.depth != p2.depth
        cfg = SimConfig(width=self.dgen.width, height=self.dgen.height, dgen_kind=self.dgen.kind,
                        start_kind=_abi.START_SEPARATED if separated else _abi.START_TOGETHER,
                        start_depth=(p1.depth, p2.depth) if separated else (p1.depth, p1.depth),
                        hp=(p1.base_max_health, p2.base_max_health), damage=(p1.base_damage, p2.base_damage),
                        armor=(p1.base_armor, p2.base_armor), seed=self._seed, n_npc=len(npcs),
                        fixed_tiles=getattr(self.dgen, 'tiles', None))
        lane = BatchedGameState(cfg, 1, self._device, game_id_base=self._game_id)
        lane.load_game_state(0, game_state)
        for k, e in enumerate(npcs):
            lane.set_npc(0, k, e.depth, e.x, e.y, e.health)
        self._lane = lane
        self._expected = self._alive(game_state)
        self._moves = torch.empty((1, 2), dtype=torch.uint8, device=lane.device)

    def _level(self, stairs):
        tiles = self._lane.level_tiles(stairs)
        return self._world.Dungeon(np.asarray(tiles, dtype='int32'))

    def update(self, game_state, player1_move, player2_move):
        """Moves the game state forward in time and returns (UpdateResult, list of GameStateUpdate),
        mutating ``game_state`` exactly as updater.py:76-162 does."""
        if self._lane is None or self._alive(game_state) != self._expected:
            self._bind(game_state)                 # first tick, or the caller added / removed entities itself
        lane, U = self._lane, self._updates
        real = self._idens
        self._sync_flat_bonuses(game_state)
        self._moves[0, 0], self._moves[0, 1] = int(player1_move), int(player2_move)
        result, ev = self._batched.update(lane, self._moves, want_events=True)
        recs = _updates.unpack_events(ev)[0]
        out = []
        world = game_state.world
        for kind, iden, a, b, depth in recs.tolist():
            if kind == _abi.EV_NONE:
                break
            order = self.get_incr_upd_order()
            if kind != _abi.EV_DUNGEON:
                iden = real[iden - 1]              # lane idens are 1, 2, 3 + slot
            if kind == _abi.EV_DUNGEON:
                dung = self._level((a, b))
                world.set_at_depth(depth, dung)                                  # updater.py:276-277
                out.append(U.DungeonCreatedUpdate(order, depth, dung))
            elif kind in (_abi.EV_MOVE, _abi.EV_DESCEND):
                ent = game_state.iden_lookup[iden]
                old_depth = ent.depth
                out.append(U.EntityPositionUpdate(order, iden, depth, old_depth, a, b))
                self._move_entity(game_state, ent, depth, a, b)                   # state.py:64-76
                if kind == _abi.EV_DESCEND and self._should_despawn(game_state, old_depth):
                    world.del_at_depth(old_depth)                                # updater.py:295-296
            elif kind == _abi.EV_COMBAT:
                defender = game_state.iden_lookup[real[a - 1]]
                if depth > 0:
                    defender.health -= depth                                     # updater.py:331-332
                out.append(U.EntityCombatUpdate(order, iden, defender.iden, depth, {b}, [], []))
            elif kind == _abi.EV_DEATH:
                out.append(U.EntityDeathUpdate(order, iden))
                self._remove_entity(game_state, game_state.iden_lookup[iden])    # updater.py:137-145
        game_state.tick += 1                                                     # updater.py:148
        self._expected = self._alive(game_state)
        return self._result_enum(int(result[0])), out

    def _sync_flat_bonuses(self, game_state):
        """The players' modifiers (game/modifiers.py:92-108): their flat bonuses are what Entity.on_tick folds into
        damage.value / armor.value / max_health.value (game/attribles.py:21-43); the lane takes the sums through
        OrxState.flat. Modifier event hooks are not run (see include/orx.h)."""
        rows = []
        for iden in (game_state.player_1_iden, game_state.player_2_iden):
            mods = getattr(game_state.iden_lookup[iden], 'modifiers', None) or []
            rows.append([sum(int(m.flat_damage) for m in mods), sum(int(m.flat_armor) for m in mods),
                         sum(int(m.flat_max_health) for m in mods)])
        if self._lane.flat is None and not any(v for row in rows for v in row):
            return
        if any(not -128 <= v <= 127 for row in rows for v in row):
            raise ValueError('flat modifier bonuses must fit int8')
        if rows != self._flat_rows:
            self._lane.enable_flat_bonuses()[0] = torch.tensor(rows, dtype=torch.int8)
            self._flat_rows = rows

    @staticmethod
    def _alive(game_state):
        """The idens the lane's NPC slots stand for, as the caller's state has them now."""
        return tuple(e.iden for e in game_state.entities
                     if e.iden not in (game_state.player_1_iden, game_state.player_2_iden))

    @staticmethod
    def _remove_entity(game_state, ent):
        if hasattr(game_state, 'remove_entity'):
            game_state.remove_entity(ent)                                        # state.py:84-88
        else:
            game_state.pos_lookup.pop((ent.depth, ent.x, ent.y), None)
            game_state.iden_lookup.pop(ent.iden, None)
            game_state.entities.remove(ent)

    @staticmethod
    def _move_entity(game_state, ent, depth, x, y):
        if hasattr(game_state, 'move_entity'):
            game_state.move_entity(ent, depth, x, y)
        else:
            game_state.pos_lookup.pop((ent.depth, ent.x, ent.y), None)
            ent.depth, ent.x, ent.y = depth, x, y
            game_state.pos_lookup[(depth, x, y)] = ent

    def _should_despawn(self, game_state, depth):
        """updater.py:245-257"""
        d1 = game_state.iden_lookup[game_state.player_1_iden].depth
        d2 = game_state.iden_lookup[game_state.player_2_iden].depth
        if self.despawn_strat == DungeonDespawningStrategy.Unreachable:
            return d1 > depth and d2 > depth
        return depth not in (d1, d2)
