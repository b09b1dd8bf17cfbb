"""The replication log: ``GameStateUpdate`` records with the reference's class and attribute
names (optimax_rogue/logic/updates.py), decoded from the fixed-slot ``OrxEvent`` records the
kernel writes (include/orx.h). Ruleset R0 emits the four kinds the reference updater actually emits:
EntityPositionUpdate (:186), EntityCombatUpdate (:71), DungeonCreatedUpdate (:308),
EntityDeathUpdate (:167). Ruleset R1 (parity unpinned, docs/RULESET_R1.md) also uses the classes the
reference defines but never emits: EntitySpawnUpdate (:141), EntityHealthUpdate (:222),
EntityModifierAddedUpdate (:255) and EntityEventUpdate (:31) -- ``decode_r1_events``."""
import dataclasses
import typing

import numpy as np

from .. import _abi
from ..game.state import empty_room_tiles
from ..game.world import Dungeon


@dataclasses.dataclass
class GameStateUpdate:
    order: int


@dataclasses.dataclass
class EntityPositionUpdate(GameStateUpdate):
    entity_iden: int
    depth: int
    old_depth: int
    posx: int
    posy: int

    @property
    def depth_changed(self):
        return self.depth != self.old_depth


@dataclasses.dataclass
class EntityCombatUpdate(GameStateUpdate):
    attacker_iden: int
    defender_iden: int
    og_damage: int
    tags: typing.Set[int]
    attack_prevals: tuple = ()
    defend_prevals: tuple = ()


@dataclasses.dataclass
class DungeonCreatedUpdate(GameStateUpdate):
    depth: int
    dungeon: Dungeon


@dataclasses.dataclass
class EntityDeathUpdate(GameStateUpdate):
    entity_iden: int


@dataclasses.dataclass
class SpawnedEntity:
    """What EntitySpawnUpdate carries of the new entity (reference: a whole ``Entity``, game/entities.py)."""
    iden: int
    depth: int
    x: int
    y: int
    health: int = 0          # enemies
    item_kind: int = -1      # ground items: 0 = +1 damage, 1 = +1 armor, 2 = +2 max health


@dataclasses.dataclass
class EntitySpawnUpdate(GameStateUpdate):
    entity: SpawnedEntity


@dataclasses.dataclass
class EntityHealthUpdate(GameStateUpdate):
    entity_iden: int
    source_iden: int
    amount: int
    tags: typing.FrozenSet[str]


@dataclasses.dataclass
class FlatModifier:
    """The item bonus as the ``Modifier.flat_*`` sums it contributes (game/modifiers.py:102-108)."""
    flat_damage: int = 0
    flat_armor: int = 0
    flat_max_health: int = 0


@dataclasses.dataclass
class EntityModifierAddedUpdate(GameStateUpdate):
    entity_iden: int
    modifier: FlatModifier
    item_iden: int = 0


@dataclasses.dataclass
class EntityEventUpdate(GameStateUpdate):
    entity_iden: int
    event_name: str
    args: typing.Any
    prevals: tuple = ()


R1_HIT_TAGS = {_abi.R1_HIT_FULL: 'full', _abi.R1_HIT_HALF: 'half', _abi.R1_HIT_NEGATED: 'negated', _abi.R1_HIT_CONTEST: 'contest'}
R1_HEALTH_TAGS = {_abi.R1_HEALTH_HEAL: 'heal', _abi.R1_HEALTH_SEPARATION: 'separation'}
_ITEM_MODIFIERS = {0: FlatModifier(flat_damage=1), 1: FlatModifier(flat_armor=1), 2: FlatModifier(flat_max_health=2)}


def decode_r1_events(records: np.ndarray, first_order: int = 0) -> typing.List[GameStateUpdate]:
    """One game's unpacked ruleset-R1 records ([max_events, 5] from ``unpack_events``) -> GameStateUpdate list in
    emission order (include/orx.h "Replication log of an R1 tick"). Idens are lane + 1: players 1-2, enemy slots
    3-10, item slots 11-14. DungeonCreatedUpdate carries ``dungeon=None`` (an R1 level is its staircase and wall key,
    not a tile array); a health-raising pickup is followed by its EntityHealthUpdate."""
    out: typing.List[GameStateUpdate] = []
    order = first_order

    def add(u):
        nonlocal order
        out.append(u)
        order += 1

    for kind, iden, a, b, value in records.tolist():
        if kind == _abi.EV_NONE:
            break
        if kind == _abi.EV_MOVE:
            add(EntityPositionUpdate(order, iden, value, value, a, b))
        elif kind == _abi.EV_DESCEND:
            add(EntityPositionUpdate(order, iden, value, value - 1, a, b))
        elif kind == _abi.EV_COMBAT:
            add(EntityCombatUpdate(order, iden, a, value, {R1_HIT_TAGS[b]}))
        elif kind == _abi.EV_DUNGEON:
            add(DungeonCreatedUpdate(order, value, None))
        elif kind == _abi.EV_DEATH:
            add(EntityDeathUpdate(order, iden))
        elif kind == _abi.EV_SPAWN:
            depth, aux = value & 0xFFFF, value >> 16
            item = iden > 2 + _abi.R1_ENEMIES
            add(EntitySpawnUpdate(order, SpawnedEntity(iden, depth, a, b, 0 if item else aux, aux if item else -1)))
        elif kind == _abi.EV_HEALTH:
            add(EntityHealthUpdate(order, iden, a, value, frozenset({R1_HEALTH_TAGS[b]})))
        elif kind == _abi.EV_PICKUP:
            add(EntityModifierAddedUpdate(order, iden, _ITEM_MODIFIERS[b], a))
            if value:
                add(EntityHealthUpdate(order, iden, iden, value, frozenset({'item'})))
        elif kind == _abi.EV_XP:
            add(EntityEventUpdate(order, iden, 'xp', {'source_iden': a, 'levels_gained': b, 'xp': value}))
        else:
            raise ValueError(f'unknown event kind {kind}')
    return out


def unpack_events(events) -> np.ndarray:
    """int32[..., 2] raw records (tensor or ndarray) -> int32[..., 5] = kind, iden, a, b, depth."""
    ev = events.cpu().numpy() if hasattr(events, 'cpu') else np.asarray(events)
    w0 = ev[..., 0].astype(np.int64) & 0xFFFFFFFF
    return np.stack([w0 & 0xFF, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF, (w0 >> 24) & 0xFF,
                     ev[..., 1].astype(np.int64)], axis=-1).astype(np.int32)


def decode_events(records: np.ndarray, first_order: int = 0, width: int = 60, height: int = 10,
                  level_tiles=None) -> typing.List[GameStateUpdate]:
    """One game's unpacked records ([max_events, 5]) -> list of GameStateUpdate, orders counting
    up from ``first_order`` (Updater.get_incr_upd_order, updater.py:71-74)."""
    out = []
    order = first_order
    for kind, iden, a, b, depth in records.tolist():
        if kind == _abi.EV_NONE:
            break
        if kind == _abi.EV_MOVE:
            out.append(EntityPositionUpdate(order, iden, depth, depth, a, b))
        elif kind == _abi.EV_DESCEND:
            out.append(EntityPositionUpdate(order, iden, depth, depth - 1, a, b))
        elif kind == _abi.EV_COMBAT:
            out.append(EntityCombatUpdate(order, iden, a, depth, {b}))
        elif kind == _abi.EV_DUNGEON:
            tiles = level_tiles((a, b)) if level_tiles is not None else empty_room_tiles(width, height, (a, b))
            out.append(DungeonCreatedUpdate(order, depth, Dungeon(tiles)))
        elif kind == _abi.EV_DEATH:
            out.append(EntityDeathUpdate(order, iden))
        else:
            raise ValueError(f'unknown event kind {kind}')
        order += 1
    return out
