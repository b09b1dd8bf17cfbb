"""The replication log: ``GameStateUpdate`` records with the reference's class and attribute
names (optimax_rogue/logic/updates.py), decoded from the fixed-slot ``OrxEvent`` records the
kernel writes (include/orx.h). Only the four kinds the updater actually emits exist:
EntityPositionUpdate (:186), EntityCombatUpdate (:71), DungeonCreatedUpdate (:308),
EntityDeathUpdate (:167)."""
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
