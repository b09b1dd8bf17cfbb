"""Game state containers.

``BatchedGameState`` holds N independent games as structure-of-arrays tensors in HBM (the planes
of ``include/orx.h:OrxState``). It replaces the reference's per-game object graph
(optimax_rogue/game/state.py:10-88): positions ARE the lookup, so ``pos_lookup``/``iden_lookup``
have no counterpart. ``GameState`` is the host-side single-game view with the reference's
attribute names, used for interop with code written against the reference.
"""
import io
import json
import typing

import numpy as np
import torch

from .. import _abi
from ..config import SimConfig
from .entities import Entity
from .world import Dungeon, Tile, World


class GameState:
    """Single-game host view (state.py:10-45 attribute names)."""

    def __init__(self, is_authoritative: bool, tick: int, player_1_iden: int, player_2_iden: int,
                 world: World, entities: typing.List[Entity]):
        self.is_authoritative = is_authoritative
        self.tick = tick
        self.player_1_iden = player_1_iden
        self.player_2_iden = player_2_iden
        self.world = world
        self.entities = entities
        self.pos_lookup = dict(((ent.depth, ent.x, ent.y), ent) for ent in entities)
        self.iden_lookup = dict((ent.iden, ent) for ent in entities)

    @property
    def player_1(self) -> Entity:
        return self.iden_lookup[self.player_1_iden]

    @property
    def player_2(self) -> Entity:
        return self.iden_lookup[self.player_2_iden]

    def move_entity(self, entity, newdepth, newx, newy):
        """state.py:64-76"""
        del self.pos_lookup[(entity.depth, entity.x, entity.y)]
        entity.depth, entity.x, entity.y = newdepth, newx, newy
        self.pos_lookup[(newdepth, newx, newy)] = entity

    def view_for(self, entity: Entity, reduce_tick: bool = False) -> 'GameState':
        """state.py:53-58: only the viewer's depth is kept."""
        world = World({entity.depth: self.world.dungeons[entity.depth]})
        ents = [e for e in self.entities if e.depth == entity.depth]
        return GameState(False, self.tick - 1 if reduce_tick else self.tick,
                         self.player_1_iden, self.player_2_iden, world, ents)

    # -- the reference's binary snapshot (state.py:94-132), byte for byte --------------------------
    ENTITY_IDEN = 'optimax_rogue.game.entities.entity'     # Serializable.identifier, serializer.py:92-96

    @staticmethod
    def _entity_bytes(ent: Entity) -> bytes:
        """ser.serialize(entity): JSON, sorted keys, ASCII (serializer.py:46-52,145-151; the
        prims are Entity.to_prims, entities.py:76-88; no modifier or item class exists)."""
        prims = {'iden': ent.iden, 'x': ent.x, 'y': ent.y, 'depth': ent.depth, 'health': ent.health,
                 'base_max_health': ent.base_max_health, 'base_damage': ent.base_damage,
                 'base_armor': ent.base_armor, 'modifiers': [], 'items': {}}
        return json.dumps({'iden': GameState.ENTITY_IDEN, 'prims': prims}, sort_keys=True).encode('ASCII')

    def to_prims(self) -> bytes:
        arr = io.BytesIO()
        arr.write((1 if self.is_authoritative else 0).to_bytes(1, 'big', signed=False))
        arr.write(int(self.tick).to_bytes(4, 'big', signed=False))
        arr.write(int(self.player_1_iden).to_bytes(4, 'big', signed=False))
        arr.write(int(self.player_2_iden).to_bytes(4, 'big', signed=False))
        wserd = self.world.to_prims()
        arr.write(len(wserd).to_bytes(8, 'big', signed=False))
        arr.write(wserd)
        arr.write(len(self.entities).to_bytes(4, 'big', signed=False))
        for ent in self.entities:
            eserd = self._entity_bytes(ent)
            arr.write(len(eserd).to_bytes(4, 'big', signed=False))
            arr.write(eserd)
        return arr.getvalue()

    @classmethod
    def from_prims(cls, prims: bytes) -> 'GameState':
        arr = io.BytesIO(prims)
        auth = int.from_bytes(arr.read(1), 'big', signed=False)
        tick = int.from_bytes(arr.read(4), 'big', signed=False)
        p1 = int.from_bytes(arr.read(4), 'big', signed=False)
        p2 = int.from_bytes(arr.read(4), 'big', signed=False)
        wlen = int.from_bytes(arr.read(8), 'big', signed=False)
        world = World.from_prims(arr.read(wlen))
        ents = []
        for _ in range(int.from_bytes(arr.read(4), 'big', signed=False)):
            elen = int.from_bytes(arr.read(4), 'big', signed=False)
            d = json.loads(arr.read(elen).decode('ASCII'))
            if d['iden'] != cls.ENTITY_IDEN:
                raise ValueError(f"unexpected serialized type {d['iden']}")
            q = d['prims']
            ents.append(Entity(q['iden'], q['depth'], q['x'], q['y'], q['health'], q['base_max_health'],
                               q['base_damage'], q['base_armor']))
        return cls(auth == 1, tick, p1, p2, world, ents)


def empty_room_tiles(width: int, height: int, stairs: typing.Tuple[int, int]) -> np.ndarray:
    """The tile grid EmptyDungeonGenerator builds (worldgen.py:34-42) for a given staircase."""
    tiles = np.full((width, height), int(Tile.Ground), 'int32')
    tiles[[0, -1], :] = int(Tile.Wall)
    tiles[:, [0, -1]] = int(Tile.Wall)
    if stairs[0] != _abi.NO_STAIRS:
        tiles[stairs[0], stairs[1]] = int(Tile.StaircaseDown)
    return tiles


class BatchedGameState:
    """N games as SoA tensors on one CUDA device.

    Planes (game i = row i): ``pos`` uint8[N,4] (x1,y1,x2,y2), ``hp`` int16[N,2],
    ``depth`` int32[N,2], ``stairs`` uint8[N,4] (staircase of each player's current level),
    ``tick`` int32[N], ``episode`` int32[N] (bit pattern of a uint32), ``status`` uint8[N]
    (UpdateResult code), plus NPC slot planes when ``cfg.n_npc > 0``.
    """

    PLANES = ('pos', 'hp', 'depth', 'stairs', 'tick', 'episode', 'status',
              'npc_pos', 'npc_hp', 'npc_depth')
    _LAYOUT = frozenset(PLANES + ('sched', 'flat', 'fixed_tiles', 'fixed_ground', 'cfg', 'game_id_base', 'n'))

    def __setattr__(self, name, value):
        # (re)binding a plane, the scratch, the bonus plane or the config changes what the C structs must point at:
        # bump a version so that BatchedUpdater re-marshals them (one integer compare per tick otherwise)
        object.__setattr__(self, name, value)
        if name in self._LAYOUT:
            object.__setattr__(self, '_layout_version', getattr(self, '_layout_version', 0) + 1)

    def __init__(self, cfg: SimConfig, n: int, device='cuda', game_id_base: int = 0):
        cfg.validate()
        self.cfg = cfg
        self.n = int(n)
        self.device = torch.device(device)
        self.game_id_base = int(game_id_base)
        dev = self.device
        self._alloc_word_planes()
        self.depth = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        self.status = torch.ones((n,), dtype=torch.uint8, device=dev)
        e = max(cfg.n_npc, 1)
        self.npc_pos = torch.zeros((n, e, 2), dtype=torch.uint8, device=dev)
        self.npc_hp = torch.zeros((n, e), dtype=torch.int16, device=dev)
        self.npc_depth = torch.full((n, e), -1, dtype=torch.int32, device=dev)
        # scratch of the tick kernel (OrxState.sched): tile counter + per-tile hand-over words, never shared
        self.sched = torch.zeros((_abi.sched_words(n),), dtype=torch.int32, device=dev)
        # Modifier seam (OrxState.flat): None until enable_flat_bonuses() -- the tick then never looks for it
        self.flat = None
        # the fixed map (shared by all games) lives beside the state
        self.fixed_tiles = self.fixed_ground = None
        self._fixed_stairs = (_abi.NO_STAIRS, _abi.NO_STAIRS)
        if cfg.dgen_kind == _abi.DGEN_FIXED:
            tiles, ground, stairs = cfg.fixed_tables()
            self.fixed_tiles = torch.from_numpy(tiles).to(dev)
            # uint16 table shipped as int16 bit pattern (torch has no general uint16 support)
            self.fixed_ground = torch.from_numpy(ground.view(np.int16).copy()).to(dev)
            self._fixed_stairs = stairs

    WORD_PLANES = (('pos', torch.uint8, (4,)), ('hp', torch.int16, (2,)), ('stairs', torch.uint8, (4,)),
                   ('tick', torch.int32, ()), ('episode', torch.int32, ()))

    def _alloc_word_planes(self):
        """The five 4-byte-per-game planes are carved out of ONE zeroed allocation at a common pitch (a
        multiple of 128 bytes), in the order pos, hp, stairs, tick, episode: each stays an ordinary
        contiguous tensor, and together they form the u32[5][n] array whose {256 games x 5 planes} boxes the
        tick kernel moves with one tensor-map copy each (include/orx.h, OrxState)."""
        n = self.n
        pitch = max(128, (4 * n + 127) // 128 * 128)
        self._word_planes = torch.zeros((5 * pitch,), dtype=torch.uint8, device=self.device)
        for k, (name, dtype, shape) in enumerate(self.WORD_PLANES):
            plane = self._word_planes[k * pitch:k * pitch + 4 * n].view(dtype).view((n,) + shape)
            setattr(self, name, plane)

    # -- ABI views -----------------------------------------------------------------------------
    def c_struct(self) -> _abi.OrxState:
        st = _abi.OrxState()
        for name in self.PLANES:
            setattr(st, name, getattr(self, name).data_ptr())
        st.sched = self.sched.data_ptr() if self.sched.is_cuda else None
        st.sched_words = int(self.sched.numel())
        st.flat = self.flat.data_ptr() if self.flat is not None else None
        return st

    def c_config(self, **overrides) -> _abi.OrxConfig:
        """OrxConfig for this state; the updater overrides despawn_strat/max_ticks/auto_reset."""
        cfg = self.cfg
        if self.fixed_tiles is not None:
            c = cfg.to_c(self.fixed_tiles.data_ptr(), self.fixed_ground.data_ptr(),
                         int(self.fixed_ground.numel()), self._fixed_stairs)
        else:
            c = cfg.to_c()
        for k, v in overrides.items():
            setattr(c, k, v)
        return c

    # -- convenience views ---------------------------------------------------------------------
    @property
    def player_1(self):
        return {'x': self.pos[:, 0], 'y': self.pos[:, 1], 'depth': self.depth[:, 0], 'health': self.hp[:, 0]}

    @property
    def player_2(self):
        return {'x': self.pos[:, 2], 'y': self.pos[:, 3], 'depth': self.depth[:, 1], 'health': self.hp[:, 1]}

    def planes_cpu(self) -> typing.Dict[str, np.ndarray]:
        return {name: getattr(self, name).cpu().numpy() for name in self.PLANES}

    def clone(self) -> 'BatchedGameState':
        o = BatchedGameState.__new__(BatchedGameState)
        o.__dict__.update(self.__dict__)
        o._alloc_word_planes()
        word = {name for name, _, _ in self.WORD_PLANES}
        for name in self.PLANES:
            if name in word:
                getattr(o, name).copy_(getattr(self, name))
            else:
                setattr(o, name, getattr(self, name).clone())
        o.sched = torch.zeros_like(self.sched)
        o.flat = self.flat.clone() if self.flat is not None else None
        return o

    def _fingerprint(self):
        """What a checkpoint must agree on to be loadable into this state (plain ints / tuples)."""
        c = self.cfg
        return {'n': self.n, 'seed': int(c.seed), 'width': c.width, 'height': c.height, 'dgen_kind': c.dgen_kind,
                'start_kind': c.start_kind, 'start_depth': tuple(c.start_depth), 'n_npc': c.n_npc,
                'hp': tuple(c.hp), 'damage': tuple(c.damage), 'armor': tuple(c.armor)}

    def state_dict(self):
        """Checkpoint: a snapshot (every plane is copied, so ticking on does not change it) of plain tensors +
        scalars, ``torch.save``-able, with the batch size and the configuration the planes belong to."""
        d = {name: getattr(self, name).detach().clone() for name in self.PLANES}
        d['game_id_base'] = self.game_id_base
        d['fingerprint'] = self._fingerprint()
        if self.flat is not None:
            d['flat'] = self.flat.detach().clone()
        return d

    def load_state_dict(self, d):
        """Restores a ``state_dict()`` snapshot. The checkpoint must come from a state of the same batch size and
        configuration (seed, room, generator, start, stats, NPC slots): a Philox stream or a plane shape that does
        not match would silently give other games."""
        fp = d.get('fingerprint')
        if fp is not None:
            mine = self._fingerprint()
            bad = [k for k in mine if k in fp and (tuple(fp[k]) if isinstance(fp[k], (list, tuple)) else fp[k]) != mine[k]]
            if bad:
                raise ValueError('checkpoint does not belong to this state: ' + ', '.join(f'{k}: {fp[k]!r} != {mine[k]!r}' for k in bad))
        for name in self.PLANES:
            src, dst = d[name], getattr(self, name)
            if tuple(src.shape) != tuple(dst.shape) or src.dtype != dst.dtype:
                raise ValueError(f'checkpoint plane {name}: {tuple(src.shape)} {src.dtype} does not fit {tuple(dst.shape)} {dst.dtype}')
        for name in self.PLANES:
            getattr(self, name).copy_(d[name])
        self.game_id_base = int(d['game_id_base'])
        if d.get('flat') is not None:
            self.enable_flat_bonuses().copy_(d['flat'])
        elif self.flat is not None:
            self.flat.zero_()
        self.sched.zero_()          # no launch is in flight across a restore

    def enable_flat_bonuses(self) -> torch.Tensor:
        """Allocates (once) and returns the Modifier seam plane int8[N,2,3]: per player the sums of
        ``flat_damage``, ``flat_armor``, ``flat_max_health`` over the modifiers its entity carries
        (game/modifiers.py:102-108, folded in by Entity.on_tick, game/attribles.py:21-43). A hit then deals
        ``(base_damage + flat_damage) - (base_armor + flat_armor)`` of the attacker. Zero = no modifier."""
        if self.flat is None:
            self.flat = torch.zeros((self.n, 2, 3), dtype=torch.int8, device=self.device)
        return self.flat

    def set_npc(self, lane: int, slot: int, depth: int, x: int, y: int, health: int):
        """Places a static NPC (an extra Entity after the players, updater.py:116-128)."""
        if not 0 <= slot < self.cfg.n_npc:
            raise ValueError('slot out of range')
        self.npc_depth[lane, slot] = depth
        self.npc_pos[lane, slot, 0] = x
        self.npc_pos[lane, slot, 1] = y
        self.npc_hp[lane, slot] = health

    # -- interop with single-game objects ------------------------------------------------------
    def level_tiles(self, stairs: typing.Tuple[int, int]) -> np.ndarray:
        if self.cfg.dgen_kind == _abi.DGEN_FIXED:
            return np.asarray(self.cfg.fixed_tiles, dtype='int32').copy()
        return empty_room_tiles(self.cfg.width, self.cfg.height, stairs)

    def to_game_state(self, i: int, planes=None) -> GameState:
        """Materialises game ``i`` as a host GameState holding the levels the players stand on
        (the levels in between are re-derivable but not stored). Pass ``planes=planes_cpu()``
        when converting many lanes."""
        p = self.planes_cpu() if planes is None else planes
        cfg = self.cfg
        ents = []
        dungeons = {}
        for k in range(2):
            d = int(p['depth'][i, k])
            ents.append(Entity(k + 1, d, int(p['pos'][i, 2 * k]), int(p['pos'][i, 2 * k + 1]),
                               int(p['hp'][i, k]), cfg.hp[k], cfg.damage[k], cfg.armor[k]))
            st = (int(p['stairs'][i, 2 * k]), int(p['stairs'][i, 2 * k + 1]))
            dungeons[d] = Dungeon(self.level_tiles(st))
        for k in range(cfg.n_npc):
            d = int(p['npc_depth'][i, k])
            if d >= 0:
                hpv = int(p['npc_hp'][i, k])
                ents.append(Entity(3 + k, d, int(p['npc_pos'][i, k, 0]), int(p['npc_pos'][i, k, 1]),
                                   hpv, hpv, 0, 0))
        return GameState(True, int(p['tick'][i]), 1, 2, World(dungeons), ents)

    def render(self, i: int, player: int = 1, planes=None) -> str:
        """Text dump of the level ``player`` (1 or 2) of game ``i`` stands on, with the glyphs of the
        reference's curses spectator (optimax_rogue_cmdspec/map.py:14-24): '.' ground, '#' wall,
        '\\' staircase, 'o' player 1, 'x' player 2, 'e' NPC."""
        gs = self.to_game_state(i, planes)
        me = gs.player_1 if player == 1 else gs.player_2
        tiles = gs.world.dungeons[me.depth].tiles
        glyph = {int(Tile.Ground): '.', int(Tile.Wall): '#', int(Tile.StaircaseDown): '\\'}
        rows = [[glyph[int(tiles[x, y])] for x in range(tiles.shape[0])] for y in range(tiles.shape[1])]
        for ent in gs.entities:
            if ent.depth == me.depth:
                rows[ent.y][ent.x] = 'o' if ent.iden == 1 else 'x' if ent.iden == 2 else 'e'
        head = f'game {self.game_id_base + i} tick {gs.tick} depth {me.depth} hp {gs.player_1.health}/{gs.player_2.health}'
        return head + '\n' + '\n'.join(''.join(r) for r in rows)

    def load_game_state(self, i: int, gs: GameState):
        """Writes a host GameState into lane ``i`` (players + their levels' staircases)."""
        for k, ent in enumerate((gs.player_1, gs.player_2)):
            self.pos[i, 2 * k] = ent.x
            self.pos[i, 2 * k + 1] = ent.y
            self.depth[i, k] = ent.depth
            self.hp[i, k] = ent.health
            dung = gs.world.dungeons[ent.depth]
            hits = np.argwhere(np.asarray(dung.tiles) == int(Tile.StaircaseDown))
            sx, sy = (int(hits[0][0]), int(hits[0][1])) if len(hits) else (_abi.NO_STAIRS, _abi.NO_STAIRS)
            self.stairs[i, 2 * k] = sx
            self.stairs[i, 2 * k + 1] = sy
        self.tick[i] = gs.tick
        self.status[i] = 1
