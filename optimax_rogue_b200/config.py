"""Simulation configuration: the knobs the reference spreads over
``server/main.py:18-40`` (``--width``, ``--height``, ``--dsunused``, ``--maxticks``,
``--gamestart``), ``logic/worldgen.py:85-86`` (entity stats) and the Philox seed."""
import dataclasses
import typing

import numpy as np

from . import _abi


@dataclasses.dataclass
class SimConfig:
    width: int = 60                     # server/main.py:27
    height: int = 10                    # server/main.py:28
    dgen_kind: int = _abi.DGEN_EMPTY
    start_kind: int = _abi.START_TOGETHER
    start_depth: typing.Tuple[int, int] = (0, 0)
    despawn_strat: int = 1              # DungeonDespawningStrategy.Unreachable (updater.py:48)
    max_ticks: int = 0                  # None/0 = never (updater.py:158)
    hp: typing.Tuple[int, int] = (10, 10)
    damage: typing.Tuple[int, int] = (2, 2)
    armor: typing.Tuple[int, int] = (1, 1)
    auto_reset: bool = False
    n_npc: int = 0
    seed: int = 0
    fixed_tiles: typing.Optional[np.ndarray] = None   # uint8[W,H] Tile codes (DGEN_FIXED)
    path_flags: int = 0                 # _abi.PATH_* bits: pins a kernel path for tests / A-B runs, never changes results
    overlap_ticks: bool = False         # throughput mode (ORX_PATH_TILE_FLAGS): ticks enqueued back to back on this state and
                                        # on others in the same stream overlap run by run (a run = the tiles of one CTA).
                                        # For several states in flight; 2-3x slower when a tick has to wait for something
                                        # (other kernels between two ticks, one small state ticked again and again)

    def __setattr__(self, name, value):
        # every assignment bumps a version number: holders of structs marshalled from this config (BatchedUpdater)
        # notice a later mutation with one integer compare instead of re-reading fifteen fields per tick
        object.__setattr__(self, name, value)
        object.__setattr__(self, '_version', getattr(self, '_version', 0) + 1)

    def validate(self):
        if not (4 <= self.width <= _abi.MAX_DIM and 4 <= self.height <= _abi.MAX_DIM):
            raise ValueError(f'width/height must be in [4, {_abi.MAX_DIM}]')
        if not 0 <= self.n_npc <= _abi.MAX_NPC:
            raise ValueError(f'n_npc must be in [0, {_abi.MAX_NPC}]')
        if self.despawn_strat not in (1, 2):
            raise ValueError(f'Unknown despawn strat {self.despawn_strat}')
        if self.start_kind == _abi.START_SEPARATED and self.start_depth[0] == self.start_depth[1]:
            # worldgen.py:112-114
            raise ValueError('cannot use SeparatedGameStartGenerator for '
                             f'p1_depth=p2_depth={self.start_depth[0]}')
        if self.dgen_kind == _abi.DGEN_FIXED:
            t = self.fixed_tiles
            if t is None or t.shape != (self.width, self.height):
                raise ValueError('fixed_tiles must be uint8[width, height]')
            if ((t < 1) | (t > 3)).any():
                raise ValueError('fixed_tiles holds codes outside Tile (1..3)')
            if int((t == 1).sum()) < 2 + self.n_npc:
                raise ValueError('fixed_tiles needs at least one Ground tile per entity')
        for v in (*self.hp, *self.damage, *self.armor):
            if not -32768 <= v <= 32767:
                raise ValueError('entity stats must fit int16')
        # what include/orx.h's check_common enforces: health lives in int16 planes and a hit is subtracted from them
        if not all(1 <= v <= 32767 for v in self.hp):
            raise ValueError('hp must be in [1, 32767]')
        if not all(abs(d - a) <= 32767 - 254 for d, a in zip(self.damage, self.armor)):
            raise ValueError('|damage - armor| must leave room for the int8 flat bonuses in an int16 hit')

    def fixed_tables(self):
        """(tiles uint8[W*H] x-major, ground uint16[#Ground], stairs (x, y)) for DGEN_FIXED."""
        t = np.ascontiguousarray(self.fixed_tiles, dtype=np.uint8)
        flat = t.reshape(-1)                         # x-major: flat = x*H + y (world.py:60)
        ground = np.flatnonzero(flat == 1).astype(np.uint16)
        hits = np.argwhere(t == 3)                   # world.py:54: first argwhere
        stairs = (int(hits[0][0]), int(hits[0][1])) if len(hits) else (_abi.NO_STAIRS, _abi.NO_STAIRS)
        return flat.copy(), ground, stairs

    def to_c(self, tiles_ptr=0, ground_ptr=0, n_ground=0, stairs=(255, 255)) -> _abi.OrxConfig:
        c = _abi.OrxConfig()
        c.struct_size = _abi.C.sizeof(_abi.OrxConfig)
        c.width, c.height = self.width, self.height
        c.dgen_kind, c.start_kind = self.dgen_kind, self.start_kind
        c.start_depth[0], c.start_depth[1] = self.start_depth
        c.despawn_strat = self.despawn_strat
        c.max_ticks = int(self.max_ticks or 0)
        for k in range(2):
            c.hp[k], c.damage[k], c.armor[k] = self.hp[k], self.damage[k], self.armor[k]
        c.auto_reset = 1 if self.auto_reset else 0
        c.n_npc = self.n_npc
        c.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        c.fixed_tiles = tiles_ptr or None
        c.fixed_ground = ground_ptr or None
        c.fixed_n_ground = n_ground
        c.fixed_stairs[0], c.fixed_stairs[1] = stairs
        c.path_flags = int(self.path_flags) | (_abi.PATH_TILE_FLAGS if self.overlap_ticks else 0)
        return c
