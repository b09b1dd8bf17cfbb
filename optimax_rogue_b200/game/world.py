"""Tiles, levels and the world-of-levels, mirroring optimax_rogue/game/world.py.

On the device a level is never stored: an ``EmptyDungeonGenerator`` level is a pure function of
(seed, game, episode, depth) and is re-derived from Philox, a fixed map is shared by all games.
``Dungeon``/``World`` here are host-side views used for interop (``BatchedGameState.to_game_state``)
and for the reference's binary wire format (world.py:73-92, :142-165).
"""
import enum
import io
import typing

import numpy as np


class Tile(enum.IntEnum):
    """world.py:10-17"""
    Ground = 1
    Wall = 2
    StaircaseDown = 3


class Dungeon:
    """The map of one level; ``tiles`` is int32[width, height] (x-major), world.py:19-66."""

    def __init__(self, tiles: np.ndarray) -> None:
        self.tiles = tiles

    @property
    def width(self):
        return self.tiles.shape[0]

    @property
    def height(self):
        return self.tiles.shape[1]

    def is_blocked(self, x: int, y: int) -> bool:
        if x < 0 or x >= self.width or y < 0 or y >= self.height:
            return True
        return bool(self.tiles[x, y] == Tile.Wall)

    def get_unblocked(self) -> np.ndarray:
        return self.tiles != Tile.Wall

    def staircase(self) -> typing.Tuple[int, int]:
        resx, resy = tuple(np.argwhere(self.tiles == Tile.StaircaseDown)[0])
        return int(resx), int(resy)

    def to_prims(self) -> bytes:
        """Reference wire format: u32be width, u32be height, uint8 tiles x-major (world.py:73-83)."""
        arr = io.BytesIO()
        arr.write(int(self.width).to_bytes(4, 'big', signed=False))
        arr.write(int(self.height).to_bytes(4, 'big', signed=False))
        arr.write(self.tiles.astype('uint8').reshape(self.width * self.height).tobytes())
        return arr.getvalue()

    @classmethod
    def from_prims(cls, prims: bytes) -> 'Dungeon':
        wid = int.from_bytes(prims[0:4], 'big', signed=False)
        hei = int.from_bytes(prims[4:8], 'big', signed=False)
        tmp = np.frombuffer(prims[8:8 + wid * hei], dtype='uint8').reshape(wid, hei)
        return cls(tmp.astype('int32'))

    def __eq__(self, other):
        return isinstance(other, Dungeon) and self.tiles.shape == other.tiles.shape \
            and bool((self.tiles != other.tiles).sum() == 0)


class World:
    """dict depth -> Dungeon (world.py:101-180)."""

    def __init__(self, dungeons: typing.Dict[int, Dungeon]) -> None:
        self.dungeons = dungeons

    def get_at_depth(self, ind: int) -> Dungeon:
        return self.dungeons[ind]

    def set_at_depth(self, ind: int, dung: Dungeon) -> None:
        self.dungeons[ind] = dung

    def del_at_depth(self, ind: int) -> None:
        del self.dungeons[ind]

    def to_prims(self) -> bytes:
        """u32be count, then per level u32be depth, u64be size, Dungeon bytes (world.py:142-151)."""
        arr = io.BytesIO()
        arr.write(len(self.dungeons).to_bytes(4, 'big', signed=False))
        for depth, dung in self.dungeons.items():
            serd = dung.to_prims()
            arr.write(int(depth).to_bytes(4, 'big', signed=False))
            arr.write(len(serd).to_bytes(8, 'big', signed=False))
            arr.write(serd)
        return arr.getvalue()

    @classmethod
    def from_prims(cls, prims: bytes) -> 'World':
        arr = io.BytesIO(prims)
        num = int.from_bytes(arr.read(4), 'big', signed=False)
        dungeons = dict()
        for _ in range(num):
            depth = int.from_bytes(arr.read(4), 'big', signed=False)
            size = int.from_bytes(arr.read(8), 'big', signed=False)
            dungeons[depth] = Dungeon.from_prims(arr.read(size))
        return cls(dungeons)

    def __eq__(self, other):
        if not isinstance(other, World) or len(self.dungeons) != len(other.dungeons):
            return False
        return all(d in other.dungeons and dung == other.dungeons[d] for d, dung in self.dungeons.items())
