"""Level and episode generators with the reference's class names
(optimax_rogue/logic/worldgen.py). They carry configuration only; the draws happen on the device
inside ``orx_reset`` / ``orx_step`` from the shared Philox schedule."""
import numpy as np

from .. import _abi
from ..config import SimConfig


class DungeonGenerator:
    """worldgen.py:9-26"""
    kind = None

    def __init__(self, width: int, height: int) -> None:
        self.width = width
        self.height = height


class EmptyDungeonGenerator(DungeonGenerator):
    """Empty room with a wall border and a random staircase (worldgen.py:28-43)."""
    kind = _abi.DGEN_EMPTY


class FixedDungeonGenerator(DungeonGenerator):
    """Plugin generator: the same tile grid (uint8[width, height] of Tile codes) at every depth."""
    kind = _abi.DGEN_FIXED

    def __init__(self, tiles) -> None:
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        super().__init__(tiles.shape[0], tiles.shape[1])
        self.tiles = tiles


class GameStartGenerator:
    """worldgen.py:47-58; ``setup_game`` returns a BatchedGameState of ``n`` fresh games."""
    start_kind = None

    def __init__(self, dgen: DungeonGenerator = None):
        self.dgen = dgen if dgen is not None else EmptyDungeonGenerator(60, 10)
        self.hp = (10, 10)        # worldgen.py:85-86
        self.damage = (2, 2)
        self.armor = (1, 1)

    def _depths(self):
        return (0, 0)

    def sim_config(self, seed: int = 0, n_npc: int = 0) -> SimConfig:
        return SimConfig(width=self.dgen.width, height=self.dgen.height, dgen_kind=self.dgen.kind,
                         start_kind=self.start_kind, start_depth=self._depths(), hp=self.hp,
                         damage=self.damage, armor=self.armor, seed=seed, n_npc=n_npc,
                         fixed_tiles=getattr(self.dgen, 'tiles', None))

    def setup_game(self, n: int = 1, *, seed: int = 0, device='cuda', game_id_base: int = 0,
                   n_npc: int = 0, episodes=None):
        """Creates ``n`` games; game i draws from Philox stream (seed, game_id_base + i, episode)."""
        from ..game.state import BatchedGameState
        from .updater import reset_games
        gs = BatchedGameState(self.sim_config(seed, n_npc), n, device, game_id_base)
        if episodes is not None:
            gs.episode.copy_(episodes)
        reset_games(gs)
        return gs


class TogetherGameStartGenerator(GameStartGenerator):
    """Both players on depth 0 of one level, distinct random Ground tiles (worldgen.py:60-87)."""
    start_kind = _abi.START_TOGETHER


class SeparatedGameStartGenerator(GameStartGenerator):
    """Players on separate levels (worldgen.py:91-135)."""
    start_kind = _abi.START_SEPARATED

    def __init__(self, dgen: DungeonGenerator = None, p1_depth: int = 0, p2_depth: int = 1000):
        if p1_depth == p2_depth:
            raise ValueError('cannot use SeparatedGameStartGenerator for '
                             + f'p1_depth=p2_depth={p1_depth}')
        super().__init__(dgen)
        self.p1_depth = p1_depth
        self.p2_depth = p2_depth

    def _depths(self):
        return (self.p1_depth, self.p2_depth)
