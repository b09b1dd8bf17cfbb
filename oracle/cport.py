"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

numpy + ctypes front end of ``oracle/orx_oracle.c`` (the plain-C restatement of the
reference tick). Same SoA planes and the same ``OrxConfig`` as the CUDA library, but host
memory, one game at a time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from optimax_rogue_b200 import _abi
from optimax_rogue_b200.config import SimConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liborx_oracle.so')
_SRC = os.path.join(_HERE, 'orx_oracle.c')
_SRC_R1 = os.path.join(_HERE, 'orx_r1_oracle.c')
_HDR = os.path.join(_HERE, '..', 'include', 'orx.h')

_PROTOS = {
    'oro_philox': (None, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'oro_set_threads': (C.c_int, [C.c_int]),
    'oro_reset': (C.c_int, [C.POINTER(_abi.OrxConfig), C.POINTER(_abi.OrxState), C.c_void_p,
                            C.c_int, C.c_int64, C.c_uint64]),
    'oro_step': (C.c_int, [C.POINTER(_abi.OrxConfig), C.POINTER(_abi.OrxState), C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64]),
    'oro_bot_moves': (C.c_int, [C.POINTER(_abi.OrxConfig), C.POINTER(_abi.OrxState), C.c_int,
                                C.c_int, C.c_void_p, C.c_int64, C.c_uint64]),
    'oro_rollout': (C.c_int, [C.POINTER(_abi.OrxConfig), C.POINTER(_abi.OrxState), C.c_int,
                              C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_uint64]),
    'oro_r1_reset': (C.c_int, [C.POINTER(_abi.OrxR1Config), C.POINTER(_abi.OrxR1State), C.c_void_p,
                               C.c_int, C.c_int64, C.c_uint64]),
    'oro_r1_step': (C.c_int, [C.POINTER(_abi.OrxR1Config), C.POINTER(_abi.OrxR1State), C.c_void_p,
                              C.c_void_p, C.c_int64, C.c_uint64]),
    'oro_r1_rollout': (C.c_int, [C.POINTER(_abi.OrxR1Config), C.POINTER(_abi.OrxR1State), C.c_int,
                                 C.c_void_p, C.c_int64, C.c_uint64]),
    'oro_r1_observe': (C.c_int, [C.POINTER(_abi.OrxR1Config), C.POINTER(_abi.OrxR1State), C.c_void_p, C.c_int,
                                 C.c_int64]),
    'oro_r1_step_events': (C.c_int, [C.POINTER(_abi.OrxR1Config), C.POINTER(_abi.OrxR1State), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_uint64]),
    'oro_r1_bot_moves': (C.c_int, [C.POINTER(_abi.OrxR1Config), C.POINTER(_abi.OrxR1State), C.c_int, C.c_int,
                                   C.c_void_p, C.c_int64, C.c_uint64]),
}


def build(force=False):
    """gcc build of the oracle into oracle/_build/ (OpenMP when the compiler has it)."""
    stale = (not os.path.exists(_SO)
             or os.path.getmtime(_SO) < max(os.path.getmtime(_SRC), os.path.getmtime(_SRC_R1), os.path.getmtime(_HDR)))
    if not (force or stale):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    base = ['gcc', '-O2', '-fPIC', '-std=c11', '-Wall', '-shared', '-o', _SO, _SRC, _SRC_R1]
    for extra in (['-fopenmp'], []):
        r = subprocess.run(base + extra, capture_output=True, text=True)
        if r.returncode == 0:
            return _SO
    raise RuntimeError('oracle build failed:\n' + r.stderr)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = _abi.bind(C.CDLL(_SO), _PROTOS)
    return _lib


def set_threads(n=0):
    """Sets (n > 0) and returns the OpenMP thread count oro_rollout will use."""
    return lib().oro_set_threads(int(n))


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oro_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return tuple(int(x) for x in out)


class HostState:
    """The SoA planes of include/orx.h:OrxState as numpy arrays."""

    def __init__(self, n, n_npc=0):
        self.n, self.n_npc = n, n_npc
        self.pos = np.zeros((n, 4), np.uint8)
        self.hp = np.zeros((n, 2), np.int16)
        self.depth = np.zeros((n, 2), np.int32)
        self.stairs = np.zeros((n, 4), np.uint8)
        self.tick = np.zeros(n, np.int32)
        self.episode = np.zeros(n, np.uint32)
        self.status = np.ones(n, np.uint8)
        e = max(n_npc, 1)
        self.npc_pos = np.zeros((n, e, 2), np.uint8)
        self.npc_hp = np.zeros((n, e), np.int16)
        self.npc_depth = np.full((n, e), -1, np.int32)
        self.flat = None            # Modifier seam: int8[n, 2, 3] once enable_flat_bonuses() was called

    def enable_flat_bonuses(self):
        if self.flat is None:
            self.flat = np.zeros((self.n, 2, 3), np.int8)
        return self.flat

    PLANES = ('pos', 'hp', 'depth', 'stairs', 'tick', 'episode', 'status',
              'npc_pos', 'npc_hp', 'npc_depth')

    def c_struct(self):
        st = _abi.OrxState()
        for name in self.PLANES:
            setattr(st, name, getattr(self, name).ctypes.data)
        st.flat = self.flat.ctypes.data if self.flat is not None else None
        return st

    def copy(self):
        o = HostState.__new__(HostState)
        o.n, o.n_npc = self.n, self.n_npc
        for name in self.PLANES:
            setattr(o, name, getattr(self, name).copy())
        o.flat = None if self.flat is None else self.flat.copy()
        return o


class Oracle:
    """Batched driver over the C oracle; mirrors the product's BatchedUpdater calls."""

    def __init__(self, cfg: SimConfig, n: int, game_id_base: int = 0):
        cfg.validate()
        self.cfg, self.n, self.game_id_base = cfg, n, game_id_base
        self.state = HostState(n, cfg.n_npc)
        if cfg.dgen_kind == _abi.DGEN_FIXED:
            self._tiles, self._ground, stairs = cfg.fixed_tables()
            self.c_cfg = cfg.to_c(self._tiles.ctypes.data, self._ground.ctypes.data,
                                  len(self._ground), stairs)
        else:
            self.c_cfg = cfg.to_c()
        self.max_events = _abi.MAX_EVENTS_BASE + cfg.n_npc

    def reset(self, mask=None, bump_episode=False):
        st = self.state.c_struct()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().oro_reset(C.byref(self.c_cfg), C.byref(st), None if m is None else m.ctypes.data,
                        int(bump_episode), self.n, self.game_id_base)

    def step(self, moves, want_events=False):
        moves = np.ascontiguousarray(moves, np.uint8).reshape(self.n, 2)
        result = np.zeros(self.n, np.uint8)
        events = np.zeros((self.n, self.max_events, 2), np.int32) if want_events else None
        st = self.state.c_struct()
        lib().oro_step(C.byref(self.c_cfg), C.byref(st), moves.ctypes.data, result.ctypes.data,
                       None if events is None else events.ctypes.data, self.n, self.game_id_base)
        return result, events

    def bot_moves(self, bot_p1, bot_p2, out=None):
        moves = np.full((self.n, 2), 5, np.uint8) if out is None else out
        st = self.state.c_struct()
        lib().oro_bot_moves(C.byref(self.c_cfg), C.byref(st), bot_p1, bot_p2, moves.ctypes.data,
                            self.n, self.game_id_base)
        return moves

    def rollout(self, bot_p1, bot_p2, n_ticks, stats=None):
        stats = np.zeros(_abi.STAT_COUNT, np.uint64) if stats is None else stats
        st = self.state.c_struct()
        lib().oro_rollout(C.byref(self.c_cfg), C.byref(st), bot_p1, bot_p2, n_ticks,
                          stats.ctypes.data, self.n, self.game_id_base)
        return stats


def decode_events(events_i32):
    """int32[..., 2] raw OrxEvent words -> uint8/int32 fields (kind, iden, a, b, depth)."""
    w0 = events_i32[..., 0].view(np.uint32) if events_i32.dtype == np.int32 else events_i32[..., 0]
    w0 = w0.astype(np.uint32)
    kind = (w0 & 0xFF).astype(np.int32)
    iden = ((w0 >> 8) & 0xFF).astype(np.int32)
    a = ((w0 >> 16) & 0xFF).astype(np.int32)
    b = ((w0 >> 24) & 0xFF).astype(np.int32)
    return np.stack([kind, iden, a, b, events_i32[..., 1].astype(np.int32)], axis=-1)


# ---- ruleset R1 (parity unpinned: pins the CUDA kernel to docs/RULESET_R1.md, not to the reference)
def r1_config(width=60, height=10, max_ticks=0, auto_reset=False, wall_density=26, seed=0):
    c = _abi.OrxR1Config()
    c.struct_size = C.sizeof(_abi.OrxR1Config)
    c.width, c.height, c.max_ticks = width, height, int(max_ticks or 0)
    c.auto_reset, c.wall_density, c.seed = int(auto_reset), wall_density, seed & 0xFFFFFFFFFFFFFFFF
    return c


class R1HostState:
    def __init__(self, n):
        self.n = n
        for name, dt, shape in _abi.R1_PLANES:
            arr = np.zeros((n,) + shape, dtype=dt)
            setattr(self, name, arr)
        self.status[:] = 1

    def c_struct(self):
        st = _abi.OrxR1State()
        for name, _, _ in _abi.R1_PLANES:
            setattr(st, name, getattr(self, name).ctypes.data)
        return st


class R1Oracle:
    def __init__(self, n, game_id_base=0, **cfg):
        self.n, self.game_id_base = n, game_id_base
        self.c_cfg = r1_config(**cfg)
        self.state = R1HostState(n)

    def reset(self, mask=None, bump_episode=False):
        st = self.state.c_struct()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().oro_r1_reset(C.byref(self.c_cfg), C.byref(st), None if m is None else m.ctypes.data,
                           int(bump_episode), self.n, self.game_id_base)

    def step(self, moves):
        moves = np.ascontiguousarray(moves, np.uint8).reshape(self.n, 2)
        result = np.zeros(self.n, np.uint8)
        st = self.state.c_struct()
        lib().oro_r1_step(C.byref(self.c_cfg), C.byref(st), moves.ctypes.data, result.ctypes.data, self.n,
                          self.game_id_base)
        return result

    def step_events(self, moves, max_events=_abi.R1_MAX_EVENTS):
        """The tick and its replication log: (result uint8[n], records int32[n, max_events, 2] as the kernel writes them;
        slots behind a game's terminator are zero here)."""
        moves = np.ascontiguousarray(moves, np.uint8).reshape(self.n, 2)
        result = np.zeros(self.n, np.uint8)
        events = np.zeros((self.n, max_events, 2), np.int32)
        st = self.state.c_struct()
        lib().oro_r1_step_events(C.byref(self.c_cfg), C.byref(st), moves.ctypes.data, result.ctypes.data,
                                 events.ctypes.data, max_events, self.n, self.game_id_base)
        return result, events

    def bot_moves(self, bot1, bot2, moves=None):
        moves = np.zeros((self.n, 2), np.uint8) if moves is None else moves
        st = self.state.c_struct()
        lib().oro_r1_bot_moves(C.byref(self.c_cfg), C.byref(st), bot1, bot2, moves.ctypes.data, self.n, self.game_id_base)
        return moves

    def observe(self, radius=4):
        obs = np.zeros((self.n, 2, _abi.R1_OBS_LEN), np.int16)
        st = self.state.c_struct()
        lib().oro_r1_observe(C.byref(self.c_cfg), C.byref(st), obs.ctypes.data, radius, self.n)
        return obs

    def rollout(self, n_ticks, stats=None):
        stats = np.zeros(_abi.STAT_COUNT, np.uint64) if stats is None else stats
        st = self.state.c_struct()
        lib().oro_r1_rollout(C.byref(self.c_cfg), C.byref(st), n_ticks, stats.ctypes.data, self.n,
                             self.game_id_base)
        return stats
