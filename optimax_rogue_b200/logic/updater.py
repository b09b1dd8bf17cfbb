"""``BatchedUpdater``: the reference's ``Updater`` (optimax_rogue/logic/updater.py:52-162) over N
games. Same constructor arguments and ``update(game_state, player1_move, player2_move)`` call;
moves and results are tensors with one element per game. All work is done by the sm_100a kernels
in ``liborx.so`` through the C ABI -- there is no CPU path."""
import ctypes as C
import enum
import typing
import weakref

import torch

from .. import _abi, _lib
from ..game.state import BatchedGameState
from .worldgen import DungeonGenerator


class UpdateResult(enum.IntEnum):
    """updater.py:16-21"""
    InProgress = 1
    Player1Win = 2
    Player2Win = 3
    Tie = 4


class DungeonDespawningStrategy(enum.IntEnum):
    """updater.py:47-50"""
    Unreachable = 1
    Unused = 2


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """``torch.cuda.device(dev)`` only when dev is not already current (the context manager costs
    several microseconds per call, which matters for the per-step host path)."""

    def __init__(self, device):
        self.ctx = None if torch.cuda.current_device() == (device.index or 0) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _require_cuda(gs: BatchedGameState):
    if gs.device.type != 'cuda':
        raise RuntimeError('BatchedGameState must live on a CUDA device: there is no CPU fallback')


def reset_games(gs: BatchedGameState, mask: typing.Optional[torch.Tensor] = None,
                bump_episode: bool = False, **cfg_overrides):
    """Episode reset for the masked lanes (setup_game, worldgen.py:77-87 / :124-135)."""
    _require_cuda(gs)
    cfg = gs.c_config(**cfg_overrides)
    st = gs.c_struct()
    mptr = None
    if mask is not None:
        mask = mask.to(device=gs.device, dtype=torch.uint8).contiguous()
        mptr = mask.data_ptr()
    with _on_device(gs.device):
        rc = _lib.lib().orx_reset(C.byref(cfg), C.byref(st), mptr, int(bump_episode), gs.n,
                                  gs.game_id_base, _stream_ptr(gs.device))
    _lib.check(rc, 'orx_reset')


class BatchedUpdater:
    """Moves N games forward in time.

    Attributes mirror the reference: ``dgen``, ``despawn_strat``, ``max_ticks`` (None = never),
    and ``current_update_order`` (here a per-game int64 tensor, created on first use).
    ``auto_reset=True`` re-initialises a finished game inside the same tick (episode + 1) so a
    rollout never stalls; otherwise finished lanes are frozen until ``reset_games``.
    """

    def __init__(self, dgen: DungeonGenerator, despawn_strat: DungeonDespawningStrategy,
                 max_ticks: typing.Optional[int] = None, *, auto_reset: bool = False):
        if int(despawn_strat) not in (1, 2):
            raise ValueError(f'Unknown despawn strat {despawn_strat} (type={type(despawn_strat)})')
        self.dgen = dgen
        self.despawn_strat = DungeonDespawningStrategy(int(despawn_strat))
        self.max_ticks = max_ticks
        self.auto_reset = auto_reset
        self.current_update_order = None
        self.track_order = True        # False: want_events=True returns the records without counting them (saves two passes over them)
        self.path_flags = 0            # _abi.PATH_* bits OR-ed into the state's: pins a kernel path (tests, A/B runs)
        self._cache = None

    # -- plumbing ------------------------------------------------------------------------------
    def _cfg(self, gs: BatchedGameState):
        # The marshalled structs are kept per game state, behind a WEAK reference (CPython reuses ids of freed objects:
        # identity alone is not a key) and two version numbers that the state and its config bump whenever a plane,
        # the scratch, the bonus plane or a config field is (re)assigned.
        cached = self._cache
        key = (gs._layout_version, gs.cfg._version, int(self.despawn_strat), int(self.max_ticks or 0), int(self.auto_reset),
               int(self.path_flags))
        if cached is None or cached[0]() is not gs or cached[1] != key:
            c = gs.cfg
            if (c.width, c.height, c.dgen_kind) != (self.dgen.width, self.dgen.height, self.dgen.kind):
                raise ValueError('updater.dgen does not match the generator the game state was built with')
            cfg = gs.c_config(despawn_strat=int(self.despawn_strat), max_ticks=int(self.max_ticks or 0),
                              auto_reset=int(self.auto_reset), path_flags=int(self.path_flags) | int(c.path_flags) | (_abi.PATH_TILE_FLAGS if c.overlap_ticks else 0))
            cached = self._cache = (weakref.ref(gs), key, cfg, gs.c_struct())
        return cached[2], cached[3]

    @staticmethod
    def _as_moves(gs, player1_move, player2_move):
        if (player2_move is None and isinstance(player1_move, torch.Tensor) and player1_move.dtype == torch.uint8
                and player1_move.dim() == 2 and player1_move.shape[0] == gs.n and player1_move.shape[1] == 2
                and player1_move.is_contiguous()):
            return player1_move                       # fast path: already a uint8[N,2] command block
        if player2_move is None:
            mv = player1_move
        else:
            mv = torch.stack((torch.as_tensor(player1_move), torch.as_tensor(player2_move)), dim=1)
        if mv.dtype != torch.uint8:
            mv = mv.to(torch.uint8)
        if tuple(mv.shape) != (gs.n, 2):
            raise ValueError(f'moves must have shape ({gs.n}, 2), got {tuple(mv.shape)}')
        return mv.contiguous()

    # -- the tick --------------------------------------------------------------------------------
    def update(self, game_state: BatchedGameState, player1_move, player2_move=None, *,
               want_events: bool = False, out: typing.Optional[torch.Tensor] = None, packed: bool = False):
        """One tick for every game (updater.py:76-162).

        ``player1_move``/``player2_move``: uint8[N] Move codes, or pass a single uint8[N,2]
        tensor as ``player1_move`` (fast path, no stacking). CUDA tensors are consumed in place and the
        tick is only ENQUEUED on the current stream (the result tensor is ordered after it on that stream).
        CPU tensors go through ``orx_step_host_sync``: the call returns when the tick is done, the result is a
        CPU tensor that can be read at once, and the moves tensor may be overwritten at once (for an
        asynchronous host loop use ``host_stepper(..., sync=False)``). Without ``out`` the CPU result is a
        pinned buffer owned by the updater and overwritten by its next host-side ``update``. Returns ``(result uint8[N] of UpdateResult codes, events)`` where
        events is None or an int32[N, max_events, 2] tensor of raw OrxEvent records
        (see logic/updates.py:decode_events).

        ``packed=True``: ``player1_move`` is one uint8[N] tensor holding both commands of a game,
        ``p1 | p2 << 4`` (``logic.moves.pack_moves``); same tick, half the command bytes.
        """
        gs = game_state
        _require_cuda(gs)
        if packed:
            mv = player1_move
            if (player2_move is not None or not isinstance(mv, torch.Tensor) or mv.dtype != torch.uint8
                    or tuple(mv.shape) != (gs.n,) or not mv.is_contiguous()):
                raise ValueError(f'packed commands must be one contiguous uint8 tensor of shape ({gs.n},)')
        else:
            mv = self._as_moves(gs, player1_move, player2_move)
        if not mv.is_cuda:
            if want_events:
                raise ValueError('events are only produced for device-resident moves')
            return self._update_host(gs, mv, out, packed), None
        cfg, st = self._cfg(gs)
        result = out if out is not None else torch.empty((gs.n,), dtype=torch.uint8, device=gs.device)
        events = None
        if want_events:
            events = torch.empty((gs.n, _abi.MAX_EVENTS_BASE + gs.cfg.n_npc, 2), dtype=torch.int32,
                                 device=gs.device)
        fn = _lib.lib().orx_step_packed if packed else _lib.lib().orx_step
        with _on_device(gs.device):
            rc = fn(C.byref(cfg), C.byref(st), mv.data_ptr(), result.data_ptr(),
                    events.data_ptr() if events is not None else None, gs.n,
                    gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_step_packed' if packed else 'orx_step')
        if events is not None and self.track_order:
            self._advance_order(gs, events)
        return result, events

    def update_bits(self, game_state: BatchedGameState, cmd5: torch.Tensor, out: typing.Optional[torch.Tensor] = None):
        """One tick told and answered in the fewest bytes (``orx_step_bits``): ``cmd5`` is the CUDA uint8 tensor of
        ``_abi.cmd5_bytes(N)`` bytes ``logic.moves.pack_moves5`` produces (5 bits per game), the return value the
        CUDA uint8 tensor of ``_abi.res2_bytes(N)`` bytes ``logic.moves.unpack_results2`` reads (2 bits per game).
        Same tick as ``update``; no NPC slots."""
        gs = game_state
        _require_cuda(gs)
        nb_in, nb_out = _abi.cmd5_bytes(gs.n), _abi.res2_bytes(gs.n)
        if (not isinstance(cmd5, torch.Tensor) or not cmd5.is_cuda or cmd5.dtype != torch.uint8 or tuple(cmd5.shape) != (nb_in,)
                or not cmd5.is_contiguous()):
            raise ValueError(f'cmd5 must be a contiguous CUDA uint8 tensor of {nb_in} bytes')
        cfg, st = self._cfg(gs)
        res2 = out if out is not None else torch.empty((nb_out,), dtype=torch.uint8, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_step_bits(C.byref(cfg), C.byref(st), cmd5.data_ptr(), res2.data_ptr(), gs.n,
                                          gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_step_bits')
        return res2

    def _update_host(self, gs, moves_host, out, packed=False):
        cfg, st = self._cfg(gs)
        if not hasattr(self, '_stage') or self._stage[0].shape[0] != gs.n or self._stage[0].device != gs.device:
            self._stage = (torch.empty((gs.n, 2), dtype=torch.uint8, device=gs.device),
                           torch.empty((gs.n,), dtype=torch.uint8, device=gs.device))
        if out is not None:
            result_host = out
        else:
            if getattr(self, '_host_result', None) is None or self._host_result.shape[0] != gs.n:
                self._host_result = torch.empty((gs.n,), dtype=torch.uint8, pin_memory=True)
            result_host = self._host_result
        name = 'orx_step_host_packed_sync' if packed else 'orx_step_host_sync'     # synchronous: see update()
        with _on_device(gs.device):
            rc = getattr(_lib.lib(), name)(C.byref(cfg), C.byref(st), moves_host.data_ptr(), result_host.data_ptr(),
                                           self._stage[0].data_ptr(), self._stage[1].data_ptr(), gs.n, gs.game_id_base,
                                           _stream_ptr(gs.device))
        _lib.check(rc, name)
        return result_host

    def host_stepper(self, game_state: BatchedGameState, moves_host: torch.Tensor, result_host: torch.Tensor,
                     sync: bool = True, bits: bool = False):
        """Binds one game state and a pair of HOST buffers (pin them: ``pin_memory()``) and returns
        ``step()``: one call = one tick with the commands currently in ``moves_host``; when it returns,
        ``result_host`` holds the tick's UpdateResult codes (``orx_step_host_sync``). This is
        ``update(gs, moves_host, out=result_host)`` + stream synchronise with the per-call argument
        marshalling done once, for host loops that tick every few tens of microseconds.

        ``moves_host`` is uint8[N,2] (p1, p2) or, nibble-packed, uint8[N] with ``p1 | p2 << 4``
        (``orx_step_host_packed_sync``: half the PCIe bytes). ``sync=False`` only enqueues the tick
        (``orx_step_host`` / ``orx_step_host_packed``): the caller synchronises, e.g. with an event,
        before it reads ``result_host`` -- for loops that keep several independent batches in flight.

        ``bits=True``: the bit-packed streams of ``orx_step_host_bits[_sync]`` -- ``moves_host`` holds
        ``_abi.cmd5_bytes(N)`` bytes from ``logic.moves.pack_moves5`` (5 bits per game), ``result_host`` receives
        ``_abi.res2_bytes(N)`` bytes for ``logic.moves.unpack_results2`` (2 bits per game): 0.875 bytes per game
        over PCIe instead of 2 or 3, in one transaction per CTA each way."""
        gs = game_state
        _require_cuda(gs)
        if bits:
            return self._host_stepper_bits(gs, moves_host, result_host, sync)
        packed = moves_host.dim() == 1
        if (moves_host.is_cuda or result_host.is_cuda or moves_host.dtype != torch.uint8 or result_host.dtype != torch.uint8
                or tuple(moves_host.shape) not in ((gs.n, 2), (gs.n,)) or tuple(result_host.shape) != (gs.n,)
                or not moves_host.is_contiguous() or not result_host.is_contiguous()):
            raise ValueError(f'need contiguous CPU uint8 tensors of shape ({gs.n}, 2) or ({gs.n},), and ({gs.n},)')
        cfg, st = self._cfg(gs)
        stage = (torch.empty((gs.n, 2), dtype=torch.uint8, device=gs.device),
                 torch.empty((gs.n,), dtype=torch.uint8, device=gs.device))
        name = 'orx_step_host' + ('_packed' if packed else '') + ('_sync' if sync else '')
        fn = getattr(_lib.lib(), name)
        args = (C.byref(cfg), C.byref(st), C.c_void_p(moves_host.data_ptr()), C.c_void_p(result_host.data_ptr()),
                C.c_void_p(stage[0].data_ptr()), C.c_void_p(stage[1].data_ptr()), C.c_int64(gs.n),
                C.c_uint64(gs.game_id_base), C.c_void_p(_stream_ptr(gs.device)))
        keep = (cfg, st, stage, moves_host, result_host, gs)
        dev_index = gs.device.index or 0

        def step():
            if torch.cuda.current_device() != dev_index:
                torch.cuda.set_device(dev_index)
            rc = fn(*args)
            if rc != 0:
                _lib.check(rc, name)
            return keep[4]
        return step

    def _host_stepper_bits(self, gs, cmd5_host, res2_host, sync):
        nb_in, nb_out = _abi.cmd5_bytes(gs.n), _abi.res2_bytes(gs.n)
        if (cmd5_host.is_cuda or res2_host.is_cuda or cmd5_host.dtype != torch.uint8 or res2_host.dtype != torch.uint8
                or tuple(cmd5_host.shape) != (nb_in,) or tuple(res2_host.shape) != (nb_out,)
                or not cmd5_host.is_contiguous() or not res2_host.is_contiguous()):
            raise ValueError(f'need contiguous CPU uint8 tensors of {nb_in} and {nb_out} bytes')
        cfg, st = self._cfg(gs)
        stage = (torch.empty((nb_in,), dtype=torch.uint8, device=gs.device), torch.empty((nb_out,), dtype=torch.uint8, device=gs.device))
        name = 'orx_step_host_bits' + ('_sync' if sync else '')
        fn = getattr(_lib.lib(), name)
        args = (C.byref(cfg), C.byref(st), C.c_void_p(cmd5_host.data_ptr()), C.c_void_p(res2_host.data_ptr()),
                C.c_void_p(stage[0].data_ptr()), C.c_void_p(stage[1].data_ptr()), C.c_int64(gs.n),
                C.c_uint64(gs.game_id_base), C.c_void_p(_stream_ptr(gs.device)))
        keep = (cfg, st, stage, cmd5_host, res2_host, gs)
        dev_index = gs.device.index or 0

        def step():
            if torch.cuda.current_device() != dev_index:
                torch.cuda.set_device(dev_index)
            rc = fn(*args)
            if rc != 0:
                _lib.check(rc, name)
            return keep[4]
        return step

    def device_stepper(self, game_state: BatchedGameState, moves: torch.Tensor, result: torch.Tensor, *,
                       packed: bool = False, obs: typing.Optional[torch.Tensor] = None, stairs_radius: int = -1,
                       bots: typing.Tuple[int, int] = (0, 0), stream: typing.Optional[torch.cuda.Stream] = None):
        """The device-resident counterpart of ``host_stepper``: binds one game state, a CUDA command buffer
        (uint8[N,2], or uint8[N] with ``packed=True``), a CUDA result buffer and optionally an observation
        buffer (int16[N,2,OBS_LEN]) and returns ``step()``: one call = one tick (``orx_step`` /
        ``orx_step_packed`` / ``orx_step_observe`` / ``orx_step_bots``) on ``stream`` (default: the stream
        current NOW) with the commands that are in ``moves`` at that point of the stream. All argument
        marshalling happens here, once, so an eager loop pays one ctypes call per tick -- what a loop that
        alternates a small policy network with the environment needs. The bound tensors must stay alive and
        in place (the closure keeps references)."""
        gs = game_state
        _require_cuda(gs)
        shape = (gs.n,) if packed else (gs.n, 2)
        if (not moves.is_cuda or moves.dtype != torch.uint8 or tuple(moves.shape) != shape or not moves.is_contiguous()
                or not result.is_cuda or result.dtype != torch.uint8 or tuple(result.shape) != (gs.n,)):
            raise ValueError(f'need contiguous CUDA uint8 tensors of shape {shape} and ({gs.n},)')
        if obs is not None and (not obs.is_cuda or obs.dtype != torch.int16 or tuple(obs.shape) != (gs.n, 2, _abi.OBS_LEN)
                                or not obs.is_contiguous()):
            raise ValueError(f'obs must be a contiguous CUDA int16 tensor of shape ({gs.n}, 2, {_abi.OBS_LEN})')
        if tuple(bots) != (0, 0) and packed:
            raise ValueError('scripted players take the uint8[N,2] command format')
        cfg, st = self._cfg(gs)
        sptr = C.c_void_p((stream or torch.cuda.current_stream(gs.device)).cuda_stream)
        n, gid = C.c_int64(gs.n), C.c_uint64(gs.game_id_base)
        mv, res = C.c_void_p(moves.data_ptr()), C.c_void_p(result.data_ptr())
        lib = _lib.lib()
        if tuple(bots) != (0, 0):
            name, fn = 'orx_step_bots', lib.orx_step_bots
            args = (C.byref(cfg), C.byref(st), mv, C.c_int(int(bots[0])), C.c_int(int(bots[1])), res, None,
                    C.c_void_p(obs.data_ptr()) if obs is not None else None, C.c_int(int(stairs_radius)), n, gid, sptr)
        elif obs is not None:
            name, fn = 'orx_step_observe', lib.orx_step_observe
            args = (C.byref(cfg), C.byref(st), mv, C.c_int(int(bool(packed))), res, C.c_void_p(obs.data_ptr()),
                    C.c_int(int(stairs_radius)), n, gid, sptr)
        else:
            name = 'orx_step_packed' if packed else 'orx_step'
            fn = getattr(lib, name)
            args = (C.byref(cfg), C.byref(st), mv, res, None, n, gid, sptr)
        keep = (cfg, st, moves, result, obs, gs, stream)
        dev_index = gs.device.index or 0

        def step():
            if torch.cuda.current_device() != dev_index:
                torch.cuda.set_device(dev_index)
            rc = fn(*args)
            if rc != 0:
                _lib.check(rc, name)
            return keep[3]
        return step

    def _advance_order(self, gs, events):
        if self.current_update_order is None:
            self.current_update_order = torch.zeros((gs.n,), dtype=torch.int64, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_event_count_add(events.data_ptr(), int(events.shape[1]),
                                                self.current_update_order.data_ptr(), gs.n, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_event_count_add')

    def get_incr_upd_order(self):
        """Per-game count of GameStateUpdates emitted so far (updater.py:71-74)."""
        return self.current_update_order

    # -- fused paths -------------------------------------------------------------------------------
    def bot_moves(self, game_state: BatchedGameState, bot_p1: int, bot_p2: int,
                  out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """Commands of the scripted bots for the current tick -> uint8[N,2]."""
        gs = game_state
        _require_cuda(gs)
        cfg, st = self._cfg(gs)
        moves = out if out is not None else torch.full((gs.n, 2), 5, dtype=torch.uint8, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_bot_moves(C.byref(cfg), C.byref(st), int(bot_p1), int(bot_p2),
                                          moves.data_ptr(), gs.n, gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_bot_moves')
        return moves

    def rollout(self, game_state: BatchedGameState, bot_p1: int, bot_p2: int, n_ticks: int,
                stats: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """``n_ticks`` fused ticks with both bots on device (the loop of server/main.py:110-113).
        Returns the int64[STAT_COUNT] device tensor of accumulated counters."""
        gs = game_state
        _require_cuda(gs)
        cfg, st = self._cfg(gs)
        if stats is None:
            stats = torch.zeros((_abi.STAT_COUNT,), dtype=torch.int64, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_rollout(C.byref(cfg), C.byref(st), int(bot_p1), int(bot_p2), int(n_ticks),
                                        stats.data_ptr(), gs.n, gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_rollout')
        return stats

    def replay(self, game_state: BatchedGameState, moves: torch.Tensor,
               out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """Plays ``T`` ticks of queued commands in one launch: ``moves`` uint8[T,N,2] on the device,
        returns uint8[T,N] results. Same outcome as T ``update`` calls (``orx_replay``)."""
        gs = game_state
        _require_cuda(gs)
        if moves.dim() != 3 or tuple(moves.shape[1:]) != (gs.n, 2) or moves.dtype != torch.uint8 or not moves.is_cuda:
            raise ValueError(f'moves must be a CUDA uint8 tensor of shape (T, {gs.n}, 2)')
        moves = moves.contiguous()
        cfg, st = self._cfg(gs)
        results = out if out is not None else torch.empty((moves.shape[0], gs.n), dtype=torch.uint8, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_replay(C.byref(cfg), C.byref(st), moves.data_ptr(), results.data_ptr(),
                                       int(moves.shape[0]), gs.n, gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_replay')
        return results

    def observe(self, game_state: BatchedGameState, stairs_radius: int = -1,
                out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-player observation int16[N,2,OBS_LEN] (GameState.view_for, state.py:53-58, plus the
        README's "ladder visible when near" as an optional Chebyshev radius; -1 = always)."""
        gs = game_state
        _require_cuda(gs)
        cfg, st = self._cfg(gs)
        obs = out if out is not None else torch.empty((gs.n, 2, _abi.OBS_LEN), dtype=torch.int16, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_observe(C.byref(cfg), C.byref(st), obs.data_ptr(), int(stairs_radius),
                                        gs.n, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_observe')
        return obs

    def observe_npc(self, game_state: BatchedGameState, out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """The entities besides the players that each player sees (GameState.view_for, state.py:53-58, keeps the
        viewer's depth): int16[N,2,n_npc,4] = on_my_depth, x, y, health per NPC slot (0, -1, -1, 0 when the slot is empty
        or on another depth). ``orx_observe_npc``; a state without NPC slots returns an empty tensor."""
        gs = game_state
        _require_cuda(gs)
        e = gs.cfg.n_npc
        obs = out if out is not None else torch.empty((gs.n, 2, e, 4), dtype=torch.int16, device=gs.device)
        if e == 0:
            return obs
        if tuple(obs.shape) != (gs.n, 2, e, 4) or obs.dtype != torch.int16 or not obs.is_cuda or not obs.is_contiguous():
            raise ValueError(f'out must be a contiguous CUDA int16 tensor of shape ({gs.n}, 2, {e}, 4)')
        cfg, st = self._cfg(gs)
        with _on_device(gs.device):
            rc = _lib.lib().orx_observe_npc(C.byref(cfg), C.byref(st), obs.data_ptr(), gs.n, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_observe_npc')
        return obs

    def update_observe(self, game_state: BatchedGameState, moves: torch.Tensor, *, packed: bool = False,
                       stairs_radius: int = -1, out: typing.Optional[torch.Tensor] = None,
                       obs_out: typing.Optional[torch.Tensor] = None):
        """One tick plus the observations of the resulting state in one pass over the planes
        (``orx_step_observe``): ``update`` followed by ``observe``, as a self-play loop needs them.
        ``moves``: CUDA uint8[N,2], or uint8[N] of ``p1 | p2 << 4`` with ``packed=True``.
        Returns ``(result uint8[N], obs int16[N,2,OBS_LEN])``."""
        gs = game_state
        _require_cuda(gs)
        shape = (gs.n,) if packed else (gs.n, 2)
        if (not isinstance(moves, torch.Tensor) or not moves.is_cuda or moves.dtype != torch.uint8
                or tuple(moves.shape) != shape or not moves.is_contiguous()):
            raise ValueError(f'moves must be a contiguous CUDA uint8 tensor of shape {shape}')
        cfg, st = self._cfg(gs)
        result = out if out is not None else torch.empty((gs.n,), dtype=torch.uint8, device=gs.device)
        obs = obs_out if obs_out is not None else torch.empty((gs.n, 2, _abi.OBS_LEN), dtype=torch.int16, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_step_observe(C.byref(cfg), C.byref(st), moves.data_ptr(), int(bool(packed)),
                                             result.data_ptr(), obs.data_ptr(), int(stairs_radius), gs.n,
                                             gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_step_observe')
        return result, obs

    def update_with_bots(self, game_state: BatchedGameState, moves: torch.Tensor, bot_p1: int = 0, bot_p2: int = 0, *,
                         want_events: bool = False, observe: bool = False, stairs_radius: int = -1,
                         out: typing.Optional[torch.Tensor] = None, obs_out: typing.Optional[torch.Tensor] = None):
        """One tick in which scripted players move themselves (``orx_step_bots``): ``bot_p1`` / ``bot_p2`` are
        ``_abi.BOT_NONE`` = 0 (the player's command comes from ``moves``), ``_abi.BOT_RANDOM`` or ``_abi.BOT_STAIRCASE``
        (``optimax_rogue_bots/randombot.py``, ``staircasebot.py``; computed inside the tick kernel, exactly the
        command ``bot_moves`` would return). ``moves``: CUDA uint8[N,2]. For training a policy against a scripted
        opponent without a second launch. Returns ``(result, events or None, obs or None)``."""
        gs = game_state
        _require_cuda(gs)
        if (not isinstance(moves, torch.Tensor) or not moves.is_cuda or moves.dtype != torch.uint8
                or tuple(moves.shape) != (gs.n, 2) or not moves.is_contiguous()):
            raise ValueError(f'moves must be a contiguous CUDA uint8 tensor of shape ({gs.n}, 2)')
        cfg, st = self._cfg(gs)
        result = out if out is not None else torch.empty((gs.n,), dtype=torch.uint8, device=gs.device)
        events = obs = None
        if want_events:
            events = torch.empty((gs.n, _abi.MAX_EVENTS_BASE + gs.cfg.n_npc, 2), dtype=torch.int32, device=gs.device)
        if observe or obs_out is not None:
            obs = obs_out if obs_out is not None else torch.empty((gs.n, 2, _abi.OBS_LEN), dtype=torch.int16, device=gs.device)
        with _on_device(gs.device):
            rc = _lib.lib().orx_step_bots(C.byref(cfg), C.byref(st), moves.data_ptr(), int(bot_p1), int(bot_p2),
                                          result.data_ptr(), events.data_ptr() if events is not None else None,
                                          obs.data_ptr() if obs is not None else None, int(stairs_radius), gs.n,
                                          gs.game_id_base, _stream_ptr(gs.device))
        _lib.check(rc, 'orx_step_bots')
        if events is not None and self.track_order:
            self._advance_order(gs, events)
        return result, events, obs

    def reset(self, game_state: BatchedGameState, mask=None, bump_episode: bool = True):
        reset_games(game_state, mask, bump_episode)
