"""Ruleset R1 -- the README-only rules (mana, heal, cooldowns, half / negated damage, enemies,
items, separation damage; readme.md:44-48,69-74), specified in docs/RULESET_R1.md.

PARITY UNPINNED: the reference has no code for these rules; the kernels are bit-exact against
oracle/orx_r1_oracle.c (the written spec), not against the reference. The reference-exact ruleset
R0 lives in ``optimax_rogue_b200.logic``; this module does not touch it.
"""
import ctypes as C
import typing

import torch

from . import _abi, _lib


class R1GameState:
    """N games of ruleset R1: entity planes [N,16] (lanes 0-1 players, 2-9 enemies, 10-13 items)
    plus per-player and per-game planes (include/orx.h:OrxR1State)."""

    def __init__(self, n: int, *, width=60, height=10, max_ticks=0, auto_reset=False, wall_density=26,
                 seed=0, device='cuda', game_id_base=0, path_flags=0, overlap_ticks=False):
        self.n, self.device, self.game_id_base = int(n), torch.device(device), int(game_id_base)
        if self.device.type != 'cuda':
            raise RuntimeError('R1GameState must live on a CUDA device: there is no CPU fallback')
        self.cfg = _abi.OrxR1Config()
        self.cfg.struct_size = C.sizeof(_abi.OrxR1Config)
        self.cfg.width, self.cfg.height, self.cfg.max_ticks = width, height, int(max_ticks or 0)
        self.cfg.auto_reset, self.cfg.wall_density = int(auto_reset), wall_density
        self.cfg.seed = seed & 0xFFFFFFFFFFFFFFFF
        # _abi.R1_PATH_*: pins a kernel path (tests, A/B runs); overlap_ticks = throughput mode (ORX_R1_PATH_BLOCK_FLAGS):
        # ticks enqueued back to back overlap block by block -- slower when other kernels run between two ticks
        self.cfg.path_flags = int(path_flags) | (_abi.R1_PATH_BLOCK_FLAGS if overlap_ticks else 0)
        for name, dt, shape in _abi.R1_PLANES:
            setattr(self, name, torch.zeros((self.n,) + shape, dtype=getattr(torch, dt), device=self.device))
        self.status.fill_(1)
        self._st = _abi.OrxR1State()
        for name, _, _ in _abi.R1_PLANES:
            setattr(self._st, name, getattr(self, name).data_ptr())
        # hand-over words of orx_r1_step (OrxR1State.sched): consecutive ticks overlap block by block
        self.sched = torch.zeros((_abi.r1_sched_words(self.n),), dtype=torch.int32, device=self.device)
        self._st.sched, self._st.sched_words = self.sched.data_ptr(), int(self.sched.numel())

    @property
    def alg_bytes_per_game_tick(self) -> int:
        """State in + state out + 2 command bytes + 1 result byte (ruleset R1)."""
        return 2 * _abi.R1_STATE_BYTES + 3

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mask: typing.Optional[torch.Tensor] = None, bump_episode: bool = False):
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_reset(C.byref(self.cfg), C.byref(self._st), mptr, int(bump_episode), self.n,
                                         self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_reset')
        return self

    def update(self, moves: torch.Tensor, out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """One tick; moves uint8[N,2] with codes 1..6 (6 = Heal). Returns uint8[N] UpdateResult codes."""
        if tuple(moves.shape) != (self.n, 2) or moves.dtype != torch.uint8 or not moves.is_cuda:
            raise ValueError(f'moves must be a CUDA uint8 tensor of shape ({self.n}, 2)')
        result = out if out is not None else torch.empty((self.n,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_step(C.byref(self.cfg), C.byref(self._st), moves.contiguous().data_ptr(),
                                        result.data_ptr(), self.n, self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_step')
        return result

    def update_events(self, moves: torch.Tensor, out: typing.Optional[torch.Tensor] = None,
                      events: typing.Optional[torch.Tensor] = None, max_events: int = _abi.R1_MAX_EVENTS):
        """One tick plus its replication log (orx_r1_step_events): returns (result uint8[N], records int32[N, max_events, 2]).
        A game's list ends at the first record of kind 0; slots behind it keep whatever the buffer held.
        ``logic.updates.decode_r1_events`` turns one game's records into GameStateUpdate objects."""
        if tuple(moves.shape) != (self.n, 2) or moves.dtype != torch.uint8 or not moves.is_cuda:
            raise ValueError(f'moves must be a CUDA uint8 tensor of shape ({self.n}, 2)')
        result = out if out is not None else torch.empty((self.n,), dtype=torch.uint8, device=self.device)
        if events is None:
            events = torch.zeros((self.n, int(max_events), 2), dtype=torch.int32, device=self.device)
        if events.dtype != torch.int32 or events.dim() != 3 or events.shape[0] != self.n or events.shape[2] != 2 or not events.is_contiguous():
            raise ValueError(f'events must be a contiguous CUDA int32 tensor of shape ({self.n}, max_events, 2)')
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_step_events(C.byref(self.cfg), C.byref(self._st), moves.contiguous().data_ptr(),
                                               result.data_ptr(), events.data_ptr(), int(events.shape[1]), self.n,
                                               self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_step_events')
        return result, events

    def bot_moves(self, bot1: int, bot2: int, out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """Commands of scripted players (BOT_RANDOM over the six R1 commands, BOT_STAIRCASE; BOT_NONE leaves the byte)."""
        moves = out if out is not None else torch.full((self.n, 2), 5, dtype=torch.uint8, device=self.device)   # Move.Stay
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_bot_moves(C.byref(self.cfg), C.byref(self._st), int(bot1), int(bot2), moves.data_ptr(),
                                             self.n, self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_bot_moves')
        return moves

    def replay(self, moves: torch.Tensor, out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """T ticks with queued commands (uint8[T,N,2]) in one launch, state in registers in between; uint8[T,N] results."""
        if moves.dim() != 3 or tuple(moves.shape[1:]) != (self.n, 2) or moves.dtype != torch.uint8 or not moves.is_cuda:
            raise ValueError(f'moves must be a CUDA uint8 tensor of shape (T, {self.n}, 2)')
        t = int(moves.shape[0])
        results = out if out is not None else torch.empty((t, self.n), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_replay(C.byref(self.cfg), C.byref(self._st), moves.contiguous().data_ptr(),
                                          results.data_ptr(), t, self.n, self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_replay')
        return results

    def update_host(self, host_moves: torch.Tensor, host_result: torch.Tensor) -> torch.Tensor:
        """One tick told and answered through PINNED host tensors (uint8[N,2] in, uint8[N] out); returns after the
        stream has been synchronised, i.e. ``host_result`` is readable."""
        for t, shape in ((host_moves, (self.n, 2)), (host_result, (self.n,))):
            if t.is_cuda or not t.is_pinned() or t.dtype != torch.uint8 or tuple(t.shape) != shape or not t.is_contiguous():
                raise ValueError('host buffers must be pinned contiguous uint8 CPU tensors of shape (N, 2) / (N,)')
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_step_host_sync(C.byref(self.cfg), C.byref(self._st), host_moves.data_ptr(),
                                                  host_result.data_ptr(), self.n, self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_step_host_sync')
        return host_result

    def rollout(self, n_ticks: int, stats: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """n_ticks fused ticks, both players uniform over the six commands."""
        if stats is None:
            stats = torch.zeros((_abi.STAT_COUNT,), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_rollout(C.byref(self.cfg), C.byref(self._st), int(n_ticks), stats.data_ptr(),
                                           self.n, self.game_id_base, self._stream())
        _lib.check(rc, 'orx_r1_rollout')
        return stats

    def observe(self, stairs_radius: int = 4, out: typing.Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-player observation int16[N,2,R1_OBS_LEN] (layout: include/orx.h:orx_r1_observe); the
        staircase is reported only within ``stairs_radius`` (Chebyshev; < 0 = always)."""
        obs = out if out is not None else torch.empty((self.n, 2, _abi.R1_OBS_LEN), dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().orx_r1_observe(C.byref(self.cfg), C.byref(self._st), obs.data_ptr(), int(stairs_radius),
                                           self.n, self._stream())
        _lib.check(rc, 'orx_r1_observe')
        return obs

    def planes_cpu(self):
        return {name: getattr(self, name).cpu().numpy() for name, _, _ in _abi.R1_PLANES}
