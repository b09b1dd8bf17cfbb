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
