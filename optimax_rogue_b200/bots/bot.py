"""Base class for batched bots (optimax_rogue_bots/bot.py:6-38): one bot object controls the
same player in every game of the batch; ``move`` returns uint8[N] Move codes on the device."""
import ctypes as C

import torch

from .. import _abi, _lib


class Bot:
    kind = _abi.BOT_NONE

    def __init__(self, entity_iden: int):
        if entity_iden not in (1, 2):
            raise ValueError('batched bots control player entity 1 or 2')
        self.entity_iden = entity_iden

    def think(self, max_time: float) -> None:
        pass

    def started(self, game_state) -> None:
        pass

    def move(self, game_state, out: torch.Tensor = None) -> torch.Tensor:
        """Move codes for this bot's player in every game. ``out`` may be a uint8[N,2] command
        buffer: only this player's column is written."""
        gs = game_state
        buf = out if out is not None else torch.full((gs.n, 2), 5, dtype=torch.uint8, device=gs.device)
        cfg, st = gs.c_config(), gs.c_struct()
        kinds = (self.kind, _abi.BOT_NONE) if self.entity_iden == 1 else (_abi.BOT_NONE, self.kind)
        with torch.cuda.device(gs.device):
            rc = _lib.lib().orx_bot_moves(C.byref(cfg), C.byref(st), kinds[0], kinds[1], buf.data_ptr(),
                                          gs.n, gs.game_id_base,
                                          torch.cuda.current_stream(gs.device).cuda_stream)
        _lib.check(rc, 'orx_bot_moves')
        return buf[:, self.entity_iden - 1]

    def on_move(self, game_state, move) -> None:
        pass

    def finished(self, game_state, result) -> None:
        pass
