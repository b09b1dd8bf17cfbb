"""Command set; same integer codes as the reference (optimax_rogue/logic/moves.py:6-12)."""
import enum


class Move(enum.IntEnum):
    """A particular action that an entity can take"""
    Up = 1
    Right = 2
    Down = 3
    Left = 4
    Stay = 5


def pack_moves(player1_move, player2_move):
    """Both players' Move codes of every game in one byte, ``p1 | p2 << 4`` -- the nibble-packed
    command format of ``orx_step_packed`` / ``orx_step_host_packed``. Works on torch tensors and
    numpy arrays of uint8; codes must be below 16 (anything outside 1..5 is played as Stay)."""
    return (player1_move & 15) | ((player2_move & 15) << 4)


def unpack_moves(cmds):
    """Inverse of ``pack_moves``: ``(p1, p2)``."""
    return cmds & 15, cmds >> 4
