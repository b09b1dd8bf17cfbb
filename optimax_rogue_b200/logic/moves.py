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


def pack_moves5(player1_move, player2_move):
    """The bit-packed command stream of ``orx_step_bits`` / ``orx_step_host_bits`` (include/orx.h): five bits per
    game, value ``(p1 - 1) * 5 + (p2 - 1)``, game i in bits [5i, 5i+5) of a little-endian bit stream. Takes numpy
    arrays or CPU torch tensors of Move codes (anything outside 1..5 is packed as Stay, which is how the tick plays
    it) and returns a numpy uint8 array of ``(5 n + 7) // 8`` bytes."""
    import numpy as np
    p1 = np.asarray(player1_move).astype(np.int64)
    p2 = np.asarray(player2_move).astype(np.int64)
    p1 = np.where((p1 < 1) | (p1 > 5), 5, p1)
    p2 = np.where((p2 < 1) | (p2 > 5), 5, p2)
    v = ((p1 - 1) * 5 + (p2 - 1)).astype(np.uint8)
    bits = ((v[:, None] >> np.arange(5, dtype=np.uint8)) & 1).astype(np.uint8).reshape(-1)
    return np.packbits(bits, bitorder='little')


def unpack_moves5(cmd5, n):
    """Inverse of ``pack_moves5``: ``(p1, p2)`` numpy uint8 arrays of n Move codes (25..31 decode as Stay, Stay)."""
    import numpy as np
    bits = np.unpackbits(np.asarray(cmd5, dtype=np.uint8), bitorder='little')[:5 * n].reshape(n, 5)
    v = (bits * (1 << np.arange(5))).sum(axis=1)
    p1, p2 = v // 5 + 1, v % 5 + 1
    bad = v >= 25
    return np.where(bad, 5, p1).astype(np.uint8), np.where(bad, 5, p2).astype(np.uint8)


def unpack_results2(res2, n):
    """UpdateResult codes (1..4) of n games from the 2-bit result stream of ``orx_step_bits``
    (game i in bits [2i, 2i+2), value ``UpdateResult - 1``). numpy in, numpy uint8 out."""
    import numpy as np
    b = np.asarray(res2, dtype=np.uint8)
    out = np.empty((b.shape[0], 4), np.uint8)
    for k in range(4):
        out[:, k] = (b >> (2 * k)) & 3
    return out.reshape(-1)[:n] + 1
