"""Command set; same integer codes as the reference (optimax_rogue/logic/moves.py:6-12)."""
import enum


class Move(enum.IntEnum):
    """A particular action that an entity can take"""
    Up = 1
    Right = 2
    Down = 3
    Left = 4
    Stay = 5
