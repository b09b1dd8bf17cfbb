"""Batched scripted bots (the reference's optimax_rogue_bots package)."""
from .bot import Bot
from .randombot import RandomBot
from .staircasebot import StaircaseBot

__all__ = ['Bot', 'RandomBot', 'StaircaseBot']
