"""RandomBot: uniform over the five Moves (optimax_rogue_bots/randombot.py:14-21), one Philox
word per game per tick (draw schedule: TICK main block, word = player index)."""
from .. import _abi
from .bot import Bot


class RandomBot(Bot):
    kind = _abi.BOT_RANDOM
