"""StaircaseBot: walks towards its level's staircase, horizontal first when |dx| > |dy|
(optimax_rogue_bots/staircasebot.py:9-20)."""
from .. import _abi
from .bot import Bot


class StaircaseBot(Bot):
    kind = _abi.BOT_STAIRCASE
