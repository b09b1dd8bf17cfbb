"""optimax_rogue_b200 -- B200-native batched simulator of Optimax Rogue's turn dynamics."""
from . import _abi
from .config import SimConfig

__all__ = ['_abi', 'SimConfig']
