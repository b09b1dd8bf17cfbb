"""optimax_rogue_b200 -- B200-native batched simulator of Optimax Rogue's turn dynamics.

``import optimax_rogue_b200`` is torch-free (ABI + config only) so the CPU test-suite can check
the library boundary; the simulator itself lives in ``optimax_rogue_b200.logic`` / ``.game`` /
``.bots`` and needs CUDA + the in-tree ``liborx.so``.
"""
from . import _abi
from .config import SimConfig

__all__ = ['_abi', 'SimConfig']
