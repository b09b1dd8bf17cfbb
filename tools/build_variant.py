"""Tuning aid: builds liborx variants with extra -D flags next to the shipped library
(optimax_rogue_b200/liborx_<tag>.so; git-ignored) for A/B runs with ORX_LIB=<path>.
    python tools/build_variant.py tag -DORX_FOO=1 ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimax_rogue_b200 import build as B

tag, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(B.OUT), f'liborx_{tag}.so')
cmd = ['nvcc'] + B.NVCC_FLAGS + extra + ['-o', out] + B.SOURCES
subprocess.run(cmd, check=True)
print(out)
