"""The C-ABI library loads (no GPU needed) and exports every symbol include/orx.h declares; the
ctypes mirror has the same struct layouts as the header."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from optimax_rogue_b200 import _abi, _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'orx.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(orx_[a-z0-9_]+)\s*\(', src)))


def test_header_and_ctypes_mirror_list_the_same_functions():
    assert declared_functions() == sorted(_abi.PROTOTYPES)


def test_library_builds_and_exports_every_symbol():
    build.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f'{name} missing from liborx.so'
    _abi.bind(lib)
    assert lib.orx_abi_version() == _abi.ABI_VERSION
    assert lib.orx_strerror(0) == b'ok'
    assert b'bad argument' in lib.orx_strerror(_abi.ERR_BAD_ARG)


def test_pure_host_queries():
    lib = _lib.lib()
    from optimax_rogue_b200 import SimConfig
    c = SimConfig().to_c()
    assert lib.orx_state_bytes(C.byref(c)) == 29
    assert lib.orx_max_events(C.byref(c)) == 4
    c = SimConfig(n_npc=3).to_c()
    assert lib.orx_state_bytes(C.byref(c)) == 29 + 24
    assert lib.orx_max_events(C.byref(c)) == 7


def test_bad_arguments_rejected_before_any_launch():
    lib = _lib.lib()
    from optimax_rogue_b200 import SimConfig
    c = SimConfig().to_c()
    st = _abi.OrxState()
    assert lib.orx_reset(None, C.byref(st), None, 0, 1, 0, None) == _abi.ERR_BAD_ARG
    assert lib.orx_reset(C.byref(c), C.byref(st), None, 0, 1, 0, None) == _abi.ERR_BAD_ARG   # null planes
    assert lib.orx_step(C.byref(c), C.byref(st), None, None, None, -1, 0, None) == _abi.ERR_BAD_ARG


def test_struct_layout_matches_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "orx.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(OrxConfig), sizeof(OrxState), sizeof(OrxEvent));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(OrxConfig, start_depth), offsetof(OrxConfig, hp),
         offsetof(OrxConfig, seed), offsetof(OrxConfig, fixed_tiles), offsetof(OrxConfig, fixed_n_ground),
         offsetof(OrxConfig, fixed_stairs));
  printf("%zu %zu\n", offsetof(OrxState, status), offsetof(OrxState, npc_depth));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, 'l.c')
        open(src, 'w').write(prog)
        exe = os.path.join(d, 'l')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', exe, src])
        out = subprocess.check_output([exe], text=True).split()
    vals = [int(x) for x in out]
    cfg, st = _abi.OrxConfig, _abi.OrxState
    assert vals[:3] == [C.sizeof(cfg), C.sizeof(st), C.sizeof(_abi.OrxEvent)]
    assert vals[3:9] == [cfg.start_depth.offset, cfg.hp.offset, cfg.seed.offset, cfg.fixed_tiles.offset,
                         cfg.fixed_n_ground.offset, cfg.fixed_stairs.offset]
    assert vals[9:] == [st.status.offset, st.npc_depth.offset]


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import reset_games
    gs = BatchedGameState(SimConfig(), 4, 'cpu')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        reset_games(gs)


def test_graft_entry_build_runs():
    """The driver's "does it build" hook: compiles (or finds current) liborx + the oracle and checks the ABI version."""
    import __graft_entry__ as g
    g.build()
