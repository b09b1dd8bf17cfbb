"""The boundary is a C ABI: include/orx.h must be valid C (not only C++), and a plain C host program
(examples/c_abi_demo.c: cudaMalloc'ed planes, no Python, no PyTorch) must link against liborx.so and -- on a GPU --
play exactly the games the Python host plays (the loop of optimax_rogue/server/main.py:110-113 with StaircaseBot vs
RandomBot, optimax_rogue_bots/)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')
LIBDIR = os.path.join(ROOT, 'optimax_rogue_b200')

needs_gcc = pytest.mark.skipif(shutil.which('gcc') is None or not os.path.exists(os.path.join(CUDA, 'include', 'cuda_runtime.h')),
                               reason='needs gcc and the CUDA runtime headers')


def build_demo(tmp_path):
    from optimax_rogue_b200 import build
    build.build()
    exe = str(tmp_path / 'c_abi_demo')
    subprocess.run(['gcc', '-O2', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(CUDA, 'include'),
                    os.path.join(ROOT, 'examples', 'c_abi_demo.c'), '-o', exe, '-L', LIBDIR, '-l:liborx.so',
                    '-L', os.path.join(CUDA, 'lib64'), '-lcudart', '-Wl,-rpath,' + LIBDIR], check=True)
    return exe


@needs_gcc
def test_header_is_valid_c99(tmp_path):
    src = tmp_path / 'hdr.c'
    src.write_text('#include "orx.h"\nint main(void) { return (int)sizeof(OrxConfig) + (int)sizeof(OrxState) + (int)sizeof(OrxR1State) == 0; }\n')
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'), str(src)], check=True)


@needs_gcc
def test_c_host_program_links_against_the_library(tmp_path):
    exe = build_demo(tmp_path)
    assert os.access(exe, os.X_OK)


@pytest.mark.gpu
@needs_gcc
def test_c_host_program_plays_the_same_games_as_the_python_host(tmp_path):
    import torch
    from optimax_rogue_b200 import SimConfig
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
    n, ticks, seed, max_ticks = 5000, 60, 11, 37
    out = subprocess.run([build_demo(tmp_path), str(n), str(ticks), str(seed), str(max_ticks)], check=True, capture_output=True, text=True).stdout.splitlines()
    step = dict(kv.split('=') for kv in out[1].split()[1:])
    rollout = [int(x) for x in out[2].split()[1:]]

    cfg = SimConfig(max_ticks=max_ticks, seed=seed, auto_reset=True)
    gs = BatchedGameState(cfg, n, 'cuda')
    reset_games(gs)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, max_ticks, auto_reset=True)
    counts, checksum = np.zeros(5, np.int64), 0
    for _ in range(ticks):
        res, _ = upd.update(gs, upd.bot_moves(gs, 2, 1))
        r = res.cpu().numpy()
        counts += np.bincount(r, minlength=5)[:5]
        for v in r.tolist():
            checksum = (checksum * 1099511628211 + v) & 0xFFFFFFFFFFFFFFFF
    stats = upd.rollout(gs, 2, 1, ticks)
    torch.cuda.synchronize()
    assert [int(step['in_progress']), int(step['p1_wins']), int(step['p2_wins']), int(step['ties'])] == counts[1:5].tolist()
    assert int(step['checksum']) == checksum
    assert rollout == [int(x) for x in stats.cpu().tolist()]
    assert counts[2] + counts[3] + counts[4] > 0          # games did end (StaircaseBot descends, max_ticks, deaths)
