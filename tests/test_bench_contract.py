"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the agreed keys and
runs without a GPU; the GPU arm refuses loudly without CUDA."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                          timeout=300, env=env)


def test_reference_arm_prints_one_json_line():
    """With a reference tree on the machine (/root/reference or the shipped oracle/_ref) the arm times the live
    Python updater loop (kind "reference"), else the C port (kind "port")."""
    from oracle import ref_harness
    out = run(['--impl', 'reference', '--steps', '3', '--warmup', '1', '--games', '8192'])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'game-ticks/s' and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == ('reference' if ref_harness.reference_available() else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['vs_baseline'] is None and d['value'] > 0
    games = d['sample_games_per_step']
    assert d['steps'] == 3 and abs(d['value'] - games * 3 / (d['ms_per_step'] * 3e-3)) / d['value'] < 1e-6


def test_both_arms_print_the_same_config_object():
    """The driver compares the two arms' ``config``: it is built from the flags alone, by one function."""
    import argparse
    import bench
    args = argparse.Namespace(games=1 << 20)
    out = run(['--impl', 'reference', '--steps', '1', '--warmup', '1', '--games', str(1 << 20)], {'WORLD_SIZE': '4', 'RANK': '0', 'LOCAL_RANK': '0'})
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d['config'] == bench.workload_config(args, 4) and d['n_gpus'] == 4
    assert d['metric'] == json.load(open(os.path.join(ROOT, 'BASELINE.json')))['metric']


def test_reference_arm_other_ranks_exit_quietly():
    out = run(['--impl', 'reference', '--steps', '2', '--warmup', '1', '--games', '4096'],
              {'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    out = run(['--steps', '2', '--warmup', '1', '--no-cpu-baseline'])
    assert out.returncode != 0 and 'no CPU fallback' in out.stderr
