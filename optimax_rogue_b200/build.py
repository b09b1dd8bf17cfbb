"""Builds ``liborx.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(_HERE, 'csrc', 'orx_api.cu'), os.path.join(_HERE, 'csrc', 'orx_r1.cu')]
HEADERS = [os.path.join(_HERE, 'csrc', 'orx_rng.cuh'), os.path.join(_HERE, 'csrc', 'orx_rules.cuh'),
           os.path.join(_HERE, 'csrc', 'orx_pipe.cuh'),
           os.path.join(_HERE, '..', 'include', 'orx.h')]
OUT = os.path.join(_HERE, 'liborx.so')

NVCC_FLAGS = ['-std=c++20', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-shared', '-Xcompiler', '-fPIC', '--cudart', 'shared']


def build(force=False, verbose=False):
    newest = max(os.path.getmtime(p) for p in SOURCES + HEADERS)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
