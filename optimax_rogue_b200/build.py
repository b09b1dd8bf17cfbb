"""Builds ``liborx.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(_HERE, 'csrc', 'orx_api.cu'), os.path.join(_HERE, 'csrc', 'orx_r1.cu')]
HEADERS = [os.path.join(_HERE, 'csrc', 'orx_rng.cuh'), os.path.join(_HERE, 'csrc', 'orx_rules.cuh'),
           os.path.join(_HERE, 'csrc', 'orx_pipe.cuh'), os.path.join(_HERE, 'csrc', 'orx_r1t.cuh'),
           os.path.join(_HERE, '..', 'include', 'orx.h')]
OUT = os.path.join(_HERE, 'liborx.so')

NVCC_FLAGS = ['-std=c++20', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-shared', '-Xcompiler', '-fPIC', '--cudart', 'shared']


HASH_FILE = OUT + '.srchash'


def source_hash() -> str:
    h = hashlib.sha256()
    for p in sorted(SOURCES + HEADERS):
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    """True when liborx.so was built from exactly the sources in the tree."""
    try:
        with open(HASH_FILE) as f:
            return os.path.exists(OUT) and f.read().strip() == source_hash()
    except OSError:
        return False


def have_nvcc() -> bool:
    return shutil.which(os.environ.get('NVCC', 'nvcc')) is not None


def build(force=False, verbose=False):
    if not force and is_current():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    with open(HASH_FILE, 'w') as f:
        f.write(source_hash())
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
