"""Loader for ``liborx.so`` (the CUDA kernels behind the C ABI of ``include/orx.h``).

There is no CPU fallback: if the library is missing or no CUDA device is usable every
product call raises. Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``python -m optimax_rogue_b200.build``.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ORX_LIB') or os.path.join(_HERE, 'liborx.so')   # ORX_LIB: tuning builds only


class OrxError(RuntimeError):
    """A negative return code from the C ABI."""

    def __init__(self, code, where):
        self.code = code
        msg = _lib.orx_strerror(code).decode() if _lib is not None else '?'
        super().__init__(f'{where} failed with code {code}: {msg}')


_lib = None


def lib():
    """The bound CDLL; raises RuntimeError (loudly) when the extension is not built."""
    global _lib
    if _lib is None:
        if 'ORX_LIB' not in os.environ:
            from . import build as _build
            if not _build.is_current():
                # never run kernels that do not match the sources in the tree
                if _build.have_nvcc():
                    _build.build()
                elif os.path.exists(LIB_PATH):
                    raise RuntimeError(f'{LIB_PATH} is stale (sources changed since it was built) and nvcc is '
                                       'not available to rebuild it. There is no CPU fallback.')
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: the CUDA extension has not been built. '
                'Run `python -m optimax_rogue_b200.build` (needs nvcc). There is no CPU fallback.')
        handle = _abi.bind(C.CDLL(LIB_PATH))
        ver = handle.orx_abi_version()
        if ver != _abi.ABI_VERSION:
            raise RuntimeError(f'liborx.so ABI version {ver} != expected {_abi.ABI_VERSION}; rebuild')
        _lib = handle
    return _lib


def check(code, where):
    if code != _abi.OK:
        raise OrxError(code, where)
