"""ctypes binding of libhevc_b200.so (include/hevc_b200.h).

There is no fallback: if the CUDA library is missing or no device is present, loading / context creation
raises ``BackendUnavailable``."""
from __future__ import annotations

import ctypes as C
import re
import subprocess
from pathlib import Path
from typing import List

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / 'libhevc_b200.so'
HEADER = PKG.parent / 'include' / 'hevc_b200.h'


class BackendUnavailable(RuntimeError):
    pass


class HbError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f'hevc_b200 error {code}: {text}')
        self.code = code


_lib = None


def build(verbose: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(['make', '-j8', '-C', str(PKG / 'csrc')], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('building libhevc_b200.so failed:\n' + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout)
    return LIB_PATH


def declared_symbols() -> List[str]:
    """Every function name declared in include/hevc_b200.h."""
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(hb_[a-z0-9_]+)\s*\(', text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise BackendUnavailable(f'{LIB_PATH} not built (run `python -c "import __graft_entry__ as g; g.build()"`)')
        L = C.CDLL(str(LIB_PATH))
        L.hb_last_error.restype = C.c_char_p
        L.hb_last_error.argtypes = [C.c_void_p]
        L.hb_stream.restype = C.c_uint64
        L.hb_stream.argtypes = [C.c_void_p]
        L.hb_launch_count.restype = C.c_uint64
        L.hb_launch_count.argtypes = [C.c_void_p]
        L.hb_destroy.restype = None
        L.hb_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class Context:
    """One device + one stream (``hb_ctx``).  Thin: every method is one C call."""

    def __init__(self, device: int = 0):
        L = lib()
        h = C.c_void_p()
        rc = L.hb_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise BackendUnavailable(f'hb_create(device={device}) failed with {rc}: no usable CUDA device')
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, 'h', None):
            lib().hb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def call(self, name: str, *args):
        rc = getattr(lib(), name)(self.h, *args)
        if rc != 0:
            raise HbError(rc, lib().hb_last_error(self.h).decode('utf-8', 'replace'))

    def sync(self):
        self.call('hb_sync')

    @property
    def stream(self) -> int:
        return lib().hb_stream(self.h)

    @property
    def launches(self) -> int:
        return lib().hb_launch_count(self.h)

    def timer_start(self):
        self.call('hb_timer_start')

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.call('hb_timer_stop', C.byref(ms))
        return ms.value


def dp(t) -> C.c_uint64:
    """Device pointer of a torch tensor (or a raw integer address)."""
    return C.c_uint64(t if isinstance(t, int) else t.data_ptr())


# ---------------------------------------------------------------- page-locked host buffers for frame readers
import threading as _threading

_PINNED_LOCK = _threading.Lock()
_PINNED_FREE: dict = {}            # size in bytes -> [address]
_PINNED_CAP = 24 << 30             # keep at most this many idle bytes


def pinned_array(nbytes: int):
    """uint8 NumPy array of ``nbytes`` backed by page-locked memory from a process-wide pool (``hb_host_alloc``); falls back to
    ordinary memory when the CUDA library or a device is missing (reading files does not need a GPU).  Return it with
    ``pinned_release`` so that the next file reuses the pages instead of pinning fresh ones (~0.3 s per GB)."""
    import numpy as np
    nbytes = int(nbytes)
    with _PINNED_LOCK:
        lst = _PINNED_FREE.get(nbytes)
        addr = lst.pop() if lst else None
    if addr is None:
        try:
            L = lib()
            L.hb_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
            out = C.c_void_p()
            if L.hb_host_alloc(nbytes, C.byref(out)) != 0 or not out.value:
                return np.empty(nbytes, np.uint8)
            addr = out.value
        except Exception:
            return np.empty(nbytes, np.uint8)
    arr = np.frombuffer((C.c_uint8 * nbytes).from_address(addr), dtype=np.uint8)
    arr = arr.view(_PinnedArray)
    arr.hb_addr, arr.hb_bytes = addr, nbytes
    return arr


def pinned_release(arr):
    addr = getattr(arr, 'hb_addr', None)
    if addr is None:
        return
    arr.hb_addr = None
    with _PINNED_LOCK:
        idle = sum(k * len(v) for k, v in _PINNED_FREE.items())
        if idle + arr.hb_bytes <= _PINNED_CAP:
            _PINNED_FREE.setdefault(arr.hb_bytes, []).append(addr)
            return
    L = lib()
    L.hb_host_free.argtypes = [C.c_void_p]
    L.hb_host_free(C.c_void_p(addr))


def _make_pinned_cls():
    import numpy as np

    class PinnedArray(np.ndarray):
        hb_addr = None
        hb_bytes = 0
    return PinnedArray


try:
    _PinnedArray = _make_pinned_cls()
except Exception:          # NumPy missing: pinned_array is never reached either
    _PinnedArray = None
