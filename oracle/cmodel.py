"""TEST INFRASTRUCTURE -- ctypes access to the plain-C oracle (oracle/*.c -> oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / 'liboracle.so'
    srcs = list(_HERE.glob('*.c')) + list(_HERE.glob('*.h'))
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(['make', '-s', '-C', str(_HERE)], check=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / 'liboracle.so'
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u16(a):
    return np.ascontiguousarray(a, dtype=np.uint16)


def _i16(a):
    return np.ascontiguousarray(a, dtype=np.int16)


def sad(a, b):
    """a, b: [n, H, W] uint16 -> int32[n]"""
    a, b = _u16(a), _u16(b)
    n, h, w = a.shape
    f = lib().orc_sad
    f.restype = C.c_int
    return np.array([f(_p(a[i]), w, _p(b[i]), w, w, h) for i in range(n)], np.int32)


def satd(a, b):
    a, b = _u16(a), _u16(b)
    n, h, w = a.shape
    f = lib().orc_satd
    f.restype = C.c_int
    return np.array([f(_p(a[i]), w, _p(b[i]), w, w, h) for i in range(n)], np.int32)


def sa8d(a, b):
    a, b = _u16(a), _u16(b)
    n, h, w = a.shape
    f = lib().orc_sa8d
    f.restype = C.c_int
    return np.array([f(_p(a[i]), w, _p(b[i]), w, w) for i in range(n)], np.int32)


def transform_matrix(size: int, is_dst: bool = False) -> np.ndarray:
    m = np.zeros((size, size), np.int16)
    lib().orc_transform_matrix(size, int(is_dst), _p(m))
    return m


def fwd_transform(res, bit_depth: int, is_dst: bool = False):
    res = _i16(res)
    n, size, _ = res.shape
    out = np.zeros_like(res)
    f = lib().orc_fwd_transform
    for i in range(n):
        f(_p(res[i]), size, _p(out[i]), size, bit_depth, int(is_dst))
    return out


def inv_transform(coef, bit_depth: int, is_dst: bool = False):
    coef = _i16(coef)
    n, size, _ = coef.shape
    out = np.zeros_like(coef)
    f = lib().orc_inv_transform
    for i in range(n):
        f(_p(coef[i]), _p(out[i]), size, size, bit_depth, int(is_dst))
    return out


def quant(coef, qp: int, bit_depth: int, is_intra: bool):
    coef = _i16(coef)
    n, size, _ = coef.shape
    out = np.zeros_like(coef)
    f = lib().orc_quant
    f.restype = C.c_int
    ns = np.array([f(_p(coef[i]), _p(out[i]), size, qp, bit_depth, int(is_intra)) for i in range(n)], np.int32)
    return out, ns


def dequant(level, qp: int, bit_depth: int):
    level = _i16(level)
    n, size, _ = level.shape
    out = np.zeros_like(level)
    f = lib().orc_dequant
    for i in range(n):
        f(_p(level[i]), _p(out[i]), size, qp, bit_depth)
    return out


def intra_pred_all(nb, size: int, is_luma: bool, strong: bool, bit_depth: int):
    """nb: [n, 4*size+1] uint16 -> [n, 35, size, size] uint16"""
    nb = _u16(nb)
    n = nb.shape[0]
    out = np.zeros((n, 35, size, size), np.uint16)
    f = lib().orc_intra_pred_all
    for i in range(n):
        f(_p(nb[i]), _p(out[i]), size, int(is_luma), int(strong), bit_depth)
    return out


def interp_luma(ref_plane, x, y, w, h, mvx, mvy, bit_depth):
    """Prediction block for quarter-sample motion vector (mvx, mvy) at integer position (x, y)."""
    ref_plane = _u16(ref_plane)
    rs = ref_plane.shape[1]
    out = np.zeros((h, w), np.uint16)
    base = ref_plane.ctypes.data + 2 * ((y + (mvy >> 2)) * rs + x + (mvx >> 2))
    lib().orc_interp_luma(C.c_void_p(base), rs, _p(out), w, w, h, mvx & 3, mvy & 3, bit_depth)
    return out
