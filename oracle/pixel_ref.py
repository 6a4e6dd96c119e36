"""TEST INFRASTRUCTURE -- NumPy oracle for the pre-encode pixel pipeline (not on the product path).

The reference's CPU path gets these conversions from libswscale inside the ffmpeg child
(``-pix_fmt`` at core/transcoder.py:464); the upscale path from ``upscale_gui_final.py:81-87`` geometry.

* ``pack_p010`` / ``to_10bit`` are exactly what libswscale does (<<8 / <<2) -- pinned against the
  bundled libswscale in tests/test_pixel_oracle.py.
* ``rgb_to_yuv420`` and ``scale_plane`` are this project's own fixed-point definitions (DESIGN.md
  "pixel kernels"); the NumPy model is the bit-exact oracle, libswscale / cv2 are +-LSB sanity bounds.
"""
from __future__ import annotations

import numpy as np

KR_KB = {'bt709': (0.2126, 0.0722), 'bt2020': (0.2627, 0.0593), 'bt601': (0.299, 0.114)}
CSC_SHIFT = 14


def pack_p010(y: np.ndarray, u: np.ndarray, v: np.ndarray):
    """8-bit planar 4:2:0 -> P010: 16-bit containers, sample in the top bits, UV interleaved."""
    y16 = y.astype(np.uint16) << 8
    uv = np.empty((u.shape[0], u.shape[1] * 2), np.uint16)
    uv[:, 0::2] = u.astype(np.uint16) << 8
    uv[:, 1::2] = v.astype(np.uint16) << 8
    return y16, uv


def to_10bit(plane: np.ndarray) -> np.ndarray:
    return plane.astype(np.uint16) << 2


def csc_coefficients(matrix: str, depth: int):
    """Q14 integer RGB->YCbCr limited-range coefficients.  Returns (cy[3], ccb[3], ccr[3], yoff, coff)."""
    kr, kb = KR_KB[matrix]
    kg = 1.0 - kr - kb
    sy = (219 << (depth - 8)) / 255.0 * (1 << CSC_SHIFT)
    sc = (224 << (depth - 8)) / 255.0 * (1 << CSC_SHIFT)
    cy = [int(round(k * sy)) for k in (kr, kg, kb)]
    ccb = [int(round(-kr / (2 * (1 - kb)) * sc)), int(round(-kg / (2 * (1 - kb)) * sc)), int(round(0.5 * sc))]
    ccr = [int(round(0.5 * sc)), int(round(-kg / (2 * (1 - kr)) * sc)), int(round(-kb / (2 * (1 - kr)) * sc))]
    return cy, ccb, ccr, 16 << (depth - 8), 128 << (depth - 8)


def rgb_to_yuv420(img: np.ndarray, matrix: str = 'bt709', depth: int = 8, bgr: bool = False):
    """Packed 8-bit full-range RGB (H, W, 3) -> limited-range planar 4:2:0 at ``depth`` bits (uint16 planes).

    Y per pixel; chroma from the 2x2 sum of each colour channel (centre-sited box)."""
    h, w, _ = img.shape
    assert h % 2 == 0 and w % 2 == 0
    c = img.astype(np.int64)
    if bgr:
        c = c[:, :, ::-1]
    cy, ccb, ccr, yoff, coff = csc_coefficients(matrix, depth)
    maxv = (1 << depth) - 1
    y = ((c[:, :, 0] * cy[0] + c[:, :, 1] * cy[1] + c[:, :, 2] * cy[2] + (1 << (CSC_SHIFT - 1))) >> CSC_SHIFT) + yoff
    s4 = c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]
    rnd, sh = 1 << (CSC_SHIFT + 1), CSC_SHIFT + 2
    cb = ((s4[:, :, 0] * ccb[0] + s4[:, :, 1] * ccb[1] + s4[:, :, 2] * ccb[2] + rnd) >> sh) + coff
    cr = ((s4[:, :, 0] * ccr[0] + s4[:, :, 1] * ccr[1] + s4[:, :, 2] * ccr[2] + rnd) >> sh) + coff
    return (np.clip(y, 0, maxv).astype(np.uint16), np.clip(cb, 0, maxv).astype(np.uint16),
            np.clip(cr, 0, maxv).astype(np.uint16))


# ---------------------------------------------------------------- polyphase scaler

SCALE_PHASES = 64


def bicubic_table() -> np.ndarray:
    """[64][4] Catmull-Rom (a = -0.5) taps in Q14; each row sums to 16384 (residue folded into tap 1)."""
    a = -0.5
    tab = np.zeros((SCALE_PHASES, 4), np.int32)
    for p in range(SCALE_PHASES):
        t = p / SCALE_PHASES
        d = np.array([1 + t, t, 1 - t, 2 - t])
        wgt = np.where(d <= 1, (a + 2) * d ** 3 - (a + 3) * d ** 2 + 1, a * d ** 3 - 5 * a * d ** 2 + 8 * a * d - 4 * a)
        q = np.round(wgt * 16384).astype(np.int32)
        q[1 if t < 0.5 else 2] += 16384 - q.sum()
        tab[p] = q
    return tab


def scale_positions(src: int, dst: int):
    """Centre-aligned mapping of each destination index to (first tap index, phase)."""
    d = np.arange(dst, dtype=np.int64)
    num = (2 * d + 1) * src - dst                      # position in units of 1/(2*dst) source samples
    ix = np.floor_divide(num, 2 * dst)
    frac = num - ix * 2 * dst
    phase = (frac * SCALE_PHASES + dst) // (2 * dst)
    ix = ix + (phase == SCALE_PHASES)
    phase = np.where(phase == SCALE_PHASES, 0, phase)
    return (ix - 1).astype(np.int32), phase.astype(np.int32)


def scale_plane(src: np.ndarray, dw: int, dh: int, out_depth: int = 8) -> np.ndarray:
    """8-bit plane -> (dh, dw) plane at ``out_depth`` bits.  Horizontal pass to Q6 int16, then vertical."""
    sh, sw = src.shape
    tab = bicubic_table().astype(np.int64)
    x0, xp = scale_positions(sw, dw)
    y0, yp = scale_positions(sh, dh)
    s = src.astype(np.int64)
    tmp = np.zeros((sh, dw), np.int64)
    for t in range(4):
        tmp += s[:, np.clip(x0 + t, 0, sw - 1)] * tab[xp, t][None, :]
    tmp = (tmp + 128) >> 8
    acc = np.zeros((dh, dw), np.int64)
    for t in range(4):
        acc += tmp[np.clip(y0 + t, 0, sh - 1), :] * tab[yp, t][:, None]
    shift = 20 - (out_depth - 8)
    out = (acc + (1 << (shift - 1))) >> shift
    return np.clip(out, 0, (1 << out_depth) - 1).astype(np.uint16)


def upscale_geometry(width: int, height: int, target_height: int = 0):
    """Target size rule of the reference's upscale path (upscale_gui_final.py:81-87)."""
    if target_height == 0:
        target_height = 1080 if height < 1080 else 2160 if height < 2160 else height
    scale = target_height / height
    return int(width * scale), target_height
