"""TEST INFRASTRUCTURE -- ctypes wrapper of the CPU encoder model (oracle/hevc_encode.c)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import cmodel


class EncParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        'width', 'height', 'fps_num', 'fps_den', 'bit_depth', 'profile_idc', 'level_idc', 'tier', 'qp_i', 'qp_p', 'keyint',
        'colour_primaries', 'transfer_characteristics', 'matrix_coeffs', 'vui_colour', 'chroma_loc', 'full_range',
        'aud', 'repeat_headers', 'hrd', 'hdr10', 'vbv_maxrate_kbps', 'vbv_bufsize_kbit')] + [
        ('master_display', C.c_uint32 * 10), ('max_cll', C.c_int), ('max_fall', C.c_int), ('hash_sei', C.c_int), ('deblock', C.c_int), ('rate_control', C.c_int),
        ('min_keyint', C.c_int), ('scenecut', C.c_int), ('intra_in_p', C.c_int), ('sao', C.c_int), ('qp_cascade', C.c_int)]


class FrameInfo(C.Structure):
    _fields_ = [('is_idr', C.c_int), ('poc', C.c_int), ('qp', C.c_int), ('bytes', C.c_int), ('n_skip', C.c_int),
                ('n_merge', C.c_int), ('n_intra', C.c_int), ('psnr_y', C.c_double), ('est_bits16', C.c_longlong)]


def make_params(width, height, bit_depth=8, qp_i=24, qp_p=26, keyint=60, fps=(30, 1), hdr10=False, hash_sei=True, **kw) -> EncParams:
    p = EncParams()
    p.width, p.height, p.bit_depth = width, height, bit_depth
    p.fps_num, p.fps_den = fps
    p.profile_idc = 2 if bit_depth > 8 else 1
    p.level_idc, p.tier = kw.get('level_idc', 120), kw.get('tier', 0)
    p.qp_i, p.qp_p, p.keyint = qp_i, qp_p, keyint
    if hdr10:
        p.colour_primaries, p.transfer_characteristics, p.matrix_coeffs = 9, 16, 9
        p.aud = p.repeat_headers = p.hrd = p.hdr10 = 1
        p.chroma_loc = 0
        p.master_display[:] = [13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50]
        p.max_cll, p.max_fall = 1000, 400
    else:
        p.colour_primaries = p.transfer_characteristics = p.matrix_coeffs = 1
        p.chroma_loc = -1
    p.vui_colour = 1
    p.vbv_maxrate_kbps, p.vbv_bufsize_kbit = kw.get('vbv_maxrate_kbps', 2940), kw.get('vbv_bufsize_kbit', 3528)
    p.hash_sei = int(hash_sei)
    p.deblock = int(kw.get('deblock', 1))
    p.rate_control = int(kw.get('rate_control', 0))
    p.min_keyint = int(kw.get('min_keyint', max(2, keyint // 2)))
    p.scenecut, p.intra_in_p, p.sao, p.qp_cascade = (int(kw.get(k, 1)) for k in ('scenecut', 'intra_in_p', 'sao', 'qp_cascade'))
    for k in ('aud', 'repeat_headers', 'hrd'):
        if k in kw:
            setattr(p, k, int(kw[k]))
    return p


class ModelEncoder:
    def __init__(self, params: EncParams):
        L = cmodel.lib()
        L.orc_enc_create.restype = C.c_void_p
        L.orc_enc_create.argtypes = [C.POINTER(EncParams)]
        L.orc_enc_destroy.argtypes = [C.c_void_p]
        L.orc_enc_headers.restype = C.c_size_t
        L.orc_enc_headers.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_enc_frame.restype = C.c_long
        L.orc_enc_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                    C.POINTER(FrameInfo)]
        L.orc_enc_get_recon.argtypes = [C.c_void_p] * 4
        L.orc_enc_coded_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_enc_last_cus.restype = C.c_void_p
        L.orc_enc_last_cus.argtypes = [C.c_void_p]
        L.orc_enc_last_coefs.restype = C.c_void_p
        L.orc_enc_last_coefs.argtypes = [C.c_void_p]
        L.orc_enc_last_coarse_mv.restype = C.c_void_p
        L.orc_enc_last_coarse_mv.argtypes = [C.c_void_p]
        self.L, self.p = L, params
        self.h = L.orc_enc_create(C.byref(params))
        wc, hc = C.c_int(), C.c_int()
        L.orc_enc_coded_size(self.h, C.byref(wc), C.byref(hc))
        self.wc, self.hc = wc.value, hc.value
        self.buf = np.zeros(self.wc * self.hc * 4 + (1 << 16), np.uint8)

    def close(self):
        if self.h:
            self.L.orc_enc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def headers(self) -> bytes:
        n = self.L.orc_enc_headers(self.h, self.buf.ctypes.data, self.buf.size)
        return self.buf[:n].tobytes()

    def encode(self, y, u, v, force_idr: bool = False):
        """y/u/v: display-size planes (any integer dtype, values already at the encoder bit depth)."""
        y, u, v = (np.ascontiguousarray(a, dtype=np.uint16) for a in (y, u, v))
        info = FrameInfo()
        n = self.L.orc_enc_frame(self.h, y.ctypes.data, y.shape[1], u.ctypes.data, v.ctypes.data, u.shape[1], int(force_idr),
                                 self.buf.ctypes.data, self.buf.size, C.byref(info))
        if n < 0:
            raise RuntimeError('model encoder failed')
        return self.buf[:n].tobytes(), info

    def recon(self):
        y = np.zeros((self.hc, self.wc), np.uint16)
        u = np.zeros((self.hc // 2, self.wc // 2), np.uint16)
        v = np.zeros_like(u)
        self.L.orc_enc_get_recon(self.h, y.ctypes.data, u.ctypes.data, v.ctypes.data)
        return y, u, v

    def last_cus(self) -> np.ndarray:
        """structured view of the per-CU decisions of the last frame"""
        dt = np.dtype([('pred_mode', 'u1'), ('intra_mode', 'u1'), ('cbf', 'u1'), ('skip', 'u1'), ('mvx', '<i2'), ('mvy', '<i2')])
        n = (self.wc // 16) * (self.hc // 16)
        raw = (C.c_uint8 * (n * dt.itemsize)).from_address(self.L.orc_enc_last_cus(self.h))
        return np.frombuffer(raw, dtype=dt).reshape(self.hc // 16, self.wc // 16).copy()

    def last_coefs(self) -> np.ndarray:
        n = (self.wc // 16) * (self.hc // 16)
        raw = (C.c_int16 * (n * 384)).from_address(self.L.orc_enc_last_coefs(self.h))
        return np.frombuffer(raw, dtype=np.int16).reshape(n, 384).copy()

    def last_coarse_mv(self) -> np.ndarray:
        tw, th = (self.wc + 31) // 32, (self.hc + 31) // 32
        raw = (C.c_int16 * (tw * th * 2)).from_address(self.L.orc_enc_last_coarse_mv(self.h))
        return np.frombuffer(raw, dtype=np.int16).reshape(th, tw, 2).copy()
