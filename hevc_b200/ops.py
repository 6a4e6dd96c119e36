"""Tensor-level wrappers over the C ABI for the pixel pipeline and the batched primitives.

torch is used only as the device-memory allocator; every operation is one call into libhevc_b200.so."""
from __future__ import annotations

import ctypes as C

import torch

from ._cabi import Context, dp

MATRIX_IDS = {'bt709': 1, 'bt601': 6, 'bt2020': 9}


def _dev(ctx: Context):
    return torch.device('cuda', ctx.device)


def pack_p010(ctx: Context, y: torch.Tensor, u: torch.Tensor, v: torch.Tensor):
    h, w = y.shape
    dy = torch.empty((h, w), dtype=torch.int16, device=y.device)
    duv = torch.empty(((h + 1) // 2, 2 * ((w + 1) // 2)), dtype=torch.int16, device=y.device)
    ctx.call('hb_pack_p010', dp(y), y.stride(0), dp(u), u.stride(0), dp(v), v.stride(0), w, h,
             dp(dy), dy.stride(0) * 2, dp(duv), duv.stride(0) * 2)
    return dy, duv


def rgb_to_yuv420(ctx: Context, img: torch.Tensor, matrix: str = 'bt709', depth: int = 8, bgr: bool = False):
    h, w, _ = img.shape
    if depth == 8:
        dy = torch.empty((h, w), dtype=torch.uint8, device=img.device)
        du = torch.empty((h // 2, w // 2), dtype=torch.uint8, device=img.device)
        dv = torch.empty_like(du)
        ctx.call('hb_rgb_to_yuv420', dp(img), img.stride(0), int(bgr), MATRIX_IDS[matrix], 8, w, h,
                 dp(dy), dy.stride(0), dp(du), du.stride(0), dp(dv), dv.stride(0))
        return dy, du, dv
    dy = torch.empty((h, w), dtype=torch.int16, device=img.device)
    duv = torch.empty((h // 2, w), dtype=torch.int16, device=img.device)
    ctx.call('hb_rgb_to_yuv420', dp(img), img.stride(0), int(bgr), MATRIX_IDS[matrix], 10, w, h,
             dp(dy), dy.stride(0) * 2, dp(duv), duv.stride(0) * 2, C.c_uint64(0), 0)
    return dy, duv


def scale_plane(ctx: Context, src: torch.Tensor, dw: int, dh: int, out_depth: int = 8):
    sh, sw = src.shape
    dst = torch.empty((dh, dw), dtype=torch.uint8 if out_depth == 8 else torch.int16, device=src.device)
    ctx.call('hb_scale_plane', dp(src), src.stride(0), sw, sh, dp(dst), dst.stride(0) * dst.element_size(), dw, dh,
             out_depth, 0, 1)
    return dst


def scale_yuv420_to_p010(ctx: Context, y, u, v, dw: int, dh: int):
    sh, sw = y.shape
    dy = torch.empty((dh, dw), dtype=torch.int16, device=y.device)
    duv = torch.empty((dh // 2, dw), dtype=torch.int16, device=y.device)
    ctx.call('hb_scale_yuv420_to_p010', dp(y), y.stride(0), dp(u), u.stride(0), dp(v), v.stride(0), sw, sh,
             dp(dy), dy.stride(0) * 2, dp(duv), duv.stride(0) * 2, dw, dh)
    return dy, duv


# ---------------------------------------------------------------- batched primitives (uint16 samples carried as int16 tensors)

def _cost(ctx, name, a, b, *shape_args):
    n = a.shape[0]
    out = torch.empty((n,), dtype=torch.int32, device=a.device)
    ctx.call(name, dp(a), dp(b), n, *shape_args, dp(out))
    return out


def sad(ctx, a, b):
    return _cost(ctx, 'hb_sad', a, b, a.shape[2], a.shape[1])


def sad_multi(ctx, fenc, refs):
    """x265 sad_x3 / sad_x4: ``fenc`` [n, H, W] against 3 or 4 reference batches -> int32 [n, len(refs)]"""
    n = fenc.shape[0]
    out = torch.empty((n, len(refs)), dtype=torch.int32, device=fenc.device)
    r = list(refs) + [None] * (4 - len(refs))
    ctx.call('hb_sad_multi', dp(fenc), dp(r[0]), dp(r[1]), dp(r[2]), dp(r[3]) if r[3] is not None else C.c_uint64(0), len(refs), n,
             fenc.shape[2], fenc.shape[1], dp(out))
    return out


def satd(ctx, a, b):
    return _cost(ctx, 'hb_satd', a, b, a.shape[2], a.shape[1])


def sa8d(ctx, a, b):
    return _cost(ctx, 'hb_sa8d', a, b, a.shape[1])


def fwd_transform(ctx, res, bit_depth, is_dst=False):
    out = torch.empty_like(res)
    ctx.call('hb_fwd_transform', dp(res), res.shape[0], res.shape[1], bit_depth, int(is_dst), dp(out))
    return out


def inv_transform(ctx, coef, bit_depth, is_dst=False):
    out = torch.empty_like(coef)
    ctx.call('hb_inv_transform', dp(coef), coef.shape[0], coef.shape[1], bit_depth, int(is_dst), dp(out))
    return out


def quant(ctx, coef, qp, bit_depth, is_intra):
    out = torch.empty_like(coef)
    ns = torch.empty((coef.shape[0],), dtype=torch.int32, device=coef.device)
    ctx.call('hb_quant', dp(coef), coef.shape[0], coef.shape[1], qp, bit_depth, int(is_intra), dp(out), dp(ns))
    return out, ns


def dequant(ctx, level, qp, bit_depth):
    out = torch.empty_like(level)
    ctx.call('hb_dequant', dp(level), level.shape[0], level.shape[1], qp, bit_depth, dp(out))
    return out


def intra_pred_all(ctx, nb, size, is_luma, strong, bit_depth):
    n = nb.shape[0]
    out = torch.empty((n, 35, size, size), dtype=torch.int16, device=nb.device)
    ctx.call('hb_intra_pred_all', dp(nb), n, size, int(is_luma), int(strong), bit_depth, dp(out))
    return out
