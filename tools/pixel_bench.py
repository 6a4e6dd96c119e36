"""Micro-benchmarks of the HBM-bound pixel kernels and the batched primitives (BASELINE config 5), CUDA-event timed on the
context stream, inputs larger than L2 cycled between iterations.  Prints JSON lines; copy into profiles/."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def timeit(ctx, fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    ctx.sync()
    ctx.timer_start()
    for i in range(iters):
        fn(i)
    return ctx.timer_stop() / iters


def main():
    import torch
    from hevc_b200 import _cabi, ops
    peaks = json.loads((Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json').read_text()) if (Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json').exists() else {'hbm_gbs': 6650.0}
    ctx = _cabi.Context(0)
    dev = torch.device('cuda', 0)
    out = []
    # ---- batched launches (one launch per 32-frame read batch, the way a worker feeds the device): every launch moves > 1 GB,
    #      several times the 126 MB L2, so nothing is served from cache
    import ctypes as C
    w, h, nb = 3840, 2160, 32
    lw, cw = w * h, (w // 2) * (h // 2)
    src = torch.randint(0, 256, (nb, lw + 2 * cw), dtype=torch.uint8, device=dev)
    dst = torch.empty((nb, 3 * lw), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ms = timeit(ctx, lambda i: ctx.call('hb_pack_p010_batch', _cabi.dp(src), C.c_size_t(lw + 2 * cw), _cabi.dp(dst), C.c_size_t(3 * lw), w, h, nb), iters=10)
    out.append({'kernel': 'k_pack_p010', 'shape': f'{nb} x 3840x2160 yuv420p8 -> P010, one launch', 'ms': ms, 'bytes': int(nb * w * h * 4.5)})
    del src
    img = torch.randint(0, 256, (nb, 3 * lw), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ms = timeit(ctx, lambda i: ctx.call('hb_rgb_to_p010_batch', _cabi.dp(img), C.c_size_t(3 * lw), 1, 9, w, h, _cabi.dp(dst), C.c_size_t(3 * lw), nb), iters=10)
    out.append({'kernel': 'k_rgb_to_yuv420<10>', 'shape': f'{nb} x 3840x2160 BGR24 -> P010 BT.2020, one launch', 'ms': ms, 'bytes': nb * w * h * 6})
    del img
    sw, sh = 1920, 1080
    sl, sc = sw * sh, (sw // 2) * (sh // 2)
    small = torch.randint(0, 256, (nb, sl + 2 * sc), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ms = timeit(ctx, lambda i: ctx.call('hb_scale_yuv420_to_p010_batch', _cabi.dp(small), C.c_size_t(sl + 2 * sc), sw, sh, _cabi.dp(dst), C.c_size_t(3 * lw),
                                        w, h, nb), iters=10)
    out.append({'kernel': 'k_scale8<1,1> + <2,2>', 'shape': f'{nb} x 1920x1080 yuv420p8 -> 3840x2160 P010 (fused scale + pack), two launches', 'ms': ms,
                'bytes': int(nb * (sw * sh * 1.5 + w * h * 3))})
    # single-frame launches of the same kernels, for comparison with round 1
    ms = timeit(ctx, lambda i: ctx.call('hb_scale_yuv420_to_p010_batch', _cabi.dp(small[i % nb]), C.c_size_t(sl + 2 * sc), sw, sh, _cabi.dp(dst[i % nb]),
                                        C.c_size_t(3 * lw), w, h, 1), iters=32)
    out.append({'kernel': 'k_scale8<1,1> + <2,2>', 'shape': '1920x1080 -> 3840x2160 P010, one frame per call', 'ms': ms, 'bytes': int(sw * sh * 1.5 + w * h * 3)})
    for r in out:
        r['GB/s'] = round(r['bytes'] / r['ms'] / 1e6, 1)
        r['frac_of_measured_hbm'] = round(r['GB/s'] / peaks['hbm_gbs'], 3)
        r['ms'] = round(r['ms'], 4)
        print(json.dumps(r), flush=True)
    del small, dst
    if '--pixel-only' in sys.argv:
        ctx.close()
        return
    # ---- primitives, 2^20 blocks per launch
    n = 1 << 20
    g = torch.Generator(device=dev).manual_seed(1)
    for (bw, bh) in ((8, 8), (16, 16), (32, 32), (64, 64)):
        nb = n if bw * bh <= 1024 else n // 4
        a = torch.randint(0, 1024, (nb, bh, bw), device=dev, generator=g, dtype=torch.int16)
        b = torch.randint(0, 1024, (nb, bh, bw), device=dev, generator=g, dtype=torch.int16)
        o = torch.empty((nb,), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for name in ('hb_sad', 'hb_satd'):
            ms = timeit(ctx, lambda i: ctx.call(name, _cabi.dp(a), _cabi.dp(b), nb, bw, bh, _cabi.dp(o)), iters=10)
            print(json.dumps({'kernel': name, 'block': f'{bw}x{bh}', 'blocks': nb, 'ms': round(ms, 4), 'Gpixel/s': round(nb * bw * bh / ms / 1e6, 1),
                              'GB/s': round(nb * bw * bh * 4 / ms / 1e6, 1)}))
        del a, b, o
    for size in (4, 8, 16, 32):
        nb = n if size <= 16 else n // 4
        res = torch.randint(-1023, 1024, (nb, size, size), device=dev, generator=g, dtype=torch.int16)
        co = torch.empty_like(res)
        torch.cuda.synchronize()
        for name, args in (('hb_fwd_transform', (10, 0)), ('hb_inv_transform', (10, 0))):
            ms = timeit(ctx, lambda i: ctx.call(name, _cabi.dp(res), nb, size, *args, _cabi.dp(co)), iters=10)
            print(json.dumps({'kernel': name, 'size': size, 'blocks': nb, 'ms': round(ms, 4), 'Gcoef/s': round(nb * size * size / ms / 1e6, 1),
                              'GMAC/s (2N per coef, matrix form)': round(nb * size * size * 2 * size / ms / 1e6, 1)}))
        lv = torch.empty_like(res)
        ns = torch.empty((nb,), dtype=torch.int32, device=dev)
        ms = timeit(ctx, lambda i: ctx.call('hb_quant', _cabi.dp(res), nb, size, 34, 10, 0, _cabi.dp(lv), _cabi.dp(ns)), iters=10)
        print(json.dumps({'kernel': 'hb_quant', 'size': size, 'blocks': nb, 'ms': round(ms, 4), 'GB/s': round(nb * size * size * 4 / ms / 1e6, 1)}))
        del res, co, lv, ns
    for size in (8, 16, 32):
        nb = 1 << 16
        nbuf_ = torch.randint(0, 1024, (nb, 4 * size + 1), device=dev, generator=g, dtype=torch.int16)
        pred = torch.empty((nb, 35, size, size), dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        ms = timeit(ctx, lambda i: ctx.call('hb_intra_pred_all', _cabi.dp(nbuf_), nb, size, 1, 0, 10, _cabi.dp(pred)), iters=5)
        print(json.dumps({'kernel': 'hb_intra_pred_all', 'size': size, 'blocks': nb, 'ms': round(ms, 4), 'Gpixel/s': round(nb * 35 * size * size / ms / 1e6, 1),
                          'GB/s written': round(nb * 35 * size * size * 2 / ms / 1e6, 1)}))
        del nbuf_, pred
    ctx.close()


if __name__ == '__main__':
    main()
