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
    # ---- pack P010, 4K, 16 distinct frames (16 x 12.4 MB in + 16 x 24.9 MB out > L2)
    w, h, nbuf = 3840, 2160, 16
    ys = [torch.randint(0, 256, (h, w), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    us = [torch.randint(0, 256, (h // 2, w // 2), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    vs = [torch.randint(0, 256, (h // 2, w // 2), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    dy = [torch.empty((h, w), dtype=torch.int16, device=dev) for _ in range(nbuf)]
    duv = [torch.empty((h // 2, w), dtype=torch.int16, device=dev) for _ in range(nbuf)]
    torch.cuda.synchronize()

    def pack(i):
        k = i % nbuf
        ctx.call('hb_pack_p010', _cabi.dp(ys[k]), w, _cabi.dp(us[k]), w // 2, _cabi.dp(vs[k]), w // 2, w, h, _cabi.dp(dy[k]), 2 * w, _cabi.dp(duv[k]), 2 * w)
    ms = timeit(ctx, pack)
    out.append({'kernel': 'k_pack_p010', 'shape': '3840x2160 yuv420p8 -> P010', 'ms': ms, 'bytes': int(w * h * 4.5)})
    # ---- BGR -> P010 BT.2020 (6 B/px)
    imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=dev) for _ in range(8)]
    torch.cuda.synchronize()

    def csc(i):
        k = i % 8
        ctx.call('hb_rgb_to_yuv420', _cabi.dp(imgs[k]), 3 * w, 1, 9, 10, w, h, _cabi.dp(dy[k]), 2 * w, _cabi.dp(duv[k]), 2 * w, 0, 0)
    ms = timeit(ctx, csc)
    out.append({'kernel': 'k_rgb_to_yuv420<10>', 'shape': '3840x2160 BGR24 -> P010 BT.2020', 'ms': ms, 'bytes': w * h * 6})
    # ---- fused upscale 1080p -> 4K P010
    sw, sh = 1920, 1080
    sy = [torch.randint(0, 256, (sh, sw), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    su = [torch.randint(0, 256, (sh // 2, sw // 2), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    sv = [torch.randint(0, 256, (sh // 2, sw // 2), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    torch.cuda.synchronize()

    def scale(i):
        k = i % nbuf
        ctx.call('hb_scale_yuv420_to_p010', _cabi.dp(sy[k]), sw, _cabi.dp(su[k]), sw // 2, _cabi.dp(sv[k]), sw // 2, sw, sh,
                 _cabi.dp(dy[k]), 2 * w, _cabi.dp(duv[k]), 2 * w, w, h)
    ms = timeit(ctx, scale)
    out.append({'kernel': 'k_scale<1>+<2>', 'shape': '1920x1080 yuv420p8 -> 3840x2160 P010 (fused scale + pack)', 'ms': ms,
                'bytes': int(sw * sh * 1.5 + w * h * 3)})
    for r in out:
        r['GB/s'] = round(r['bytes'] / r['ms'] / 1e6, 1)
        r['frac_of_measured_hbm'] = round(r['GB/s'] / peaks['hbm_gbs'], 3)
        r['ms'] = round(r['ms'], 4)
        print(json.dumps(r))
    del ys, us, vs, imgs, sy, su, sv, dy, duv
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
