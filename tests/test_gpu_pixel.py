"""GPU parity: pixel pipeline kernels (through the C ABI) vs the NumPy oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    import torch
    from hevc_b200 import _cabi
    ctx = _cabi.Context(0)
    yield ctx, torch
    ctx.close()


def _gpu(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('w,h', [(1920, 1080), (3840, 2160), (1280, 720), (66, 34), (16, 2), (130, 70)])
def test_pack_p010(env, w, h):
    from hevc_b200 import ops
    from oracle import pixel_ref
    ctx, torch = env
    rng = np.random.default_rng(w * 7 + h)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    u = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
    v = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
    dy, duv = ops.pack_p010(ctx, _gpu(torch, y), _gpu(torch, u), _gpu(torch, v))
    ctx.sync()
    ry, ruv = pixel_ref.pack_p010(y, u, v)
    assert (dy.cpu().numpy().view(np.uint16) == ry).all()
    assert (duv.cpu().numpy().view(np.uint16) == ruv).all()


def test_pack_p010_unaligned_views(env):
    from hevc_b200 import ops
    from oracle import pixel_ref
    ctx, torch = env
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (70, 200), dtype=np.uint8)
    y, u, v = big[1:65, 3:131], big[2:34, 5:69], big[30:62, 101:165]     # odd base offsets, pitch 200
    g = _gpu(torch, big)
    dy, duv = ops.pack_p010(ctx, g[1:65, 3:131], g[2:34, 5:69], g[30:62, 101:165])
    ctx.sync()
    ry, ruv = pixel_ref.pack_p010(y, u, v)
    assert (dy.cpu().numpy().view(np.uint16) == ry).all() and (duv.cpu().numpy().view(np.uint16) == ruv).all()


@pytest.mark.parametrize('matrix', ['bt709', 'bt2020'])
@pytest.mark.parametrize('depth', [8, 10])
@pytest.mark.parametrize('w,h,bgr', [(1920, 1080, True), (64, 32, False), (70, 18, True)])
def test_rgb_to_yuv420(env, matrix, depth, w, h, bgr):
    from hevc_b200 import ops
    from oracle import pixel_ref
    ctx, torch = env
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[:2, :8] = 255
    img[2:4, :8] = 0
    ry, rcb, rcr = pixel_ref.rgb_to_yuv420(img, matrix, depth, bgr)
    out = ops.rgb_to_yuv420(ctx, _gpu(torch, img), matrix, depth, bgr)
    ctx.sync()
    if depth == 8:
        gy, gu, gv = (t.cpu().numpy() for t in out)
        assert (gy == ry).all() and (gu == rcb).all() and (gv == rcr).all()
    else:
        gy, guv = (t.cpu().numpy().view(np.uint16) for t in out)
        assert (gy == ry << 6).all() and (guv[:, 0::2] == rcb << 6).all() and (guv[:, 1::2] == rcr << 6).all()


@pytest.mark.parametrize('sw,sh,dw,dh,depth', [(1920, 1080, 3840, 2160, 8), (960, 540, 1920, 1080, 10), (1280, 720, 1920, 1080, 8),
                                              (100, 60, 37, 23, 8), (64, 48, 64, 48, 8), (640, 360, 64, 36, 10)])
def test_scale_plane(env, sw, sh, dw, dh, depth):
    from hevc_b200 import ops
    from oracle import pixel_ref
    ctx, torch = env
    rng = np.random.default_rng(sw + dw)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    out = ops.scale_plane(ctx, _gpu(torch, src), dw, dh, depth)
    ctx.sync()
    ref = pixel_ref.scale_plane(src, dw, dh, depth)
    got = out.cpu().numpy()
    got = got.astype(np.uint16) if depth == 8 else got.view(np.uint16)
    assert (got == ref).all()


def test_fused_upscale_to_p010(env):
    """upscale_gui_final geometry (1080p -> 2160p) fused with the 10-bit P010 pack."""
    from hevc_b200 import ops
    from oracle import pixel_ref
    ctx, torch = env
    rng = np.random.default_rng(5)
    sw, sh = 1920, 1080
    dw, dh = pixel_ref.upscale_geometry(sw, sh)
    y = rng.integers(16, 236, (sh, sw), dtype=np.uint8)
    u = rng.integers(16, 241, (sh // 2, sw // 2), dtype=np.uint8)
    v = rng.integers(16, 241, (sh // 2, sw // 2), dtype=np.uint8)
    dy, duv = ops.scale_yuv420_to_p010(ctx, _gpu(torch, y), _gpu(torch, u), _gpu(torch, v), dw, dh)
    ctx.sync()
    gy, guv = dy.cpu().numpy().view(np.uint16), duv.cpu().numpy().view(np.uint16)
    assert (gy == pixel_ref.scale_plane(y, dw, dh, 10) << 6).all()
    assert (guv[:, 0::2] == pixel_ref.scale_plane(u, dw // 2, dh // 2, 10) << 6).all()
    assert (guv[:, 1::2] == pixel_ref.scale_plane(v, dw // 2, dh // 2, 10) << 6).all()


@pytest.mark.parametrize('w,h,n', [(1920, 1080, 3), (130, 70, 4), (64, 32, 2)])
def test_batched_forms_match_the_oracle_frame_by_frame(env, w, h, n):
    """hb_pack_p010_batch / hb_rgb_to_p010_batch / hb_scale_yuv420_to_p010_batch: one launch over n tightly packed frames"""
    import ctypes as C

    from hevc_b200 import _cabi
    from oracle import pixel_ref
    ctx, torch = env
    rng = np.random.default_rng(w + h + n)
    lw, cw = w * h, (w // 2) * (h // 2)
    src = rng.integers(0, 256, (n, lw + 2 * cw), dtype=np.uint8)
    d_src = _gpu(torch, src)
    dst = torch.zeros((n, 3 * lw), dtype=torch.uint8, device='cuda')
    ctx.call('hb_pack_p010_batch', _cabi.dp(d_src), C.c_size_t(lw + 2 * cw), _cabi.dp(dst), C.c_size_t(3 * lw), w, h, n)
    ctx.sync()
    got = dst.cpu().numpy().view(np.uint16).reshape(n, -1)
    for i in range(n):
        ry, ruv = pixel_ref.pack_p010(src[i, :lw].reshape(h, w), src[i, lw:lw + cw].reshape(h // 2, w // 2), src[i, lw + cw:].reshape(h // 2, w // 2))
        assert (got[i, :lw] == ry.reshape(-1)).all() and (got[i, lw:] == ruv.reshape(-1)).all()
    # ---- BGR -> P010
    img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    d_img = _gpu(torch, img)
    dst.zero_()
    ctx.call('hb_rgb_to_p010_batch', _cabi.dp(d_img), C.c_size_t(3 * lw), 1, 9, w, h, _cabi.dp(dst), C.c_size_t(3 * lw), n)
    ctx.sync()
    got = dst.cpu().numpy().view(np.uint16).reshape(n, -1)
    for i in range(n):
        y, u, v = pixel_ref.rgb_to_yuv420(img[i], 'bt2020', 10, bgr=True)
        uv = np.empty((h // 2, w), np.uint16)
        uv[:, 0::2], uv[:, 1::2] = u << 6, v << 6
        assert (got[i, :lw] == (y << 6).reshape(-1)).all() and (got[i, lw:] == uv.reshape(-1)).all()
    # ---- scale x2 -> P010
    dw, dh = 2 * w, 2 * h
    big = torch.zeros((n, 3 * dw * dh), dtype=torch.uint8, device='cuda')
    ctx.call('hb_scale_yuv420_to_p010_batch', _cabi.dp(d_src), C.c_size_t(lw + 2 * cw), w, h, _cabi.dp(big), C.c_size_t(3 * dw * dh), dw, dh, n)
    ctx.sync()
    got = big.cpu().numpy().view(np.uint16).reshape(n, -1)
    for i in range(n):
        y = pixel_ref.scale_plane(src[i, :lw].reshape(h, w), dw, dh, 10) << 6
        u = pixel_ref.scale_plane(src[i, lw:lw + cw].reshape(h // 2, w // 2), dw // 2, dh // 2, 10) << 6
        v = pixel_ref.scale_plane(src[i, lw + cw:].reshape(h // 2, w // 2), dw // 2, dh // 2, 10) << 6
        uv = np.empty((dh // 2, dw), np.uint16)
        uv[:, 0::2], uv[:, 1::2] = u, v
        assert (got[i, :dw * dh] == y.reshape(-1)).all() and (got[i, dw * dh:] == uv.reshape(-1)).all()
