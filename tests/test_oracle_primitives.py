"""CPU oracle self-checks: known answers and algebraic properties of oracle/primitives.c and oracle/pixel_ref.py,
and pixel_ref pinned against the bundled libswscale (what the reference's ffmpeg child runs for -pix_fmt)."""
import numpy as np
import pytest

from oracle import cmodel, fforacle, pixel_ref

RNG = np.random.default_rng(7)

# H.265 table of the 8-point core transform (rows 0..7)
DCT8 = np.array([[64, 64, 64, 64, 64, 64, 64, 64], [89, 75, 50, 18, -18, -50, -75, -89], [83, 36, -36, -83, -83, -36, 36, 83],
                 [75, -18, -89, -50, 50, 89, 18, -75], [64, -64, -64, 64, 64, -64, -64, 64], [50, -89, 18, 75, -75, -18, 89, -50],
                 [36, -83, 83, -36, -36, 83, -83, 36], [18, -50, 75, -89, 89, -75, 50, -18]])


def test_transform_matrices():
    assert (cmodel.transform_matrix(8) == DCT8).all()
    m32 = cmodel.transform_matrix(32).astype(int)
    assert list(m32[1, :16]) == [90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4]
    for n in (4, 8, 16, 32):
        m = cmodel.transform_matrix(n).astype(int)
        g = m @ m.T                      # near-orthogonal: off-diagonal energy tiny vs 64*64*n
        assert np.abs(g - np.diag(np.diag(g))).max() < 0.004 * 64 * 64 * n
        if n > 4:                        # even rows embed the half-size transform
            assert (m[::2, : n // 2] == cmodel.transform_matrix(n // 2).astype(int)).all()


@pytest.mark.parametrize('size', [4, 8, 16, 32])
@pytest.mark.parametrize('depth', [8, 10])
def test_dct_roundtrip(size, depth):
    lim = (1 << depth) - 1
    res = RNG.integers(-lim, lim + 1, (16, size, size)).astype(np.int16)
    back = cmodel.inv_transform(cmodel.fwd_transform(res, depth), depth)
    assert np.abs(back.astype(int) - res).max() <= 6 << (depth - 8)      # integer transform pair is only near-orthogonal
    if size == 4:
        back = cmodel.inv_transform(cmodel.fwd_transform(res, depth, True), depth, True)
        assert np.abs(back.astype(int) - res).max() <= 6 << (depth - 8)


def test_dct_dc_known_answer():
    # constant residual c: only the DC coefficient, value c * N * 64 * 64 >> (shift1 + shift2) = c * 2^(6 - log2N) for 8 bit ... = c << (15 - 8 - log2 N) / 2
    for size, l2 in ((4, 2), (8, 3), (16, 4), (32, 5)):
        res = np.full((1, size, size), 10, np.int16)
        co = cmodel.fwd_transform(res, 8)[0]
        assert co[0, 0] == 10 * size * size * 64 * 64 >> (l2 - 1 + l2 + 6) and np.count_nonzero(co) == 1


def _hadamard(n):
    h = np.array([[1]])
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h


def test_satd_sa8d_against_matrix_hadamard():
    for (w, h) in ((4, 4), (8, 4), (8, 8), (16, 16), (16, 12), (12, 16), (32, 8), (64, 64)):
        a = RNG.integers(0, 1024, (3, h, w)).astype(np.uint16)
        b = RNG.integers(0, 1024, (3, h, w)).astype(np.uint16)
        d = a.astype(int) - b
        h4 = _hadamard(4)
        want = [sum(np.abs(h4 @ d[i, y:y + 4, x:x + 4] @ h4).sum() // 2 for y in range(0, h, 4) for x in range(0, w, 4)) for i in range(3)]
        assert list(cmodel.satd(a, b)) == want
        assert list(cmodel.sad(a, b)) == list(np.abs(d).sum((1, 2)))
    h8 = _hadamard(8)
    for n in (8, 16, 32):
        a = RNG.integers(0, 256, (2, n, n)).astype(np.uint16)
        b = RNG.integers(0, 256, (2, n, n)).astype(np.uint16)
        d = a.astype(int) - b
        want = []
        for i in range(2):
            if n == 8:
                want.append((np.abs(h8 @ d[i] @ h8).sum() + 2) >> 2)
            else:
                want.append(sum((sum(np.abs(h8 @ d[i, y + j:y + j + 8, x + k:x + k + 8] @ h8).sum() for j in (0, 8) for k in (0, 8)) + 2) >> 2
                                for y in range(0, n, 16) for x in range(0, n, 16)))
        assert list(cmodel.sa8d(a, b)) == want


def test_sad_matches_libavutil_pixelutils():
    """SAD pinned against an independent implementation: av_pixelutils_get_sad_fn of the FFmpeg libavutil bundled with the OpenCV
    wheel (square 8-bit blocks 4x4 ... 32x32, the x265 `sad[]` sizes on the CU grid).  libx265 itself is not in the image."""
    import ctypes as C
    from hevc_b200 import avreader
    try:
        av = avreader._lib('avutil')
        get = av.av_pixelutils_get_sad_fn
    except (avreader.Unsupported, AttributeError, OSError):
        pytest.skip('bundled libavutil without pixelutils')
    get.restype = C.c_void_p
    get.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t)
    rng = np.random.default_rng(5)
    checked = 0
    for bits in (2, 3, 4, 5):
        addr = get(bits, bits, 0, None)
        if not addr:
            continue
        fn, n = proto(addr), 1 << bits
        for trial in range(20):
            a = rng.integers(0, 256, (n, n), dtype=np.uint8)
            b = (255 - a if trial == 0 else rng.integers(0, 256, (n, n), dtype=np.uint8))
            want = fn(a.ctypes.data, n, b.ctypes.data, n)
            got = int(cmodel.sad(a.astype(np.uint16)[None], b.astype(np.uint16)[None])[0])
            assert got == want == int(np.abs(a.astype(int) - b.astype(int)).sum()), (n, trial)
            checked += 1
    assert checked >= 20


def test_quant_dequant_properties():
    for depth in (8, 10):
        for size in (4, 8, 16, 32):
            coef = RNG.integers(-32768, 32768, (4, size, size)).astype(np.int16)
            for qp in (0, 17, 22 + 6 * (depth - 8), 51 + 6 * (depth - 8)):
                lvl, ns = cmodel.quant(coef, qp, depth, True)
                assert (ns == np.count_nonzero(lvl.reshape(4, -1), axis=1)).all()
                assert (np.sign(lvl) * np.sign(coef) >= 0).all()
                lvl_p, _ = cmodel.quant(coef, qp, depth, False)
                assert (np.abs(lvl_p.astype(int)) <= np.abs(lvl.astype(int))).all()   # smaller dead-zone offset for inter
                rec = cmodel.dequant(lvl, qp, depth).astype(int)
                step = ([40, 45, 51, 57, 64, 72][qp % 6] << (qp // 6)) / 2.0 ** (6 - (15 - depth - int(np.log2(size))))
                ok = (np.abs(rec) < 32767) & (np.abs(lvl.astype(int)) < 32767)
                assert (np.abs(rec - coef)[ok] <= step + 1).all()


def test_intra_dc_planar_flat():
    for size in (4, 8, 16, 32):
        nb = np.full((1, 4 * size + 1), 300, np.uint16)
        pred = cmodel.intra_pred_all(nb, size, True, True, 10)
        assert (pred == 300).all()            # every mode reproduces a flat neighbourhood
    # vertical mode copies the top row (no edge filter for chroma)
    nb = np.arange(4 * 8 + 1, dtype=np.uint16).reshape(1, -1) + 100
    pred = cmodel.intra_pred_all(nb, 8, False, False, 10)[0]
    assert (pred[26] == nb[0, 1:9][None, :]).all()
    assert (pred[10] == nb[0, 17:25][:, None]).all()
    assert (pred[34][0] == nb[0, 2:10]).all()          # mode 34: 45 degrees down-left from the top-right samples


def test_pack_and_depth_match_swscale():
    h, w = 32, 48
    y = RNG.integers(0, 256, (h, w)).astype(np.uint8)
    u = RNG.integers(0, 256, (h // 2, w // 2)).astype(np.uint8)
    v = RNG.integers(0, 256, (h // 2, w // 2)).astype(np.uint8)
    py, puv = pixel_ref.pack_p010(y, u, v)
    sy, suv = fforacle.sws_convert([y, u, v], fforacle.AV_PIX_FMT_YUV420P, fforacle.AV_PIX_FMT_P010LE, w, h,
                                   [(h, 2 * w, np.uint16), (h // 2, 2 * w, np.uint16)])
    assert (py == sy).all() and (puv == suv).all()
    t = fforacle.sws_convert([y, u, v], fforacle.AV_PIX_FMT_YUV420P, fforacle.AV_PIX_FMT_YUV420P10LE, w, h,
                             [(h, 2 * w, np.uint16), (h // 2, w, np.uint16), (h // 2, w, np.uint16)])
    assert (pixel_ref.to_10bit(y) == t[0]).all() and (pixel_ref.to_10bit(u) == t[1]).all()


@pytest.mark.parametrize('matrix,cs', [('bt709', fforacle.SWS_CS_ITU709), ('bt2020', fforacle.SWS_CS_BT2020)])
def test_csc_close_to_swscale(matrix, cs):
    h, w = 64, 96
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 2 + yy) % 256, (yy * 3) % 256, (xx + 2 * yy) % 256], -1).astype(np.uint8)   # smooth ramps
    y, cb, cr = pixel_ref.rgb_to_yuv420(img, matrix, 8)
    sy, su, sv = fforacle.sws_convert([img.reshape(h, w * 3)], fforacle.AV_PIX_FMT_RGB24, fforacle.AV_PIX_FMT_YUV420P, w, h,
                                      [(h, w, np.uint8), (h // 2, w // 2, np.uint8), (h // 2, w // 2, np.uint8)],
                                      algo=fforacle.SWS_BILINEAR, colorspace=cs)
    assert np.abs(y.astype(int) - sy).max() <= 1
    assert np.abs(cb.astype(int) - su).max() <= 2 and np.abs(cr.astype(int) - sv).max() <= 2
    assert y.min() >= 16 and y.max() <= 235
    # 10-bit output is the same quantity at 4x resolution
    y10, _, _ = pixel_ref.rgb_to_yuv420(img, matrix, 10)
    assert np.abs(y10.astype(int) - 4 * y.astype(int)).max() <= 3


def test_scaler_properties():
    tab = pixel_ref.bicubic_table()
    assert (tab.sum(1) == 16384).all() and tab[0].tolist() == [0, 16384, 0, 0]
    flat = np.full((20, 30), 77, np.uint8)
    assert (pixel_ref.scale_plane(flat, 60, 40) == 77).all()
    assert (pixel_ref.scale_plane(flat, 60, 40, 10) == 77 * 4).all()
    src = RNG.integers(0, 256, (24, 36)).astype(np.uint8)
    assert (pixel_ref.scale_plane(src, 36, 24) == src).all()          # identity ratio is exact
    import cv2
    yy, xx = np.mgrid[0:54, 0:96]
    smooth = ((np.sin(xx / 9.0) + np.cos(yy / 7.0)) * 60 + 128).astype(np.uint8)
    mine = pixel_ref.scale_plane(smooth, 192, 108).astype(int)
    ref = cv2.resize(smooth, (192, 108), interpolation=cv2.INTER_CUBIC).astype(int)
    assert np.abs(mine - ref)[4:-4, 4:-4].max() <= 3       # same kernel family, different fixed point
    assert pixel_ref.upscale_geometry(1920, 1080) == (3840, 2160)
    assert pixel_ref.upscale_geometry(1280, 720) == (1920, 1080)
    assert pixel_ref.upscale_geometry(3840, 2160) == (3840, 2160)
