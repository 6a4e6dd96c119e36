"""GPU parity (BASELINE config 5): batched primitives through the C ABI vs oracle/primitives.c, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PU_SIZES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64),
            (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32), (64, 48), (48, 64), (64, 16), (16, 64)]


@pytest.fixture(scope='module')
def env():
    import torch
    from hevc_b200 import _cabi
    ctx = _cabi.Context(0)
    yield ctx, torch
    ctx.close()


def _g16(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).cuda()


def _blocks(rng, n, h, w, depth):
    a = rng.integers(0, 1 << depth, (n, h, w)).astype(np.uint16)
    b = rng.integers(0, 1 << depth, (n, h, w)).astype(np.uint16)
    a[0], b[0] = 0, (1 << depth) - 1                    # extremes: overflow check
    a[1] = b[1]
    a[2, ::2], a[2, 1::2], b[2] = 0, (1 << depth) - 1, (1 << depth) - 1
    return a, b


@pytest.mark.parametrize('w,h', PU_SIZES)
def test_sad_satd_all_pu_sizes(env, w, h):
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    rng = np.random.default_rng(w * 100 + h)
    for depth in (8, 10):
        n = 67
        a, b = _blocks(rng, n, h, w, depth)
        ga, gb = _g16(torch, a), _g16(torch, b)
        sad = ops.sad(ctx, ga, gb)
        satd = ops.satd(ctx, ga, gb)
        ctx.sync()
        assert (sad.cpu().numpy() == cmodel.sad(a, b)).all()
        assert (satd.cpu().numpy() == cmodel.satd(a, b)).all()


@pytest.mark.parametrize('w,h', [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (16, 12), (12, 16), (64, 48), (8, 32)])
def test_sad_x3_x4(env, w, h):
    """x265 sad_x3 / sad_x4: one source block against three / four references = the single SADs"""
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    rng = np.random.default_rng(w * 31 + h)
    n = 53
    a, _ = _blocks(rng, n, h, w, 10)
    refs = [rng.integers(0, 1024, (n, h, w)).astype(np.uint16) for _ in range(4)]
    ga, gr = _g16(torch, a), [_g16(torch, r) for r in refs]
    for k in (3, 4):
        got = ops.sad_multi(ctx, ga, gr[:k])
        ctx.sync()
        want = np.stack([cmodel.sad(a, r) for r in refs[:k]], axis=1)
        assert (got.cpu().numpy() == want).all()


@pytest.mark.parametrize('size', [4, 8, 16, 32, 64])
def test_sa8d(env, size):
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    rng = np.random.default_rng(size)
    a, b = _blocks(rng, 45, size, size, 10)
    out = ops.sa8d(ctx, _g16(torch, a), _g16(torch, b))
    ctx.sync()
    assert (out.cpu().numpy() == cmodel.sa8d(a, b)).all()


@pytest.mark.parametrize('size', [4, 8, 16, 32])
@pytest.mark.parametrize('depth', [8, 10])
def test_transforms(env, size, depth):
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    rng = np.random.default_rng(size + depth)
    lim = (1 << depth) - 1
    res = rng.integers(-lim, lim + 1, (53, size, size)).astype(np.int16)
    res[0], res[1], res[2] = lim, -lim, 0
    res[3, ::2], res[3, 1::2] = lim, -lim
    for is_dst in ((False, True) if size == 4 else (False,)):
        want = cmodel.fwd_transform(res, depth, is_dst)
        got = ops.fwd_transform(ctx, torch.from_numpy(res).cuda(), depth, is_dst)
        ctx.sync()
        assert (got.cpu().numpy() == want).all()
        coef = rng.integers(-32768, 32768, (53, size, size)).astype(np.int16)
        coef[0], coef[1] = 32767, -32768
        coef[5:] = want[5:]
        want_i = cmodel.inv_transform(coef, depth, is_dst)
        got_i = ops.inv_transform(ctx, torch.from_numpy(coef).cuda(), depth, is_dst)
        ctx.sync()
        assert (got_i.cpu().numpy() == want_i).all()


@pytest.mark.parametrize('size', [4, 8, 16, 32])
def test_quant_dequant_all_qp(env, size):
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    rng = np.random.default_rng(size)
    coef = rng.integers(-32768, 32768, (9, size, size)).astype(np.int16)
    coef[0], coef[1], coef[2] = 32767, -32768, 0
    gc = torch.from_numpy(coef).cuda()
    for depth in (8, 10):
        for qp in range(0, 52 + 6 * (depth - 8)):
            for intra in (True, False):
                lvl, ns = ops.quant(ctx, gc, qp, depth, intra)
                ctx.sync()
                wl, wn = cmodel.quant(coef, qp, depth, intra)
                assert (lvl.cpu().numpy() == wl).all() and (ns.cpu().numpy() == wn).all(), (depth, qp, intra)
            deq = ops.dequant(ctx, lvl, qp, depth)
            ctx.sync()
            assert (deq.cpu().numpy() == cmodel.dequant(wl, qp, depth)).all(), (depth, qp)


@pytest.mark.parametrize('size', [4, 8, 16, 32])
@pytest.mark.parametrize('is_luma', [True, False])
def test_intra_all_modes(env, size, is_luma):
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    rng = np.random.default_rng(size * 2 + is_luma)
    for depth in (8, 10):
        nb = rng.integers(0, 1 << depth, (21, 4 * size + 1)).astype(np.uint16)
        nb[0] = (1 << depth) - 1
        nb[1] = 0
        nb[2] = np.linspace(100, 120, 4 * size + 1).astype(np.uint16)      # smooth: triggers 32x32 strong smoothing
        nb[3] = np.linspace(60 << (depth - 8), 64 << (depth - 8), 4 * size + 1).astype(np.uint16)
        for strong in (False, True):
            got = ops.intra_pred_all(ctx, _g16(torch, nb), size, is_luma, strong, depth)
            ctx.sync()
            want = cmodel.intra_pred_all(nb, size, is_luma, strong, depth)
            g = got.cpu().numpy().view(np.uint16)
            bad = np.argwhere(g != want)
            assert bad.size == 0, (depth, strong, bad[:4])


def test_large_batch_checksum(env):
    """2^20 blocks per launch (config 5 size): compare a checksum of all outputs with the oracle on a strided sample."""
    from hevc_b200 import ops
    from oracle import cmodel
    ctx, torch = env
    n = 1 << 20
    g = torch.Generator(device='cuda').manual_seed(1)
    a = torch.randint(0, 1024, (n, 8, 8), device='cuda', generator=g, dtype=torch.int16)
    b = torch.randint(0, 1024, (n, 8, 8), device='cuda', generator=g, dtype=torch.int16)
    torch.cuda.synchronize()          # inputs were produced on torch's stream; the context has its own
    satd = ops.satd(ctx, a, b)
    sad = ops.sad(ctx, a, b)
    ctx.sync()
    idx = torch.arange(0, n, 4099, device='cuda')
    sa, sb = a[idx].cpu().numpy().view(np.uint16), b[idx].cpu().numpy().view(np.uint16)
    assert (satd[idx].cpu().numpy() == cmodel.satd(sa, sb)).all()
    assert (sad[idx].cpu().numpy() == cmodel.sad(sa, sb)).all()
    assert int(sad.sum()) == int((a.int() - b.int()).abs().sum())      # SAD total is linear: checksum over all 2^20 blocks
