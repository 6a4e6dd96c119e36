"""GPU parity of the stream encoder, through the C ABI:
  1. every decision, level, reconstructed sample and output byte equals the CPU model (oracle/hevc_encode.c);
  2. every stream decodes under the FFmpeg hevc decoder to exactly the encoder's own reconstruction
     (decoded-picture-hash verification on)."""
import numpy as np
import pytest

from tests import enc_common as ec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from hevc_b200 import _cabi
    c = _cabi.Context(0)
    yield c
    c.close()


def _gpu_encode(ctx, p, frames8, qp, batch, hash_sei=True, split=None, read_back=True, rate_control=False, deblock=True, **tools):
    from hevc_b200 import encoder as E
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=qp, hash_sei=hash_sei, keep_recon=True, rate_control=rate_control, deblock=deblock, **tools),
                        max_batch=batch)
    stream, recs, decs = b'', [], []
    _gpu_encode.last_qps = []
    _gpu_encode.last_stats = []
    pos = 0
    for n in (split or [len(frames8)]):
        out, stats = enc.encode(E.pack_yuv420p8(frames8[pos:pos + n]), n)
        stream += out
        _gpu_encode.last_qps = getattr(_gpu_encode, 'last_qps', []) + [s.qp for s in stats]
        _gpu_encode.last_stats += [(int(s.is_idr), s.poc) for s in stats]
        if read_back and n <= batch:          # the encoder keeps the reconstructions of its last batch
            for i in range(n):
                recs.append(enc.read_recon(i))
                decs.append(enc.read_decisions(i))
        pos += n
    enc.close()
    return stream, recs, decs


CASES = [(192, 112, 8, 6, (26, 28)), (200, 120, 10, 5, (22, 24)), (64, 64, 8, 3, (40, 42)), (416, 240, 8, 5, (10, 12)),
         (32, 16, 10, 3, (30, 30)), (328, 184, 10, 6, (35, 37)),
         # geometry edge cases of the CABAC staging ring / wavefronts: one CU, odd CU counts, exactly three CTUs in one row
         (16, 16, 8, 3, (30, 32)), (80, 48, 8, 4, (24, 26)), (96, 32, 10, 4, (20, 22))]


@pytest.mark.parametrize('w,h,depth,n,qp', CASES)
def test_matches_cpu_model_bit_for_bit(ctx, w, h, depth, n, qp):
    p = ec.b200_params(w, h, depth, keyint=4)
    frames = ec.clip_frames(w, h, n)
    m_stream, m_aus, m_recs, m_decs = ec.run_model(p, frames, qp[0], qp[1], hash_sei=True)
    g_stream, g_recs, g_decs = _gpu_encode(ctx, p, frames, qp, batch=8)
    for i in range(n):
        mc, ml = m_decs[i]
        gc, gl = g_decs[i]
        for field in ('pred_mode', 'intra_mode', 'mvx', 'mvy', 'cbf'):
            bad = np.argwhere(mc[field] != gc[field])
            assert bad.size == 0, f'frame {i} {field} differs first at CU {bad[0].tolist()}: model {mc[field][tuple(bad[0])]} gpu {gc[field][tuple(bad[0])]}'
        coded = (mc['cbf'].reshape(-1) != 0)
        assert (ml[coded] == gl[coded]).all(), f'frame {i} levels differ'
        for c in range(3):
            assert (m_recs[i][c] == g_recs[i][c]).all(), f'frame {i} plane {c} reconstruction differs'
    assert g_stream == m_stream


@pytest.mark.parametrize('w,h,depth,n,qp', CASES[:4])
def test_decoder_reproduces_encoder_reconstruction(ctx, w, h, depth, n, qp):
    from oracle import fforacle
    p = ec.b200_params(w, h, depth, keyint=4)
    frames = ec.clip_frames(w, h, n, seed=3)
    stream, recs, _ = _gpu_encode(ctx, p, frames, qp, batch=8)
    decoded = fforacle.decode_hevc(stream, verify_hash=True)
    assert len(decoded) == n
    for dec, rec in zip(decoded, recs):
        for c in range(3):
            assert (dec[c] == rec[c][:dec[c].shape[0], :dec[c].shape[1]]).all()


def test_batches_and_calls_do_not_change_the_stream(ctx):
    """the same 10 frames through one call, through batches of 3, and through three calls give identical bytes"""
    w, h = 192, 112
    p = ec.b200_params(w, h, 8, keyint=5)
    frames = ec.clip_frames(w, h, 10, seed=5)
    a, _, _ = _gpu_encode(ctx, p, frames, (27, 29), batch=16, hash_sei=False, read_back=False)
    b, _, _ = _gpu_encode(ctx, p, frames, (27, 29), batch=3, hash_sei=False, read_back=False)
    c, _, _ = _gpu_encode(ctx, p, frames, (27, 29), batch=4, hash_sei=False, split=[4, 1, 5], read_back=False)
    assert a == b == c
    m_stream, _, _, _ = ec.run_model(p, frames, 27, 29, hash_sei=False)
    assert a == m_stream


def test_1080p_frame_pair(ctx):
    """full-size geometry (1080 is not a multiple of 16: conformance window, partial last CTU row)"""
    from oracle import fforacle
    w, h = 1920, 1080
    p = ec.b200_params(w, h, 8, keyint=30)
    frames = ec.clip_frames(w, h, 2, seed=2)
    stream, recs, _ = _gpu_encode(ctx, p, frames, (24, 26), batch=2)
    decoded = fforacle.decode_hevc(stream, verify_hash=True)
    assert len(decoded) == 2 and decoded[0][0].shape == (1080, 1920)
    for dec, rec in zip(decoded, recs):
        for c in range(3):
            assert (dec[c] == rec[c][:dec[c].shape[0], :dec[c].shape[1]]).all()
    m_stream, _, _, _ = ec.run_model(p, frames, 24, 26, hash_sei=True)
    assert stream == m_stream


@pytest.mark.parametrize('depth,kbps', [(8, 300), (10, 250), (8, 50000)])
def test_rate_control_matches_cpu_model(ctx, depth, kbps):
    """VBV-constrained mode: the QP trajectory is decided on the device from size estimates and must equal the model's,
    including the two-pass first key frame, across batch boundaries and calls"""
    from oracle import fforacle
    w, h, n = 320, 192, 14
    p = ec.b200_params(w, h, depth, keyint=6)
    p.vbv_maxrate_kbps, p.vbv_bufsize_kbit = kbps, int(kbps * 1.2)
    frames = ec.clip_frames(w, h, n, seed=7)
    m_stream, _, m_recs, _ = ec.run_model(p, frames, 20, 22, hash_sei=True, rate_control=True)
    m_qps = list(ec.run_model.last_qps)
    g_stream, g_recs, _ = _gpu_encode(ctx, p, frames, (20, 22), batch=4, rate_control=True, split=[5, 9], read_back=False)
    assert _gpu_encode.last_qps == m_qps
    if kbps < 1000:
        assert max(m_qps) > 24            # the cap really binds in these cases
    else:
        assert set(m_qps) == {20, 22, 24, 26}          # the quality ceiling: key frame, anchor P frames, the cascade between them
    assert g_stream == m_stream
    assert len(fforacle.decode_hevc(g_stream, verify_hash=True)) == n


def test_4k60_hdr10_full_size(ctx):
    """BASELINE configs[1] at full size (3840x2160 Main10, level 5.1, HDR10 SEI, VBV rate control): the stream decodes
    under the FFmpeg decoder with picture-hash verification on, to exactly the encoder's reconstruction; the first
    access unit equals the CPU model's byte for byte"""
    from hevc_b200 import derive, encoder as E
    from hevc_b200.probe import VideoInfo
    from hevc_b200.synth import TorchSynthClip
    from oracle import fforacle
    info = VideoInfo(3840, 2160, 60.0, 'bt2020', 'smpte2084', 'bt2020nc', 'yuv420p', '', '', 0, True, None, None, 5.0)
    p = derive.derive_b200_params(info)
    assert (p.level_idc, p.keyint, p.vbv_maxrate_kbps, p.crf) == (153, 120, 23520, 19)
    clip = TorchSynthClip(3840, 2160, seed=11, device='cuda:0')
    n = 4
    frames = clip.frames(0, n).cpu().numpy()
    enc = E.B200Encoder(ctx, E.to_c_params(p, hash_sei=True, keep_recon=True), max_batch=n)
    stream, stats = enc.encode(frames, n)
    recs = [enc.read_recon(i) for i in range(n)]
    enc.close()
    dec = fforacle.decode_hevc(stream, verify_hash=True)
    assert len(dec) == n and dec[0][0].shape == (2160, 3840) and dec[0][0].dtype == np.uint16
    for d, r in zip(dec, recs):
        for c in range(3):
            assert (d[c] == r[c][:d[c].shape[0], :d[c].shape[1]]).all()
    types = [t for _, t, _ in fforacle.iter_nals(stream)]
    assert types[:4] == [35, 32, 33, 34] and types.count(19) == 1 and types.count(1) == n - 1
    # first access unit (two-pass rate-controlled IDR) against the CPU model
    lw, cw = 3840 * 2160, 1920 * 1080
    f0 = frames[0]
    qi, qp = E.crf_to_qp(p.crf)
    m_stream, m_aus, _, _ = ec.run_model(p, [(f0[:lw].reshape(2160, 3840), f0[lw:lw + cw].reshape(1080, 1920), f0[lw + cw:].reshape(1080, 1920))],
                                         qi, qp, hash_sei=True, rate_control=True)
    assert stream[:stats[0].bytes] == m_aus[0]


def test_deblocking_off_also_matches(ctx):
    w, h, n = 200, 120, 5
    p = ec.b200_params(w, h, 8, keyint=4)
    frames = ec.clip_frames(w, h, n, seed=9)
    m_stream, _, m_recs, _ = ec.run_model(p, frames, 36, 38, hash_sei=True, deblock=False)
    g_stream, g_recs, _ = _gpu_encode(ctx, p, frames, (36, 38), batch=8, deblock=False)
    assert g_stream == m_stream
    on_stream, _, _ = _gpu_encode(ctx, p, frames, (36, 38), batch=8, deblock=True)
    assert on_stream != g_stream


def test_closed_gop_segments_concatenate(ctx):
    """config 4 in miniature: segments encoded by independent encoders (two worker threads on device 0) concatenate into
    one stream that the decoder plays through, IDR at every segment start"""
    from hevc_b200 import batch, encoder as E
    from oracle import fforacle
    w, h, n = 192, 112, 14
    p = ec.b200_params(w, h, 8, keyint=4)
    frames = ec.clip_frames(w, h, n, seed=12)
    buf = E.pack_yuv420p8(frames).reshape(n, -1)
    stream = batch.encode_clip_segmented(buf, n, p, devices=[0, 0], c_params_kwargs={'hash_sei': True, 'qp': (28, 30), 'rate_control': False})
    dec = fforacle.decode_hevc(stream, verify_hash=True)
    assert len(dec) == n
    types = [t for _, t, _ in fforacle.iter_nals(stream)]
    assert types.count(19) == 4 and types.count(33) == 4          # 4 segments, each with its own IDR + SPS
    # identical to one encoder running through the clip with the same GOP cadence (constant QP: no state crosses a GOP)
    whole, _, _ = _gpu_encode(ctx, p, frames, (28, 30), batch=16, read_back=False)
    assert [len(d[0]) for d in fforacle.decode_hevc(whole, verify_hash=True)] == [len(d[0]) for d in dec]
    for a, b in zip(fforacle.decode_hevc(whole, verify_hash=False), dec):
        assert all((a[c] == b[c]).all() for c in range(3))


def test_delayed_pipeline_matches_synchronous(ctx):
    """hb_enc_encode_delayed returns the access units one batch late (the next batch's frame chain overlaps the previous
    batch's CABAC tail and host assembly); the bytes are those of the synchronous call."""
    from hevc_b200 import encoder as E
    w, h, n = 192, 112, 23
    p = ec.b200_params(w, h, 8, keyint=6)
    frames = ec.clip_frames(w, h, n, seed=5)
    ref_stream, _, _ = _gpu_encode(ctx, p, frames, (28, 30), batch=4, hash_sei=False, read_back=False, rate_control=True)
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=(28, 30), hash_sei=False, keep_recon=False, rate_control=True), max_batch=4)
    out, counts, pos = b'', [], 0
    for k in (4, 7, 1, 8, 3):                         # batches of mixed size, some calls spanning several batches
        chunk, stats = enc.encode_delayed(E.pack_yuv420p8(frames[pos:pos + k]), k)
        out += chunk
        counts.append(len(stats))
        pos += k
    chunk, stats = enc.flush()
    out += chunk
    counts.append(len(stats))
    enc.close()
    assert pos == n and sum(counts) == n
    assert counts[0] == 0 and counts[-1] > 0          # the first call has nothing to return yet, the flush returns the rest
    assert out == ref_stream


def test_empty_call_stop_and_argument_errors(ctx):
    """Edge cases of the C ABI: zero frames is a no-op, a flush with nothing in flight returns nothing, request_stop makes the
    next call fail with HB_ERR_STOPPED instead of encoding, bad geometry is rejected at creation."""
    from hevc_b200 import _cabi
    from hevc_b200 import encoder as E
    p = ec.b200_params(64, 64, 8, keyint=4)
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=(30, 32), hash_sei=False, keep_recon=False, rate_control=False), max_batch=4)
    out, stats = enc.encode(np.zeros(1, np.uint8), 0)
    assert out == b'' and stats == []
    out, stats = enc.flush()
    assert out == b'' and stats == []
    frames = ec.clip_frames(64, 64, 2)
    first, _ = enc.encode(E.pack_yuv420p8(frames), 2)
    assert first[:5] == b'\x00\x00\x00\x01\x40'            # VPS first: the stream starts with its parameter sets
    enc.request_stop()
    with pytest.raises(_cabi.HbError):
        enc.encode(E.pack_yuv420p8(frames), 2)
    enc.close()
    bad = ec.b200_params(63, 64, 8, keyint=4)
    with pytest.raises(_cabi.HbError):
        E.B200Encoder(ctx, E.to_c_params(bad, qp=(30, 32)), max_batch=2)


def test_8k_maximum_size(ctx):
    """Largest supported geometry class (7680x4320, 135 CTU rows, 240 CTUs per row): IDR + P through both batch sets; the
    stream decodes under the FFmpeg decoder (picture hash on) to the encoder's reconstruction."""
    from hevc_b200 import encoder as E
    from hevc_b200.synth import TorchSynthClip
    from oracle import fforacle
    w, h, n = 7680, 4320, 2
    p = ec.b200_params(w, h, 8, keyint=4)
    p.level_idc = 183
    clip = TorchSynthClip(w, h, seed=3, device='cuda:0')
    frames = clip.frames(0, n).cpu().numpy()
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=(30, 32), hash_sei=True, keep_recon=True, rate_control=False), max_batch=1)
    stream, stats = enc.encode(frames, n)                 # two batches of one frame: exercises the set hand-over at full width
    rec1 = enc.read_recon(0)                              # the last drained batch holds frame 1
    enc.close()
    assert [s.is_idr for s in stats] == [True, False]
    dec = fforacle.decode_hevc(stream, verify_hash=True)
    assert len(dec) == n and dec[1][0].shape == (h, w)
    for c in range(3):
        assert (dec[1][c] == rec1[c][:dec[1][c].shape[0], :dec[1][c].shape[1]]).all()


def test_parallel_segment_streams_are_deterministic_and_decodable(ctx):
    """ParallelSegmentEncoder: two encoder streams on one GPU fed with closed-GOP segments round-robin; the threaded, pipelined
    result equals the same segment-to-encoder assignment run sequentially, and the concatenation decodes."""
    from hevc_b200 import encoder as E
    from oracle import fforacle
    w, h, seg, nseg = 192, 112, 4, 5
    p = ec.b200_params(w, h, 8, keyint=seg)
    cp = E.to_c_params(p, qp=(28, 30), hash_sei=True, keep_recon=False, rate_control=True)
    frames = ec.clip_frames(w, h, seg * nseg, seed=9)
    packs = [E.pack_yuv420p8(frames[i * seg:(i + 1) * seg]) for i in range(nseg)]
    pse = E.ParallelSegmentEncoder(0, cp, streams=2, max_batch=seg)
    got, n_stats = b'', 0
    for pk in packs:
        out, stats = pse.submit(pk, seg)
        got += out
        n_stats += len(stats)
    out, stats = pse.finish()
    got += out
    n_stats += len(stats)
    pse.close()
    assert n_stats == seg * nseg
    encs = [E.B200Encoder(ctx, cp, max_batch=seg) for _ in range(2)]
    want = b''.join(encs[i % 2].encode(pk, seg, force_idr=True)[0] for i, pk in enumerate(packs))
    for e in encs:
        e.close()
    assert got == want
    dec = fforacle.decode_hevc(got, verify_hash=True)
    assert len(dec) == seg * nseg


@pytest.mark.parametrize('fmt,enc_depth', [('p16_10', 10), ('p010', 10), ('p16_10', 8), ('p16_12', 10)])
def test_source_formats_match_cpu_model(ctx, fmt, enc_depth):
    """encoder-level parity for the non-8-bit source formats: HB_PIX_YUV420P16 carrying 10- or 12-bit samples (src_bit_depth)
    and HB_PIX_P010; the ingest stage's depth conversion ((v + round) >> shift, clamp) feeds the same samples the model gets"""
    from hevc_b200 import encoder as E
    w, h, n, qp = 176, 112, 4, (24, 26)
    p = ec.b200_params(w, h, enc_depth, keyint=4, hdr10=False)
    rng = np.random.default_rng(11)
    src_depth = 12 if fmt == 'p16_12' else 10
    base = ec.clip_frames(w, h, n, seed=9)
    deep = [tuple(np.minimum((pl.astype(np.uint16) << (src_depth - 8)) + rng.integers(0, 1 << (src_depth - 8), pl.shape, dtype=np.uint16),
                             (1 << src_depth) - 1).astype(np.uint16) for pl in f) for f in base]
    if fmt == 'p010':
        def pack(f):
            y, u, v = f
            uv = np.empty((h // 2, w), np.uint16)
            uv[:, 0::2], uv[:, 1::2] = u << 6, v << 6
            return np.concatenate([(y << 6).reshape(-1), uv.reshape(-1)])
        data = np.concatenate([pack(f) for f in deep]).view(np.uint8)
        pix, kw = E.PIX_P010, {}
        want_in = deep
    else:
        data = np.concatenate([np.concatenate([pl.reshape(-1) for pl in f]) for f in deep]).view(np.uint8)
        pix, kw = E.PIX_YUV420P16, {'src_bit_depth': src_depth}
        sh = src_depth - enc_depth
        want_in = [tuple(np.minimum((pl.astype(np.int32) + (1 << (sh - 1))) >> sh, (1 << enc_depth) - 1).astype(np.uint16) if sh > 0 else pl
                         for pl in f) for f in deep]
    from oracle import encoder_model as em
    model = em.ModelEncoder(ec.model_params(p, qp[0], qp[1], True))
    want = b''.join(model.encode(*f)[0] for f in want_in)
    model.close()
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=qp, hash_sei=True, keep_recon=True, rate_control=False), max_batch=4)
    got, _ = enc.encode(data, n, fmt=pix, **kw)
    enc.close()
    assert got == want


def test_rgb_and_scaled_ingest_feed_the_pixel_oracle_samples(ctx):
    """fused ingest: BGR24 frames (matrix conversion) and a scaled 8-bit 4:2:0 source must give the encoder exactly the
    samples oracle/pixel_ref.py defines -- checked by encoding the oracle's planes through the CPU model"""
    from hevc_b200 import encoder as E
    from oracle import encoder_model as em
    from oracle import pixel_ref
    rng = np.random.default_rng(5)
    qp = (22, 24)
    # ---- BGR24 -> Main10
    w, h, n = 144, 80, 3
    p = ec.b200_params(w, h, 10, keyint=4, hdr10=False)
    imgs = [np.ascontiguousarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for _ in range(n)]
    model = em.ModelEncoder(ec.model_params(p, qp[0], qp[1], True))
    want = b''.join(model.encode(*pixel_ref.rgb_to_yuv420(im, 'bt709', 10, bgr=True))[0] for im in imgs)
    model.close()
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=qp, hash_sei=True, rate_control=False), max_batch=4)
    got, _ = enc.encode(np.stack(imgs).reshape(n, -1), n, fmt=E.PIX_BGR24)
    enc.close()
    assert got == want
    # ---- 8-bit 4:2:0 at 96x64 -> 192x128 Main10 and -> 8-bit Main (both scaler output depths), odd ratio 1.5 as well
    for (sw, sh_, dw, dh, depth) in [(96, 64, 192, 128, 10), (96, 64, 144, 96, 8), (320, 180, 640, 360, 10)]:
        p = ec.b200_params(dw, dh, depth, keyint=4, hdr10=False)
        frames = ec.clip_frames(sw, sh_, n, seed=13)
        model = em.ModelEncoder(ec.model_params(p, qp[0], qp[1], True))
        want = b''
        for y, u, v in frames:
            want += model.encode(pixel_ref.scale_plane(y, dw, dh, depth), pixel_ref.scale_plane(u, dw // 2, dh // 2, depth),
                                 pixel_ref.scale_plane(v, dw // 2, dh // 2, depth))[0]
        model.close()
        enc = E.B200Encoder(ctx, E.to_c_params(p, qp=qp, hash_sei=True, rate_control=False), max_batch=4)
        got, _ = enc.encode(E.pack_yuv420p8(frames), n, fmt=E.PIX_YUV420P8, src_size=(sw, sh_))
        enc.close()
        assert got == want, (sw, sh_, dw, dh, depth)


def test_4k60_full_gop_matches_cpu_model_and_vbv(ctx):
    """BASELINE configs[1], one whole key-frame interval plus the start of the next (122 frames, crf 19, VBV 23520 / 28224 on):
    every byte equals the CPU model's (the model runs its CU loops on all host cores), and the ACTUAL access-unit sizes of
    240 frames (two GOPs) keep the signalled HRD buffer from underflowing (hrd=1, reference core/utils.py:65)."""
    from hevc_b200 import derive, encoder as E
    from hevc_b200.probe import VideoInfo
    from hevc_b200.synth import TorchSynthClip
    from tests.test_oracle_encoder import vbv_underflows
    info = VideoInfo(3840, 2160, 60.0, 'bt2020', 'smpte2084', 'bt2020nc', 'yuv420p', '', '', 0, True, None, None, 5.0)
    p = derive.derive_b200_params(info)
    clip = TorchSynthClip(3840, 2160, seed=5, device='cuda:0')
    n_model, n = 122, 240
    enc = E.B200Encoder(ctx, E.to_c_params(p, hash_sei=False), max_batch=120)
    stream, stats = b'', []
    for s in range(0, n, 60):
        fr = clip.frames(s, 60).cpu().numpy()
        out, st = enc.encode_delayed(fr, 60)
        stream += out
        stats += st
    out, st = enc.flush()
    stream += out
    stats += st
    enc.close()
    assert len(stats) == n and [i for i, s in enumerate(stats) if s.is_idr] == [0, 120]
    assert vbv_underflows([8 * s.bytes for s in stats], p.vbv_maxrate_kbps, p.vbv_bufsize_kbit, 60.0) == []
    lw, cw = 3840 * 2160, 1920 * 1080
    qi, qp = E.crf_to_qp(p.crf)
    from oracle import encoder_model as em
    model = em.ModelEncoder(ec.model_params(p, qi, qp, False, rate_control=True))
    pos = 0
    for i in range(n_model):
        f = clip.frames(i, 1).cpu().numpy()[0]
        au, minfo = model.encode(*(a.astype(np.uint16) << 2 for a in (f[:lw].reshape(2160, 3840), f[lw:lw + cw].reshape(1080, 1920),
                                                                      f[lw + cw:].reshape(1080, 1920))))
        assert stream[pos:pos + stats[i].bytes] == au, f'access unit {i} differs (model qp {minfo.qp}, gpu qp {stats[i].qp})'
        pos += stats[i].bytes
    model.close()


def hard_cut_frames(w, h, n, cut, seed=21, noise=2.0):
    """two unrelated synthetic scenes joined at frame `cut`"""
    from hevc_b200.synth import SynthClip
    a, b = SynthClip(w, h, seed=seed, noise=noise), SynthClip(w, h, seed=seed + 100, noise=noise)
    return [a.frame(i) if i < cut else tuple(np.ascontiguousarray(p[::-1, ::-1]) for p in b.frame(i + 37)) for i in range(n)]


@pytest.mark.parametrize('depth,scenecut', [(8, True), (10, True), (8, False), (10, False)])
def test_hard_cut_clip_matches_cpu_model(ctx, depth, scenecut):
    """a clip with a scene change in the middle of a GOP.  scenecut on: the device-side detector makes the frame after the cut
    a key frame (it is past min-keyint) and the key-frame cadence restarts there; scenecut off: the frame stays a P frame and
    most of its CUs go intra.  Either way: same bytes as the model, same frame types, decodable, and the picture after the
    cut does not fall apart."""
    from oracle import fforacle
    w, h, n, cut = 320, 192, 16, 9
    p = ec.b200_params(w, h, depth, keyint=12)           # min-keyint 6
    frames = hard_cut_frames(w, h, n, cut)
    m_stream, m_aus, m_recs, _ = ec.run_model(p, frames, 26, 28, hash_sei=True, scenecut=scenecut)
    infos = ec.run_model.last_infos
    g_stream, g_recs, _ = _gpu_encode(ctx, p, frames, (26, 28), batch=8, split=[8, 8], scenecut=scenecut)
    assert _gpu_encode.last_stats == [(i, pc) for i, pc, _ in infos]
    if scenecut:
        assert [k for k, (i, _, _) in enumerate(infos) if i] == [0, cut]              # no key frame at 12: the cadence restarted
    else:
        assert [k for k, (i, _, _) in enumerate(infos) if i] == [0, 12] and infos[cut][2] > (w // 16) * (h // 16) // 3
    assert g_stream == m_stream
    dec = fforacle.decode_hevc(g_stream, verify_hash=True)
    assert len(dec) == n
    sh = depth - 8
    src = frames[cut][0].astype(np.float64) * (1 << sh)
    mse = np.mean((dec[cut][0].astype(np.float64) - src) ** 2)
    assert 10 * np.log10(((255 << sh) ** 2) / mse) > 30


def test_intra_cus_in_p_frames_match_cpu_model(ctx):
    """uncovered background / fast content: P frames carry intra CUs (sparse wavefront pass after the inter kernel); every
    decision, level, sample and byte equals the model, with the tool on and off"""
    from hevc_b200.synth import content_clip
    w, h, n = 352, 208, 7
    p = ec.b200_params(w, h, 8, keyint=30)
    frames = content_clip('pan', w, h, n, seed=3)
    for on in (True, False):
        m_stream, _, m_recs, m_decs = ec.run_model(p, frames, 30, 32, hash_sei=True, intra_in_p=on, scenecut=False)
        n_intra = sum(x[2] for x in ec.run_model.last_infos[1:])
        assert (n_intra > 0) == on
        g_stream, g_recs, g_decs = _gpu_encode(ctx, p, frames, (30, 32), batch=8, intra_in_p=on, scenecut=False)
        for i in range(n):
            for field in ('pred_mode', 'intra_mode', 'mvx', 'mvy', 'cbf'):
                assert (m_decs[i][0][field] == g_decs[i][0][field]).all(), (on, i, field)
            for c in range(3):
                assert (m_recs[i][c] == g_recs[i][c]).all(), (on, i, c)
        assert g_stream == m_stream


@pytest.mark.parametrize('w,h,depth,n,qp', [(192, 112, 8, 6, (30, 32)), (200, 120, 10, 5, (22, 24)), (416, 240, 8, 5, (12, 14)),
                                            (328, 184, 10, 6, (38, 40)), (16, 16, 8, 3, (30, 32)), (96, 32, 10, 4, (26, 28))])
def test_sao_matches_cpu_model(ctx, w, h, depth, n, qp):
    """sample adaptive offset: per-CTU statistics, parameter decision (edge classes + band offset, luma and chroma), merge
    signalling and application equal the model's -- parameters, reconstruction (= the next frame's reference) and bytes; the
    FFmpeg decoder reproduces the reconstruction (picture hash on)"""
    from oracle import fforacle
    p = ec.b200_params(w, h, depth, keyint=4)
    frames = ec.clip_frames(w, h, n, seed=4)
    m_stream, _, m_recs, m_decs = ec.run_model(p, frames, qp[0], qp[1], hash_sei=True, sao=True)
    g_stream, g_recs, g_decs = _gpu_encode(ctx, p, frames, qp, batch=8, sao=True)
    for i in range(n):
        for c in range(3):
            assert (m_recs[i][c] == g_recs[i][c]).all(), f'frame {i} plane {c} reconstruction differs'
    assert g_stream == m_stream
    off_stream, _, _ = _gpu_encode(ctx, p, frames, qp, batch=8, sao=False, read_back=False)
    assert off_stream != g_stream
    dec = fforacle.decode_hevc(g_stream, verify_hash=True)
    assert len(dec) == n


def test_reset_starts_an_independent_stream(ctx):
    """hb_enc_reset: a reused encoder (the batch worker's pool) produces exactly the bytes a fresh encoder would, including the
    rate-control trajectory and the two-pass first key frame"""
    from hevc_b200 import encoder as E
    w, h = 192, 112
    p = ec.b200_params(w, h, 8, keyint=5)
    p.vbv_maxrate_kbps, p.vbv_bufsize_kbit = 300, 360
    a, b = ec.clip_frames(w, h, 7, seed=31), ec.clip_frames(w, h, 6, seed=32)
    cp = E.to_c_params(p, qp=(22, 24), hash_sei=False, rate_control=True)
    fresh = []
    for frames in (a, b):
        enc = E.B200Encoder(ctx, cp, max_batch=4)
        fresh.append(enc.encode(E.pack_yuv420p8(frames), len(frames))[0])
        enc.close()
    enc = E.B200Encoder(ctx, cp, max_batch=4)
    first = enc.encode(E.pack_yuv420p8(a), len(a))[0]
    enc.reset()
    second = enc.encode(E.pack_yuv420p8(b), len(b))[0]
    enc.close()
    assert first == fresh[0] and second == fresh[1]
    m_stream, _, _, _ = ec.run_model(p, b, 22, 24, hash_sei=False, rate_control=True)
    assert second == m_stream


@pytest.mark.parametrize('seed', range(14))
def test_randomised_configurations_match_cpu_model(ctx, seed):
    """fuzz over geometry (odd CU counts, conformance-window crops), bit depth, quantiser, key-frame cadence, content class
    (incl. hard cuts and fast pans -> scene-cut key frames, intra CUs in P frames), rate control and tool flags"""
    from hevc_b200.synth import content_clip
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(9, 60)) * 8, int(rng.integers(5, 34)) * 8
    depth = int(rng.choice([8, 10]))
    qp = int(rng.integers(8, 46))
    keyint = int(rng.choice([3, 5, 8, 30]))
    kind = str(rng.choice(['base', 'hardcut', 'pan', 'static', 'grain', 'clean']))
    n = int(rng.integers(4, 10))
    rc = bool(rng.integers(0, 2))
    tools = {'scenecut': bool(rng.integers(0, 4) > 0), 'intra_in_p': bool(rng.integers(0, 4) > 0), 'sao': bool(rng.integers(0, 4) > 0)}
    deblock = bool(rng.integers(0, 5) > 0)
    p = ec.b200_params(w, h, depth, keyint=keyint)
    if rc:
        p.vbv_maxrate_kbps = int(rng.choice([150, 600, 4000]))
        p.vbv_bufsize_kbit = int(p.vbv_maxrate_kbps * 1.2)
    frames = content_clip(kind, w, h, n, seed=seed)
    m_stream, m_aus, _, _ = ec.run_model(p, frames, qp, min(51, qp + 2), hash_sei=True, rate_control=rc, deblock=deblock, **tools)
    split = [n] if n < 6 else [n // 2, n - n // 2]
    g_stream, _, _ = _gpu_encode(ctx, p, frames, (qp, min(51, qp + 2)), batch=int(rng.choice([2, 3, 8])), split=split, read_back=False,
                                 rate_control=rc, deblock=deblock, **tools)
    assert g_stream == m_stream, (w, h, depth, qp, keyint, kind, n, rc, tools, deblock)
