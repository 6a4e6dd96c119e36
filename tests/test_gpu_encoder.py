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


def _gpu_encode(ctx, p, frames8, qp, batch, hash_sei=True, split=None, read_back=True, rate_control=False, deblock=True):
    from hevc_b200 import encoder as E
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=qp, hash_sei=hash_sei, keep_recon=True, rate_control=rate_control, deblock=deblock), max_batch=batch)
    stream, recs, decs = b'', [], []
    _gpu_encode.last_qps = []
    pos = 0
    for n in (split or [len(frames8)]):
        out, stats = enc.encode(E.pack_yuv420p8(frames8[pos:pos + n]), n)
        stream += out
        _gpu_encode.last_qps = getattr(_gpu_encode, 'last_qps', []) + [s.qp for s in stats]
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
        assert set(m_qps) == {20, 22}
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
