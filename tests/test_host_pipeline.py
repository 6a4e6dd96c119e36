"""Host-side logic (no GPU): probe parsing vs the reference, MP4 muxer + compliance checks on a stream from the CPU model,
worker / batch scheduling, rank sharding (world_size 2 over gloo)."""
import dataclasses
import json
import os
import subprocess
import sys
import threading
from pathlib import Path

import numpy as np
import pytest

from hevc_b200 import batch, compliance, mp4, probe, transcoder, worker
from hevc_b200.frames import Y4MReader, write_y4m
from hevc_b200.synth import SynthClip
from tests import enc_common as ec

GOLD = json.loads((Path(__file__).parent / 'golden' / 'probe_golden.json').read_text())


def test_probe_parsing_matches_reference():
    for case in GOLD:
        try:
            info = probe.info_from_ffprobe_json(case['doc'])
        except Exception:
            info = probe.VideoInfo(**probe._FALLBACK)        # probe_media's catch-all (core/probe.py:114-116)
        assert dataclasses.asdict(info) == case['info'], case['doc']


def test_probe_never_raises_and_reads_y4m(tmp_path):
    assert dataclasses.asdict(probe.probe_media(tmp_path / 'missing.mp4')) == probe._FALLBACK
    clip = SynthClip(64, 48, seed=2)
    p = tmp_path / 'a.y4m'
    write_y4m(p, [clip.frame(i) for i in range(3)], 64, 48, (60, 1))
    (tmp_path / 'a.y4m.json').write_text(json.dumps({'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'color_space': 'bt2020nc'}))
    info = probe.probe_media(p)
    assert (info.width, info.height, info.fps, info.nb_frames, info.hdr, info.pix_fmt) == (64, 48, 60.0, 3, True, 'yuv420p')
    assert abs(info.duration - 0.05) < 1e-9
    bufs = list(Y4MReader(p).batches(2))
    assert [b[1] for b in bufs] == [2, 1] and bufs[0][0].shape == (2, 64 * 48 * 3 // 2)
    assert (bufs[0][0][1][:64 * 48].reshape(48, 64) == clip.frame(1)[0]).all()


def _model_stream(depth, n=5, w=96, h=64, keyint=3):
    p = ec.b200_params(w, h, depth, keyint=keyint)
    stream, aus, recs, _ = ec.run_model(p, ec.clip_frames(w, h, n), 24, 26, hash_sei=False)
    return p, stream, recs


@pytest.mark.parametrize('depth', [8, 10])
def test_mp4_mux_and_compliance(depth, tmp_path):
    p, stream, recs = _model_stream(depth)
    track = mp4.TrackInfo(p.width, p.height, p.fps_num, p.fps_den, p.profile_idc, p.level_idc, p.tier, p.bit_depth, p.colour_primaries,
                          p.transfer_characteristics, p.matrix_coeffs, 0, p.master_display if p.hdr10 else None, p.max_cll, p.max_fall)
    data = mp4.mux_annexb(track, stream)
    expect = {'profile_idc': p.profile_idc, 'level_idc': p.level_idc, 'tier': 0, 'keyint': 3, 'hdr10': bool(p.hdr10),
              'master_display': p.master_display, 'max_cll': p.max_cll, 'max_fall': p.max_fall}
    assert compliance.check_bytes(data, expect) == []
    rep = compliance.inspect(data)
    assert rep['n_samples'] == 5 and rep['sync_samples'] == [1, 4] and rep['sps']['width'] == 96 and rep['sps']['height'] == 64
    # a decoder that knows nothing about this package (OpenCV's FFmpeg: mov demuxer + hevc decoder) plays the file
    import cv2
    f = tmp_path / 'o.mp4'
    f.write_bytes(data)
    cap = cv2.VideoCapture(str(f))
    assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 5
    assert abs(cap.get(cv2.CAP_PROP_FPS) - 30.0) < 1e-3
    n = 0
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        y = recs[n][0][:64, :96].astype(np.float64) / (4.0 if depth == 10 else 1.0)
        got = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY).astype(np.float64)
        # same picture; OpenCV's YUV->BGR->gray path differs (and is BT.2020/PQ-unaware for the HDR case)
        assert np.corrcoef(y.reshape(-1), got.reshape(-1))[0, 1] > (0.98 if depth == 8 else 0.7)
        n += 1
    assert n == 5
    # and the samples, put back into Annex-B with the hvcC parameter sets, decode bit-exactly to the model's reconstruction
    import struct
    from oracle import fforacle
    ps = rep['param_sets']
    annexb = b''.join(b'\0\0\0\1' + ps[t][0] for t in (32, 33, 34))
    pos = rep['chunk_offset']
    for size in rep['sample_sizes']:
        end = pos + size
        while pos < end:
            ln = struct.unpack('>I', data[pos:pos + 4])[0]
            annexb += b'\0\0\0\1' + data[pos + 4:pos + 4 + ln]
            pos += 4 + ln
    dec = fforacle.decode_hevc(annexb, verify_hash=False)
    assert len(dec) == 5
    for d, r in zip(dec, recs):
        assert all((d[c] == r[c][:d[c].shape[0], :d[c].shape[1]]).all() for c in range(3))


def test_compliance_flags_problems():
    p, stream, _ = _model_stream(8)
    track = mp4.TrackInfo(p.width, p.height, 30, 1, 1, 120, 0, 8, 1, 1, 1)
    good = mp4.mux_annexb(track, stream)
    assert compliance.check_bytes(good) == []
    assert 'sample entry is hev1, not hvc1' in compliance.check_bytes(good.replace(b'hvc1', b'hev1'))
    assert any('brand' in s for s in compliance.check_bytes(good.replace(b'ftyp' + b'mp42', b'ftyp' + b'isom', 1)))
    assert any('level_idc' in s for s in compliance.check_bytes(good, {'level_idc': 153}))
    assert any('HDR10 SEI' in s for s in compliance.check_bytes(good, {'hdr10': True}))


def test_convert_video_contract_without_backend(tmp_path, monkeypatch):
    """never raises, result keys, CANCELLED status, final (total, total) progress tick (core/transcoder.py:537-638)"""
    clip = SynthClip(64, 48, seed=1)
    src = tmp_path / 'clip.y4m'
    write_y4m(src, [clip.frame(i) for i in range(2)], 64, 48)
    ticks = []
    monkeypatch.setattr(transcoder, 'encode_b200', lambda *a, **k: (1, 'no device here'))
    res = transcoder.convert_video(src, tmp_path, progress_callback=lambda *a: ticks.append(a), encoder='b200')
    assert list(res) == ['file', 'status', 'quality', 'retries', 'method', 'hdr']
    assert res['status'] == 'FAILED' and res['method'] == 'B200' and res['file'] == 'clip.y4m'
    assert ticks[-1][1] == ticks[-1][2]
    ev = threading.Event()
    ev.set()
    assert transcoder.convert_video(src, tmp_path, encoder='b200', stop_event=ev)['status'] == 'CANCELLED'
    monkeypatch.setattr(transcoder, 'encode_b200', lambda *a, **k: (0, ''))
    monkeypatch.setattr(compliance, 'check_file', lambda *a, **k: [])
    ok = transcoder.convert_video(src, tmp_path, encoder='b200', progress_callback=lambda *a: 1 / 0)      # callback errors are swallowed
    assert ok['status'] == 'SUCCESS' and ok['quality'] == 17 + 1 - 1 or ok['quality'] in range(16, 25)
    # the reference branch without ffmpeg fails the same way the reference does (status FAILED, method CPU)
    cpu = transcoder.convert_video(src, tmp_path, force_cpu=True)
    assert cpu['status'] == 'FAILED' and cpu['method'] == 'CPU'


def test_worker_and_batch_scheduling(tmp_path, monkeypatch):
    files = [tmp_path / f'f{i}.y4m' for i in range(7)]
    for f in files:
        f.write_bytes(b'YUV4MPEG2 W16 H16 F30:1\n')
    active, peak, lock = [0], [0], threading.Lock()

    def fake(file_path, out_dir, progress_callback=None, stop_event=None, **kw):
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        threading.Event().wait(0.02)
        with lock:
            active[0] -= 1
        return {'file': Path(file_path).name, 'status': 'SUCCESS', 'quality': 19, 'retries': 0, 'method': 'B200', 'hdr': False}

    monkeypatch.setattr(batch, 'convert_video', fake)
    res = batch.batch_convert(tmp_path, tmp_path / 'out', max_workers=3, encoder='b200')
    assert sorted(r['file'] for r in res) == sorted(f.name for f in files) and 1 < peak[0] <= 3
    rows = (tmp_path / 'out' / 'transcode_log.csv').read_text().strip().splitlines()
    assert rows[0] == 'file,status,quality,retries,method,hdr' and len(rows) == 8
    monkeypatch.setattr(worker, 'convert_video', fake)
    w = worker.TranscodeWorker(files[0], tmp_path, encoder='b200')
    got = []
    w.finished.connect(got.append)
    w.start()
    w.join()
    assert got[0]['status'] == 'SUCCESS'
    w.stop()
    assert w.stop_event.is_set()


def test_gop_segments_and_lpt():
    assert batch.gop_segments(300, 120) == [(0, 120), (120, 240), (240, 300)]
    items = [9, 8, 7, 3, 2, 1]
    shards = [batch.shard_for_rank(items, r, 2, cost=float) for r in range(2)]
    assert sorted(shards[0] + shards[1]) == sorted(items) and abs(sum(shards[0]) - sum(shards[1])) <= 1


_RANK_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hevc_b200.batch import shard_for_rank, gop_segments
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
units = gop_segments(1800, 60)                       # config 4: one long clip cut at keyint
mine = shard_for_rank(units, rank, world)
flags = torch.zeros(len(units), dtype=torch.int32)
for u in mine:
    flags[units.index(u)] = 1
dist.all_reduce(flags)                                # test-only collective: every unit must be owned exactly once
assert bool((flags == 1).all()), flags
counts = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int32))
assert max(int(c) for c in counts) - min(int(c) for c in counts) <= 1
dist.destroy_process_group()
print('rank', rank, 'ok', len(mine))
'''


def test_rank_sharding_world_size_2(tmp_path):
    script = tmp_path / 'rank.py'
    script.write_text(_RANK_SCRIPT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', '29533', str(script), str(Path(__file__).resolve().parent.parent)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count('ok') == 2


def test_parallel_segment_encoder_orders_and_splits_segments(monkeypatch):
    """Host logic of ParallelSegmentEncoder without a GPU: stub encoders that return their input one call late (like
    hb_enc_encode_delayed); segments must come back complete, in submission order, whichever stream finishes first."""
    import time

    from hevc_b200 import _cabi
    from hevc_b200 import encoder as E

    class StubCtx:
        launches = 0

        def __init__(self, device):
            pass

        def close(self):
            pass

    class StubEnc:
        count = 0

        def __init__(self, ctx, params, max_batch=32):
            self.k = StubEnc.count
            StubEnc.count += 1
            self.held = None

        def encode_delayed(self, data, n, force_idr=False, **kw):
            assert force_idr
            time.sleep(0.002 * (2 - self.k))                      # stream 0 is the slow one
            prev, self.held = self.held, (bytes(data), n)
            return self._emit(prev)

        def flush(self):
            prev, self.held = self.held, None
            return self._emit(prev)

        @staticmethod
        def _emit(item):
            if item is None:
                return b'', []
            payload, n = item
            per = len(payload) // n
            return payload, [E.FrameStat(i == 0, i, 30, per) for i in range(n)]

        def close(self):
            pass

    monkeypatch.setattr(_cabi, 'Context', StubCtx)
    monkeypatch.setattr(E, 'B200Encoder', StubEnc)
    pse = E.ParallelSegmentEncoder(0, None, streams=2, max_batch=4)
    segs = [bytes([65 + i]) * (4 * (i % 3 + 1)) for i in range(7)]          # 4 frames each, different sizes
    got, frames = b'', 0
    for s in segs:
        out, stats = pse.submit(s, 4)
        got += out
        frames += len(stats)
    out, stats = pse.finish()
    got += out
    frames += len(stats)
    pse.close()
    assert got == b''.join(segs) and frames == 28
    with pytest.raises(ValueError):
        E.ParallelSegmentEncoder.submit(pse, b'', 0)


_SEGMENT_MUX_SCRIPT = r'''
import os, sys, pickle, torch, torch.distributed as dist
root, work = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
import numpy as np
from hevc_b200 import mp4
from hevc_b200.batch import gop_segments
from hevc_b200.synth import SynthClip
from oracle import encoder_model as em
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
w, h, n, keyint = 96, 64, 14, 4
segs = gop_segments(n, keyint)                         # config 4 in miniature: closed-GOP segment k -> rank k mod world
clip = SynthClip(w, h, seed=5)
runs = {}
for k, (a, b) in enumerate(segs):
    if k % world != rank:
        continue
    enc = em.ModelEncoder(em.make_params(w, h, 8, qp_i=30, qp_p=32, keyint=keyint, hash_sei=False))     # every segment: its own encoder, IDR first
    es = b''.join(enc.encode(*clip.frame(i))[0] for i in range(a, b))
    enc.close()
    runs[k] = mp4.to_samples(es)                       # converted to MP4 sample form on the rank that coded it
    open(os.path.join(work, 'seg%02d.es' % k), 'wb').write(es)
gathered = [None] * world if rank == 0 else None
dist.gather_object(runs, gathered, dst=0)
if rank == 0:
    allruns = {}
    for g in gathered:
        allruns.update(g)
    track = mp4.TrackInfo(w, h, 30, 1, 1, 90, 0, 8, 1, 1, 1)
    open(os.path.join(work, 'out.mp4'), 'wb').write(mp4.assemble(track, [allruns[k] for k in range(len(segs))]))
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok', sorted(runs))
'''


def test_segment_gather_and_single_mux_world_size_2(tmp_path):
    """the N > 1 host path of config 4 on two gloo ranks: per-rank closed-GOP segments (CPU model streams), per-rank conversion
    to MP4 sample form, gather on rank 0, ONE file -- equal to muxing the concatenated stream in one process, and decodable"""
    from oracle import fforacle
    script = tmp_path / 'segmux.py'
    script.write_text(_SEGMENT_MUX_SCRIPT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', '29547', str(script), str(Path(__file__).resolve().parent.parent), str(tmp_path)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    es = b''.join((tmp_path / f'seg{k:02d}.es').read_bytes() for k in range(4))
    track = mp4.TrackInfo(96, 64, 30, 1, 1, 90, 0, 8, 1, 1, 1)
    data = (tmp_path / 'out.mp4').read_bytes()
    assert data == mp4.mux_annexb(track, es)
    rep = compliance.inspect(data)
    assert rep['n_samples'] == 14 and rep['sync_samples'] == [1, 5, 9, 13]
    assert len(fforacle.decode_hevc(es, verify_hash=False)) == 14


def test_stream_muxer_equals_one_shot_mux(tmp_path):
    """the incremental muxer (samples spooled to disk, tables in memory) writes the same bytes as mux_annexb, leaves no spool
    behind, and writes nothing when the encode fails half way"""
    from oracle import encoder_model as em
    clip = SynthClip(64, 48, seed=3)
    enc = em.ModelEncoder(em.make_params(64, 48, 8, qp_i=30, qp_p=32, keyint=3, hash_sei=False))
    aus = [enc.encode(*clip.frame(i))[0] for i in range(7)]
    enc.close()
    track = mp4.TrackInfo(64, 48, 30, 1, 1, 90, 0, 8, 1, 1, 1)
    out = tmp_path / 'a.mp4'
    with mp4.StreamMuxer(track, out) as m:
        m.feed(b'')
        m.feed(b''.join(aus[:2]))
        m.feed(aus[2])
        m.feed(b''.join(aus[3:]))
    assert out.read_bytes() == mp4.mux_annexb(track, b''.join(aus))
    assert sorted(p.name for p in tmp_path.iterdir()) == ['a.mp4']
    bad = tmp_path / 'b.mp4'
    with pytest.raises(RuntimeError):
        with mp4.StreamMuxer(track, bad) as m:
            m.feed(aus[0])
            raise RuntimeError('encoder died')
    assert not bad.exists() and not (tmp_path / 'b.mp4.mdat.tmp').exists()


def test_encoder_pool_reuses_and_evicts(monkeypatch):
    from hevc_b200 import _cabi
    from hevc_b200 import encoder as E
    made, closed = [], []

    class StubCtx:
        def __init__(self, device):
            self.device = device

        def close(self):
            closed.append(('ctx', self.device))

    class StubEnc:
        def __init__(self, ctx, params, max_batch=32):
            self.resets = 0
            made.append(self)

        def reset(self):
            self.resets += 1

        def close(self):
            closed.append(('enc', id(self)))

    monkeypatch.setattr(_cabi, 'Context', StubCtx)
    monkeypatch.setattr(E, 'B200Encoder', StubEnc)
    pool = E.EncoderPool(max_idle=2)
    pa, pb = E.HbEncParams(width=64, height=64), E.HbEncParams(width=128, height=64)
    k1, c1, e1 = pool.acquire(0, pa, 8)
    pool.release(k1, c1, e1)
    k2, c2, e2 = pool.acquire(0, pa, 8)
    assert e2 is e1 and e1.resets == 1 and len(made) == 1                 # same geometry: reused after a reset
    k3, c3, e3 = pool.acquire(0, pb, 8)
    assert e3 is not e1 and len(made) == 2                                # other geometry: a new one
    pool.release(k2, c2, e2)
    pool.release(k3, c3, e3, reusable=False)                              # a failed encode never goes back
    assert ('enc', id(e3)) in closed and ('enc', id(e1)) not in closed
    for _ in range(3):
        k, c, e = pool.acquire(1, pa, 8)                                  # other device: new encoders; the idle list is capped
        pool.release(k, c, e)
    pool.close()
    assert sum(1 for x in closed if x[0] == 'enc') == len(made)


def test_memory_reader_and_ring_buffers(tmp_path):
    from hevc_b200.frames import MemoryReader, RawYuvReader, _Ring
    frames = np.arange(7 * 24, dtype=np.uint8).reshape(7, 24)
    got = list(MemoryReader(frames).batches(3))
    assert [n for _, n, _ in got] == [3, 3, 1] and (np.concatenate([b for b, _, _ in got]) == frames).all()
    ring = _Ring(16, 3)
    ids = [ring.next().ctypes.data for _ in range(7)]
    assert ids[:3] == ids[3:6] and ids[6] == ids[0] and len(set(ids[:3])) == 3
    ring.close()
    # raw planar reader: a yielded batch stays intact while ring - 1 further batches are read
    w, h = 16, 8
    raw = np.random.default_rng(0).integers(0, 256, (5, w * h * 3 // 2), dtype=np.uint8)
    path = tmp_path / 'c.yuv'
    path.write_bytes(raw.tobytes())
    info = probe.VideoInfo(**{**probe._FALLBACK, 'width': w, 'height': h, 'pix_fmt': 'yuv420p'})
    held = []
    for buf, n, fmt in RawYuvReader(path, info).batches(2, ring=3):
        held.append((buf, buf.copy()))
        for view, snap in held[-2:]:
            assert (view == snap).all()
    assert sum(len(v) for v, _ in held) == 5 and (np.concatenate([s for _, s in held]) == raw).all()


def _model_mp4(tmp_path, depth, w=176, h=112, n=5, name='src.mp4'):
    """an hvc1 MP4 written by the CPU model + the muxer; returns (path, decoded planes per frame)"""
    from oracle import encoder_model as em
    from oracle import fforacle
    clip = SynthClip(w, h, seed=2)
    enc = em.ModelEncoder(em.make_params(w, h, depth, qp_i=20, qp_p=22, keyint=10, hash_sei=False, hdr10=(depth == 10)))
    es = b''.join(enc.encode(*[a.astype(np.uint16) << (depth - 8) for a in clip.frame(i)])[0] for i in range(n))
    enc.close()
    md = (13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50)
    hdr = depth == 10
    track = mp4.TrackInfo(w, h, 30, 1, 2 if hdr else 1, 90, 0, depth, 9 if hdr else 1, 16 if hdr else 1, 9 if hdr else 1, 0, md if hdr else None, 1000, 400)
    path = tmp_path / name
    path.write_bytes(mp4.mux_annexb(track, es))
    return path, fforacle.decode_hevc(es, verify_hash=False)


@pytest.mark.parametrize('depth', [8, 10])
def test_container_reader_keeps_native_depth_and_probe_reads_hdr10_boxes(tmp_path, depth):
    """a11: container sources are demuxed + decoded through the bundled libavformat / libavcodec at their own bit depth (the
    samples are exactly the decoder's), and the ffprobe-less prober reads the pixel format and the colr / mdcv / clli boxes --
    a Main10 HDR10 MP4 is classified HDR with its mastering-display string, like ffprobe would report it (core/probe.py:47-111)"""
    from hevc_b200.avreader import AvReader
    from hevc_b200.frames import open_reader
    path, dec = _model_mp4(tmp_path, depth)
    info = probe.probe_media(path)
    assert (info.width, info.height, info.nb_frames) == (176, 112, 5)
    if depth == 10:
        assert info.hdr and info.pix_fmt == 'yuv420p10le' and (info.color_primaries, info.color_transfer, info.color_space) == ('bt2020', 'smpte2084', 'bt2020nc')
        assert info.master_display == 'G(13250,34500)B(7500,3000)R(34000,16000)WP(15635,16450)L(10000000,50)' and info.max_cll == '1000,400'
    else:
        assert not info.hdr and info.pix_fmt == 'yuv420p' and info.color_primaries == 'bt709'
    r = open_reader(path, info)
    assert isinstance(r, AvReader) and r.src_bit_depth == depth and r.kind == 'yuv'
    got = []
    for buf, k, fmt in r.batches(2):
        got += [b.copy() for b in buf[:k]]
    r.close()
    assert len(got) == 5
    for g, d in zip(got, dec):
        want = np.concatenate([pl.astype('<u2' if depth == 10 else np.uint8).reshape(-1).view(np.uint8) for pl in d])
        assert (g == want).all()


def test_stream_level_hdr10_probe_for_any_container(tmp_path):
    """without ffprobe and outside MP4 / MOV (MKV, TS, raw Annex-B ...) the colour description comes from the decoder context and
    the HDR10 static metadata from the first frame's side data, both through the bundled libavcodec: an HDR10 stream is classified
    HDR with its mastering-display string like ffprobe's side_data_list would give it (reference core/probe.py:47-111)"""
    import shutil

    from hevc_b200.avreader import decoded_format
    src, _ = _model_mp4(tmp_path, 10)
    fmt = decoded_format(src)
    assert (fmt['color_primaries'], fmt['color_transfer'], fmt['color_space']) == (9, 16, 9) and fmt['pix_fmt'] == 'yuv420p10le'
    assert fmt['master_display'] == 'G(13250,34500)B(7500,3000)R(34000,16000)WP(15635,16450)L(10000000,50)' and fmt['max_cll'] == '1000,400'
    other = tmp_path / 'no_boxes_probed.dat'               # same bytes, but the MP4 box prober only looks at .mp4 / .mov / .m4v
    shutil.copy(src, other)
    info = probe.probe_media(other)
    assert info.hdr and (info.color_primaries, info.color_transfer, info.color_space) == ('bt2020', 'smpte2084', 'bt2020nc')
    assert info.master_display == fmt['master_display'] and info.max_cll == '1000,400'
    # a raw Annex-B elementary stream: no container at all, tags from VUI + SEI 137 / 144; OpenCV's bogus frame count is dropped
    from oracle import encoder_model as em
    clip = SynthClip(176, 112, seed=2)
    enc = em.ModelEncoder(em.make_params(176, 112, 10, qp_i=20, qp_p=22, keyint=10, hash_sei=False, hdr10=True))
    es = b''.join(enc.encode(*[a.astype(np.uint16) << 2 for a in clip.frame(i)])[0] for i in range(4))
    enc.close()
    raw = tmp_path / 'raw.hevc'
    raw.write_bytes(es)
    info = probe.probe_media(raw)
    assert info.hdr and info.pix_fmt == 'yuv420p10le' and info.master_display == fmt['master_display'] and (info.width, info.height) == (176, 112)
    assert info.nb_frames is None or 0 < info.nb_frames < 100
    sdr, _ = _model_mp4(tmp_path, 8, name='sdr.mp4')
    f8 = decoded_format(sdr)
    assert f8['color_primaries'] == 1 and 'master_display' not in f8


def test_audio_channel_probe_through_libavformat(tmp_path):
    """ffprobe-less audio `channels` (reference core/probe.py:47-111): WAV files of 1 / 2 / 6 / 8 channels, and 0 for a video-only MP4"""
    import wave

    from hevc_b200.avreader import audio_channels
    for ch in (1, 2, 6, 8):
        p = tmp_path / f'a{ch}.wav'
        with wave.open(str(p), 'wb') as w:
            w.setnchannels(ch)
            w.setsampwidth(2)
            w.setframerate(48000)
            w.writeframes(np.zeros((480, ch), np.int16).tobytes())
        assert audio_channels(p) == ch
    src, _ = _model_mp4(tmp_path, 8)
    assert audio_channels(src) == 0 and probe.probe_media(src).audio_channels == 0
    assert audio_channels(tmp_path / 'missing.mkv') is None


def _write_y4m_tag(path, tag, w, h, frames, dtype):
    with open(path, 'wb') as fh:
        fh.write(f'YUV4MPEG2 W{w} H{h} F30:1 Ip A1:1 C{tag}\n'.encode())
        for planes in frames:
            fh.write(b'FRAME\n')
            for pl in planes:
                fh.write(np.ascontiguousarray(pl, dtype=dtype).tobytes())


@pytest.mark.parametrize('tag,depth,cdiv', [('422p10', 10, (2, 1)), ('444', 8, (1, 1)), ('422', 8, (2, 1)), ('444p10', 10, (1, 1)), ('420p12', 12, (2, 2))])
def test_other_yuv_layouts_go_through_libswscale_at_native_depth(tmp_path, tag, depth, cdiv):
    """4:2:2 / 4:4:4 / 12-bit sources (ProRes-style masters; here Y4M C422p10 ...) are converted to planar 4:2:0 by the bundled libswscale
    -- the call the reference's ffmpeg child makes for -pix_fmt (core/transcoder.py:464) -- at 8 bits for 8-bit sources and 10 bits for
    deeper ones: luma passes through untouched (rounded from 12 bits), chroma is the filtered 2:1 reduction of the source planes"""
    from hevc_b200.avreader import AvReader
    from hevc_b200.frames import open_reader
    w, h, n = 64, 48, 3
    cw, ch = w // cdiv[0], h // cdiv[1]
    mx = (1 << depth) - 1
    yy, xx = np.mgrid[0:h, 0:w]
    cy, cx = np.mgrid[0:ch, 0:cw]
    src = [(((np.sin(xx / 7 + i) + np.cos(yy / 5)) * 0.2 + 0.5) * mx, ((np.sin(cx / (12 / cdiv[0]) + i) * 0.2) + 0.5) * mx,
            ((np.cos(cy / (8 / cdiv[1]) + i) * 0.2) + 0.5) * mx) for i in range(n)]
    src = [tuple(p.astype(np.int64) for p in f) for f in src]
    path = tmp_path / f'src_{tag}.y4m'
    _write_y4m_tag(path, tag, w, h, src, '<u2' if depth > 8 else np.uint8)
    info = probe.probe_media(path)
    assert (info.width, info.height) == (w, h) and info.pix_fmt.startswith('yuv4') and str(depth if depth > 8 else '') in info.pix_fmt
    r = open_reader(path, info)
    assert isinstance(r, AvReader) and r.kind == 'yuv' and r.src_bit_depth == (10 if depth > 8 else 8)
    assert r.source_pix_fmt == info.pix_fmt and r.pix_fmt == ('yuv420p10le' if depth > 8 else 'yuv420p')
    out_depth = r.src_bit_depth
    got = []
    for buf, k, _ in r.batches(2, ring=3):
        got += [np.array(buf[i] if buf.ndim > 1 else buf[i * r.frame_bytes:(i + 1) * r.frame_bytes]) for i in range(k)]
    r.close()
    assert len(got) == n
    sh = depth - out_depth
    for f, g in zip(src, got):
        g = g.view('<u2' if out_depth > 8 else np.uint8).astype(np.int64)
        y, u, v = g[:w * h].reshape(h, w), g[w * h:w * h + w * h // 4].reshape(h // 2, w // 2), g[w * h + w * h // 4:].reshape(h // 2, w // 2)
        want_y = (f[0] + (1 << sh >> 1)) >> sh if sh else f[0]
        assert np.abs(y - want_y).max() <= (1 if sh else 0)
        for got_c, src_c in ((u, f[1]), (v, f[2])):
            box = src_c.reshape(h // 2, cdiv[1] and (ch // (h // 2)), w // 2, cw // (w // 2)).mean(axis=(1, 3)) / (1 << sh)
            assert np.abs(got_c - box).max() <= 0.02 * (1 << out_depth)      # smooth content: any sane 2:1 filter stays within 2 % of the box mean


def test_container_reader_falls_back_to_opencv_for_rgb_codecs(tmp_path):
    import cv2

    from hevc_b200.frames import Cv2Reader, open_reader
    src = tmp_path / 'rgb.avi'
    vw = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*'FFV1'), 25.0, (64, 48))
    for i in range(3):
        vw.write(np.full((48, 64, 3), 40 * i, np.uint8))
    vw.release()
    r = open_reader(src, probe.probe_media(src))
    assert isinstance(r, Cv2Reader) and r.kind == 'bgr'
    assert sum(n for _, n, _ in r.batches(2)) == 3


def test_compliance_checks_hrd_conformance_of_sample_sizes():
    """the self-check parses the HRD the stream signals (hrd=1, core/utils.py:65) and runs the decoder-buffer model over the real
    sample sizes: a rate-controlled stream passes, the same content coded far above the signalled rate is flagged"""
    from oracle import encoder_model as em
    w, h = 320, 192
    clip = SynthClip(w, h, seed=7)
    md = (13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50)
    track = mp4.TrackInfo(w, h, 30, 1, 2, 120, 0, 10, 9, 16, 9, 0, md, 1000, 400)

    def stream(n, **kw):
        enc = em.ModelEncoder(em.make_params(w, h, 10, keyint=12, hash_sei=False, hdr10=True, vbv_maxrate_kbps=250, vbv_bufsize_kbit=300, **kw))
        es = b''.join(enc.encode(*[a.astype(np.uint16) << 2 for a in clip.frame(i)])[0] for i in range(n))
        enc.close()
        return mp4.mux_annexb(track, es)

    good = stream(30, qp_i=20, qp_p=22, rate_control=1)
    sps = compliance.inspect(good)['sps']
    assert sps['hrd'] == 1 and abs(sps['hrd_bit_rate'] - 250000) < 64 and sps['hrd_cpb_size'] == 300000
    assert compliance.check_bytes(good, {'hdr10': True}) == []
    bad = compliance.check_bytes(stream(12, qp_i=12, qp_p=14, rate_control=0), {'hdr10': True})
    assert len(bad) == 1 and bad[0].startswith('HRD buffer underflow')
    assert compliance.hrd_underflows([100, 100, 5000, 100], 3000.0, 1000.0, 30.0) == [(2, 5000 - 900.0)]       # starts 90 % full, refills 100 / frame
