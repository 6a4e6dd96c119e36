"""Host-side session of the B200 HEVC encoder: thin ctypes layer over ``hb_enc_*`` (include/hevc_b200.h).

The parameters come from ``derive.derive_b200_params`` -- i.e. from the same derivation the reference feeds to
``-x265-params`` (core/transcoder.py:398-412) -- so this object stands where the reference's ffmpeg child stands
(``run_ffmpeg``, core/transcoder.py:497-535) for the video-encode step."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from .derive import B200Params

PIX_YUV420P8, PIX_P010, PIX_YUV420P16, PIX_BGR24, PIX_RGB24 = 0, 1, 2, 3, 4


class HbEncParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        'width', 'height', 'fps_num', 'fps_den', 'bit_depth', 'profile_idc', 'level_idc', 'tier', 'qp_i', 'qp_p', 'keyint', 'min_keyint',
        'vbv_maxrate_kbps', 'vbv_bufsize_kbit', 'colour_primaries', 'transfer_characteristics', 'matrix_coeffs', 'vui_colour',
        'chroma_loc', 'full_range', 'aud', 'repeat_headers', 'hrd', 'hdr10')] + [
        ('master_display', C.c_uint32 * 10), ('max_cll', C.c_int), ('max_fall', C.c_int), ('hash_sei', C.c_int),
        ('keep_recon', C.c_int), ('rate_control', C.c_int), ('deblock', C.c_int), ('scenecut', C.c_int), ('intra_in_p', C.c_int),
        ('sao', C.c_int), ('qp_cascade', C.c_int), ('reserved', C.c_int * 3)]


class HbFrames(C.Structure):
    _fields_ = [('data', C.c_void_p), ('on_device', C.c_int), ('format', C.c_int), ('n_frames', C.c_int), ('src_bit_depth', C.c_int),
                ('frame_bytes', C.c_size_t), ('src_width', C.c_int), ('src_height', C.c_int), ('matrix', C.c_int), ('reserved0', C.c_int)]


class HbFrameStat(C.Structure):
    _fields_ = [('is_idr', C.c_int), ('poc', C.c_int), ('qp', C.c_int), ('bytes', C.c_uint32), ('n_skip', C.c_uint32), ('n_merge', C.c_uint32)]


CU_DTYPE = np.dtype([('pred_mode', 'u1'), ('intra_mode', 'u1'), ('cbf', 'u1'), ('skip', 'u1'), ('mvx', '<i2'), ('mvy', '<i2')])


def crf_to_qp(crf: int) -> Tuple[int, int]:
    """Base QPs for the constant-quality operating point ``crf=`` (reference core/transcoder.py:399): the CRF value
    is taken as the P-frame quantiser minus 2, with key frames 2 below that (x265's default ipratio is ~ -2.9 QP)."""
    qp_p = max(0, min(51, crf + 2))
    return max(0, min(51, qp_p - 2)), qp_p


def to_c_params(p: B200Params, qp: Optional[Tuple[int, int]] = None, hash_sei: bool = False, keep_recon: bool = False,
                rate_control: bool = True, deblock: bool = True, scenecut: bool = True, intra_in_p: bool = True,
                sao: bool = True, qp_cascade: bool = True) -> HbEncParams:
    c = HbEncParams()
    qp_i, qp_p = qp if qp is not None else crf_to_qp(p.crf)
    for name in ('width', 'height', 'fps_num', 'fps_den', 'bit_depth', 'profile_idc', 'level_idc', 'tier', 'keyint', 'min_keyint',
                 'vbv_maxrate_kbps', 'vbv_bufsize_kbit', 'colour_primaries', 'transfer_characteristics', 'matrix_coeffs',
                 'vui_colour', 'full_range', 'aud', 'repeat_headers', 'hrd', 'hdr10', 'max_cll', 'max_fall'):
        setattr(c, name, int(getattr(p, name)))
    c.chroma_loc = int(p.chroma_loc) if p.hdr10 else -1       # only the HDR path passes chromaloc= (core/utils.py:67)
    c.qp_i, c.qp_p = qp_i, qp_p
    md = tuple(p.master_display) if p.hdr10 else (0,) * 10
    for i in range(10):
        c.master_display[i] = int(md[i])
    c.hash_sei, c.keep_recon = int(hash_sei), int(keep_recon)
    # crf= with vbv-maxrate= / vbv-bufsize= (core/transcoder.py:398-403): CRF is the quality ceiling, the VBV model caps the rate
    c.rate_control = int(rate_control)
    c.deblock = int(deblock)          # libx265 runs with the in-loop deblocking filter on unless told otherwise
    c.scenecut, c.intra_in_p, c.sao = int(scenecut), int(intra_in_p), int(sao)      # likewise on in libx265 (scenecut=40, sao)
    c.qp_cascade = int(qp_cascade)    # P-frame QP cascade, in place of x265's P / B quantiser ratio (pbratio)
    return c


@dataclass
class FrameStat:
    is_idr: bool
    poc: int
    qp: int
    bytes: int


class B200Encoder:
    """One encoder = one stream state (reference picture, POC, GOP position) on one ``Context``."""

    def __init__(self, ctx: _cabi.Context, params: HbEncParams, max_batch: int = 32):
        self.ctx, self.params, self.max_batch = ctx, params, max_batch
        L = _cabi.lib()
        L.hb_enc_create.argtypes = [C.c_void_p, C.POINTER(HbEncParams), C.c_int, C.POINTER(C.c_void_p)]
        L.hb_enc_destroy.argtypes = [C.c_void_p]
        L.hb_enc_destroy.restype = None
        L.hb_enc_headers.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.hb_enc_coded_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hb_enc_encode.argtypes = [C.c_void_p, C.POINTER(HbFrames), C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        L.hb_enc_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.hb_enc_request_stop.argtypes = [C.c_void_p]
        L.hb_enc_poll_progress.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.hb_enc_read_recon.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hb_enc_read_decisions.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.L = L
        h = C.c_void_p()
        rc = L.hb_enc_create(ctx.h, C.byref(params), max_batch, C.byref(h))
        if rc != 0:
            raise _cabi.HbError(rc, L.hb_last_error(ctx.h).decode('utf-8', 'replace'))
        self.h = h
        wc, hc = C.c_int(), C.c_int()
        L.hb_enc_coded_size(h, C.byref(wc), C.byref(hc))
        self.coded_w, self.coded_h = wc.value, hc.value
        self._out = np.empty(max(1 << 22, params.width * params.height * 3), np.uint8)

    def close(self):
        if getattr(self, 'h', None):
            self.L.hb_enc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise _cabi.HbError(rc, self.L.hb_last_error(self.ctx.h).decode('utf-8', 'replace'))

    def reset(self):
        """start a new, independent stream on this encoder (``hb_enc_reset``): nothing may be in flight"""
        self.L.hb_enc_reset.argtypes = [C.c_void_p]
        self._check(self.L.hb_enc_reset(self.h))

    def headers(self) -> bytes:
        buf = (C.c_uint8 * 1024)()
        n = C.c_size_t()
        self._check(self.L.hb_enc_headers(self.h, buf, 1024, C.byref(n)))
        return bytes(buf[:n.value])

    def frame_bytes(self, fmt: int, src_size: Optional[Tuple[int, int]] = None) -> int:
        w, h = src_size if src_size else (self.params.width, self.params.height)
        if fmt in (PIX_BGR24, PIX_RGB24):
            return 3 * w * h
        n = w * h + 2 * (w // 2) * (h // 2)
        return n if fmt == PIX_YUV420P8 else 2 * n

    def _fill(self, fr: 'HbFrames', data, n_frames, fmt, on_device, frame_bytes, src_bit_depth, src_size, matrix):
        if on_device:
            fr.data = int(data)
            arr = None
        else:
            arr = np.ascontiguousarray(data)
            fr.data = arr.ctypes.data
        fr.on_device, fr.format, fr.n_frames = int(on_device), fmt, n_frames
        fr.src_bit_depth = int(src_bit_depth) if fmt == PIX_YUV420P16 else 0
        fr.frame_bytes = frame_bytes or self.frame_bytes(fmt, src_size)
        fr.src_width, fr.src_height = (int(src_size[0]), int(src_size[1])) if src_size else (0, 0)
        fr.matrix = int(matrix)
        return arr

    def encode(self, data, n_frames: int, fmt: int = PIX_YUV420P8, force_idr: bool = False, on_device: bool = False,
               frame_bytes: Optional[int] = None, src_bit_depth: int = 0, src_size: Optional[Tuple[int, int]] = None,
               matrix: int = 0) -> Tuple[bytes, List[FrameStat]]:
        """``data``: numpy array (host; pinned memory avoids a staging copy in the driver) or an integer device address.
        ``src_bit_depth``: significant bits of a PIX_YUV420P16 source (0 = the encoder's depth).  ``src_size``: (w, h) of a
        PIX_YUV420P8 source that the ingest stage resamples to the encoder's size.  ``matrix``: HB_MATRIX_* for RGB sources."""
        fr = HbFrames()
        arr = self._fill(fr, data, n_frames, fmt, on_device, frame_bytes, src_bit_depth, src_size, matrix)   # noqa: F841 (keeps the buffer alive)
        need = n_frames * self.frame_bytes(PIX_YUV420P8) + (1 << 20)
        if self._out.size < need:
            self._out = np.empty(need, np.uint8)
        stats = (HbFrameStat * max(1, n_frames))()
        n = C.c_size_t()
        self._check(self.L.hb_enc_encode(self.h, C.byref(fr), int(force_idr), self._out.ctypes.data, self._out.size, C.byref(n), stats))
        out = self._out[:n.value].tobytes()
        return out, [FrameStat(bool(s.is_idr), s.poc, s.qp, s.bytes) for s in stats[:n_frames]]

    def encode_delayed(self, data, n_frames: int, fmt: int = PIX_YUV420P8, force_idr: bool = False, on_device: bool = False,
                       frame_bytes: Optional[int] = None, src_bit_depth: int = 0, src_size: Optional[Tuple[int, int]] = None,
                       matrix: int = 0) -> Tuple[bytes, List[FrameStat]]:
        """Pipelined form: enqueue ``n_frames`` frames, return the access units of frames submitted by earlier calls (possibly
        none).  ``data is None`` flushes.  The input buffer must stay alive until its frames have been returned."""
        if data is None:
            frp = None
        else:
            fr = HbFrames()
            arr = self._fill(fr, data, n_frames, fmt, on_device, frame_bytes, src_bit_depth, src_size, matrix)
            if arr is not None:
                self._inflight = (getattr(self, '_inflight', ()) + (arr,))[-3:]      # keep the host buffers of the batches in flight
            frp = C.byref(fr)
        need = 2 * self.max_batch * self.frame_bytes(PIX_YUV420P8) + max(0, n_frames) * self.frame_bytes(PIX_YUV420P8) + (1 << 20)
        if self._out.size < need:
            self._out = np.empty(need, np.uint8)
        cap_stats = max(1, 2 * self.max_batch + max(0, n_frames))
        stats = (HbFrameStat * cap_stats)()
        n, k = C.c_size_t(), C.c_int()
        self.L.hb_enc_encode_delayed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.L.hb_enc_encode_delayed(self.h, frp, int(force_idr), self._out.ctypes.data, self._out.size, C.byref(n), stats, C.byref(k)))
        return self._out[:n.value].tobytes(), [FrameStat(bool(s.is_idr), s.poc, s.qp, s.bytes) for s in stats[:k.value]]

    def flush(self) -> Tuple[bytes, List[FrameStat]]:
        return self.encode_delayed(None, 0)

    def mark(self):
        self.L.hb_enc_mark.argtypes = [C.c_void_p]
        self._check(self.L.hb_enc_mark(self.h))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        self.L.hb_enc_elapsed.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self.L.hb_enc_elapsed(self.h, C.byref(ms)))
        return ms.value

    def last_timing(self) -> Tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        self.L.hb_enc_last_timing(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def profile(self, enable: int = -1):
        """-> ({class: ms}, {class: launches}) accumulated since the last reset; enable=1/0 switches recording and resets."""
        ms, ln = (C.c_float * 8)(), (C.c_int * 8)()
        self.L.hb_enc_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.L.hb_enc_profile(self.h, enable, ms, ln)
        names = ('inter', 'intra', 'coarse', 'entropy', 'ingest', 'chain', 'me', 'spare')
        return dict(zip(names, ms)), dict(zip(names, ln))

    def request_stop(self):
        self.L.hb_enc_request_stop(self.h)

    def progress(self) -> int:
        n = C.c_int()
        self.L.hb_enc_poll_progress(self.h, C.byref(n))
        return n.value

    def read_recon(self, i: int):
        y = np.empty((self.coded_h, self.coded_w), np.uint16)
        u = np.empty((self.coded_h // 2, self.coded_w // 2), np.uint16)
        v = np.empty_like(u)
        self._check(self.L.hb_enc_read_recon(self.h, i, y.ctypes.data, u.ctypes.data, v.ctypes.data))
        return y, u, v

    def read_decisions(self, i: int):
        ncu = (self.coded_w // 16) * (self.coded_h // 16)
        cus = np.empty(ncu, CU_DTYPE)
        coefs = np.empty((ncu, 384), np.int16)
        self._check(self.L.hb_enc_read_decisions(self.h, i, cus.ctypes.data, coefs.ctypes.data))
        return cus.reshape(self.coded_h // 16, self.coded_w // 16), coefs


class EncoderPool:
    """Idle encoders kept for reuse, keyed by (device, parameter block, batch size): a batch worker that transcodes many
    short files of the same geometry resets one encoder per file (``hb_enc_reset``) instead of allocating and pinning
    gigabytes per file.  At most ``max_idle`` encoders are kept; the oldest is closed first."""

    def __init__(self, max_idle: int = 4):
        import threading
        self.max_idle = max_idle
        self._idle: list = []                  # [(key, ctx, enc)], oldest first
        self._lock = threading.Lock()

    @staticmethod
    def _key(device: int, params: HbEncParams, max_batch: int):
        return (int(device), bytes(params), int(max_batch))

    def acquire(self, device: int, params: HbEncParams, max_batch: int):
        key = self._key(device, params, max_batch)
        with self._lock:
            for i, (k, ctx, enc) in enumerate(self._idle):
                if k == key:
                    del self._idle[i]
                    enc.reset()
                    return key, ctx, enc
        ctx = _cabi.Context(device)
        try:
            return key, ctx, B200Encoder(ctx, params, max_batch=max_batch)
        except BaseException:
            ctx.close()
            raise

    def release(self, key, ctx, enc, reusable: bool = True):
        drop = []
        with self._lock:
            if reusable and self.max_idle > 0:
                self._idle.append((key, ctx, enc))
                while len(self._idle) > self.max_idle:
                    drop.append(self._idle.pop(0))
            else:
                drop.append((key, ctx, enc))
        for _, c, e in drop:
            e.close()
            c.close()

    def close(self):
        with self._lock:
            drop, self._idle = self._idle, []
        for _, c, e in drop:
            e.close()
            c.close()


def pack_yuv420p8(frames: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]]) -> np.ndarray:
    """[(y, u, v)] uint8 planes -> one contiguous buffer in the layout ``hb_frames`` expects."""
    return np.concatenate([np.concatenate([p.reshape(-1) for p in f]) for f in frames]).astype(np.uint8, copy=False)


class ParallelSegmentEncoder:
    """Several independent encoders on ONE device, each on its own context / CUDA stream / host thread, fed with closed-GOP
    segments round-robin (segment k -> encoder k mod S).  This is the per-GPU form of the reference's concurrency model -- N
    worker threads on independent files or segments (gui/mainwindow.py:289-301): the frame chain of one
    stream is serial (every P frame predicts from the previous reconstruction), so a second stream fills the SMs during the
    first one's kernel tails, key frames and small kernels.  Every segment starts with an IDR frame and its parameter sets
    and is rate-controlled by its own encoder, so segments concatenate into one valid stream in submission order.

    ``submit`` returns the segments that have completed, in order (possibly none); ``finish`` returns the rest."""

    def __init__(self, device, params: HbEncParams, streams: int = 2, max_batch: int = 32):
        """``device``: one device index, or a sequence of them -- then ``streams`` encoders run on EACH device and segment k
        goes to worker k mod (devices x streams), device-major, i.e. the reference's ``cycle(gpu_list)`` fan-out
        (upscale_gui_final.py:25-30,123-126) with in-order collection (:164-178)."""
        import queue
        import threading
        devices = [int(device)] if isinstance(device, int) else [int(d) for d in device]
        per = max(1, int(streams))
        self.devices = devices
        self.streams = per * len(devices)
        self.max_batch = max_batch
        self._ctxs = [_cabi.Context(devices[k % len(devices)]) for k in range(self.streams)]
        self._encs = [B200Encoder(c, params, max_batch=max_batch) for c in self._ctxs]
        self._queues = [queue.Queue() for _ in range(self.streams)]
        self._lock = threading.Lock()
        self._done: dict = {}                  # segment number -> (bytes, stats)
        self._frames: dict = {}                # segment number -> frame count
        self._error: Optional[BaseException] = None
        self._submitted = 0
        self._emitted = 0
        self._finished = 0                     # segments whose access units have come back from an encoder
        self.max_outstanding = 2 * self.streams      # submit() blocks beyond this: callers may recycle input buffers behind it
        self._timeline = [[] for _ in range(self.streams)]      # per stream: device time (ms since mark) at which each drain completed
        self._threads = [threading.Thread(target=self._work, args=(k,), daemon=True) for k in range(self.streams)]
        for t in self._threads:
            t.start()

    @property
    def launches(self) -> int:
        return sum(c.launches for c in self._ctxs)

    def _work(self, k: int):
        enc, q = self._encs[k], self._queues[k]
        pending: List[int] = []                # segment numbers inside the encoder's pipeline, oldest first
        try:
            while True:
                job = q.get()
                if job is None:
                    break
                if job == 'flush':
                    out, stats = enc.flush()
                    self._stamp(k, enc, stats)
                    self._publish(pending, out, stats)
                    pending = []
                    with self._lock:
                        self._done[('flushed', k)] = True
                    continue
                seq, data, n, kw = job
                out, stats = enc.encode_delayed(data, n, force_idr=True, **kw)
                self._stamp(k, enc, stats)
                pending.append(seq)
                done_now = pending[:-1] if stats else []
                if done_now:
                    self._publish(done_now, out, stats)
                    pending = pending[-1:]
        except BaseException as exc:           # surfaced by the next submit / finish
            with self._lock:
                self._error = exc

    def _stamp(self, k: int, enc: 'B200Encoder', stats):
        if stats:
            try:
                self._timeline[k].append(enc.elapsed_ms())
            except (_cabi.HbError, AttributeError):      # no mark recorded yet
                pass

    def timeline(self) -> List[List[float]]:
        """per stream: device time in ms (since ``mark``) of the end of every completed segment's bitstream download"""
        return [list(t) for t in self._timeline]

    def _publish(self, seqs: List[int], out: bytes, stats: List[FrameStat]):
        # one or more whole segments came back concatenated: split them at their frame counts
        pos = fpos = 0
        with self._lock:
            self._finished += len(seqs)
            for s in seqs:
                n = self._frames[s]
                size = sum(st.bytes for st in stats[fpos:fpos + n])
                self._done[s] = (out[pos:pos + size], stats[fpos:fpos + n])
                pos += size
                fpos += n

    def _collect(self) -> Tuple[bytes, List[FrameStat]]:
        out, stats = [], []
        with self._lock:
            if self._error is not None:
                raise self._error
            while self._emitted in self._done:
                b, s = self._done.pop(self._emitted)
                out.append(b)
                stats += s
                self._emitted += 1
        return b''.join(out), stats

    def submit(self, data, n_frames: int, **kw) -> Tuple[bytes, List[FrameStat]]:
        """One closed-GOP segment of at most ``max_batch`` frames (same arguments as ``B200Encoder.encode_delayed``)."""
        if n_frames < 1 or n_frames > self.max_batch:
            raise ValueError('a segment is 1..max_batch frames')
        import time
        while True:                            # back-pressure: at most max_outstanding segments between submit and completion
            with self._lock:
                if self._error is not None:
                    raise self._error
                if self._submitted - self._finished < self.max_outstanding:
                    break
            time.sleep(0.0002)      # (queued segments always drain: an encoder holds at most one segment back, so 2 x streams is never idle)
        seq = self._submitted
        self._frames[seq] = n_frames
        self._submitted += 1
        self._queues[seq % self.streams].put((seq, data, n_frames, kw))
        return self._collect()

    def reset(self):
        """start a new clip on the same encoders (nothing may be in flight: call after ``finish``)"""
        for e in self._encs:
            e.reset()
        with self._lock:
            self._done.clear()
            self._frames.clear()
            self._submitted = self._emitted = self._finished = 0

    def finish(self) -> Tuple[bytes, List[FrameStat]]:
        """Drain every stream; returns the remaining segments in order."""
        import time
        for k, q in enumerate(self._queues):
            with self._lock:
                self._done.pop(('flushed', k), None)
            q.put('flush')
        while True:
            with self._lock:
                if self._error is not None:
                    raise self._error
                if all(('flushed', k) in self._done for k in range(self.streams)):
                    for k in range(self.streams):
                        self._done.pop(('flushed', k))
                    break
            time.sleep(0.0002)
        return self._collect()

    def mark(self):
        for k, e in enumerate(self._encs):
            e.mark()
            self._timeline[k] = []

    def elapsed_ms(self) -> float:
        """Device time from ``mark`` to the last completed download, maximum over the streams."""
        spans = []
        for e in self._encs:
            try:
                spans.append(e.elapsed_ms())
            except _cabi.HbError:              # a stream that received no segment since the mark
                pass
        return max(spans) if spans else 0.0

    def profile(self, enable: int = -1):
        ms, ln = {}, {}
        for e in self._encs:
            a, b = e.profile(enable)
            for k in a:
                ms[k] = ms.get(k, 0.0) + a[k]
                ln[k] = ln.get(k, 0) + b[k]
        return ms, ln

    def close(self):
        for q in self._queues:
            q.put(None)
        for t in self._threads:
            t.join(timeout=30)
        for e in self._encs:
            e.close()
        for c in self._ctxs:
            c.close()
