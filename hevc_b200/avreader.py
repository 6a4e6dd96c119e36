"""Container sources at their native bit depth: demux + decode through the FFmpeg libraries bundled with the OpenCV wheel
(libavformat / libavcodec via ctypes), planar 4:2:0 out -- the decode step of the reference's ffmpeg child.

``cv2.VideoCapture`` hands over 8-bit BGR: a Main10 / 10-bit source loses two bits and takes a YUV -> BGR -> YUV round trip.  This
reader keeps the decoder's own samples (yuv420p -> HB_PIX_YUV420P8, yuv420p10le -> HB_PIX_YUV420P16 with 10 significant bits).
Every other YUV layout -- 4:2:2 / 4:4:4 (ProRes, DNxHR, v210, Y4M C422 / C444), full-range yuvj*, 12- / 16-bit, packed or
semi-planar -- goes through the bundled libswscale to yuv420p (8-bit sources) or yuv420p10le (deeper ones): the same library call
the reference gets from the scaler ffmpeg inserts for ``-pix_fmt`` (core/transcoder.py:464), so 10-bit 4:2:2 masters keep their
precision.  RGB / palette / Bayer formats and library layouts this module does not recognise raise ``Unsupported`` and the caller
falls back to the OpenCV reader (the RGB -> YUV matrix then follows the probe, hevc_b200/transcoder.py).

No FFmpeg headers are available, so only long-stable, documented structure heads are touched, each guarded by a sanity check:
``AVFormatContext.nb_streams / .streams`` (offsets 44 / 48 since libavformat 58), ``AVStream.codecpar`` (offset 16 since
libavformat 59, where ``av_class`` became the first member), ``AVPacket.stream_index`` (offset 36), and the head of ``AVFrame``
(data, linesize, width, height, format).  Everything else goes through functions.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from pathlib import Path
from typing import Iterator, Optional, Tuple

import numpy as np

AV_PIX_FMT_YUV420P, AV_PIX_FMT_YUVJ420P, AV_PIX_FMT_YUV420P10LE = 0, 12, 62
_PIX_FLAG_PAL, _PIX_FLAG_RGB, _PIX_FLAG_BAYER = 1 << 1, 1 << 5, 1 << 8
_SWS_BICUBIC = 4                   # ffmpeg's default for the scaler it auto-inserts
_EAGAIN, _EOF = -11, -541478725


class Unsupported(RuntimeError):
    pass


class _AVFrameHead(C.Structure):
    _fields_ = [('data', C.c_void_p * 8), ('linesize', C.c_int * 8), ('extended_data', C.c_void_p),
                ('width', C.c_int), ('height', C.c_int), ('nb_samples', C.c_int), ('format', C.c_int)]


class _AVComponentDescriptor(C.Structure):
    _fields_ = [('plane', C.c_int), ('step', C.c_int), ('offset', C.c_int), ('shift', C.c_int), ('depth', C.c_int)]


class _AVPixFmtDescriptorHead(C.Structure):          # libavutil >= 58: five ints per component, no deprecated members
    _fields_ = [('name', C.c_char_p), ('nb_components', C.c_uint8), ('log2_chroma_w', C.c_uint8), ('log2_chroma_h', C.c_uint8),
                ('flags', C.c_uint64), ('comp', _AVComponentDescriptor * 4)]


class _AVPacketHead(C.Structure):
    _fields_ = [('buf', C.c_void_p), ('pts', C.c_int64), ('dts', C.c_int64), ('data', C.c_void_p), ('size', C.c_int),
                ('stream_index', C.c_int)]


_LIBS: dict = {}


def _libdir() -> str:
    import cv2
    base = os.path.dirname(os.path.dirname(cv2.__file__))
    for name in ('opencv_python_headless.libs', 'opencv_python.libs', 'opencv_contrib_python_headless.libs', 'opencv_contrib_python.libs'):
        d = os.path.join(base, name)
        if os.path.isdir(d):
            return d
    raise Unsupported('OpenCV-bundled FFmpeg libraries not found')


def _lib(stem: str) -> C.CDLL:
    if stem not in _LIBS:
        d = _libdir()
        for dep in {'avcodec': ('avutil', 'swresample'), 'avformat': ('avutil', 'swresample', 'avcodec'), 'swresample': ('avutil',), 'swscale': ('avutil',)}.get(stem, ()):
            _lib(dep)
        hits = sorted(glob.glob(os.path.join(d, f'lib{stem}-*.so*')))
        if not hits:
            raise Unsupported(f'lib{stem} is not bundled')
        _LIBS[stem] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
    return _LIBS[stem]


class AvReader:
    """Yields batches as (uint8 buffer in the hb_frames layout, n_frames, pix_fmt), like the readers in ``frames.py``."""

    kind = 'yuv'

    def __init__(self, path: Path):
        from .encoder import PIX_YUV420P8, PIX_YUV420P16
        self._fmt_ids = (PIX_YUV420P8, PIX_YUV420P16)
        try:
            self.avf, self.avc, self.avu = _lib('avformat'), _lib('avcodec'), _lib('avutil')
        except OSError as exc:
            raise Unsupported(str(exc))
        f, a, u = self.avf, self.avc, self.avu
        f.avformat_version.restype = C.c_uint
        if (f.avformat_version() >> 16) < 59:
            raise Unsupported('libavformat older than 59: AVStream layout differs')
        f.avformat_open_input.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_void_p, C.c_void_p]
        f.avformat_find_stream_info.argtypes = [C.c_void_p, C.c_void_p]
        f.av_find_best_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int]
        f.av_read_frame.argtypes = [C.c_void_p, C.c_void_p]
        f.avformat_close_input.argtypes = [C.POINTER(C.c_void_p)]
        a.avcodec_alloc_context3.restype = C.c_void_p
        a.avcodec_alloc_context3.argtypes = [C.c_void_p]
        a.avcodec_parameters_to_context.argtypes = [C.c_void_p, C.c_void_p]
        a.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        a.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
        a.av_packet_alloc.restype = C.c_void_p
        a.av_packet_free.argtypes = [C.POINTER(C.c_void_p)]
        a.av_packet_unref.argtypes = [C.c_void_p]
        a.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
        a.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
        u.av_frame_alloc.restype = C.c_void_p
        u.av_frame_free.argtypes = [C.POINTER(C.c_void_p)]
        u.av_frame_unref.argtypes = [C.c_void_p]
        u.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        u.av_log_set_level.argtypes = [C.c_int]
        u.av_log_set_level(16)
        self.fmt = C.c_void_p()
        self.ctx = C.c_void_p()
        self.pkt = C.c_void_p()
        self.frame = C.c_void_p()
        self._ring = None
        self.stream_tags = None        # colour description + HDR10 static metadata of the first decoded frame (_read_stream_tags)
        self._sws = None               # ((format, w, h), SwsContext, bytes per sample, library, format name) of the conversion in use
        if f.avformat_open_input(C.byref(self.fmt), str(path).encode(), None, None) < 0:
            self.fmt = C.c_void_p()
            raise Unsupported(f'avformat cannot open {path}')
        try:
            if f.avformat_find_stream_info(self.fmt, None) < 0:
                raise Unsupported('no stream info')
            dec = C.c_void_p()
            self.stream = f.av_find_best_stream(self.fmt, 0, -1, -1, C.byref(dec), 0)        # AVMEDIA_TYPE_VIDEO
            if self.stream < 0 or not dec:
                raise Unsupported('no decodable video stream')
            nb = C.c_uint.from_address(self.fmt.value + 44).value
            if not (0 < nb <= 256 and self.stream < nb):
                raise Unsupported('unexpected AVFormatContext layout')
            streams = C.c_void_p.from_address(self.fmt.value + 48).value
            st = C.c_void_p.from_address(streams + 8 * self.stream).value
            index = C.c_int.from_address(st + 8).value
            codecpar = C.c_void_p.from_address(st + 16).value
            if index != self.stream or not codecpar or C.c_int.from_address(codecpar).value != 0:      # codec_type == VIDEO
                raise Unsupported('unexpected AVStream layout')
            self.ctx = C.c_void_p(a.avcodec_alloc_context3(dec))
            if a.avcodec_parameters_to_context(self.ctx, codecpar) < 0:
                raise Unsupported('avcodec_parameters_to_context failed')
            u.av_opt_set(self.ctx, b'threads', b'auto', 0)
            if a.avcodec_open2(self.ctx, dec, None) < 0:
                raise Unsupported('avcodec_open2 failed')
            self.pkt = C.c_void_p(a.av_packet_alloc())
            self.frame = C.c_void_p(u.av_frame_alloc())
            self._eof = self._flushed = False
            self._pending = self._next_frame()           # decode one frame now: geometry and pixel format
            if self._pending is None:
                raise Unsupported('no frames decoded')
        except Exception:
            self.close()
            raise
        self.width, self.height, bps = self._pending[1], self._pending[2], self._pending[3]
        self.fmt_id = self._fmt_ids[0] if bps == 1 else self._fmt_ids[1]
        self.src_bit_depth = 8 if bps == 1 else 10
        self.pix_fmt = 'yuv420p' if bps == 1 else 'yuv420p10le'          # what this reader DELIVERS
        self.source_pix_fmt = self._sws[4] if self._sws else self.pix_fmt     # what the decoder produced
        self.frame_bytes = (self.width * self.height + 2 * (self.width // 2) * (self.height // 2)) * bps

    # ---- what ffprobe reports per stream: colour description (codec context, filled from the container or the bitstream's VUI once a
    #      frame is decoded; read through the AVOption API, no structure offsets) and the HDR10 static metadata carried as frame side
    #      data (SEI 137 / 144, or the container's equivalent): AVFrameSideData = {type, data, size, ...}; AVMasteringDisplayMetadata =
    #      {AVRational display_primaries[3][2] (R, G, B), white_point[2], min_luminance, max_luminance, int has_primaries, has_luminance};
    #      AVContentLightMetadata = {unsigned MaxCLL, MaxFALL} -- all unchanged since libavutil 55
    def _read_stream_tags(self) -> dict:
        u = self.avu
        out: dict = {}
        try:
            u.av_opt_get_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
            for key, opt in (('color_primaries', b'color_primaries'), ('color_transfer', b'color_trc'), ('color_space', b'colorspace'),
                             ('color_range', b'color_range')):
                v = C.c_int64(-1)
                if u.av_opt_get_int(self.ctx, opt, 0, C.byref(v)) >= 0 and 0 <= v.value < 256:
                    out[key] = int(v.value)
            u.av_frame_get_side_data.restype = C.c_void_p
            u.av_frame_get_side_data.argtypes = [C.c_void_p, C.c_int]

            def side(kind, need):
                sd = u.av_frame_get_side_data(self.frame, kind)
                if not sd:
                    return None
                data, size = C.c_void_p.from_address(sd + 8).value, C.c_size_t.from_address(sd + 16).value
                return data if data and size >= need else None
            md = side(11, 88)              # AV_FRAME_DATA_MASTERING_DISPLAY_METADATA
            if md:
                q = (C.c_int * 20).from_address(md)
                has_prim, has_lum = C.c_int.from_address(md + 80).value, C.c_int.from_address(md + 84).value

                def scaled(i, unit):
                    num, den = q[2 * i], q[2 * i + 1]
                    return int(round(num * unit / den)) if den else 0
                if has_prim == 1 and has_lum == 1:
                    r, g, b = ((scaled(2 * c, 50000), scaled(2 * c + 1, 50000)) for c in range(3))
                    wp = (scaled(6, 50000), scaled(7, 50000))
                    lmin, lmax = scaled(8, 10000), scaled(9, 10000)
                    out['master_display'] = 'G(%d,%d)B(%d,%d)R(%d,%d)WP(%d,%d)L(%d,%d)' % (g + b + r + wp + (lmax, lmin))
            cl = side(14, 8)               # AV_FRAME_DATA_CONTENT_LIGHT_LEVEL
            if cl:
                v = (C.c_uint * 2).from_address(cl)
                out['max_cll'] = '%d,%d' % (v[0], v[1])
        except Exception:                  # a probe nicety, never a reason to lose the reader
            pass
        return out

    # ---- a decoded frame in any YUV layout -> planar 4:2:0 at 8 bits (8-bit sources) or 10 bits (deeper ones) through libswscale
    def _sws_to_420(self, fr: _AVFrameHead, w: int, h: int):
        u = self.avu
        if self._sws is None or self._sws[0] != (fr.format, w, h):
            u.av_pix_fmt_desc_get.restype = C.POINTER(_AVPixFmtDescriptorHead)
            u.av_pix_fmt_desc_get.argtypes = [C.c_int]
            d = u.av_pix_fmt_desc_get(fr.format)
            if not d:
                raise Unsupported(f'pixel format {fr.format} unknown to libavutil')
            d = d.contents
            depth = d.comp[0].depth
            if not (d.name and 1 <= d.nb_components <= 4 and 8 <= depth <= 16 and d.log2_chroma_w <= 2 and d.log2_chroma_h <= 2):
                raise Unsupported(f'pixel format {fr.format}: unexpected descriptor layout')
            if d.flags & (_PIX_FLAG_RGB | _PIX_FLAG_PAL | _PIX_FLAG_BAYER):
                raise Unsupported(f'pixel format {d.name.decode()} is not YUV')
            try:
                sws = _lib('swscale')
            except OSError as exc:
                raise Unsupported(str(exc))
            sws.sws_getContext.restype = C.c_void_p
            sws.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
            sws.sws_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
            sws.sws_freeContext.argtypes = [C.c_void_p]
            bps = 1 if depth <= 8 else 2
            ctx = sws.sws_getContext(w, h, fr.format, w, h, AV_PIX_FMT_YUV420P if bps == 1 else AV_PIX_FMT_YUV420P10LE, _SWS_BICUBIC, None, None, None)
            if not ctx:
                raise Unsupported(f'libswscale cannot convert {d.name.decode()}')
            if self._sws is not None:
                sws.sws_freeContext(self._sws[1])
            self._sws = ((fr.format, w, h), ctx, bps, sws, d.name.decode())
        _, ctx, bps, sws, _ = self._sws
        luma, chroma = w * h * bps, (w // 2) * (h // 2) * bps
        out = np.empty(luma + 2 * chroma, np.uint8)
        base = out.ctypes.data
        dst = (C.c_void_p * 4)(base, base + luma, base + luma + chroma, None)
        dst_ls = (C.c_int * 4)(w * bps, (w // 2) * bps, (w // 2) * bps, 0)
        src = (C.c_void_p * 4)(*[fr.data[i] for i in range(4)])
        src_ls = (C.c_int * 4)(*[fr.linesize[i] for i in range(4)])
        if sws.sws_scale(ctx, src, src_ls, 0, h, dst, dst_ls) <= 0:
            raise Unsupported('sws_scale failed')
        return out, bps

    # ---- one decoded frame as (packed planes uint8, width, height, bytes per sample), or None at the end of the stream
    def _next_frame(self):
        a, f, u = self.avc, self.avf, self.avu
        while True:
            rc = a.avcodec_receive_frame(self.ctx, self.frame)
            if rc >= 0:
                fr = _AVFrameHead.from_address(self.frame.value)
                if self.stream_tags is None:
                    self.stream_tags = self._read_stream_tags()
                w, h = fr.width & ~1, fr.height & ~1
                if fr.format == AV_PIX_FMT_YUV420P:
                    bps = 1
                elif fr.format == AV_PIX_FMT_YUV420P10LE:
                    bps = 2
                else:                                        # any other YUV layout: libswscale, as the reference's ffmpeg child does
                    try:
                        out, bps = self._sws_to_420(fr, w, h)
                    finally:
                        u.av_frame_unref(self.frame)
                    return out, w, h, bps
                parts = []
                for i, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
                    ls = fr.linesize[i]
                    raw = np.frombuffer((C.c_uint8 * (ls * ph)).from_address(fr.data[i]), dtype=np.uint8).reshape(ph, ls)
                    parts.append(np.ascontiguousarray(raw[:, :pw * bps]).reshape(-1))
                out = np.concatenate(parts)
                u.av_frame_unref(self.frame)
                return out, w, h, bps
            if rc not in (_EAGAIN, _EOF):
                raise Unsupported(f'avcodec_receive_frame -> {rc}')
            if self._flushed:
                return None
            if self._eof:
                a.avcodec_send_packet(self.ctx, None)
                self._flushed = True
                continue
            rc = f.av_read_frame(self.fmt, self.pkt)
            if rc < 0:
                self._eof = True
                continue
            if _AVPacketHead.from_address(self.pkt.value).stream_index == self.stream:
                a.avcodec_send_packet(self.ctx, self.pkt)
            a.av_packet_unref(self.pkt)

    def batches(self, batch: int, ring: int = 4) -> Iterator[Tuple[np.ndarray, int, int]]:
        from .frames import _Ring
        bufs = self._ring = _Ring(batch * self.frame_bytes, ring)
        while True:
            buf = bufs.next().reshape(batch, self.frame_bytes)
            n = 0
            while n < batch:
                item, self._pending = (self._pending, None) if self._pending is not None else (self._next_frame(), None)
                if item is None:
                    break
                if item[0].size != self.frame_bytes:
                    raise Unsupported('geometry changes in mid-stream')
                buf[n] = item[0]
                n += 1
            if n == 0:
                return
            yield buf[:n], n, self.fmt_id
            if n < batch:
                return

    def close(self):
        ring, self._ring = getattr(self, '_ring', None), None
        if ring is not None:
            ring.close()
        sws, self._sws = getattr(self, '_sws', None), None
        if sws is not None:
            sws[3].sws_freeContext(sws[1])
        if getattr(self, 'frame', None) and self.frame.value:
            self.avu.av_frame_free(C.byref(self.frame))
        if getattr(self, 'pkt', None) and self.pkt.value:
            self.avc.av_packet_free(C.byref(self.pkt))
        if getattr(self, 'ctx', None) and self.ctx.value:
            self.avc.avcodec_free_context(C.byref(self.ctx))
        if getattr(self, 'fmt', None) and self.fmt.value:
            self.avf.avformat_close_input(C.byref(self.fmt))
        self.frame = self.pkt = self.ctx = self.fmt = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decoded_format(path: Path) -> Optional[dict]:
    """geometry, the decoder's pixel format and the stream's colour description / HDR10 static metadata (numeric H.273 codes, x265-style
    strings) as of the first decoded frame, or None when this reader cannot handle the file"""
    try:
        r = AvReader(path)
    except Exception:
        return None
    try:
        return {'width': r.width, 'height': r.height, 'pix_fmt': r.source_pix_fmt, **(r.stream_tags or {})}
    finally:
        r.close()


_LAYOUT_CHANNELS = {'mono': 1, 'stereo': 2, '2.1': 3, '3.0': 3, '3.0(back)': 3, '4.0': 4, 'quad': 4, 'quad(side)': 4, '3.1': 4, '5.0': 5,
                    '5.0(side)': 5, '4.1': 5, '5.1': 6, '5.1(side)': 6, '6.0': 6, '6.0(front)': 6, 'hexagonal': 6, '6.1': 7, '6.1(back)': 7,
                    '6.1(front)': 7, '7.0': 7, '7.0(front)': 7, '7.1': 8, '7.1(wide)': 8, '7.1(wide-side)': 8, 'octagonal': 8}


def audio_channels(path: Path) -> Optional[int]:
    """channel count of the best audio stream (0: the file has no audio), or None when libavformat cannot tell -- what ffprobe reports as
    ``channels`` (reference core/probe.py:47-111).  The count is read through the AVOption API of a scratch codec context
    (``ch_layout`` as a string: a layout name or "<n> channels"), not from structure offsets."""
    import re
    try:
        f, a, u = _lib('avformat'), _lib('avcodec'), _lib('avutil')
    except (OSError, Unsupported):
        return None
    f.avformat_version.restype = C.c_uint
    if (f.avformat_version() >> 16) < 59:
        return None
    f.avformat_open_input.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_void_p, C.c_void_p]
    f.avformat_find_stream_info.argtypes = [C.c_void_p, C.c_void_p]
    f.av_find_best_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int]
    f.avformat_close_input.argtypes = [C.POINTER(C.c_void_p)]
    a.avcodec_alloc_context3.restype = C.c_void_p
    a.avcodec_alloc_context3.argtypes = [C.c_void_p]
    a.avcodec_parameters_to_context.argtypes = [C.c_void_p, C.c_void_p]
    a.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
    u.av_opt_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    u.av_free.argtypes = [C.c_void_p]
    u.av_log_set_level.argtypes = [C.c_int]
    u.av_log_set_level(16)
    fmt, ctx = C.c_void_p(), C.c_void_p()
    if f.avformat_open_input(C.byref(fmt), str(path).encode(), None, None) < 0:
        return None
    try:
        if f.avformat_find_stream_info(fmt, None) < 0:
            return None
        idx = f.av_find_best_stream(fmt, 1, -1, -1, None, 0)              # AVMEDIA_TYPE_AUDIO
        if idx < 0:
            return 0
        nb = C.c_uint.from_address(fmt.value + 44).value
        if not (0 < nb <= 256 and idx < nb):
            return None
        st = C.c_void_p.from_address(C.c_void_p.from_address(fmt.value + 48).value + 8 * idx).value
        codecpar = C.c_void_p.from_address(st + 16).value
        if C.c_int.from_address(st + 8).value != idx or not codecpar or C.c_int.from_address(codecpar).value != 1:     # codec_type == AUDIO
            return None
        ctx = C.c_void_p(a.avcodec_alloc_context3(None))
        if not ctx or a.avcodec_parameters_to_context(ctx, codecpar) < 0:
            return None
        s = C.c_void_p()
        if u.av_opt_get(ctx, b'ch_layout', 0, C.byref(s)) < 0 or not s:
            return None
        name = C.cast(s, C.c_char_p).value.decode('ascii', 'replace')
        u.av_free(s)
        if name in _LAYOUT_CHANNELS:
            return _LAYOUT_CHANNELS[name]
        m = re.match(r'(\d+) channels', name)
        return int(m.group(1)) if m else name.count('+') + 1 if '+' in name else None
    finally:
        if ctx:
            a.avcodec_free_context(C.byref(ctx))
        f.avformat_close_input(C.byref(fmt))
