"""TEST INFRASTRUCTURE -- not part of the product path.

ctypes bindings to the FFmpeg 8 shared libraries bundled inside the OpenCV wheel
(SURVEY.md section 0.4 / 8c):

* ``HevcDecoder``  -- the native ``hevc`` decoder: the normative oracle.  "decoder output ==
  encoder reconstruction" pins inverse transform, dequantisation, intra/inter prediction and all
  of the syntax / CABAC at once.  ``crccheck+explode`` makes the decoder verify the
  decoded-picture-hash SEI itself.
* ``sws_convert``  -- libswscale with SWS_ACCURATE_RND|SWS_BITEXACT: the code the reference's CPU
  path runs for ``-pix_fmt`` (core/transcoder.py:464), used as the pixel-pipeline oracle.

Only the documented, layout-stable head of AVFrame / AVPacket is read; options go through
``av_opt_set``.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Iterator, List, Optional, Tuple

import numpy as np

AV_PIX_FMT_YUV420P = 0
AV_PIX_FMT_RGB24 = 2
AV_PIX_FMT_BGR24 = 3
AV_PIX_FMT_YUV444P = 5
AV_PIX_FMT_NV12 = 23
AV_PIX_FMT_YUV420P10LE = 62
AV_PIX_FMT_P010LE = 158
_PAD = 64  # AV_INPUT_BUFFER_PADDING_SIZE


def _libdir() -> str:
    import cv2
    base = os.path.dirname(os.path.dirname(cv2.__file__))
    for name in ('opencv_python_headless.libs', 'opencv_python.libs', 'opencv_contrib_python_headless.libs'):
        d = os.path.join(base, name)
        if os.path.isdir(d):
            return d
    raise OSError('OpenCV-bundled FFmpeg libraries not found')


_LIBS = {}


def _lib(stem: str) -> C.CDLL:
    if stem not in _LIBS:
        d = _libdir()
        if stem != 'avutil':
            _lib('avutil')
        if stem in ('avcodec',):
            _lib('swresample')
        hits = sorted(glob.glob(os.path.join(d, f'lib{stem}-*.so*')))
        if not hits:
            raise OSError(f'lib{stem} not bundled')
        _LIBS[stem] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
    return _LIBS[stem]


def available() -> bool:
    try:
        _lib('avcodec')
        _lib('swscale')
        return True
    except OSError:
        return False


class _AVFrameHead(C.Structure):
    _fields_ = [('data', C.c_void_p * 8), ('linesize', C.c_int * 8), ('extended_data', C.c_void_p),
                ('width', C.c_int), ('height', C.c_int), ('nb_samples', C.c_int), ('format', C.c_int)]


class _AVPacketHead(C.Structure):
    _fields_ = [('buf', C.c_void_p), ('pts', C.c_int64), ('dts', C.c_int64), ('data', C.c_void_p), ('size', C.c_int)]


class DecodeError(RuntimeError):
    pass


class HevcDecoder:
    """Decode an Annex-B HEVC elementary stream to planar frames (numpy, uint8 or uint16)."""

    def __init__(self, verify_hash: bool = True, threads: int = 1):
        self.avc = _lib('avcodec')
        self.avu = _lib('avutil')
        a, u = self.avc, self.avu
        a.avcodec_find_decoder_by_name.restype = C.c_void_p
        a.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
        a.avcodec_alloc_context3.restype = C.c_void_p
        a.avcodec_alloc_context3.argtypes = [C.c_void_p]
        a.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        a.avcodec_free_context.argtypes = [C.c_void_p]
        a.av_packet_alloc.restype = C.c_void_p
        a.av_packet_free.argtypes = [C.c_void_p]
        a.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
        a.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
        a.av_parser_init.restype = C.c_void_p
        a.av_parser_init.argtypes = [C.c_int]
        a.av_parser_close.argtypes = [C.c_void_p]
        a.av_parser_parse2.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                       C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64]
        u.av_frame_alloc.restype = C.c_void_p
        u.av_frame_free.argtypes = [C.c_void_p]
        u.av_frame_unref.argtypes = [C.c_void_p]
        u.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        u.av_log_set_level.argtypes = [C.c_int]
        u.av_log_set_level(16)  # AV_LOG_ERROR

        codec = a.avcodec_find_decoder_by_name(b'hevc')
        if not codec:
            raise OSError('hevc decoder missing from bundled libavcodec')
        self.ctx = a.avcodec_alloc_context3(codec)
        if verify_hash:
            if u.av_opt_set(self.ctx, b'err_detect', b'crccheck+explode', 0) < 0:
                raise OSError('err_detect option rejected')
        u.av_opt_set(self.ctx, b'threads', str(threads).encode(), 0)
        if a.avcodec_open2(self.ctx, codec, None) < 0:
            raise OSError('avcodec_open2 failed')
        self.parser = a.av_parser_init(173)  # AV_CODEC_ID_HEVC
        if not self.parser:
            raise OSError('hevc parser missing')
        self.pkt = a.av_packet_alloc()
        self.frame = u.av_frame_alloc()

    def close(self):
        if getattr(self, 'ctx', None):
            self.avc.av_parser_close(self.parser)
            p = C.c_void_p(self.pkt)
            self.avc.av_packet_free(C.byref(p))
            f = C.c_void_p(self.frame)
            self.avu.av_frame_free(C.byref(f))
            c = C.c_void_p(self.ctx)
            self.avc.avcodec_free_context(C.byref(c))
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _drain(self) -> Iterator[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        while True:
            rc = self.avc.avcodec_receive_frame(self.ctx, self.frame)
            if rc < 0:
                if rc in (-11, -541478725):  # EAGAIN, AVERROR_EOF
                    return
                raise DecodeError(f'avcodec_receive_frame -> {rc}')
            fr = _AVFrameHead.from_address(self.frame)
            if fr.format == AV_PIX_FMT_YUV420P:
                dt, bps = np.uint8, 1
            elif fr.format == AV_PIX_FMT_YUV420P10LE:
                dt, bps = np.uint16, 2
            else:
                raise DecodeError(f'unexpected pixel format {fr.format}')
            planes = []
            for i, (w, h) in enumerate(((fr.width, fr.height), ((fr.width + 1) // 2, (fr.height + 1) // 2),
                                        ((fr.width + 1) // 2, (fr.height + 1) // 2))):
                ls = fr.linesize[i]
                raw = (C.c_uint8 * (ls * h)).from_address(fr.data[i])
                arr = np.frombuffer(raw, dtype=np.uint8).reshape(h, ls)[:, :w * bps]
                planes.append(np.ascontiguousarray(arr).view(dt).reshape(h, w).copy())
            self.avu.av_frame_unref(self.frame)
            yield tuple(planes)

    def decode(self, stream: bytes) -> List[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        """Decode a whole Annex-B elementary stream; returns frames in output (display) order.
        Access units are split here (one packet per AU) rather than through the libavcodec parser."""
        out: List[Tuple[np.ndarray, np.ndarray, np.ndarray]] = []
        for au in split_access_units(stream):
            buf = C.create_string_buffer(au + b'\0' * _PAD, len(au) + _PAD)
            pk = _AVPacketHead.from_address(self.pkt)
            pk.data, pk.size = C.addressof(buf), len(au)
            rc = self.avc.avcodec_send_packet(self.ctx, self.pkt)
            pk.data, pk.size = None, 0
            if rc < 0:
                raise DecodeError(f'avcodec_send_packet -> {rc} (bitstream rejected or picture-hash mismatch)')
            out.extend(self._drain())
        self.avc.avcodec_send_packet(self.ctx, None)
        out.extend(self._drain())
        return out


def iter_nals(stream: bytes):
    """Yield (offset_of_start_code, nal_type, payload_offset) for every NAL unit of an Annex-B stream."""
    i, n = 0, len(stream)
    while True:
        j = stream.find(b'\x00\x00\x01', i)
        if j < 0 or j + 5 > n:
            return
        sc = j - 1 if j > 0 and stream[j - 1] == 0 else j
        yield sc, (stream[j + 3] >> 1) & 0x3f, j + 3
        i = j + 3


def split_access_units(stream: bytes) -> List[bytes]:
    """Cut an Annex-B stream into access units (H.265 7.4.2.4.4, restricted to the NAL types used here)."""
    cuts, seen_vcl = [], False
    for sc, typ, pay in iter_nals(stream):
        is_vcl = typ < 32
        first_slice = is_vcl and pay + 2 < len(stream) and (stream[pay + 2] & 0x80) != 0
        starts = typ in (32, 33, 34, 35, 39) or first_slice
        if not cuts or (seen_vcl and starts):
            cuts.append(sc)
            seen_vcl = False
        if is_vcl:
            seen_vcl = True
    cuts.append(len(stream))
    return [bytes(stream[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]


def decode_hevc(stream: bytes, verify_hash: bool = True):
    dec = HevcDecoder(verify_hash=verify_hash)
    try:
        return dec.decode(stream)
    finally:
        dec.close()


# ------------------------------------------------------------------ libswscale

_SWS_FLAGS = 0x10 | 0x40000 | 0x80000 | 0x2000 | 0x4000  # POINT|ACCURATE_RND|BITEXACT|FULL_CHR_H_INT|FULL_CHR_H_INP
SWS_BILINEAR, SWS_BICUBIC, SWS_POINT, SWS_LANCZOS = 2, 4, 0x10, 0x200
SWS_CS_ITU709, SWS_CS_BT2020 = 1, 9


def sws_convert(src_planes, src_fmt: int, dst_fmt: int, width: int, height: int, dst_shapes,
                dst_w: Optional[int] = None, dst_h: Optional[int] = None, algo: int = SWS_POINT,
                colorspace: Optional[int] = None):
    """Run libswscale on numpy planes.  ``dst_shapes``: [(rows, row_bytes, dtype)] per output plane."""
    s = _lib('swscale')
    s.sws_getContext.restype = C.c_void_p
    s.sws_getContext.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    s.sws_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    s.sws_freeContext.argtypes = [C.c_void_p]
    s.sws_getCoefficients.restype = C.POINTER(C.c_int)
    s.sws_getCoefficients.argtypes = [C.c_int]
    s.sws_setColorspaceDetails.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
    dst_w, dst_h = dst_w or width, dst_h or height
    flags = (algo | 0x40000 | 0x80000 | 0x2000 | 0x4000)
    ctx = s.sws_getContext(width, height, src_fmt, dst_w, dst_h, dst_fmt, flags, None, None, None)
    if not ctx:
        raise OSError('sws_getContext failed')
    try:
        if colorspace is not None:
            coef = s.sws_getCoefficients(colorspace)
            s.sws_setColorspaceDetails(ctx, coef, 1, coef, 0, 0, 1 << 16, 1 << 16)  # src full-range RGB, dst limited
        src = [np.ascontiguousarray(p) for p in src_planes]
        sp = (C.c_void_p * 4)(*[p.ctypes.data for p in src], *([None] * (4 - len(src))))
        ss = (C.c_int * 4)(*[p.strides[0] for p in src], *([0] * (4 - len(src))))
        dst = [np.zeros((rows, rb // np.dtype(dt).itemsize), dtype=dt) for rows, rb, dt in dst_shapes]
        dp = (C.c_void_p * 4)(*[p.ctypes.data for p in dst], *([None] * (4 - len(dst))))
        ds = (C.c_int * 4)(*[p.strides[0] for p in dst], *([0] * (4 - len(dst))))
        s.sws_scale(ctx, sp, ss, 0, height, dp, ds)
        return dst
    finally:
        s.sws_freeContext(ctx)
