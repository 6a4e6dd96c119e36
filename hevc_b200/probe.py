"""Media probing for the B200 encode backend.

Mirrors the reference's ``core/probe.py`` interface (``VideoInfo`` :9-24,
``probe_media`` :47-116): same 14 fields, same HDR rule (>= 2 of 4 features,
:76-82), same "never raises, falls back to 1080p30 SDR" convention (:114-116).

The reference shells out to ``ffprobe``; this image has none, so three probers
are chained: ffprobe (if on PATH) -> Y4M / raw sidecar header -> OpenCV
container probe.  A ``<file>.json`` sidecar can carry colour tags for raw or
OpenCV-decoded inputs.
"""
from __future__ import annotations

import json
import logging
import shutil
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

logger = logging.getLogger(__name__)


@dataclass
class VideoInfo:
    # field order and defaults follow reference core/probe.py:9-24
    width: int
    height: int
    fps: float
    color_primaries: str
    color_transfer: str
    color_space: str
    pix_fmt: str
    master_display: str
    max_cll: str
    audio_channels: int
    hdr: bool = False
    audio_language: Optional[str] = 'eng'
    nb_frames: Optional[int] = None
    duration: Optional[float] = None


# reference core/probe.py:26-29
HDR_PIXFMTS = frozenset({'yuv420p10le', 'p010le', 'yuv444p10le'})
HDR_COLOR_SPACES = frozenset({'bt2020', 'bt2020-ncl', 'bt2020nc'})
HDR_TRANSFERS = frozenset({'smpte2084', 'pq'})
HDR_PRIMARIES = frozenset({'bt2020', 'bt2020-ncl'})

_FALLBACK = dict(width=1920, height=1080, fps=30.0, color_primaries='bt709',
                 color_transfer='bt709', color_space='bt709', pix_fmt='yuv420p',
                 master_display='', max_cll='', audio_channels=2, hdr=False,
                 audio_language='eng', nb_frames=None, duration=None)


def parse_fps(rate_str: str) -> float:
    """'num/den' -> float; anything malformed -> 30.0 (reference core/probe.py:31-38)."""
    try:
        if not rate_str or '/' not in rate_str:
            return 30.0
        num, den = (int(t) for t in rate_str.split('/'))
        return num / den if den else 30.0
    except Exception:
        return 30.0


def classify_hdr(color_primaries: str, color_transfer: str, color_space: str, pix_fmt: str) -> bool:
    """HDR when at least two of the four features match (reference core/probe.py:76-82)."""
    hits = ((color_primaries in HDR_PRIMARIES) + (color_transfer in HDR_TRANSFERS)
            + (color_space in HDR_COLOR_SPACES) + (pix_fmt in HDR_PIXFMTS))
    return hits >= 2


def _first_tag(tags: dict, *names: str, default: str = '') -> str:
    for n in names:
        val = tags.get(n)
        if val:
            return val
    return default


def info_from_ffprobe_json(doc: dict) -> VideoInfo:
    """Interpret an ``ffprobe -show_streams -show_format`` document exactly as the
    reference does (core/probe.py:53-111)."""
    streams = doc.get('streams', [])
    v = next((s for s in streams if s.get('codec_type') == 'video'), None)
    if not v:
        raise ValueError('no video stream')
    width = int(v.get('width') or 1920)
    height = int(v.get('height') or 1080)
    rate = v.get('avg_frame_rate') or v.get('r_frame_rate') or '30/1'
    if rate.strip() == '0/0' or not rate.strip():
        fps = 30.0
    else:
        try:
            num, den = (int(t) for t in rate.split('/'))
            fps = num / den if den else 30.0
        except Exception:
            fps = 30.0
    fmt = doc.get('format', {})
    tags = fmt.get('tags', {}) or {}

    def colour(key: str) -> str:
        return (v.get(key) or tags.get(key.upper()) or tags.get(key) or 'bt709').lower()

    prim, trc, spc = colour('color_primaries'), colour('color_transfer'), colour('color_space')
    pix_fmt = (v.get('pix_fmt') or '').lower()
    hdr = classify_hdr(prim, trc, spc, pix_fmt)
    master_display = _first_tag(tags, 'master-display', 'MASTER_DISPLAY', 'master_display', 'mastering_display')
    max_cll = _first_tag(tags, 'max-cll', 'MAX_CLL', 'max_cll')

    a = next((s for s in streams if s.get('codec_type') == 'audio'), None)
    if a:
        atags = a.get('tags', {}) or {}
        lang = atags.get('language') or atags.get('LANGUAGE') or 'eng'
        channels = int(a.get('channels', a.get('CHANNELS', 2)))
    else:
        lang, channels = None, 0

    try:
        nb_frames = int(v.get('nb_frames')) if v.get('nb_frames') else None
    except Exception:
        nb_frames = None
    try:
        duration = float(fmt.get('duration')) if fmt.get('duration') else None
    except Exception:
        duration = None
    return VideoInfo(width, height, fps, prim, trc, spc, pix_fmt, master_display, max_cll,
                     channels, hdr, lang, nb_frames, duration)


# ---------------------------------------------------------------- ffprobe-less probers

_Y4M_CSP = {'420': 'yuv420p', '420jpeg': 'yuv420p', '420mpeg2': 'yuv420p', '420paldv': 'yuv420p',
            '420p10': 'yuv420p10le', '444': 'yuv444p', '444p10': 'yuv444p10le', '422': 'yuv422p', '422p10': 'yuv422p10le',
            '420p12': 'yuv420p12le', '422p12': 'yuv422p12le', '444p12': 'yuv444p12le'}


def _probe_y4m(path: Path) -> dict:
    with open(path, 'rb') as fh:
        head = fh.readline(512)
    if not head.startswith(b'YUV4MPEG2'):
        raise ValueError('not y4m')
    out = {'pix_fmt': 'yuv420p'}
    for tok in head.decode('ascii', 'replace').split()[1:]:
        k, val = tok[0], tok[1:]
        if k == 'W':
            out['width'] = int(val)
        elif k == 'H':
            out['height'] = int(val)
        elif k == 'F':
            n, d = val.split(':')
            out['fps'] = int(n) / int(d) if int(d) else 30.0
        elif k == 'C':
            out['pix_fmt'] = _Y4M_CSP.get(val, 'gray' if val.startswith('mono') else 'unknown')
    out['header_len'] = len(head)
    bps = 2 if out['pix_fmt'].endswith('le') else 1
    w, h = out['width'], out['height']
    if out['pix_fmt'].startswith('yuv444'):
        fsz = w * h * 3 * bps
    elif out['pix_fmt'].startswith('yuv422'):
        fsz = w * h * 2 * bps
    elif out['pix_fmt'].startswith('yuv420'):
        fsz = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)) * bps
    else:                         # mono / unknown tags: the size of a frame is not known here, the container reader finds out
        out['frame_bytes'] = 0
        out['nb_frames'] = None
        return out
    out['frame_bytes'] = fsz
    out['nb_frames'] = (path.stat().st_size - len(head)) // (fsz + 6)  # 'FRAME\n' per frame
    return out


def _sane_count(v) -> Optional[int]:
    """OpenCV reports garbage (negative, or astronomically large) frame counts for streams without an index, e.g. raw Annex-B files"""
    try:
        n = int(v)
    except (TypeError, ValueError, OverflowError):
        return None
    return n if 0 < n < (1 << 31) else None


def _probe_cv2(path: Path) -> dict:
    import cv2  # deferred: optional dependency
    cap = cv2.VideoCapture(str(path))
    try:
        if not cap.isOpened():
            raise ValueError('cv2 cannot open')
        out = {'width': int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), 'height': int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)),
               'fps': float(cap.get(cv2.CAP_PROP_FPS)) or 30.0, 'nb_frames': _sane_count(cap.get(cv2.CAP_PROP_FRAME_COUNT)),
               'pix_fmt': 'yuv420p'}
        if out['width'] <= 0 or out['height'] <= 0:
            raise ValueError('cv2 reports no video')
        return out
    finally:
        cap.release()


_PRIM_NAMES = {1: 'bt709', 5: 'bt470bg', 6: 'smpte170m', 9: 'bt2020'}
_TRC_NAMES = {1: 'bt709', 6: 'smpte170m', 14: 'bt2020-10', 15: 'bt2020-12', 16: 'smpte2084', 18: 'arib-std-b67'}
_SPC_NAMES = {1: 'bt709', 5: 'bt470bg', 6: 'smpte170m', 9: 'bt2020nc', 10: 'bt2020c'}


def _probe_mp4_colour(path: Path) -> dict:
    """colour description and HDR10 static metadata from an ISO-BMFF file's video sample entry: ``colr`` (nclx), ``mdcv``, ``clli``
    -- what ffprobe reports as color_primaries / color_transfer / color_space and the mastering-display side data
    (reference core/probe.py:47-111).  Only ``moov`` is read; returns {} for anything unexpected."""
    import struct
    out: dict = {}
    containers = (b'moov', b'trak', b'mdia', b'minf', b'stbl')

    def walk(buf: bytes, start: int, end: int, depth: int):
        pos = start
        while pos + 8 <= end:
            size, kind = struct.unpack('>I4s', buf[pos:pos + 8])
            hdr = 8
            if size == 1:
                size = struct.unpack('>Q', buf[pos + 8:pos + 16])[0]
                hdr = 16
            if size < hdr or pos + size > end:
                return
            if kind in containers:
                walk(buf, pos + hdr, pos + size, depth + 1)
            elif kind == b'stsd':
                entry = pos + hdr + 8                                  # version/flags + entry_count
                esize, etype = struct.unpack('>I4s', buf[entry:entry + 8])
                if etype in (b'hvc1', b'hev1', b'avc1', b'avc3', b'av01', b'vp09', b'apch', b'apcn', b'ap4h') and esize >= 86:
                    inner(buf, entry + 86, min(entry + esize, end))
            pos += size

    def inner(buf: bytes, pos: int, end: int):
        while pos + 8 <= end:
            size, kind = struct.unpack('>I4s', buf[pos:pos + 8])
            if size < 8 or pos + size > end:
                return
            body = buf[pos + 8:pos + size]
            if kind == b'colr' and body[:4] in (b'nclx', b'nclc') and len(body) >= 10:
                p_, t_, m_ = struct.unpack('>HHH', body[4:10])
                out['color_primaries'] = _PRIM_NAMES.get(p_, 'unknown')
                out['color_transfer'] = _TRC_NAMES.get(t_, 'unknown')
                out['color_space'] = _SPC_NAMES.get(m_, 'unknown')
            elif kind == b'mdcv' and len(body) >= 24:
                v = struct.unpack('>8HII', body[:24])
                out['master_display'] = 'G(%d,%d)B(%d,%d)R(%d,%d)WP(%d,%d)L(%d,%d)' % v
            elif kind == b'clli' and len(body) >= 4:
                out['max_cll'] = '%d,%d' % struct.unpack('>HH', body[:4])
            pos += size

    try:
        with open(path, 'rb') as fh:
            fsize = path.stat().st_size
            pos = 0
            while pos + 8 <= fsize:
                fh.seek(pos)
                head = fh.read(16)
                size, kind = struct.unpack('>I4s', head[:8])
                hdr = 8
                if size == 1:
                    size = struct.unpack('>Q', head[8:16])[0]
                    hdr = 16
                elif size == 0:
                    size = fsize - pos
                if size < hdr:
                    break
                if kind == b'moov' and size <= (256 << 20):
                    fh.seek(pos + hdr)
                    buf = fh.read(size - hdr)
                    walk(buf, 0, len(buf), 0)
                    break
                pos += size
    except Exception:
        logger.debug('mp4 colour probe failed for %s', path, exc_info=True)
        return {}
    return out


def _sidecar(path: Path) -> dict:
    side = path.with_suffix(path.suffix + '.json')
    if side.exists():
        try:
            return json.loads(side.read_text())
        except Exception:
            logger.debug('bad sidecar %s', side, exc_info=True)
    return {}


def _probe_without_ffprobe(path: Path) -> VideoInfo:
    side = _sidecar(path)
    try:
        base = _probe_y4m(path)
    except Exception:
        if path.suffix.lower() in ('.yuv', '.raw', '.p010') and side:
            base = {}
        else:
            base = _probe_cv2(path)
            # what OpenCV does not tell: the decoder's pixel format (bit depth) and the container's colour / HDR10 boxes
            try:
                from .avreader import decoded_format
                fmt = decoded_format(path)
                if fmt:
                    base['pix_fmt'] = fmt['pix_fmt']
                    # stream-level colour description and HDR10 static metadata (any container: MKV, TS, MP4 ...), as ffprobe's
                    # color_primaries / color_transfer / color_space and side_data_list (reference core/probe.py:47-111); 2 = unspecified
                    for key, names in (('color_primaries', _PRIM_NAMES), ('color_transfer', _TRC_NAMES), ('color_space', _SPC_NAMES)):
                        if fmt.get(key) in names:
                            base[key] = names[fmt[key]]
                    for key in ('master_display', 'max_cll'):
                        if fmt.get(key):
                            base[key] = fmt[key]
            except Exception:
                logger.debug('decoded-format probe failed', exc_info=True)
            if path.suffix.lower() in ('.mp4', '.mov', '.m4v'):
                base.update(_probe_mp4_colour(path))
            try:                  # ffprobe's audio `channels` (core/probe.py:47-111): the b200 path warns that it writes no audio track
                from .avreader import audio_channels
                ach = audio_channels(path)
                if ach:
                    base['audio_channels'] = ach
            except Exception:
                logger.debug('audio probe failed', exc_info=True)
    base.update({k: v for k, v in side.items() if v is not None})
    prim = str(base.get('color_primaries', 'bt709')).lower()
    trc = str(base.get('color_transfer', 'bt709')).lower()
    spc = str(base.get('color_space', 'bt709')).lower()
    pix = str(base.get('pix_fmt', 'yuv420p')).lower()
    fps = float(base.get('fps', 30.0))
    nb = base.get('nb_frames')
    dur = base.get('duration')
    if dur is None and nb and fps:
        dur = nb / fps
    achan = int(base.get('audio_channels', 0))
    return VideoInfo(int(base['width']), int(base['height']), fps, prim, trc, spc, pix,
                     str(base.get('master_display', '')), str(base.get('max_cll', '')), achan,
                     classify_hdr(prim, trc, spc, pix), base.get('audio_language', 'eng' if achan else None),
                     int(nb) if nb else None, float(dur) if dur else None)


def probe_media(file_path: Path) -> VideoInfo:
    """Same contract as reference core/probe.py:47: returns a VideoInfo, never raises."""
    file_path = Path(file_path)
    try:
        if shutil.which('ffprobe'):
            res = subprocess.run(['ffprobe', '-v', 'quiet', '-print_format', 'json', '-show_streams',
                                  '-show_format', str(file_path)],
                                 capture_output=True, text=True, check=True, encoding='utf-8')
            return info_from_ffprobe_json(json.loads(res.stdout))
        return _probe_without_ffprobe(file_path)
    except Exception as exc:
        logger.error('probe failed: %s, %s', file_path.name, exc)
        return VideoInfo(**_FALLBACK)
