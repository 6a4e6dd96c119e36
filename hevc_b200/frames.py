"""Source-frame readers for the B200 encode path (the decode step of the reference's ffmpeg child).

The image has no ffmpeg binary, so inputs are read natively: Y4M (8- and 10-bit 4:2:0), raw planar .yuv with a JSON
sidecar, or anything OpenCV's bundled FFmpeg can decode (BGR frames, converted on the GPU by hb_rgb_to_yuv420)."""
from __future__ import annotations

from pathlib import Path
from typing import Iterator, Tuple

import numpy as np

from . import _cabi
from .encoder import PIX_YUV420P8, PIX_YUV420P16
from .probe import VideoInfo, _probe_y4m


class _Ring:
    """``count`` page-locked buffers of ``nbytes`` from the process-wide pool, handed out round-robin"""

    def __init__(self, nbytes: int, count: int):
        self.nbytes, self.count, self.bufs, self.i = nbytes, max(2, count), [], 0

    def next(self) -> np.ndarray:
        if len(self.bufs) < self.count:
            self.bufs.append(_cabi.pinned_array(self.nbytes))
            self.i = len(self.bufs) - 1
            return self.bufs[-1]
        self.i = (self.i + 1) % self.count
        return self.bufs[self.i]

    def close(self):
        for b in self.bufs:
            _cabi.pinned_release(b)
        self.bufs = []


class Y4MReader:
    """Yields batches as (uint8 buffer in hb_frames layout, n_frames, pix_fmt)."""

    def __init__(self, path: Path):
        self.path = Path(path)
        self.meta = _probe_y4m(self.path)
        if self.meta['pix_fmt'] not in ('yuv420p', 'yuv420p10le'):       # open_reader then goes through avreader (libswscale)
            raise ValueError(f"Y4M pixel format {self.meta['pix_fmt']} is not read natively")
        self.fmt = PIX_YUV420P16 if '10' in self.meta['pix_fmt'] else PIX_YUV420P8
        self.src_bit_depth = 10 if self.fmt == PIX_YUV420P16 else 8
        self.frame_bytes = self.meta['frame_bytes']
        self.kind = 'yuv'

    def batches(self, batch: int, ring: int = 4) -> Iterator[Tuple[np.ndarray, int, int]]:
        """``ring`` page-locked buffers are filled in turn: a yielded batch stays valid until ``ring - 1`` more were yielded.
        The buffers go back to the pool in ``close()`` -- call it only when the encoder no longer reads from them."""
        bufs = self._ring = _Ring(batch * self.frame_bytes, ring)
        with open(self.path, 'rb') as fh:
            fh.seek(self.meta['header_len'])
            while True:
                buf = bufs.next().reshape(batch, self.frame_bytes)
                n = 0
                while n < batch:
                    line = fh.readline()
                    if not line.startswith(b'FRAME'):
                        break
                    got = fh.readinto(memoryview(buf[n]))
                    if got != self.frame_bytes:
                        break
                    n += 1
                if n == 0:
                    return
                yield buf[:n], n, self.fmt
                if n < batch:
                    return

    def close(self):
        ring, self._ring = getattr(self, '_ring', None), None
        if ring is not None:
            ring.close()

    __del__ = close


class RawYuvReader:
    """Headerless planar 4:2:0; geometry and bit depth come from the probe (sidecar JSON)."""

    def __init__(self, path: Path, info: VideoInfo):
        self.path = Path(path)
        bps = 2 if '10' in info.pix_fmt else 1
        self.fmt = PIX_YUV420P16 if bps == 2 else PIX_YUV420P8
        self.src_bit_depth = 10 if bps == 2 else 8
        self.frame_bytes = (info.width * info.height + 2 * (info.width // 2) * (info.height // 2)) * bps
        self.kind = 'yuv'

    def batches(self, batch: int, ring: int = 4):
        bufs = self._ring = _Ring(batch * self.frame_bytes, ring)
        with open(self.path, 'rb') as fh:
            while True:
                buf = bufs.next()
                got = fh.readinto(memoryview(buf))
                n = got // self.frame_bytes
                if n == 0:
                    return
                yield buf[:n * self.frame_bytes].reshape(n, self.frame_bytes), n, self.fmt
                if n < batch:
                    return

    def close(self):
        ring, self._ring = getattr(self, '_ring', None), None
        if ring is not None:
            ring.close()

    __del__ = close


class Cv2Reader:
    """Container files through cv2.VideoCapture: yields BGR frame batches [n, h, w, 3] (kind == 'bgr')."""

    def __init__(self, path: Path):
        import cv2
        self.cap = cv2.VideoCapture(str(path))
        if not self.cap.isOpened():
            raise ValueError(f'cannot open {path}')
        self.kind = 'bgr'
        self.src_bit_depth = 8

    def batches(self, batch: int, ring: int = 0):
        while True:
            frames = []
            while len(frames) < batch:
                ok, f = self.cap.read()
                if not ok:
                    break
                frames.append(f)
            if not frames:
                self.cap.release()
                return
            yield np.stack(frames), len(frames), -1
            if len(frames) < batch:
                self.cap.release()
                return

    def close(self):
        pass


class MemoryReader:
    """Frames already in host memory (uint8 array [n, frame_bytes] in the hb_frames layout): the reader interface over a buffer,
    for callers that decode elsewhere and for benchmarks that must not time a disk."""

    def __init__(self, frames: np.ndarray, fmt: int = PIX_YUV420P8, src_bit_depth: int = 8):
        self.frames, self.fmt, self.src_bit_depth, self.kind = frames, fmt, src_bit_depth, 'yuv'

    def batches(self, batch: int, ring: int = 0):
        for s in range(0, len(self.frames), batch):
            chunk = self.frames[s:s + batch]
            yield chunk, len(chunk), self.fmt

    def close(self):
        pass


def open_reader(path: Path, info: VideoInfo):
    path = Path(path)
    with open(path, 'rb') as fh:
        magic = fh.read(9)
    if magic == b'YUV4MPEG2':
        try:
            return Y4MReader(path)
        except ValueError:        # C422 / C444 / C4xxp12 ...: libavformat's y4m demuxer + libswscale (avreader) below
            pass
    if path.suffix.lower() in ('.yuv', '.raw'):
        return RawYuvReader(path, info)
    try:                          # containers: the decoder's own 4:2:0 samples at their native depth when the bundled FFmpeg allows it
        from .avreader import AvReader
        return AvReader(path)
    except Exception:
        return Cv2Reader(path)    # 8-bit BGR through OpenCV (other chroma formats, RGB codecs, unknown library layout)


def write_y4m(path: Path, frames, width: int, height: int, fps: Tuple[int, int] = (30, 1), ten_bit: bool = False):
    """Test/fixture helper: frames = iterable of (y, u, v) planes."""
    with open(path, 'wb') as fh:
        fh.write(f'YUV4MPEG2 W{width} H{height} F{fps[0]}:{fps[1]} Ip A1:1 C{"420p10" if ten_bit else "420"}\n'.encode())
        for y, u, v in frames:
            fh.write(b'FRAME\n')
            for p in (y, u, v):
                fh.write(np.ascontiguousarray(p, dtype='<u2' if ten_bit else np.uint8).tobytes())
