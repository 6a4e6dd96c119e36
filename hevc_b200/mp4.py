"""Minimal ISO-BMFF (MP4) writer for one HEVC video track.

Stands in for the mov muxer inside the reference's ffmpeg child and reproduces the container-level contract
of its command line (core/transcoder.py:452-495): ``-tag:v hvc1`` (:466) -> ``hvc1`` sample entry with the
parameter sets in ``hvcC``; ``-brand mp42`` (:491); ``-movflags +write_colr+faststart`` (:492) -> ``colr nclx``
box and ``moov`` before ``mdat``; ``handler_name=VideoHandler`` (:414).  HDR10 streams also get ``mdcv``/``clli``.
Audio is not handled."""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

NAL_VPS, NAL_SPS, NAL_PPS, NAL_AUD = 32, 33, 34, 35


def split_nals(annexb) -> List[memoryview]:
    """Annex-B byte stream -> NAL units (without start codes, emulation prevention untouched) as zero-copy views."""
    data = bytes(annexb) if not isinstance(annexb, (bytes, bytearray)) else annexb
    view = memoryview(data)
    out, i, n = [], 0, len(data)
    starts = []
    find = data.find
    while True:
        j = find(b'\x00\x00\x01', i)
        if j < 0:
            break
        starts.append(j)
        i = j + 3
    for k, s in enumerate(starts):
        end = starts[k + 1] if k + 1 < len(starts) else n
        if k + 1 < len(starts):
            while end > s + 3 and data[end - 1] == 0:       # trailing zero_byte of the next start code
                end -= 1
        out.append(view[s + 3:end])
    return out


def nal_type(nal: bytes) -> int:
    return (nal[0] >> 1) & 0x3f


def split_access_units(annexb: bytes) -> List[List[bytes]]:
    """Group NAL units into access units (AUD / parameter set / prefix SEI / first slice after a VCL NAL starts a new one)."""
    aus: List[List[bytes]] = []
    seen_vcl = False
    for nal in split_nals(annexb):
        t = nal_type(nal)
        is_vcl = t < 32
        first_slice = is_vcl and len(nal) > 2 and (nal[2] & 0x80) != 0
        if not aus or (seen_vcl and (t in (32, 33, 34, 35, 39) or first_slice)):
            aus.append([])
            seen_vcl = False
        aus[-1].append(nal)
        seen_vcl = seen_vcl or is_vcl
    return aus


def box(kind: bytes, *payload: bytes) -> bytes:
    body = b''.join(payload)
    return struct.pack('>I4s', 8 + len(body), kind) + body


def full_box(kind: bytes, version: int, flags: int, *payload: bytes) -> bytes:
    return box(kind, struct.pack('>I', (version << 24) | flags), *payload)


@dataclass
class TrackInfo:
    width: int
    height: int
    fps_num: int
    fps_den: int
    profile_idc: int
    level_idc: int
    tier: int
    bit_depth: int
    colour_primaries: int = 2
    transfer_characteristics: int = 2
    matrix_coeffs: int = 2
    full_range: int = 0
    master_display: Optional[Sequence[int]] = None      # Gx,Gy,Bx,By,Rx,Ry,WPx,WPy,Lmax,Lmin
    max_cll: int = 0
    max_fall: int = 0


def hvcc(info: TrackInfo, vps: bytes, sps: bytes, pps: bytes) -> bytes:
    compat = (1 << (31 - info.profile_idc)) | ((1 << (31 - 2)) if info.profile_idc == 1 else 0)
    constraint = 0x9 << 44                           # progressive_source + frame_only_constraint
    rec = struct.pack('>B', 1)
    rec += struct.pack('>B', (0 << 6) | (info.tier << 5) | info.profile_idc)
    rec += struct.pack('>I', compat)
    rec += constraint.to_bytes(6, 'big')
    rec += struct.pack('>B', info.level_idc)
    rec += struct.pack('>H', 0xf000 | 0)             # min_spatial_segmentation_idc
    rec += struct.pack('>B', 0xfc | 3)               # parallelismType 3: entropy-coding-sync (WPP)
    rec += struct.pack('>B', 0xfc | 1)               # chroma_format_idc 4:2:0
    rec += struct.pack('>B', 0xf8 | (info.bit_depth - 8))
    rec += struct.pack('>B', 0xf8 | (info.bit_depth - 8))
    rec += struct.pack('>H', 0)                      # avgFrameRate unspecified
    rec += struct.pack('>B', (0 << 6) | (1 << 3) | (1 << 2) | 3)   # constantFrameRate 0, 1 temporal layer, nested, 4-byte lengths
    rec += struct.pack('>B', 3)
    for t, nal in ((NAL_VPS, vps), (NAL_SPS, sps), (NAL_PPS, pps)):
        rec += struct.pack('>BHH', 0x80 | t, 1, len(nal)) + nal
    return box(b'hvcC', rec)


def sample_entry(info: TrackInfo, vps: bytes, sps: bytes, pps: bytes) -> bytes:
    visual = struct.pack('>6xH', 1)                                  # reserved, data_reference_index
    visual += struct.pack('>HH12x', 0, 0)                            # pre_defined, reserved, pre_defined[3]
    visual += struct.pack('>HH', info.width, info.height)
    visual += struct.pack('>II', 0x00480000, 0x00480000)             # 72 dpi
    visual += struct.pack('>IH', 0, 1)                               # reserved, frame_count
    name = b'hevc_b200'
    visual += bytes([len(name)]) + name + b'\0' * (31 - len(name))   # compressorname
    visual += struct.pack('>Hh', 0x0018, -1)                         # depth, pre_defined
    extra = hvcc(info, vps, sps, pps)
    extra += box(b'colr', b'nclx', struct.pack('>HHHB', info.colour_primaries, info.transfer_characteristics, info.matrix_coeffs,
                                               0x80 if info.full_range else 0))
    if info.master_display:
        md = list(info.master_display)
        extra += box(b'mdcv', struct.pack('>8HII', *md[:8], md[8], md[9]))
        extra += box(b'clli', struct.pack('>HH', info.max_cll, info.max_fall))
    extra += box(b'pasp', struct.pack('>II', 1, 1))
    return box(b'hvc1', visual, extra)


def _head(info: TrackInfo, vps: bytes, sps: bytes, pps: bytes, sizes: Sequence[int], sync: Sequence[int]) -> bytes:
    """ftyp + moov (sample tables for one chunk holding every sample) + the 64-bit mdat header: everything before the samples."""
    n = len(sizes)
    payload_len = sum(sizes)
    timescale, delta = info.fps_num, info.fps_den
    duration = n * delta
    matrix = struct.pack('>9I', 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)

    stbl_fixed = [
        full_box(b'stsd', 0, 0, struct.pack('>I', 1), sample_entry(info, vps, sps, pps)),
        full_box(b'stts', 0, 0, struct.pack('>III', 1, n, delta)),
        full_box(b'stss', 0, 0, struct.pack('>I', len(sync)), b''.join(struct.pack('>I', s) for s in sync)),
        full_box(b'stsc', 0, 0, struct.pack('>IIII', 1, 1, n, 1)),       # one chunk holding every sample
        full_box(b'stsz', 0, 0, struct.pack('>II', 0, n), struct.pack('>%dI' % n, *sizes)),
    ]

    def moov(chunk_offset: int) -> bytes:
        stco = full_box(b'co64', 0, 0, struct.pack('>IQ', 1, chunk_offset))
        stbl = box(b'stbl', *stbl_fixed, stco)
        dinf = box(b'dinf', full_box(b'dref', 0, 0, struct.pack('>I', 1), full_box(b'url ', 0, 1)))
        minf = box(b'minf', full_box(b'vmhd', 0, 1, struct.pack('>HHHH', 0, 0, 0, 0)), dinf, stbl)
        hdlr = full_box(b'hdlr', 0, 0, struct.pack('>I4s12x', 0, b'vide'), b'VideoHandler\0')
        mdhd = full_box(b'mdhd', 0, 0, struct.pack('>IIIIHH', 0, 0, timescale, duration, 0x55c4, 0))
        mdia = box(b'mdia', mdhd, hdlr, minf)
        tkhd = full_box(b'tkhd', 0, 3, struct.pack('>IIIII8xHHHH', 0, 0, 1, 0, duration, 0, 0, 0, 0), matrix,
                        struct.pack('>II', info.width << 16, info.height << 16))
        mvhd = full_box(b'mvhd', 0, 0, struct.pack('>IIIIIH10x', 0, 0, timescale, duration, 0x10000, 0x100), matrix,
                        struct.pack('>24xI', 2))
        return box(b'moov', mvhd, box(b'trak', tkhd, mdia))

    ftyp = box(b'ftyp', b'mp42', struct.pack('>I', 0), b'mp42', b'isom', b'iso2')
    head = ftyp + moov(0)
    head = ftyp + moov(len(head) + 16)                                   # 64-bit mdat header
    return head + struct.pack('>I4sQ', 1, b'mdat', 16 + payload_len)


def _sample_pieces(nals, sets: dict):
    """length-prefixed sample of one access unit as a list of pieces (no copy) + its size; parameter sets go to ``sets``
    (``hvc1`` keeps them out of band)"""
    pieces, size = [], 0
    for nal in nals:
        t = nal_type(nal)
        if t in (NAL_VPS, NAL_SPS, NAL_PPS):
            if t not in sets:
                sets[t] = bytes(nal)
        else:
            pieces.append(struct.pack('>I', len(nal)))
            pieces.append(nal)
            size += 4 + len(nal)
    return pieces, size


def _sample_of(nals, sets: dict) -> bytes:
    return b''.join(_sample_pieces(nals, sets)[0])


def to_samples(annexb):
    """Annex-B run of whole access units -> (payload in MP4 sample form, sample sizes, 1-based indices of sync samples within
    the run, parameter sets).  Independent runs (closed-GOP segments coded elsewhere) convert in parallel and ``assemble``
    only concatenates."""
    sets: dict = {}
    pieces, sizes, sync = [], [], []
    for au in split_access_units(annexb):
        pc, size = _sample_pieces(au, sets)
        pieces += pc
        sizes.append(size)
        if any(16 <= nal_type(x) <= 23 for x in au):
            sync.append(len(sizes))
    return b''.join(pieces), sizes, sync, sets


def assemble(info: TrackInfo, runs) -> bytes:
    """runs: iterable of ``to_samples`` results in presentation order -> the complete file, moov first"""
    sets: dict = {}
    payloads, sizes, sync = [], [], []
    for payload, sz, sy, st in runs:
        for t, nal in st.items():
            sets.setdefault(t, nal)
        sync += [len(sizes) + k for k in sy]
        sizes += sz
        payloads.append(payload)
    if len(sets) != 3:
        raise ValueError('stream carries no VPS/SPS/PPS')
    return b''.join([_head(info, sets[NAL_VPS], sets[NAL_SPS], sets[NAL_PPS], sizes, sync)] + payloads)


def mux(info: TrackInfo, access_units: Iterable[Tuple[List[bytes], bool]]) -> bytes:
    """access_units: (NAL units of one frame, is_sync).  Parameter sets are taken from the first access unit that carries
    them and are removed from the samples (``hvc1`` keeps them out of band).  Returns the complete file, moov first."""
    sets: dict = {}
    pieces, sizes = [], []
    sync: List[int] = []
    for i, (nals, is_sync) in enumerate(access_units):
        pc, size = _sample_pieces(nals, sets)
        pieces += pc
        sizes.append(size)
        if is_sync:
            sync.append(i + 1)
    if len(sets) != 3:
        raise ValueError('stream carries no VPS/SPS/PPS')
    return b''.join([_head(info, sets[NAL_VPS], sets[NAL_SPS], sets[NAL_PPS], sizes, sync)] + pieces)


class StreamMuxer:
    """Incremental form of ``mux_annexb`` for long files: ``feed`` takes runs of whole access units as the encoder delivers
    them and spools the samples to ``<out>.mdat.tmp``; only the sample-size table stays in memory.  On a clean exit the file
    is written moov-first (faststart, core/transcoder.py:492) and the spool is appended and removed."""

    def __init__(self, info: TrackInfo, out_path):
        from pathlib import Path
        self.info, self.out_path = info, Path(out_path)
        self.spool_path = self.out_path.with_name(self.out_path.name + '.mdat.tmp')
        self.sets: dict = {}
        self.sizes: List[int] = []
        self.sync: List[int] = []
        self._spool = None
        self._aborted = False

    def __enter__(self):
        self._spool = open(self.spool_path, 'wb')
        return self

    def abort(self):
        """cancelled / failed: leave no output file behind"""
        self._aborted = True

    def feed(self, annexb: bytes):
        if not annexb:
            return
        payload, sizes, sync, sets = to_samples(annexb)
        for t, nal in sets.items():
            self.sets.setdefault(t, nal)
        self._spool.write(payload)
        self.sync += [len(self.sizes) + k for k in sync]
        self.sizes += sizes

    def __exit__(self, exc_type, exc, tb):
        import shutil
        self._spool.close()
        try:
            if exc_type is None and self.sizes and not self._aborted:
                if len(self.sets) != 3:
                    raise ValueError('stream carries no VPS/SPS/PPS')
                with open(self.out_path, 'wb') as out, open(self.spool_path, 'rb') as spool:
                    out.write(_head(self.info, self.sets[NAL_VPS], self.sets[NAL_SPS], self.sets[NAL_PPS], self.sizes, self.sync))
                    shutil.copyfileobj(spool, out, 1 << 24)
        finally:
            try:
                self.spool_path.unlink()
            except OSError:
                pass
        return False


def mux_annexb(info: TrackInfo, annexb: bytes) -> bytes:
    return assemble(info, [to_samples(annexb)])


def parse_boxes(data: bytes, start: int = 0, end: Optional[int] = None, depth: int = 0):
    """Tiny box walker used by the compliance checks: yields (path-depth, type, payload_start, payload_end)."""
    end = len(data) if end is None else end
    pos = start
    containers = {b'moov', b'trak', b'mdia', b'minf', b'stbl', b'dinf'}
    while pos + 8 <= end:
        size, kind = struct.unpack('>I4s', data[pos:pos + 8])
        hdr = 8
        if size == 1:
            size = struct.unpack('>Q', data[pos + 8:pos + 16])[0]
            hdr = 16
        if size < hdr:
            break
        yield depth, kind, pos + hdr, pos + size
        if kind in containers:
            yield from parse_boxes(data, pos + hdr, pos + size, depth + 1)
        pos += size
