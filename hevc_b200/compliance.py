"""Apple-compliance self-check of an output file.

Replaces the external AppleHEVCValidator the reference shells out to (core/transcoder.py:35-68; absent from every
image, and its lock is undefined there).  Checks are made with this package's own MP4 / H.265 header parser:
hvc1 sample entry with VPS/SPS/PPS in hvcC (and none in-band), colr nclx present, profile Main/Main10, level and
tier as derived, VUI colour description, chroma_sample_loc_type, SEI 137/144 contents, IDR cadence."""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Tuple, Dict, List, Optional

from . import mp4


class BitReader:
    def __init__(self, data: bytes):
        self.d, self.pos = data, 0

    def u(self, n: int) -> int:
        v = 0
        for _ in range(n):
            byte = self.d[self.pos >> 3]
            v = (v << 1) | ((byte >> (7 - (self.pos & 7))) & 1)
            self.pos += 1
        return v

    def ue(self) -> int:
        z = 0
        while self.u(1) == 0:
            z += 1
        return (1 << z) - 1 + (self.u(z) if z else 0)

    def se(self) -> int:
        k = self.ue()
        return (k + 1) // 2 if k & 1 else -(k // 2)


def unescape(nal: bytes) -> bytes:
    out = bytearray()
    zeros = 0
    for b in nal:
        if zeros >= 2 and b == 3:
            zeros = 0
            continue
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


def parse_sps(nal: bytes) -> Dict[str, int]:
    r = BitReader(unescape(nal)[2:])
    s: Dict[str, int] = {}
    r.u(4)
    max_sub = r.u(3)
    r.u(1)
    r.u(2)
    s['tier'] = r.u(1)
    s['profile_idc'] = r.u(5)
    r.u(32)
    r.u(4)
    r.u(32); r.u(11); r.u(1)
    s['level_idc'] = r.u(8)
    assert max_sub == 0
    r.ue()
    s['chroma_format_idc'] = r.ue()
    s['coded_width'] = r.ue()
    s['coded_height'] = r.ue()
    s['width'], s['height'] = s['coded_width'], s['coded_height']
    if r.u(1):
        l, rr, t, b = r.ue(), r.ue(), r.ue(), r.ue()
        s['width'] -= 2 * (l + rr)
        s['height'] -= 2 * (t + b)
    s['bit_depth'] = r.ue() + 8
    r.ue()
    s['log2_max_poc_lsb'] = r.ue() + 4
    if r.u(1):
        r.ue(); r.ue(); r.ue()
    s['log2_min_cb'] = r.ue() + 3
    s['log2_ctb'] = s['log2_min_cb'] + r.ue()
    r.ue(); r.ue(); r.ue(); r.ue()
    if r.u(1):
        raise ValueError('scaling lists not expected')
    r.u(1)
    s['sao'] = r.u(1)
    if r.u(1):
        raise ValueError('pcm not expected')
    n_rps = r.ue()
    for i in range(n_rps):
        if i and r.u(1):
            raise ValueError('inter RPS not expected')
        neg, pos = r.ue(), r.ue()
        for _ in range(neg + pos):
            r.ue(); r.u(1)
    if r.u(1):
        raise ValueError('long-term refs not expected')
    r.u(1); r.u(1)
    s['vui'] = r.u(1)
    if s['vui']:
        if r.u(1):
            if r.u(8) == 255:
                r.u(32)
        if r.u(1):
            r.u(1)
        s['video_signal'] = r.u(1)
        if s['video_signal']:
            r.u(3)
            s['full_range'] = r.u(1)
            s['colour_description'] = r.u(1)
            if s['colour_description']:
                s['colour_primaries'], s['transfer_characteristics'], s['matrix_coeffs'] = r.u(8), r.u(8), r.u(8)
        s['chroma_loc_present'] = r.u(1)
        if s['chroma_loc_present']:
            s['chroma_loc'] = r.ue()
            r.ue()
        r.u(3)
        if r.u(1):
            r.ue(); r.ue(); r.ue(); r.ue()
        s['timing'] = r.u(1)
        if s['timing']:
            s['num_units_in_tick'], s['time_scale'] = r.u(32), r.u(32)
            if r.u(1):
                r.ue()
            s['hrd'] = r.u(1)
            if s['hrd']:
                # hrd_parameters(1, 0) (E.2.2): only the NAL HRD of the single sub-layer is expected here
                nal_hrd, vcl_hrd = r.u(1), r.u(1)
                if nal_hrd or vcl_hrd:
                    if r.u(1):                       # sub_pic_hrd_params_present_flag
                        r.u(8); r.u(5); r.u(1); r.u(5)
                    bit_rate_scale, cpb_size_scale = r.u(4), r.u(4)
                    r.u(5); r.u(5); r.u(5)           # delay / output-delay lengths
                fixed_general = r.u(1)
                fixed_cvs = 1 if fixed_general else r.u(1)
                if fixed_cvs:
                    r.ue()                           # elemental_duration_in_tc_minus1
                low_delay = 0 if fixed_cvs else r.u(1)
                cpb_cnt = 0 if low_delay else r.ue()
                if nal_hrd:
                    for _ in range(cpb_cnt + 1):
                        br, cpb = r.ue(), r.ue()
                        s['hrd_cbr'] = r.u(1)
                    s['hrd_bit_rate'] = (br + 1) << (6 + bit_rate_scale)
                    s['hrd_cpb_size'] = (cpb + 1) << (4 + cpb_size_scale)
    return s


def hrd_underflows(sample_bits: List[int], bit_rate: float, cpb_bits: float, fps: float, initial_fullness: float = 0.9) -> List[Tuple[int, float]]:
    """Leaky-bucket check of the ACTUAL access-unit sizes against the HRD a stream signals (Annex C, VBR): the CPB fills at
    ``bit_rate`` up to ``cpb_bits``, the first removal happens when it holds ``initial_fullness`` of its size (the
    buffering-period SEI's initial_cpb_removal_delay), one access unit leaves per frame interval.  -> [(frame, missing bits)]"""
    per, full, bad = bit_rate / fps, initial_fullness * cpb_bits, []
    for i, bits in enumerate(sample_bits):
        if bits > full + 1e-6:
            bad.append((i, bits - full))
        full = min(cpb_bits, max(0.0, full - bits) + per)
    return bad


def parse_sei(nal: bytes) -> Dict[int, bytes]:
    d = unescape(nal)[2:]
    out, pos = {}, 0
    while pos + 2 <= len(d) and d[pos] != 0x80:
        t = 0
        while d[pos] == 255:
            t += 255
            pos += 1
        t += d[pos]
        pos += 1
        n = 0
        while d[pos] == 255:
            n += 255
            pos += 1
        n += d[pos]
        pos += 1
        out[t] = d[pos:pos + n]
        pos += n
    return out


def inspect(data: bytes) -> Dict[str, object]:
    """Collect everything the checks look at from an MP4 byte string."""
    rep: Dict[str, object] = {'boxes': [], 'sample_entry': None}
    order = []
    for depth, kind, a, b in mp4.parse_boxes(data):
        rep['boxes'].append(kind.decode('latin1'))
        if depth == 0:
            order.append(kind)
        if kind == b'ftyp':
            rep['major_brand'] = data[a:a + 4].decode('latin1')
        if kind == b'hdlr':
            rep['handler_name'] = data[a + 24:b].rstrip(b'\0').decode('latin1')
        if kind == b'stsd':
            entry = data[a + 8:b]
            rep['sample_entry'] = entry[4:8].decode('latin1')
            pos = 8 + 78
            inner = {}
            while pos + 8 <= len(entry):
                size, k = struct.unpack('>I4s', entry[pos:pos + 8])
                inner[k.decode('latin1')] = entry[pos + 8:pos + size]
                pos += size
            rep['entry_boxes'] = inner
        if kind == b'stss':
            n = struct.unpack('>I', data[a + 4:a + 8])[0]
            rep['sync_samples'] = list(struct.unpack(f'>{n}I', data[a + 8:a + 8 + 4 * n]))
        if kind == b'stsz':
            rep['n_samples'] = struct.unpack('>I', data[a + 8:a + 12])[0]
            rep['sample_sizes'] = list(struct.unpack(f'>{rep["n_samples"]}I', data[a + 12:a + 12 + 4 * rep['n_samples']]))
        if kind == b'co64':
            rep['chunk_offset'] = struct.unpack('>Q', data[a + 8:a + 16])[0]
    rep['moov_before_mdat'] = order.index(b'moov') < order.index(b'mdat') if b'moov' in order and b'mdat' in order else False
    inner = rep.get('entry_boxes') or {}
    if 'hvcC' in inner:
        c = inner['hvcC']
        rep['hvcc_profile'] = c[1] & 31
        rep['hvcc_tier'] = (c[1] >> 5) & 1
        rep['hvcc_level'] = c[12]
        pos, arrays = 23, {}
        for _ in range(c[22]):
            t = c[pos] & 0x3f
            cnt = struct.unpack('>H', c[pos + 1:pos + 3])[0]
            pos += 3
            for _ in range(cnt):
                ln = struct.unpack('>H', c[pos:pos + 2])[0]
                arrays.setdefault(t, []).append(c[pos + 2:pos + 2 + ln])
                pos += 2 + ln
        rep['param_sets'] = arrays
        if 33 in arrays:
            rep['sps'] = parse_sps(arrays[33][0])
    # walk the first samples: NAL types in-band, SEI payloads
    nal_types, sei = [], {}
    off = rep.get('chunk_offset')
    if off is not None and rep.get('sample_sizes'):
        for size in rep['sample_sizes'][:2]:
            pos, end = off, off + size
            while pos + 4 <= end:
                ln = struct.unpack('>I', data[pos:pos + 4])[0]
                nal = data[pos + 4:pos + 4 + ln]
                t = mp4.nal_type(nal)
                nal_types.append(t)
                if t == 39:
                    sei.update(parse_sei(nal))
                pos += 4 + ln
            off += size
    rep['inband_nal_types'] = nal_types
    rep['sei'] = sei
    return rep


def check_bytes(data: bytes, expect: Optional[Dict[str, int]] = None) -> List[str]:
    """-> list of problems (empty = compliant).  ``expect`` may pin profile_idc / level_idc / tier / keyint / hdr10 / colour."""
    rep = inspect(data)
    bad: List[str] = []
    if rep.get('major_brand') != 'mp42':
        bad.append('major brand is not mp42')
    if not rep.get('moov_before_mdat'):
        bad.append('moov is not before mdat (faststart)')
    if rep.get('sample_entry') != 'hvc1':
        bad.append(f"sample entry is {rep.get('sample_entry')}, not hvc1")
    inner = rep.get('entry_boxes') or {}
    if 'hvcC' not in inner:
        bad.append('no hvcC')
    if 'colr' not in inner:
        bad.append('no colr box')
    ps = rep.get('param_sets') or {}
    if not all(t in ps for t in (32, 33, 34)):
        bad.append('hvcC lacks VPS/SPS/PPS')
    if any(t in (32, 33, 34) for t in rep.get('inband_nal_types', [])):
        bad.append('parameter sets found in-band in an hvc1 track')
    sps = rep.get('sps') or {}
    if sps:
        if sps['profile_idc'] not in (1, 2):
            bad.append('profile is not Main / Main10')
        if sps['profile_idc'] == 2 and sps['bit_depth'] != 10 or sps['profile_idc'] == 1 and sps['bit_depth'] != 8:
            bad.append('bit depth does not match the profile')
        if sps['chroma_format_idc'] != 1:
            bad.append('not 4:2:0')
        if not sps.get('colour_description'):
            bad.append('VUI colour description missing')
        if sps.get('full_range'):
            bad.append('full-range flag set')
        if rep.get('hvcc_level') != sps['level_idc'] or rep.get('hvcc_profile') != sps['profile_idc']:
            bad.append('hvcC profile/level differ from the SPS')
    else:
        bad.append('SPS not parsed')
    if sps.get('hrd_bit_rate') and sps.get('num_units_in_tick') and rep.get('sample_sizes'):
        # the stream signals an HRD (hrd=1, reference core/utils.py:65): its real sample sizes must not underflow that buffer
        fps = sps['time_scale'] / sps['num_units_in_tick']
        under = hrd_underflows([8 * n for n in rep['sample_sizes']], sps['hrd_bit_rate'], sps['hrd_cpb_size'], fps)
        if under:
            bad.append(f'HRD buffer underflow at sample {under[0][0]} ({int(under[0][1])} bits short; {len(under)} samples in all)')
    if rep.get('handler_name') != 'VideoHandler':
        bad.append('handler_name is not VideoHandler')
    if rep.get('sync_samples', [0])[0] != 1:
        bad.append('first sample is not a sync sample')
    if expect:
        for key in ('profile_idc', 'level_idc', 'tier'):
            if key in expect and sps and sps.get(key) != expect[key]:
                bad.append(f'{key} {sps.get(key)} != expected {expect[key]}')
        if 'keyint' in expect:
            ss = rep.get('sync_samples', [])
            if any((s - 1) % expect['keyint'] for s in ss):
                bad.append('IDR cadence is not a multiple of keyint')
        if expect.get('hdr10'):
            if sps and (sps.get('colour_primaries'), sps.get('transfer_characteristics'), sps.get('matrix_coeffs')) != (9, 16, 9):
                bad.append('VUI is not BT.2020 / PQ / BT.2020nc')
            if sps and sps.get('chroma_loc') != 0:
                bad.append('chroma_sample_loc_type is not 0')
            if 137 not in rep['sei'] or 144 not in rep['sei']:
                bad.append('HDR10 SEI (137/144) missing from the first access unit')
            if 35 not in rep.get('inband_nal_types', []):
                bad.append('no AUD')
            if 'mdcv' not in inner or 'clli' not in inner:
                bad.append('mdcv / clli boxes missing')
            if 'master_display' in expect and 137 in rep['sei']:
                md = struct.unpack('>8HII', rep['sei'][137][:24])
                if tuple(md) != tuple(expect['master_display']):
                    bad.append('mastering-display SEI differs from the requested values')
            if 'max_cll' in expect and 144 in rep['sei']:
                if struct.unpack('>HH', rep['sei'][144][:4]) != (expect['max_cll'], expect['max_fall']):
                    bad.append('content-light-level SEI differs')
    return bad


def check_file(path: Path, expect: Optional[Dict[str, int]] = None) -> List[str]:
    return check_bytes(Path(path).read_bytes(), expect)
