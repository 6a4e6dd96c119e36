"""Apple-HEVC parameter derivation for the B200 backend.

Restates the reference's L2 derivation layer (core/transcoder.py:136-412,
core/utils.py:29-70) so that the B200 encoder signals exactly what the
reference's libx265 command line would have requested.  Every function that
shares a name with the reference returns the same value for the same
``VideoInfo`` -- including its quirks: tier is always
``main`` on the x265 path, the VBV clamp uses the mis-scaled level table, CRF
depends on clip length.  tests/test_derive.py checks this against golden
vectors produced by importing the reference (tests/golden/make_derive_golden.py).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from fractions import Fraction
from typing import List, Optional, Tuple

from .probe import VideoInfo

# level -> (max luma picture samples, max luma sample rate, "max bitrate" bps, "max cpb" bits,
#           main-tier kbps, high-tier kbps).  Values are the reference's (core/transcoder.py:158-172),
# including the columns that are ~1/4 of the real H.265 limits; they are part of the contract.
_LEVEL_ROWS = (
    ('1',   36864,     552960,      64000,      4608, 128,    128),
    ('2',   122880,    3686400,     150000,     18432, 1500,   3000),
    ('2.1', 245760,    7372800,     300000,     36864, 3000,   6000),
    ('3',   552960,    16588800,    600000,     61440, 6000,   12000),
    ('3.1', 983040,    33177600,    1200000,    122880, 10000,  20000),
    ('4',   2228224,   66846720,    3000000,    245760, 12000,  30000),
    ('4.1', 2228224,   133693440,   6000000,    491520, 20000,  50000),
    ('5',   8912896,   267386880,   12000000,   983040, 25000,  100000),
    ('5.1', 8912896,   534773760,   24000000,   1966080, 40000,  160000),
    ('5.2', 8912896,   1069547520,  48000000,   3932160, 60000,  240000),
    ('6',   35651584,  1069547520,  48000000,   3932160, 60000,  240000),
    ('6.1', 35651584,  2139095040,  96000000,   7864320, 120000, 480000),
    ('6.2', 35651584,  4278190080,  192000000,  15728640, 240000, 800000),
)
HEVC_LEVEL_LIMITS = {name: (ps, sr, br, cpb_bytes * 8, mt, ht) for name, ps, sr, br, cpb_bytes, mt, ht in _LEVEL_ROWS}


def calculate_apple_hevc_level(info: VideoInfo) -> Tuple[str, str]:
    """(level, tier) for the x265 path -- reference core/transcoder.py:174-187.

    The tier test compares luma samples/second with a kbps column, so it can
    never select ``high`` for real video; that behaviour is preserved."""
    pic = info.width * info.height
    rate = round(pic * info.fps)
    wide = max(info.width, info.height) >= 3840
    for name, (max_pic, max_rate, _br, _cpb, _main_kbps, high_kbps) in HEVC_LEVEL_LIMITS.items():
        if pic <= max_pic and rate <= max_rate:
            tier = 'main'
            if (info.hdr or wide or info.fps > 60) and rate <= high_kbps:
                tier = 'high'
            return name, tier
    return '6.2', 'main'


def calculate_nvenc_hevc_level(info: VideoInfo) -> Tuple[str, str, str, str]:
    """(level, tier, profile, pix_fmt) for the NVENC path -- reference core/transcoder.py:189-209."""
    longest = max(info.width, info.height)
    level = '4.0' if longest <= 1920 else '4.1' if longest <= 2560 else '5.1' if longest <= 3840 else '5.2'
    if info.hdr:
        return level, 'high', 'main10', 'p010le'
    return level, 'main', 'main', 'yuv420p'


def compute_aligned_gop(fps: float, preferred_gop_sec: float, max_gop_frames: int = 240) -> int:
    """Key-frame interval in frames, aligned to whole seconds (reference core/transcoder.py:211-260)."""
    fps = max(1.0, fps)
    want = max(2, min(preferred_gop_sec * fps, max_gop_frames))
    try:
        ratio = Fraction(str(fps)).limit_denominator(1001)
        num, den = ratio.numerator, ratio.denominator
    except Exception:
        num, den = int(round(fps)), 1

    best: Optional[int] = None
    gap = float('inf')
    for seconds in range(1, 9):
        frames = round(num * seconds / den)
        if 2 <= frames <= max_gop_frames and abs(frames - want) < gap:
            best, gap = frames, abs(frames - want)
    if best is None:
        best = max(2, min(int(round(want)), max_gop_frames))

    if abs(round(fps) - fps) < 1e-6:          # integer frame rate: snap to fps * n
        whole = int(round(fps))
        best = max(2, min(whole * max(1, round(best / whole)), max_gop_frames))
    else:                                      # NTSC-style rates: snap to round(fps * whole seconds)
        secs = max(1, round(best / fps))
        best = min(max_gop_frames, max(2, round(fps * secs)))
    return best


_CRF_BY_HEIGHT = ((480, 17), (720, 18), (1080, 19), (1440, 20), (2160, 21), (4320, 22))


def _target_kbps(longest: int, hdr: bool) -> int:
    if longest >= 7680:
        return 140000
    if longest >= 3840:
        return 65000 if hdr else 50000
    if longest >= 2560:
        return 30000 if hdr else 26000
    if longest >= 1920:
        return 19000 if hdr else 16000
    return 10000 if hdr else 8000


def calculate_dynamic_values(info: VideoInfo, use_nvenc: bool = True, gpu_name: str = '') -> Tuple[int, int, int, int, int]:
    """(crf, cq, vbv_maxrate_kbps, vbv_bufsize_kbit, gop_frames) -- reference core/transcoder.py:263-354."""
    longest = max(info.width, info.height)
    fps = float(info.fps) if info.fps else 30.0
    hdr = bool(info.hdr)

    crf = _CRF_BY_HEIGHT[-1][1]
    for limit, value in _CRF_BY_HEIGHT:
        if info.height <= limit:
            crf = value
            break
    if hdr:
        crf = max(8, crf - 1)

    if info.nb_frames:
        frames = info.nb_frames
    elif info.duration:
        frames = int(round(info.duration * fps))
    else:
        frames = int(round(60 * fps))
    density = frames / (info.width * info.height + 1)      # the reference's "motion density"
    busy, calm = density > 0.00025, density < 0.00006
    if busy:
        crf += 1
    elif calm:
        crf = max(8, crf - 1)
    crf = max(16, min(crf, 24))
    cq = crf + 1

    kbps = _target_kbps(longest, hdr)
    if busy:
        kbps = int(kbps * 1.15)
    elif calm:
        kbps = int(kbps * 0.92)
    maxrate = int(kbps)
    bufsize = int(maxrate * 1.5)

    level, _tier = calculate_apple_hevc_level(info)
    if level in HEVC_LEVEL_LIMITS:
        _, _, lvl_bps, lvl_cpb_bits, _, _ = HEVC_LEVEL_LIMITS[level]
        maxrate = min(maxrate, int(int(lvl_bps / 1000) * 0.98))
        bufsize = min(bufsize, max(int(maxrate * 1.2), int(int(lvl_cpb_bits / 1000) * 0.9)))

    if hdr:
        gop_sec = 2.0 if longest >= 3840 else 2.5
    else:
        gop_sec = 2.5 if longest >= 3840 else 3.0
    if fps > 60:
        gop_sec *= 1.05
    gop = compute_aligned_gop(fps, gop_sec, max_gop_frames=240)
    if abs(round(fps) - fps) < 1e-6:
        whole = int(round(fps))
        gop = max(2, min(240, whole * max(1, round(gop / whole))))
    return crf, cq, maxrate, bufsize, gop


DEFAULT_MASTER_DISPLAY = 'G(13250,34500)B(7500,3000)R(34000,16000)WP(15635,16450)L(10000000,50)'
DEFAULT_MAX_CLL = '1000,400'


def build_hdr_metadata(master_display: str, max_cll: str, use_nvenc: bool, fps: float = 30.0) -> List[str]:
    """HDR10 option list, same strings as reference core/utils.py:29-70."""
    md = (master_display or '').strip() or DEFAULT_MASTER_DISPLAY
    cll = (max_cll or '').strip() or DEFAULT_MAX_CLL
    if use_nvenc:
        out: List[str] = []
        for key, val in (('color_primaries', 'bt2020'), ('color_trc', 'smpte2084'), ('colorspace', 'bt2020nc'),
                         ('master_display', md), ('max_cll', cll)):
            out += ['-metadata:s:v:0', f'{key}={val}']
        return out + ['-color_primaries', 'bt2020', '-color_trc', 'smpte2084', '-colorspace', 'bt2020nc']
    opts = ['hdr10=1', 'colorprim=bt2020', 'transfer=smpte2084', 'colormatrix=bt2020nc',
            f'master-display={md}', f'max-cll={cll}', 'hrd=1', 'aud=1', 'chromaloc=0', 'repeat-headers=1']
    return ['-x265-params', ':'.join(opts)]


@dataclass
class FFmpegParams:
    # reference core/transcoder.py:25-33
    vcodec: str
    pix_fmt: str
    profile: str
    level: str
    color_flags: List[str]
    vparams: List[str]
    hdr_metadata: List[str]


def x265_option_list(info: VideoInfo) -> List[str]:
    """The ``-x265-params`` entries the reference's CPU branch builds (core/transcoder.py:398-410)."""
    level, tier = calculate_apple_hevc_level(info)
    crf, _cq, maxrate, bufsize, gop = calculate_dynamic_values(info, False, '')
    profile = 'main10' if info.hdr else 'main'
    opts = [f'crf={crf}', 'preset=slow', 'log-level=error', 'nal-hrd=vbr', f'vbv-maxrate={maxrate}',
            f'vbv-bufsize={bufsize}', f'tier={tier}', f'keyint={gop}', f'min-keyint={max(2, int(gop // 2))}',
            f'profile={profile}', 'level-idc=' + str(level)]
    if info.hdr:
        opts += build_hdr_metadata(info.master_display, info.max_cll, use_nvenc=False, fps=info.fps)[1].split(':')
    return opts


# ---------------------------------------------------------------- B200 encoder parameters

_PRIMARIES = {'bt709': 1, 'bt470bg': 5, 'smpte170m': 6, 'bt2020': 9, 'bt2020-ncl': 9}
_TRANSFER = {'bt709': 1, 'smpte170m': 6, 'smpte2084': 16, 'pq': 16, 'arib-std-b67': 18, 'hlg': 18,
             'bt2020-10': 14, 'bt2020-12': 15}
_MATRIX = {'bt709': 1, 'bt470bg': 5, 'smpte170m': 6, 'bt2020': 9, 'bt2020nc': 9, 'bt2020-ncl': 9, 'bt2020c': 10}

_MD_RE = re.compile(r'G\((\d+),(\d+)\)B\((\d+),(\d+)\)R\((\d+),(\d+)\)WP\((\d+),(\d+)\)L\((\d+),(\d+)\)')


def parse_master_display(text: str) -> Tuple[int, ...]:
    """x265 ``master-display`` string -> (Gx,Gy,Bx,By,Rx,Ry,WPx,WPy,Lmax,Lmin) in SEI-137 units
    (chromaticity 0.00002, luminance 0.0001 cd/m2).  Malformed text selects the reference default."""
    m = _MD_RE.search(text or '') or _MD_RE.search(DEFAULT_MASTER_DISPLAY)
    return tuple(int(g) for g in m.groups())


def parse_max_cll(text: str) -> Tuple[int, int]:
    try:
        a, b = (int(t) for t in (text or '').split(','))
        return a, b
    except Exception:
        a, b = DEFAULT_MAX_CLL.split(',')
        return int(a), int(b)


@dataclass
class B200Params:
    """What the encode step needs; mirrors ``hb_enc_params`` in include/hevc_b200.h field for field."""
    width: int
    height: int
    fps_num: int
    fps_den: int
    bit_depth: int            # 8 (Main) or 10 (Main10)
    profile_idc: int          # 1 Main, 2 Main10
    level_idc: int            # 30 * level
    tier: int                 # 0 main, 1 high
    crf: int
    vbv_maxrate_kbps: int
    vbv_bufsize_kbit: int
    keyint: int
    min_keyint: int
    colour_primaries: int
    transfer_characteristics: int
    matrix_coeffs: int
    chroma_loc: int = 0
    full_range: int = 0       # '-color_range tv' core/transcoder.py:490
    aud: int = 0
    repeat_headers: int = 0
    hrd: int = 0
    hdr10: int = 0
    master_display: Tuple[int, ...] = field(default_factory=lambda: (0,) * 10)
    max_cll: int = 0
    max_fall: int = 0
    vui_colour: int = 1

    @property
    def profile(self) -> str:
        return 'main10' if self.profile_idc == 2 else 'main'


def derive_b200_params(info: VideoInfo, force_main10: bool = False) -> B200Params:
    """Translate the reference's x265 option list for ``info`` into encoder parameters.

    ``force_main10`` raises only the coding bit depth (profile Main10, 10-bit samples) for an SDR source -- the upscale
    path's scaler writes P010 -- and leaves every colour / HDR10 signalling decision to the source's own tags: bit depth
    and HDR are separate properties.

    Equivalent of ``build_ffmpeg_params(info, use_nvenc=False, ...)`` (core/transcoder.py:357-412) for a
    backend that takes structured parameters instead of an argv.  Options the reference passes that
    libx265 ignores (``nal-hrd``, ``tier``) have no effect here either; the SDR path therefore carries no
    AUD / HRD / repeated headers while the HDR path carries all of them."""
    level, tier = calculate_apple_hevc_level(info)
    crf, _cq, maxrate, bufsize, gop = calculate_dynamic_values(info, False, '')
    ratio = Fraction(str(float(info.fps) if info.fps else 30.0)).limit_denominator(1001)
    hdr = bool(info.hdr)
    p = B200Params(
        width=info.width, height=info.height, fps_num=ratio.numerator, fps_den=ratio.denominator,
        bit_depth=10 if hdr else 8, profile_idc=2 if hdr else 1,
        level_idc=int(round(float(level) * 30)), tier=1 if tier == 'high' else 0,
        crf=crf, vbv_maxrate_kbps=maxrate, vbv_bufsize_kbit=bufsize, keyint=gop, min_keyint=max(2, int(gop // 2)),
        colour_primaries=_PRIMARIES.get(info.color_primaries, 2), transfer_characteristics=_TRANSFER.get(info.color_transfer, 2),
        matrix_coeffs=_MATRIX.get(info.color_space, 2))
    if force_main10 and not hdr:
        p.bit_depth, p.profile_idc = 10, 2
    if hdr:
        # hdr10=1:colorprim=bt2020:transfer=smpte2084:colormatrix=bt2020nc:...:hrd=1:aud=1:chromaloc=0:repeat-headers=1
        p.colour_primaries, p.transfer_characteristics, p.matrix_coeffs = 9, 16, 9
        p.hdr10 = p.hrd = p.aud = p.repeat_headers = 1
        p.chroma_loc = 0
        p.master_display = parse_master_display(info.master_display)
        p.max_cll, p.max_fall = parse_max_cll(info.max_cll)
    return p
