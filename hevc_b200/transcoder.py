"""Per-file transcode driver with the B200 backend as a third encoder choice.

Drop-in for the reference's ``core/transcoder.py``: ``convert_video`` keeps its signature, result dictionary,
progress / cancel / never-raise conventions (:537-638); ``build_ffmpeg_params`` / ``build_ffmpeg_command`` /
``run_ffmpeg`` keep building and running the reference's libx265 / NVENC command lines (:357-535) for
``encoder='cpu'|'nvenc'``; ``encoder='b200'`` replaces the ffmpeg child by the CUDA encoder and the MP4 writer.
The default ``encoder='auto'`` preserves the old behaviour unless the environment variable HEVC_B200_ENCODER says
otherwise, so existing callers (gui/worker.py:30-41) keep working unchanged."""
from __future__ import annotations

import logging
import os
import shutil
import subprocess
import threading
from collections import OrderedDict
from pathlib import Path
from typing import Sequence, Any, Callable, Dict, List, Optional, Tuple

from . import derive
from .derive import FFmpegParams, calculate_dynamic_values
from .probe import VideoInfo, probe_media

logger = logging.getLogger(__name__)

ProgressCb = Optional[Callable[[str, int, int], None]]

# ------------------------------------------------------------------ reference command-line branch (a3, a8-a10)

NVENC_RETRIES = [
    {'-bf': '3', '-b_ref_mode': 'middle'},
    {'-bf': '0', '-b_ref_mode': 'disabled'},
    {'-bf': '0', '-b_ref_mode': 'disabled', '-temporal-aq': '0'},
    {'-bf': '0', '-b_ref_mode': 'disabled', '-temporal-aq': '0', '-spatial-aq': '0'},
]
VIDEO_METADATA_FLAGS = ['-metadata:s:v:0', 'handler_name=VideoHandler']


def has_nvenc() -> bool:
    """core/utils.py:9-15"""
    try:
        res = subprocess.run(['ffmpeg', '-hide_banner', '-encoders'], capture_output=True, text=True, check=True, encoding='utf-8')
        return 'hevc_nvenc' in res.stdout
    except Exception:
        return False


def detect_gpu_type() -> str:
    """core/utils.py:17-27 (cached by the caller's lifetime here: one nvidia-smi spawn per process)"""
    if not hasattr(detect_gpu_type, '_name'):
        try:
            res = subprocess.run(['nvidia-smi', '--query-gpu=name', '--format=csv,noheader'], capture_output=True, text=True, check=True,
                                 encoding='utf-8')
            detect_gpu_type._name = res.stdout.strip().lower()
        except Exception:
            detect_gpu_type._name = 'unknown'
    return detect_gpu_type._name


def decide_encoder(info: VideoInfo, force_cpu: bool, force_gpu: bool) -> bool:
    """True -> NVENC (core/transcoder.py:70-75)"""
    return False if force_cpu else has_nvenc()


def select_nvenc_preset(info: VideoInfo, gpu_name: str) -> str:
    res = max(info.width, info.height)
    tiers = ('p7', 'p6', 'p5') if info.hdr else ('p6', 'p5', 'p4')
    return tiers[0] if res >= 3840 else tiers[1] if res >= 2560 else tiers[2]


def adjust_nvenc_params(params: List[str], attempt: int) -> List[str]:
    """Apply retry ladder step ``attempt`` (1-based) to a flat option list (core/transcoder.py:101-134)."""
    if attempt <= 0:
        return list(params)
    opts: 'OrderedDict[str, str]' = OrderedDict()
    i = 0
    while i < len(params):
        key = params[i]
        if i + 1 < len(params) and not params[i + 1].startswith('-'):
            opts[key] = params[i + 1]
            i += 2
        else:
            opts[key] = ''
            i += 1
    opts.update(NVENC_RETRIES[min(attempt, len(NVENC_RETRIES)) - 1])
    flat: List[str] = []
    for key, val in opts.items():
        flat.append(key)
        if val != '':
            flat.append(str(val))
    return flat


def build_ffmpeg_params(info: VideoInfo, use_nvenc: bool, gpu_name: str) -> FFmpegParams:
    """core/transcoder.py:357-412"""
    if not use_nvenc:
        level, _tier = derive.calculate_apple_hevc_level(info)
        return FFmpegParams('libx265', 'p010le' if info.hdr else 'yuv420p', 'main10' if info.hdr else 'main', level, [],
                            ['-x265-params', ':'.join(derive.x265_option_list(info)), '-threads', '0'], [])
    level, tier, profile, pix_fmt = derive.calculate_nvenc_hevc_level(info)
    _crf, cq, maxrate, bufsize, gop = calculate_dynamic_values(info, True, gpu_name)
    lookahead = int(min(info.fps * 1.5, 120))
    aq = 6
    longest = max(info.width, info.height)
    if info.hdr and longest >= 3840:
        aq, lookahead = 7, min(info.fps * 2, 120)
    if info.hdr and longest >= 7680:
        aq, lookahead = 8, 120
    vparams = ['-rc', 'vbr', '-tune', 'hq', '-multipass', 'fullres', '-cq', str(cq), '-b:v', '0', '-maxrate', str(maxrate * 1000),
               '-bufsize', str(bufsize * 1000), '-bf', '3', '-b_ref_mode', 'middle', '-rc-lookahead', str(lookahead), '-spatial-aq', '1',
               '-aq-strength', str(aq), '-temporal-aq', '1', '-preset', select_nvenc_preset(info, gpu_name), '-no-scenecut', '1',
               '-g', str(gop), '-tier', tier]
    joined = ' '.join(vparams)
    if 'aud=1' not in joined and '-aud' not in joined:
        vparams += ['-aud', '1']
    meta = derive.build_hdr_metadata(info.master_display, info.max_cll, True, info.fps) if info.hdr else []
    return FFmpegParams('hevc_nvenc', pix_fmt, profile, level, [], vparams, meta)


def get_audio_flags(audio_channels: int) -> List[str]:
    """core/transcoder.py:423-450"""
    if not audio_channels or audio_channels < 1:
        return []
    kbps = min(max(128, audio_channels * 64), 512)
    if audio_channels > 2:
        kbps = max(kbps, 256)
    flags = ['-c:a', 'aac', '-b:a', f'{kbps}k', '-ar', '48000']
    layouts = {1: 'mono', 2: 'stereo', 6: '5.1', 8: '7.1'}
    flags += ['-ac', str(max(1, audio_channels))]
    if audio_channels in layouts:
        flags += ['-channel_layout', layouts[audio_channels]]
    return flags


def build_ffmpeg_command(file_path: Path, out_path: Path, ff_params: FFmpegParams, audio_channels: int,
                         audio_language: Optional[str] = 'eng', extra_vparams: Optional[List[str]] = None) -> List[str]:
    """core/transcoder.py:452-495"""
    cmd = ['ffmpeg', '-hide_banner', '-y', '-i', str(file_path), '-map_metadata', '0', '-c:v', ff_params.vcodec, '-pix_fmt', ff_params.pix_fmt,
           '-profile:v', ff_params.profile, '-tag:v', 'hvc1']
    cmd += ff_params.hdr_metadata or []
    cmd += extra_vparams if extra_vparams else ff_params.vparams
    cmd += VIDEO_METADATA_FLAGS
    if audio_channels and audio_channels > 0:
        cmd += ['-metadata:s:a:0', 'handler_name=SoundHandler', '-metadata:s:a:0', f'language={audio_language or "eng"}',
                '-metadata:s:a:0', 'title="Main Audio"']
        cmd += get_audio_flags(audio_channels)
    cmd += ['-color_range', 'tv', '-brand', 'mp42', '-movflags', '+write_colr+use_metadata_tags+faststart', str(out_path)]
    return cmd


def run_ffmpeg(cmd: List[str], progress_callback: ProgressCb, file_name: str, total_frames: int,
               stop_event: Optional[threading.Event] = None, debug: bool = False) -> Tuple[int, str]:
    """Spawn ffmpeg, parse ``frame=`` for progress, honour the stop event (core/transcoder.py:497-535)."""
    if debug:
        logger.debug('ffmpeg command: %s', ' '.join(cmd))
    lines: List[str] = []
    try:
        with subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, encoding='utf-8', errors='replace') as proc:
            for line in proc.stdout:
                lines.append(line)
                if stop_event and stop_event.is_set():
                    try:
                        proc.terminate()
                    except Exception:
                        pass
                    return 1, ''.join(lines)
                if 'frame=' in line and progress_callback:
                    try:
                        frame = int(line.strip().split('frame=')[-1].split()[0])
                        progress_callback(file_name, frame, total_frames)
                    except Exception:
                        logger.debug('progress parse / callback failed', exc_info=True)
            return proc.wait(), ''.join(lines)
    except Exception as exc:
        logger.error('running ffmpeg failed: %s - %s', cmd[:3], exc)
        return 1, str(exc)


# ------------------------------------------------------------------ B200 branch

def b200_available() -> bool:
    try:
        from . import _cabi
        return _cabi.LIB_PATH.exists() and _cabi.lib().hb_device_count() > 0
    except Exception:
        return False


_GPU_CYCLE_LOCK = threading.Lock()
_GPU_CYCLE_NEXT = [0]


def next_device() -> int:
    """Round-robin device assignment under a lock, like the upscaler's cycle(gpu_list) (upscale_gui_final.py:25-30,123-126)."""
    from . import _cabi
    n = max(1, _cabi.lib().hb_device_count())
    with _GPU_CYCLE_LOCK:
        d = _GPU_CYCLE_NEXT[0] % n
        _GPU_CYCLE_NEXT[0] += 1
    return d


_ENCODER_POOL = None
_ENCODER_POOL_LOCK = threading.Lock()


def encoder_pool():
    """process-wide pool of idle encoders (``HEVC_B200_POOL`` = how many to keep, default 4, 0 disables reuse)"""
    global _ENCODER_POOL
    with _ENCODER_POOL_LOCK:
        if _ENCODER_POOL is None:
            from .encoder import EncoderPool
            _ENCODER_POOL = EncoderPool(int(os.environ.get('HEVC_B200_POOL', '4') or 0))
        return _ENCODER_POOL


_SEGMENT_POOL: Dict[Any, list] = {}


def _segment_pool_get(key):
    with _ENCODER_POOL_LOCK:
        lst = _SEGMENT_POOL.get(key)
        pse = lst.pop() if lst else None
    if pse is not None:
        try:
            pse.reset()
        except Exception:
            pse.close()
            return None
    return pse


def _segment_pool_put(key, pse, reusable: bool):
    keep = int(os.environ.get('HEVC_B200_POOL', '4') or 0) > 0
    if reusable and keep:
        with _ENCODER_POOL_LOCK:
            lst = _SEGMENT_POOL.setdefault(key, [])
            if len(lst) < 1:
                lst.append(pse)
                return
    pse.close()


def _shutdown_pools():
    """close pooled encoders while the CUDA library is still loaded (interpreter exit)"""
    global _ENCODER_POOL
    with _ENCODER_POOL_LOCK:
        pool, _ENCODER_POOL = _ENCODER_POOL, None
        segs = [p for lst in _SEGMENT_POOL.values() for p in lst]
        _SEGMENT_POOL.clear()
    for pse in segs:
        try:
            pse.close()
        except Exception:
            pass
    if pool is not None:
        try:
            pool.close()
        except Exception:
            pass


import atexit  # noqa: E402

atexit.register(_shutdown_pools)


def encode_b200(file_path: Path, out_path: Path, info: VideoInfo, progress_callback: ProgressCb, total_frames: int,
                stop_event: Optional[threading.Event], device: Optional[int] = None, batch: int = 32,
                target_size: Optional[Tuple[int, int]] = None, streams: Optional[int] = None, force_main10: bool = False,
                devices: Optional[Sequence[int]] = None) -> Tuple[int, str]:
    """Encode ``file_path`` to ``out_path`` (hvc1 MP4) on B200.  Returns (0, '') or (1, reason) like run_ffmpeg."""
    from .frames import open_reader
    try:
        src_info = info
        reader = open_reader(file_path, src_info)
    except Exception as exc:
        return 1, f'{type(exc).__name__}: {exc}'
    return encode_reader_b200(reader, Path(file_path).name, out_path, info, progress_callback, total_frames, stop_event, device, batch,
                              target_size, streams, force_main10, devices)


def encode_reader_b200(reader, name: str, out_path: Path, info: VideoInfo, progress_callback: ProgressCb, total_frames: int,
                       stop_event: Optional[threading.Event], device: Optional[int] = None, batch: int = 32,
                       target_size: Optional[Tuple[int, int]] = None, streams: Optional[int] = None, force_main10: bool = False,
                       devices: Optional[Sequence[int]] = None) -> Tuple[int, str]:
    """The encode step for frames delivered by ``reader`` (``frames.py``: file readers, or an in-memory source).

    The pre-encode pixel pipeline (BGR -> 4:2:0 matrix conversion for container sources, polyphase scaling for
    ``target_size``, bit-depth conversion) runs inside the encoder's ingest stage on the device: host frames go to
    ``hb_enc_encode_delayed`` as they were read.

    ``streams`` > 1 (or ``HEVC_B200_STREAMS``) and / or several ``devices`` encode closed-GOP segments of one key-frame interval
    on that many independent encoder streams (``ParallelSegmentEncoder``), written in order and muxed once."""
    import numpy as np

    from . import _cabi, mp4
    from .encoder import PIX_BGR24, ParallelSegmentEncoder, to_c_params
    release = None
    try:
        devs = [int(d) for d in devices] if devices else [next_device() if device is None else int(device)]
        if _cabi.lib().hb_device_count() <= max(devs):
            return 1, f'B200 backend unavailable: device {max(devs)} not present'
    except Exception as exc:
        return 1, f'B200 backend unavailable: {exc}'
    try:
        src_w, src_h = info.width, info.height
        if target_size:
            info = VideoInfo(**{**info.__dict__, 'width': target_size[0], 'height': target_size[1]})
        params = derive.derive_b200_params(info, force_main10=force_main10)
        if getattr(info, 'audio_channels', 0):
            # the reference transcodes audio to AAC inside the same ffmpeg child (core/transcoder.py:423-450); this backend writes
            # the video track only (audio stays with an external tool)
            logger.warning('%s: %d audio channel(s) in the source are NOT carried over by encoder=b200 (video-only MP4)',
                           name, info.audio_channels)
        if streams is None:
            streams = int(os.environ.get('HEVC_B200_STREAMS', '1') or 1)
        kw = {}
        if reader.kind == 'bgr':
            fmt_override = PIX_BGR24
        else:
            fmt_override = None
            if reader.src_bit_depth > 8:
                kw['src_bit_depth'] = reader.src_bit_depth
        if target_size and (src_w, src_h) != tuple(target_size):
            if fmt_override is None and reader.src_bit_depth > 8:
                return 1, 'scaling of 10-bit planar sources is not supported'
            kw['src_size'] = (src_w, src_h)
        segmented = streams > 1 or len(devs) > 1
        unit = max(1, min(int(params.keyint), 1024)) if segmented else batch
        ring = 4
        if segmented:
            cparams = to_c_params(params)
            skey = (tuple(devs), bytes(cparams), int(streams), int(unit))
            pse = _segment_pool_get(skey) or ParallelSegmentEncoder(devs, cparams, streams=streams, max_batch=unit)
            ok = [False]
            release = lambda: _segment_pool_put(skey, pse, ok[0])      # noqa: E731
            submit = lambda data, n, fmt: pse.submit(data, n, fmt=fmt, **kw)      # noqa: E731
            finish = pse.finish
            ring = pse.max_outstanding + 2           # submit() blocks beyond max_outstanding segments, + one being filled + slack
        else:
            pool = encoder_pool()
            key, ctx, single = pool.acquire(devs[0], to_c_params(params), unit)
            ok = [False]
            release = lambda: pool.release(key, ctx, single, reusable=ok[0])      # noqa: E731
            submit = lambda data, n, fmt: single.encode_delayed(data, n, fmt=fmt, **kw)      # noqa: E731
            finish = single.flush
        track = mp4.TrackInfo(params.width, params.height, params.fps_num, params.fps_den, params.profile_idc, params.level_idc, params.tier,
                              params.bit_depth, params.colour_primaries, params.transfer_characteristics, params.matrix_coeffs, params.full_range,
                              params.master_display if params.hdr10 else None, params.max_cll, params.max_fall)
        done = submitted = 0

        def tick(st):
            nonlocal done
            done += len(st)
            if progress_callback and st:
                try:
                    progress_callback(name, done, max(total_frames, done))
                except Exception:
                    logger.debug('progress callback raised', exc_info=True)

        # the elementary stream goes to the muxer as it arrives: samples are spooled to disk, the sample tables stay in memory
        with mp4.StreamMuxer(track, out_path) as mux:
            for buf, n, fmt in reader.batches(unit, ring):
                if stop_event is not None and stop_event.is_set():
                    mux.abort()
                    return 1, 'cancelled'
                data = np.ascontiguousarray(buf).reshape(n, -1)
                out, st = submit(data, n, fmt_override if fmt_override is not None else fmt)
                mux.feed(out)
                submitted += n
                tick(st)
            if submitted == 0:
                mux.abort()
                return 1, 'no frames decoded'
            out, st = finish()
            mux.feed(out)
            tick(st)
        ok[0] = True                           # drained cleanly: the encoder(s) can serve the next file
        return 0, ''
    except Exception as exc:
        logger.debug('B200 encode failed', exc_info=True)
        return 1, f'{type(exc).__name__}: {exc}'
    finally:
        try:
            if release is not None:
                release()              # (a failed or cancelled run closes its encoder: nothing reads the frame buffers after this)
        except Exception:
            pass
        try:
            reader.close()             # page-locked frame buffers back to the pool -- only now, after the encoder has let go of them
        except Exception:
            pass


# ------------------------------------------------------------------ the drop-in entry point

def convert_video(file_path: Path, out_dir: Path, progress_callback: ProgressCb = None, debug: bool = False, skip_validator: bool = False,
                  force_cpu: bool = False, force_gpu: bool = False, stop_event: Optional[threading.Event] = None,
                  encoder: str = 'auto', device: Optional[int] = None) -> Dict[str, Any]:
    """Transcode one file; never raises for encode failures (core/transcoder.py:537-638).

    encoder: 'auto' (reference behaviour: NVENC if ffmpeg offers it, else libx265), 'cpu', 'nvenc', or 'b200'."""
    file_path, out_dir = Path(file_path), Path(out_dir)
    if encoder == 'auto':
        encoder = os.environ.get('HEVC_B200_ENCODER', 'auto')
    info = probe_media(file_path)
    out_path = out_dir / (file_path.stem + '.mp4')
    use_b200 = encoder == 'b200'
    use_nvenc = False if use_b200 else (decide_encoder(info, force_cpu or encoder == 'cpu', force_gpu) if encoder != 'nvenc' else has_nvenc())
    entry: Dict[str, Any] = {'file': file_path.name, 'status': 'FAILED', 'quality': None, 'retries': 0,
                             'method': 'B200' if use_b200 else ('NVENC' if use_nvenc else 'CPU'), 'hdr': info.hdr}
    crf, _cq, _, _, _ = calculate_dynamic_values(info, use_nvenc=False)
    _, nvenc_cq, _, _, _ = calculate_dynamic_values(info, use_nvenc=True, gpu_name='')
    total_frames = max(1, int(info.duration * info.fps)) if info.duration and info.fps else 1

    if use_b200:
        rc, why = encode_b200(file_path, out_path, info, progress_callback, total_frames, stop_event, device)
        if rc == 0:
            entry.update(status='SUCCESS', quality=crf)
        else:
            logger.error('B200 encode failed: %s: %s', file_path.name, why)
    else:
        gpu_name = detect_gpu_type()
        ff_params = build_ffmpeg_params(info, use_nvenc, gpu_name)
        if use_nvenc:
            for attempt, mods in enumerate(NVENC_RETRIES + [None], 1):
                vparams = adjust_nvenc_params(ff_params.vparams, attempt) if mods else ff_params.vparams
                cmd = build_ffmpeg_command(file_path, out_path, ff_params, info.audio_channels, info.audio_language, vparams)
                rc, text = run_ffmpeg(cmd, progress_callback, file_path.name, total_frames, stop_event, debug)
                if rc == 0:
                    entry.update(status='SUCCESS', quality=nvenc_cq, retries=min(attempt, len(NVENC_RETRIES)), method='NVENC')
                    break
                logger.warning('NVENC attempt %d failed: %s | %s', attempt, file_path.name, text[:1000])
                if attempt == len(NVENC_RETRIES) + 1:
                    use_nvenc = False
        if not use_nvenc and entry['status'] != 'SUCCESS':
            cpu_params = build_ffmpeg_params(info, False, gpu_name)
            cmd = build_ffmpeg_command(file_path, out_path, cpu_params, info.audio_channels, info.audio_language)
            rc, text = run_ffmpeg(cmd, progress_callback, file_path.name, total_frames, stop_event, debug)
            if rc == 0:
                entry.update(status='SUCCESS', quality=crf, retries=0, method='CPU')
            else:
                logger.error('CPU transcode failed: %s\n%s', file_path.name, text[:2000])

    if entry['status'] == 'SUCCESS' and not skip_validator:
        try:
            from .compliance import check_file
            problems = check_file(out_path)
            if problems:
                logger.warning('compliance check: %s: %s', out_path.name, '; '.join(problems))
        except Exception:
            logger.debug('compliance check raised', exc_info=True)
    if stop_event and stop_event.is_set() and entry['status'] != 'SUCCESS':
        entry['status'] = 'CANCELLED'
    if progress_callback:
        try:
            progress_callback(file_path.name, total_frames, total_frames)
        except Exception:
            logger.debug('progress callback raised at completion', exc_info=True)
    return entry
