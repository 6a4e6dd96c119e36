"""Upscale + Main10 encode in one pass (BASELINE config 4).

The reference's upscaler (upscale_gui_final.py:72-212) decodes with OpenCV, pushes every frame through an external
Real-ESRGAN process, writes an mp4v file and muxes audio back; its output is never HEVC.  Here the same geometry rule
(:81-87) drives the device-side polyphase scaler fused with the P010 pack, and the result goes straight into the HEVC
encoder.  The neural network itself is out of scope (un-vendored script + weights)."""
from __future__ import annotations

import threading
from pathlib import Path
from typing import Any, Callable, Dict, Optional, Sequence, Tuple

from . import transcoder
from .derive import calculate_dynamic_values
from .probe import probe_media


def target_geometry(width: int, height: int, target_height: int = 0) -> Tuple[int, int]:
    """upscale_gui_final.py:81-87: auto target height (0) -> 1080 below 1080, 2160 below 2160, else unchanged"""
    if target_height == 0:
        target_height = 1080 if height < 1080 else 2160 if height < 2160 else height
    scale = target_height / height
    return int(width * scale) & ~1, target_height


def process_video(video_path: Path, output_dir: Path, target_height: int = 0, progress_callback: Optional[Callable[[str, int, int], None]] = None,
                  stop_flag: Optional[threading.Event] = None, device: Optional[int] = None,
                  devices: Optional[Sequence[int]] = None) -> Dict[str, Any]:
    """Scale ``video_path`` to the target geometry and encode it as Main10 HEVC into ``output_dir/<stem>.mp4``.

    ``devices``: several GPUs -> the clip is cut into closed-GOP segments of one key-frame interval, dealt round-robin to one
    worker per device (the reference's ``cycle(gpu_list)`` fan-out, upscale_gui_final.py:25-30,123-126), written in order
    (:164-178) and muxed once (BASELINE config 4)."""
    video_path, output_dir = Path(video_path), Path(output_dir)
    info = probe_media(video_path)
    tw, th = target_geometry(info.width, info.height, target_height)
    out_path = output_dir / (video_path.stem + '.mp4')
    total = max(1, int(info.duration * info.fps)) if info.duration and info.fps else 1
    # the fused scaler writes P010, so the encode is Main10 whatever the source depth; colour description and HDR10
    # signalling stay the source's own (an SDR BT.709 clip stays BT.709 SDR: Main10 is a bit depth, not a transfer function)
    rc, why = transcoder.encode_b200(video_path, out_path, info, progress_callback, total, stop_flag, device, target_size=(tw, th),
                                     force_main10=True, devices=devices)
    crf = calculate_dynamic_values(type(info)(**{**info.__dict__, 'width': tw, 'height': th}), False)[0]
    status = 'SUCCESS' if rc == 0 else ('CANCELLED' if stop_flag is not None and stop_flag.is_set() else 'FAILED')
    return {'file': video_path.name, 'status': status, 'quality': crf, 'retries': 0, 'method': 'B200', 'hdr': bool(info.hdr),
            'width': tw, 'height': th, 'reason': why}
