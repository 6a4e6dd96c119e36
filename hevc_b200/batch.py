"""Batch scheduling: the reference's N-concurrent-workers / refill-on-finish policy (gui/mainwindow.py:289-345,
apple_hevc_batch.py:861-882) mapped onto GPUs, plus closed-GOP segment sharding of one long clip.

The data path needs no collective: files (or segments) are independent units.  Within one process workers pick a
device round-robin; across processes (torchrun: one rank per GPU) ``shard_for_rank`` gives each rank its files."""
from __future__ import annotations

import csv
import threading
from collections import deque
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple

from .transcoder import convert_video

INPUT_EXTS = ('.mp4', '.mkv', '.mov', '.avi', '.webm', '.m4v', '.y4m', '.yuv')
CSV_FIELDS = ['file', 'status', 'quality', 'retries', 'method', 'hdr']      # gui/mainwindow.py:351


def find_inputs(input_dir: Path) -> List[Path]:
    return sorted(p for p in Path(input_dir).rglob('*') if p.suffix.lower() in INPUT_EXTS and p.is_file())


def lpt_order(files: Sequence[Path], cost: Callable[[Path], float]) -> List[Path]:
    """longest-processing-time-first, so a 4K straggler does not end the batch alone"""
    return sorted(files, key=cost, reverse=True)


def shard_for_rank(items: Sequence[Any], rank: int, world: int, cost: Optional[Callable[[Any], float]] = None) -> List[Any]:
    """Static partition of independent units over ranks: greedy LPT bin packing when costs are given, round-robin otherwise."""
    if cost is None:
        return [it for i, it in enumerate(items) if i % world == rank]
    loads = [0.0] * world
    mine: List[Any] = []
    for it in sorted(items, key=cost, reverse=True):
        k = min(range(world), key=lambda j: (loads[j], j))
        loads[k] += cost(it)
        if k == rank:
            mine.append(it)
    return mine


def gop_segments(n_frames: int, keyint: int) -> List[Tuple[int, int]]:
    """Closed-GOP segments [start, end) of one clip; each starts with an IDR, so they encode independently and their
    Annex-B outputs concatenate into one valid stream."""
    return [(s, min(n_frames, s + keyint)) for s in range(0, n_frames, keyint)]


def batch_convert(input_dir: Path, output_dir: Path, max_workers: int = 2, progress: Optional[Callable[[str, int, int], None]] = None,
                  stop_event: Optional[threading.Event] = None, files: Optional[Iterable[Path]] = None, **kwargs) -> List[Dict[str, Any]]:
    """Run ``convert_video`` over a directory with ``max_workers`` concurrent workers, refilling as each finishes, and
    rewrite ``transcode_log.csv`` after every file (gui/mainwindow.py:318-355)."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    queue = deque(files if files is not None else find_inputs(input_dir))
    results: List[Dict[str, Any]] = []
    lock = threading.Lock()

    def save_csv():
        with open(output_dir / 'transcode_log.csv', 'w', newline='', encoding='utf-8') as fh:
            w = csv.DictWriter(fh, fieldnames=CSV_FIELDS)
            w.writeheader()
            for r in results:
                w.writerow({k: r.get(k) for k in CSV_FIELDS})

    def worker():
        while True:
            with lock:
                if not queue or (stop_event is not None and stop_event.is_set()):
                    return
                f = queue.popleft()
            res = convert_video(f, output_dir, progress_callback=progress, stop_event=stop_event, **kwargs)
            with lock:
                results.append(res)
                save_csv()

    threads = [threading.Thread(target=worker, daemon=True) for _ in range(max(1, min(max_workers, len(queue))))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results


def encode_clip_segmented(frames, n_frames: int, params, devices: Sequence[int], fmt: int = 0, max_batch: int = 32,
                          c_params_kwargs: Optional[dict] = None) -> bytes:
    """Encode one clip as independent closed-GOP segments spread round-robin over ``devices`` (one worker thread and one
    encoder per device), then concatenate the Annex-B outputs in display order (BASELINE config 4).

    ``frames``: uint8 array [n_frames, frame_bytes] in the ``hb_frames`` layout.  Each segment starts with an IDR and its
    own parameter sets, so the concatenation is one valid stream.  Rate control restarts per segment (documented caveat)."""
    import numpy as np

    from . import _cabi
    from .encoder import B200Encoder, to_c_params
    segs = gop_segments(n_frames, params.keyint)
    outputs: List[Optional[bytes]] = [None] * len(segs)
    errors: List[BaseException] = []

    def work(slot: int, device: int):
        try:
            ctx = _cabi.Context(device)
            try:
                for k in range(slot, len(segs), len(devices)):
                    a, b = segs[k]
                    enc = B200Encoder(ctx, to_c_params(params, **(c_params_kwargs or {})), max_batch=min(max_batch, b - a))
                    try:
                        outputs[k], _ = enc.encode(np.ascontiguousarray(frames[a:b]), b - a, fmt=fmt, force_idr=True)
                    finally:
                        enc.close()
            finally:
                ctx.close()
        except BaseException as exc:      # surfaced to the caller below
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i, d), daemon=True) for i, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return b''.join(o for o in outputs if o is not None)
