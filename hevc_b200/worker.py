"""Thread wrapper around ``convert_video`` with the reference worker's surface (gui/worker.py:14-56): same constructor
arguments, ``progress`` / ``finished`` / ``log`` notifications, ``stop()``.  PySide6 is not a dependency of this package, so
the three Qt signals are plain callback lists with an ``emit``/``connect`` pair; a Qt front-end can connect real signals."""
from __future__ import annotations

import threading
from pathlib import Path
from typing import Callable, List

from .transcoder import convert_video


class Signal:
    def __init__(self):
        self._slots: List[Callable] = []

    def connect(self, fn: Callable):
        self._slots.append(fn)

    def emit(self, *args):
        for fn in list(self._slots):
            fn(*args)


class TranscodeWorker(threading.Thread):
    def __init__(self, file_path: Path, out_dir: Path, debug=False, skip_validator=False, force_cpu=False, force_gpu=False,
                 encoder: str = 'auto', device=None):
        super().__init__(daemon=True)
        self.file_path, self.out_dir = Path(file_path), Path(out_dir)
        self.debug, self.skip_validator, self.force_cpu, self.force_gpu = debug, skip_validator, force_cpu, force_gpu
        self.encoder, self.device = encoder, device
        self.stop_event = threading.Event()
        self.progress, self.finished, self.log = Signal(), Signal(), Signal()
        self.result = None

    def run(self):
        try:
            self.result = convert_video(self.file_path, self.out_dir, progress_callback=self.progress.emit, debug=self.debug,
                                        skip_validator=self.skip_validator, force_cpu=self.force_cpu, force_gpu=self.force_gpu,
                                        stop_event=self.stop_event, encoder=self.encoder, device=self.device)
        except Exception as exc:           # same catch-all as gui/worker.py:43-52
            self.log.emit(f'[ERROR] {self.file_path.name}: {exc}')
            self.result = {'file': self.file_path.name, 'status': 'FAILED', 'quality': None, 'retries': 0, 'method': 'UNKNOWN', 'hdr': False}
        self.finished.emit(self.result)

    def stop(self):
        self.stop_event.set()
