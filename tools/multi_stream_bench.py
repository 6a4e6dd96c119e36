#!/usr/bin/env python
"""Aggregate throughput of N independent encoders (one context + stream each, one host thread each) on ONE GPU: the
per-file sharding of BASELINE config 2 seen from a single device.  usage: multi_stream_bench.py [clip] [threads...]"""
import json
import sys
import threading
import time

import torch

from hevc_b200 import _cabi, derive
from hevc_b200 import encoder as E
from hevc_b200.probe import VideoInfo
from hevc_b200.synth import CLIP_TYPES, TorchSynthClip


def run(clip_name, n_threads, frames=120, steps=6):
    w, h, fps, hdr = CLIP_TYPES[clip_name]
    info = VideoInfo(w, h, float(fps), 'bt2020' if hdr else 'bt709', 'smpte2084' if hdr else 'bt709', 'bt2020nc' if hdr else 'bt709',
                     'yuv420p', '', '', 0, hdr, None, None, 5.0)
    p = derive.derive_b200_params(info)
    clip = TorchSynthClip(w, h, seed=1, device='cuda:0')
    dev = clip.frames(0, frames)
    torch.cuda.synchronize()
    ctxs = [_cabi.Context(0) for _ in range(n_threads)]
    encs = [E.B200Encoder(c, E.to_c_params(p), max_batch=frames) for c in ctxs]
    start = threading.Barrier(n_threads + 1)
    done = []

    def work(enc):
        for _ in range(2):
            enc.encode_delayed(dev.data_ptr(), frames, on_device=True, force_idr=True, frame_bytes=clip.frame_bytes)
        enc.flush()
        start.wait()
        for _ in range(steps):
            enc.encode_delayed(dev.data_ptr(), frames, on_device=True, force_idr=True, frame_bytes=clip.frame_bytes)
        enc.flush()
        done.append(time.perf_counter())

    ts = [threading.Thread(target=work, args=(e,)) for e in encs]
    for t in ts:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ts:
        t.join()
    dt = max(done) - t0
    for e in encs:
        e.close()
    for c in ctxs:
        c.close()
    return {'clip': clip_name, 'encoders': n_threads, 'frames_per_s': round(n_threads * steps * frames / dt, 1)}


if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else '1080p_sdr'
    for n in [int(a) for a in sys.argv[2:]] or [1, 2, 4]:
        print(json.dumps(run(name, n)))
