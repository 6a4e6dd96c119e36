"""Wall-clock rate of the drop-in itself: convert_video(encoder='b200') from a Y4M file on disk to the hvc1 MP4 on disk
(probe -> read -> ingest + encode on the device -> streaming mux -> compliance self-check), per clip type.

    python tools/dropin_bench.py [clip types ...] [--frames 120] [--streams 2]      -> one JSON line per clip type"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from hevc_b200 import transcoder
    from hevc_b200.frames import write_y4m
    from hevc_b200.synth import CLIP_TYPES, TorchSynthClip
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    opts = {a.split('=')[0]: a.split('=')[1] for a in sys.argv[1:] if a.startswith('--') and '=' in a}
    n = int(opts.get('--frames', 120))
    streams = int(opts.get('--streams', 2))
    base = '/dev/shm' if os.path.isdir('/dev/shm') else None
    for name in args or ['4k60_hdr', '1080p_sdr']:
        w, h, fps, hdr = CLIP_TYPES[name]
        with tempfile.TemporaryDirectory(dir=base) as d:
            d = Path(d)
            clip = TorchSynthClip(w, h, seed=1, device='cuda:0')
            lw, cw = w * h, (w // 2) * (h // 2)
            frames = []
            for s in range(0, n, 30):
                for f in clip.frames(s, min(30, n - s)).cpu().numpy():
                    frames.append((f[:lw].reshape(h, w), f[lw:lw + cw].reshape(h // 2, w // 2), f[lw + cw:].reshape(h // 2, w // 2)))
            src = d / f'{name}.y4m'
            write_y4m(src, frames, w, h, (fps, 1))
            if hdr:
                (d / f'{name}.y4m.json').write_text(json.dumps({'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'color_space': 'bt2020nc'}))
            del frames
            os.environ['HEVC_B200_STREAMS'] = str(streams)
            out = d / 'out'
            out.mkdir()
            res = transcoder.convert_video(src, out, encoder='b200', device=0)           # warm-up: library, encoder pool, page cache
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                res = transcoder.convert_video(src, out, encoder='b200', device=0)
                times.append(time.perf_counter() - t0)
            t = sorted(times)[1]
            print(json.dumps({'what': 'convert_video(encoder="b200") wall clock, Y4M on tmpfs -> hvc1 MP4 on tmpfs, median of 3', 'clip': name,
                              'frames': n, 'streams': streams, 'status': res['status'], 'seconds': round(t, 3), 'frames_per_s': round(n / t, 1),
                              'mp4_bytes': (out / f'{name}.mp4').stat().st_size, 'source_bytes': src.stat().st_size}), flush=True)


if __name__ == '__main__':
    main()
