"""Golden vectors for probe parsing: run the reference's probe_media (core/probe.py:47-116) with ffprobe replaced by canned
JSON documents and record the VideoInfo it builds.  Build container only (needs /root/reference)."""
import dataclasses
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, '/root/reference')
from core import probe as ref   # noqa: E402

DOCS = [
    {'streams': [{'codec_type': 'video', 'width': 3840, 'height': 2160, 'avg_frame_rate': '60/1', 'color_primaries': 'bt2020',
                  'color_transfer': 'smpte2084', 'color_space': 'bt2020nc', 'pix_fmt': 'yuv420p10le', 'nb_frames': '300'},
                 {'codec_type': 'audio', 'channels': 6, 'tags': {'language': 'fra'}}],
     'format': {'duration': '5.000000', 'tags': {'MASTER_DISPLAY': 'G(1,2)B(3,4)R(5,6)WP(7,8)L(9,10)', 'max-cll': '1000,400'}}},
    {'streams': [{'codec_type': 'video', 'width': 1920, 'height': 1080, 'avg_frame_rate': '30000/1001', 'pix_fmt': 'yuv420p'}], 'format': {}},
    {'streams': [{'codec_type': 'video', 'width': 1920, 'height': 1080, 'avg_frame_rate': '0/0', 'r_frame_rate': '25/1',
                  'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'pix_fmt': 'yuv420p'}], 'format': {'duration': 'x'}},
    {'streams': [{'codec_type': 'video', 'width': 1280, 'height': 720, 'avg_frame_rate': 'abc', 'pix_fmt': 'p010le', 'color_space': 'BT2020NC'}],
     'format': {'tags': {'COLOR_PRIMARIES': 'BT2020'}}},
    {'streams': [{'codec_type': 'audio', 'channels': 2}], 'format': {}},
    {'streams': [{'codec_type': 'video', 'avg_frame_rate': '24/1', 'pix_fmt': 'yuv444p10le', 'nb_frames': 'N/A'},
                 {'codec_type': 'audio', 'tags': {'LANGUAGE': 'deu'}}], 'format': {'duration': '12.5'}},
    {'streams': [{'codec_type': 'video', 'width': 640, 'height': 360, 'avg_frame_rate': '30/0', 'color_transfer': 'pq', 'color_space': 'bt2020'}],
     'format': {'tags': {'master_display': 'x', 'MAX_CLL': '1,2'}}},
]


class _Res:
    def __init__(self, text):
        self.stdout = text


out = []
for doc in DOCS:
    orig = subprocess.run
    subprocess.run = lambda *a, **k: _Res(json.dumps(doc))
    try:
        info = ref.probe_media(Path('x.mp4'))
    finally:
        subprocess.run = orig
    out.append({'doc': doc, 'info': dataclasses.asdict(info)})
Path(__file__).with_name('probe_golden.json').write_text(json.dumps(out, indent=1))
print(len(out), 'probe cases')
