"""Stage-by-stage comparison of the CUDA encoder with the CPU model; prints the first mismatch per stage."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests import enc_common as ec   # noqa: E402


def main():
    w, h, depth, n = (int(a) for a in sys.argv[1:5])
    qp = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (26, 28)
    from hevc_b200 import _cabi, encoder as E
    from oracle import fforacle
    ctx = _cabi.Context(0)
    p = ec.b200_params(w, h, depth, keyint=4)
    frames = ec.clip_frames(w, h, n)
    m_stream, m_aus, m_recs, m_decs = ec.run_model(p, frames, qp[0], qp[1], hash_sei=True)
    enc = E.B200Encoder(ctx, E.to_c_params(p, qp=qp, hash_sei=True, keep_recon=True), max_batch=max(n, 1))
    out, stats = enc.encode(E.pack_yuv420p8(frames), n)
    print('gpu bytes', len(out), 'model bytes', len(m_stream), 'timing', enc.last_timing())
    pos = 0
    for i in range(n):
        gc, gl = enc.read_decisions(i)
        mc, ml = m_decs[i]
        for field in ('pred_mode', 'intra_mode', 'mvx', 'mvy', 'cbf'):
            bad = np.argwhere(mc[field] != gc[field])
            if bad.size:
                b = tuple(bad[0])
                print(f'frame {i} {field}: {len(bad)} CUs differ, first {b}: model {mc[field][b]} gpu {gc[field][b]}')
        coded = mc['cbf'].reshape(-1) != 0
        badc = np.argwhere(ml[coded] != gl[coded])
        if badc.size:
            print(f'frame {i} levels: {len(badc)} differ, first {badc[0].tolist()}')
        rec = enc.read_recon(i)
        for c in range(3):
            d = np.argwhere(rec[c] != m_recs[i][c])
            if d.size:
                print(f'frame {i} plane {c}: {len(d)} samples differ, first {d[0].tolist()} model {m_recs[i][c][tuple(d[0])]} gpu {rec[c][tuple(d[0])]}')
        au = out[pos:pos + stats[i].bytes]
        pos += stats[i].bytes
        if au != m_aus[i]:
            k = next((j for j in range(min(len(au), len(m_aus[i]))) if au[j] != m_aus[i][j]), min(len(au), len(m_aus[i])))
            print(f'frame {i} AU differs: gpu {len(au)} B model {len(m_aus[i])} B, first byte {k}')
        else:
            print(f'frame {i} AU identical ({len(au)} B)')
    try:
        dec = fforacle.decode_hevc(out, verify_hash=True)
        print('decoder accepted gpu stream:', len(dec), 'frames')
    except Exception as exc:
        print('decoder rejected gpu stream:', exc)


if __name__ == '__main__':
    main()
