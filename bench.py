#!/usr/bin/env python
"""Benchmarks of the B200 HEVC encode backend on the BASELINE.json configurations.

Default (what the driver runs): BASELINE configs[1] -- 4K60 HDR10 Main10 encode frames/s.  One step = one pass of the
encode hot path over one closed-GOP segment (120 frames = keyint of that configuration) per encoder stream.  ``value`` is
device-timed with the clips resident in HBM; ``e2e`` goes through the C ABI with HOST buffers (upload, encode, bitstream
download, access-unit assembly) on the wall clock.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU; no collective on the data path)
    python bench.py --impl reference ...      the CPU restatement (oracle/) on all host cores, bounded sample
    python bench.py --config 1                1080p30 SDR Main (BASELINE configs[0], the reference's own CPU-runnable case)
    python bench.py --config 3                64 mixed clips sharded per file over the ranks, LPT order + refill-on-finish workers
    python bench.py --config 4                one 1800-frame 1080p clip -> 4K Main10, closed-GOP segments over the ranks, ONE mux
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

UNIT = 'frames/s'
METRICS = {1: '1080p30 SDR Main HEVC encode throughput', 2: '4K60 HDR10 Main10 HEVC encode throughput',
           3: 'batch of 64 mixed SDR/HDR 1080p/720p/4K clips, HEVC encode throughput (per-file sharding)',
           4: '1080p -> 4K upscale + Main10 HEVC encode throughput (closed-GOP segments of one clip)'}
CONFIG_CLIP = {1: '1080p_sdr', 2: '4k60_hdr'}
BATCH_TYPES = ['1080p_sdr', '720p_sdr', '4k_sdr', '1080p_hdr', '4k_hdr']      # reference tests/generate_test_videos.py:10-16


def clip_info(name: str, seconds: float = 5.0):
    from hevc_b200.probe import VideoInfo
    from hevc_b200.synth import CLIP_TYPES
    w, h, fps, hdr = CLIP_TYPES[name]
    tags = ('bt2020', 'smpte2084', 'bt2020nc') if hdr else ('bt709', 'bt709', 'bt709')
    # the reference's fixtures are 5 s clips (tests/generate_test_videos.py:30); CRF/VBV derive from that duration
    return VideoInfo(w, h, float(fps), *tags, 'yuv420p', '', '', 0, hdr, None, None, seconds)


def workload_config(args, params, n_gpus):
    S = getattr(args, 'streams', 1)
    return {'workload': f'{args.clip}: {params.width}x{params.height}@{params.fps_num}/{params.fps_den} '
                        f'{"Main10 HDR10" if params.hdr10 else "Main10" if params.bit_depth == 10 else "Main"} yuv420p8 source',
            'baseline_config': args.config,
            'frames_per_step': args.frames * S, 'segment_frames': args.frames, 'streams_per_gpu': S,
            'keyint': params.keyint, 'min_keyint': params.min_keyint, 'crf': params.crf,
            'vbv_maxrate_kbps': params.vbv_maxrate_kbps, 'vbv_bufsize_kbit': params.vbv_bufsize_kbit,
            'level_idc': params.level_idc,
            'coding_tools': 'CTU32/CU16, 35-mode intra, intra CUs in P frames, quarter-sample ME + merge, scene-cut key frames, P-frame QP cascade, deblocking, SAO, WPP CABAC',
            'rate_control': 'crf quality ceiling (qp_i, qp_p) = %s, VBV-constrained on the device (vbv-maxrate / vbv-bufsize)' % (str(args.qp),),
            'parallelism': ('%d GPU(s), one process each; per GPU %d independent encoder stream(s) fed with closed-GOP segments of %d frames '
                            '(the reference\'s N-worker model applied to one device), no collective on the data path'
                            % (n_gpus, S, args.frames)),
            'l2': 'inputs (%.1f GB/step) exceed the 126 MB L2' % (args.frames * S * params.width * params.height * 1.5 / 1e9)}


class ClockSampler:
    """nvidia-smi clocks line from the profiling recipe, sampled during the timed region"""

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            self.path = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False).name
            q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        try:
            rows = [r.split(', ') for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows)
            out['sm_mhz'] = sm[len(sm) // 2]
            out['sm_max_mhz'] = float(rows[0][2])
            out['samples'] = len(rows)
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            for k, name in enumerate(names):
                if any(r[5 + k].strip().lower().startswith('active') for r in rows):
                    out['reasons'].append(name)
        except Exception:
            pass
        return out


def measured_peaks():
    try:
        return json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0}, 'fallback'


def committed_rd(size: str):
    """Bjontegaard delta rate of the shipped tool set against the round-1 tool set, per content class, as committed under profiles/
    (measured on the CPU model, to which the CUDA encoder is byte-identical: tools/model_rd.py).  Not measured in this run."""
    path = ROOT / 'profiles' / f'rd_model_{size}_r2.jsonl'
    try:
        rows = [json.loads(l) for l in path.read_text().splitlines() if l.strip()]
    except (OSError, ValueError):
        return None
    bd = {r['class']: [r['bd_rate_psnr_pct'], r['bd_rate_ssim_pct']] for r in rows if r.get('anchor') == 'r1' and r.get('set') == 'r2'}
    if not bd:
        return None
    return {'bd_rate_pct_psnr_ssim': bd, 'anchor': 'round-1 tool set of this encoder (own anchor)', 'x265_anchor': None,
            'x265_anchor_note': 'libx265 is not in the image (BASELINE.md section 4): the anchor column stays empty',
            'source': f'profiles/{path.name} (committed; constant QP 22/27/32/37, PSNR-Y and SSIM of the reconstruction)'}


def dist_env():
    return int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))


def bind_rank_to_cores(local: int, world: int):
    """Give each rank its own slice of the host cores (and so of the memory controllers feeding its uploads): eight ranks
    sharing every core of one socket lost 13 % end to end in round 1 (SCALE_r01: all GPUs on NUMA node 0, CPUs 0-31)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ CPU restatement arm
def cpu_model_run(host_frames, depth, qp, params, threads=None):
    """frames/s of the CPU model (oracle/hevc_encode.c; its CU loops run on all host cores) over one closed-GOP run of frames:
    one encoder, one key frame, the rest P frames -- the same GOP shape the CUDA arm codes"""
    from oracle import encoder_model as em
    w, h = params.width, params.height
    if threads:
        os.environ['OMP_NUM_THREADS'] = str(threads)
    enc = em.ModelEncoder(em.make_params(w, h, depth, qp_i=qp[0], qp_p=qp[1], keyint=params.keyint, min_keyint=params.min_keyint,
                                         fps=(params.fps_num, params.fps_den), hdr10=bool(params.hdr10), hash_sei=False,
                                         level_idc=params.level_idc, vbv_maxrate_kbps=params.vbv_maxrate_kbps,
                                         vbv_bufsize_kbit=params.vbv_bufsize_kbit, rate_control=1))
    sh = depth - 8
    lw, cw = w * h, (w // 2) * (h // 2)
    total = 0
    t0 = time.perf_counter()
    for f in host_frames:
        y = f[:lw].reshape(h, w).astype(np.uint16) << sh
        u = f[lw:lw + cw].reshape(h // 2, w // 2).astype(np.uint16) << sh
        v = f[lw + cw:].reshape(h // 2, w // 2).astype(np.uint16) << sh
        au, _ = enc.encode(y, u, v)
        total += len(au)
    dt = time.perf_counter() - t0
    enc.close()
    return len(host_frames) / dt, dt, total


def host_sample_frames(params, count, seed=0):
    """first `count` frames of the rank-0 clip on the host (generated on the GPU when there is one)"""
    import torch
    from hevc_b200.synth import TorchSynthClip
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    clip = TorchSynthClip(params.width, params.height, seed=seed, device=dev)
    return [clip.frame(i).cpu().numpy() for i in range(count)]


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return
    from hevc_b200 import derive
    from hevc_b200.encoder import crf_to_qp
    from oracle import cmodel
    cmodel.build()
    cmodel.lib()                  # loaded in THIS process: the run below executes liboracle.so's own code, no worker processes
    params = derive.derive_b200_params(clip_info(args.clip))
    args.qp = crf_to_qp(params.crf)
    cores = os.cpu_count() or 1
    sample_n = min(args.frames, args.ref_frames)
    frames = host_sample_frames(params, sample_n)
    vals = []
    for step in range(args.warmup + args.steps):
        if step < args.warmup and step > 0:
            continue            # one warm-up pass pages the library in; each pass costs ~10 s of all cores
        fps, dt, _ = cpu_model_run(frames, params.bit_depth, args.qp, params)
        if step >= args.warmup:
            vals.append((fps, dt))
    fps = float(np.median([v[0] for v in vals]))
    ms = 1000.0 * float(np.median([v[1] for v in vals]))
    sample = (f'first {sample_n} frames of the same synthetic clip per step: one closed GOP (1 key frame + {sample_n - 1} P frames, the GOP shape '
              f'of the CUDA arm), one encoder whose CU loops run on {cores} OpenMP threads (gcc -O3 -march=x86-64-v3)')
    line = {'impl': 'reference', 'metric': METRICS[args.config], 'value': round(fps, 4), 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(ms, 2), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u16' if params.bit_depth > 8 else 'u8', 'data': 'synthetic',
            'config': workload_config(args, params, 1),
            'cpu_baseline': {'value': round(fps, 4), 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': round(fps, 4), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'libx265 (the reference CPU encoder, core/transcoder.py:398) is not in the image; this is the CPU restatement of the same '
                    'algorithm and tool set the CUDA path runs (oracle/hevc_encode.c): a stated baseline for THIS encoder, not x265'}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm, configs 1 and 2
def run_b200(args):
    import torch
    import torch.distributed as dist
    from hevc_b200 import derive
    from hevc_b200.encoder import ParallelSegmentEncoder, crf_to_qp, to_c_params
    from hevc_b200.synth import TorchSynthClip

    world, rank, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the B200 backend has no CPU fallback')
    torch.cuda.set_device(local)
    cores_bound = bind_rank_to_cores(local, world) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    params = derive.derive_b200_params(clip_info(args.clip))
    args.qp = crf_to_qp(params.crf)
    S = max(1, args.streams)
    if args.frames > args.batch:
        raise SystemExit('--frames (one closed-GOP segment) must fit one device batch (--batch)')
    enc = ParallelSegmentEncoder(local, to_c_params(params), streams=S, max_batch=args.batch)
    n = args.frames
    dev_frames, host_frames = [], []
    for k in range(S):                                               # every stream encodes its own clip
        clip = TorchSynthClip(params.width, params.height, seed=rank * 16 + k, device=f'cuda:{local}')
        d = clip.frames(0, n).contiguous()                           # resident in HBM
        hbuf = torch.empty(d.shape, dtype=torch.uint8, pin_memory=True)
        hbuf.copy_(d)
        dev_frames.append(d)
        host_frames.append(hbuf)
    torch.cuda.synchronize()
    fb = clip.frame_bytes

    # A step is one closed-GOP segment of `n` frames per stream, submitted through the pipelined entry point
    # (hb_enc_encode_delayed under ParallelSegmentEncoder): a submit enqueues a segment and returns finished ones, so that the
    # frame chain of a segment overlaps the CABAC tail, download and access-unit assembly of the previous one, and the second
    # stream fills the SMs during the first one's key frames, kernel tails and small kernels.  A flush closes every timed
    # region, so exactly K steps are inside it.
    def step_resident():
        return sum(len(enc.submit(dev_frames[k].data_ptr(), n, on_device=True, frame_bytes=fb)[0]) for k in range(S))

    def step_host():
        return sum(len(enc.submit(host_frames[k].numpy(), n, on_device=False, frame_bytes=fb)[0]) for k in range(S))

    def flush():
        return len(enc.finish()[0])

    for _ in range(args.warmup):
        step_resident()
    flush()
    step_host()
    flush()
    # ---- timed: device-resident input; device time from a CUDA event on the encoder's stream before the first step to the
    #      end of the last bitstream download.  The encoder also stamps the end of every segment on the same device timeline,
    #      which gives the median step.
    launches0 = enc.launches
    enc.profile(1)
    barrier()
    with ClockSampler(local) as clocks:
        enc.mark()
        t0 = time.perf_counter()
        bytes_out = 0
        for _ in range(args.steps):
            bytes_out += step_resident()
        bytes_out += flush()
        dev_ms = enc.elapsed_ms()
        barrier()
        wall_resident = time.perf_counter() - t0
    bytes_out //= args.steps
    prof_ms, prof_n = enc.profile(0)
    launches = enc.launches - launches0
    # steady-state steps: per stream, the distance between the ends of consecutive segments on the device timeline
    step_ms = [b - a for tl in enc.timeline() for a, b in zip(tl[:-1], tl[1:])] or [dev_ms / args.steps]
    # ---- timed: host buffers through the C ABI (upload + encode + download + access-unit assembly), wall clock
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    flush()
    barrier()
    wall_host = time.perf_counter() - t0

    times = torch.tensor([dev_ms / 1000.0, wall_resident, wall_host], device=f'cuda:{local}', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, wall_res_s, wall_host_s = (float(x) for x in times.cpu())
    total_frames = world * S * n * args.steps
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        w, h = params.width, params.height
        wc, hc = (w + 15) & ~15, (h + 15) & ~15
        clk = clocks.summary()
        # Dominant kernel: k_me, the per-CU motion search (one launch per P frame).  It is bound by integer instruction issue
        # (ncu: profiles/k_me_r2.json), so the roofline is warp instructions per second against SMs x 4 schedulers x clock;
        # the HBM view (algorithmic bytes: luma source + luma reference read once, vectors out) is kept beside it.
        ncu_, nctu_ = (wc // 16) * (hc // 16), ((wc + 31) // 32) * ((hc + 31) // 32)
        me_bytes = 2 * (wc * hc * 2) + nctu_ * 4 + ncu_ * 8
        me_n = max(1, prof_n['me'])
        me_ms = prof_ms['me'] / me_n
        cap = {}
        try:
            cap = json.loads((ROOT / 'profiles' / 'k_me_r2.json').read_text())
            if cap.get('workload') != args.clip:
                cap = {}
        except (OSError, ValueError):
            cap = {}
        inst = cap.get('warp_inst_per_launch')
        sm_mhz = clk.get('sm_mhz') or peaks.get('sm_max_mhz') or 1965.0
        sms = cap.get('sm_count', 148)
        issue_peak = sms * 4 * sm_mhz * 1e6 / 1e9                                    # G warp-instructions / s
        issue_ach = inst / (me_ms * 1e-3) / 1e9 if inst and me_ms > 0 else None
        hbm_ach = me_bytes / (me_ms * 1e-3) / 1e9 if me_ms > 0 else 0.0
        line = {
            'metric': METRICS[args.config], 'value': round(total_frames / dev_s, 3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(1000.0 * dev_s / args.steps, 3), 'ms_per_step_median': round(float(np.median(step_ms)), 3),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u16' if params.bit_depth > 8 else 'u8', 'data': 'synthetic', 'config': workload_config(args, params, world),
            'timing': 'CUDA events: mark on each encoder stream before the first step -> end of its last bitstream download (K pipelined steps + flush), '
                      'max over streams and ranks; ms_per_step_median = median of the steady-state steps on the same device timeline',
            'wall_ms_per_step_resident': round(1000.0 * wall_res_s / args.steps, 3),
            'e2e': {'value': round(total_frames / wall_host_s, 3), 'unit': UNIT, 'h2d_bytes_per_step': S * n * fb, 'd2h_bytes_per_step': bytes_out,
                    'timing': 'wall clock around K x hb_enc_encode_delayed + flush with pinned host buffers'},
            'gpu_launches': int(launches),
            'bitrate_kbps': round(bytes_out / S * 8 / 1000.0 / (n * params.fps_den / params.fps_num), 1),
            'roofline': {'kernel': 'k_me (motion search, one launch per P frame)', 'bound': 'int_issue',
                         'achieved': None if issue_ach is None else round(issue_ach, 2), 'peak': round(issue_peak, 2), 'unit': 'Gwarp-inst/s',
                         'frac': None if issue_ach is None else round(issue_ach / issue_peak, 4),
                         'peak_kind': f'{sms} SMs x 4 schedulers x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)',
                         'warp_inst_per_launch': inst, 'algorithmic_warp_inst_per_launch': cap.get('algorithmic_warp_inst_per_launch'),
                         'traffic': cap.get('dram_bytes_per_launch'), 'avg_launch_ms': round(me_ms, 4),
                         'hbm': {'achieved': round(hbm_ach, 2), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': round(hbm_ach / peaks['hbm_gbs'], 5),
                                 'algorithmic_bytes_per_launch': me_bytes, 'peak_kind': peak_kind},
                         'note': 'instruction-issue bound search kernel (SATD butterflies, dp2a interpolation), DRAM < 2 % of peak; counts from the '
                                 'committed ncu capture profiles/k_me_r2.json, launch time measured live with CUDA events on the encoder stream'},
            'kernel_ms_per_step': {k: round(v / args.steps, 3) for k, v in prof_ms.items()},
            'clocks': clk,
        }
        if cores_bound:
            line['host_cores_per_rank'] = cores_bound
        rd = committed_rd('4k' if params.height >= 2160 else '1080p')
        if rd:
            line['rd_vs_round1'] = rd
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cmodel
            cmodel.build()
            sample_n = min(n, 12)
            frames_h = [host_frames[0][i].numpy() for i in range(sample_n)]
            fps, dt, _ = cpu_model_run(frames_h, params.bit_depth, args.qp, params)
            line['cpu_baseline'] = {'value': round(fps, 4), 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
                                    'sample': f'first {sample_n} frames (1 key frame + {sample_n - 1} P) of the same clip, one encoder on all host '
                                              f'cores (OpenMP), {dt:.1f} s'}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    enc.close()


# ------------------------------------------------------------------------------------------------ config 3: batch of files
def run_batch(args):
    """64 clips cycling the reference's five fixture types (150 frames = 5 s @ 30 each), sharded per FILE over the ranks by
    longest-processing-time-first bin packing (batch.shard_for_rank), and inside a rank handed to `--workers` concurrent
    workers that refill as they finish (the GUI's policy, gui/mainwindow.py:318-341).  Every file goes through the drop-in's
    encode step (transcoder.encode_reader_b200: encoder from the pool, ingest on the device, streaming MP4 mux to disk).  The
    source frames are host buffers (one pinned clip per type: decoding is not what this measures).  Strong scaling."""
    import torch
    import torch.distributed as dist
    from hevc_b200 import batch as hb_batch
    from hevc_b200 import derive, transcoder
    from hevc_b200.frames import MemoryReader
    from hevc_b200.synth import CLIP_TYPES, TorchSynthClip
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    cores_bound = bind_rank_to_cores(local, world) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n_frames, n_files = args.clip_frames, args.files
    files = [(i, BATCH_TYPES[i % len(BATCH_TYPES)]) for i in range(n_files)]
    cost = lambda f: CLIP_TYPES[f[1]][0] * CLIP_TYPES[f[1]][1] * n_frames      # noqa: E731
    mine = hb_batch.shard_for_rank(files, rank, world, cost=cost)               # LPT order inside the shard too
    clips = {}
    for name in sorted({t for _, t in mine}):
        w, h, fps, hdr = CLIP_TYPES[name]
        c = TorchSynthClip(w, h, seed=hash(name) % 1000, device=f'cuda:{local}')
        hbuf = torch.empty((n_frames, c.frame_bytes), dtype=torch.uint8, pin_memory=True)
        for s in range(0, n_frames, 30):
            hbuf[s:s + 30].copy_(c.frames(s, min(30, n_frames - s)))
        clips[name] = hbuf.numpy()
    torch.cuda.synchronize()
    out_dir = Path(tempfile.mkdtemp(prefix='hb_batch_'))
    os.environ['HEVC_B200_POOL'] = str(len(BATCH_TYPES) * args.workers)      # one idle encoder per (geometry, worker): no re-allocation per file

    def one_pass():
        queue = list(mine)
        lock = threading.Lock()
        results = []

        def worker():
            while True:
                with lock:
                    if not queue:
                        return
                    idx, name = queue.pop(0)
                info = clip_info(name, n_frames / CLIP_TYPES[name][2])
                rc, why = transcoder.encode_reader_b200(MemoryReader(clips[name]), f'clip{idx:02d}_{name}', out_dir / f'clip{idx:02d}.mp4', info,
                                                        None, n_frames, None, device=local, batch=args.file_batch)
                with lock:
                    results.append((idx, rc, why))
        threads = [threading.Thread(target=worker) for _ in range(max(1, args.workers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        bad = [r for r in results if r[1] != 0]
        if bad:
            raise SystemExit(f'batch encode failed: {bad[:3]}')
        return sum((out_dir / f'clip{i:02d}.mp4').stat().st_size for i, _ in mine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 1))):
        one_pass()                      # one warm-up pass fills the encoder pool (allocation is per geometry, not per file)
    from hevc_b200 import _cabi
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        t0 = time.perf_counter()
        out_bytes = 0
        for _ in range(args.steps):
            out_bytes += one_pass()
        torch.cuda.synchronize()
        ev1.record()
        torch.cuda.synchronize()
        dev_s = ev0.elapsed_time(ev1) / 1000.0
        barrier()
        wall = time.perf_counter() - t0
    times = torch.tensor([dev_s, wall], device=f'cuda:{local}', dtype=torch.float64)
    loads = torch.tensor([float(sum(cost(f) for f in mine))], device=f'cuda:{local}', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        lmax = loads.clone()
        dist.all_reduce(lmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(loads, op=dist.ReduceOp.SUM)
    else:
        lmax = loads
    if rank == 0:
        total_frames = n_files * n_frames * args.steps
        in_bytes = sum(CLIP_TYPES[t][0] * CLIP_TYPES[t][1] * 3 // 2 * n_frames for _, t in files)
        line = {'metric': METRICS[3], 'value': round(total_frames / float(times[0]), 3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': 1, 'ms_per_step': round(1000.0 * float(times[0]) / args.steps, 2), 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'u8/u16', 'data': 'synthetic',
                'config': {'workload': f'{n_files} files x {n_frames} frames cycling {BATCH_TYPES} (SDR -> Main, HDR-tagged -> Main10 HDR10)',
                           'baseline_config': 3, 'sharding': 'per file, LPT bin packing over ranks (batch.shard_for_rank), '
                           f'{args.workers} refill-on-finish workers per rank (gui/mainwindow.py:318-341), encoders reused through the pool',
                           'lpt_imbalance': round(float(lmax[0]) * world / float(loads[0]), 4), 'file_batch': args.file_batch,
                           'output': 'hvc1 MP4 per file, streamed to a temporary directory'},
                'timing': 'CUDA events on the default stream around K passes over the whole batch (bracketed by device syncs), max over ranks',
                'e2e': {'value': round(total_frames / float(times[1]), 3), 'unit': UNIT, 'h2d_bytes_per_step': in_bytes, 'd2h_bytes_per_step': out_bytes // args.steps,
                        'timing': 'wall clock around the same passes: host frame buffers -> encode_reader_b200 -> MP4 files'},
                'gpu_launches': None, 'clocks': clocks.summary()}
        if cores_bound:
            line['host_cores_per_rank'] = cores_bound
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    transcoder.encoder_pool().close()
    for f in out_dir.glob('*'):
        f.unlink()
    out_dir.rmdir()


# ------------------------------------------------------------------------------------------------ config 4: upscale + segments
def run_upscale(args):
    """One long 1080p30 clip -> 3840x2160 Main10 (geometry rule of upscale_gui_final.py:81-87): the clip is cut into closed-GOP
    segments of one key-frame interval, segment k goes to rank k mod N (the reference's cycle(gpu_list), :25-30,123-126); every
    rank runs the fused scale -> encode on its segments, the bitstreams are gathered on rank 0, concatenated in order and muxed
    ONCE into an hvc1 MP4 (:164-178).  Strong scaling: the clip is fixed, N grows."""
    import torch
    import torch.distributed as dist
    from hevc_b200 import batch as hb_batch
    from hevc_b200 import derive, mp4, upscale
    from hevc_b200.encoder import ParallelSegmentEncoder, crf_to_qp, to_c_params
    from hevc_b200.synth import TorchSynthClip
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    cores_bound = bind_rank_to_cores(local, world) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    sw, sh, fps = 1920, 1080, 30
    tw, th = upscale.target_geometry(sw, sh)
    n_total = args.clip_frames
    info = clip_info('1080p_sdr', n_total / fps)
    info = type(info)(**{**info.__dict__, 'width': tw, 'height': th})
    params = derive.derive_b200_params(info, force_main10=True)
    args.qp = crf_to_qp(params.crf)
    segs = hb_batch.gop_segments(n_total, params.keyint)
    mine = [k for k in range(len(segs)) if k % world == rank]
    S = max(1, args.streams)
    enc = ParallelSegmentEncoder(local, to_c_params(params), streams=S, max_batch=params.keyint)
    clip = TorchSynthClip(sw, sh, seed=3, device=f'cuda:{local}')
    fb = clip.frame_bytes
    dev, host = {}, {}
    for k in mine:
        a, b = segs[k]
        d = clip.frames(a, b - a).contiguous()
        hbuf = torch.empty(d.shape, dtype=torch.uint8, pin_memory=True)
        hbuf.copy_(d)
        dev[k], host[k] = d, hbuf
    torch.cuda.synchronize()
    track = mp4.TrackInfo(params.width, params.height, params.fps_num, params.fps_den, params.profile_idc, params.level_idc, params.tier,
                          params.bit_depth, params.colour_primaries, params.transfer_characteristics, params.matrix_coeffs, params.full_range, None, 0, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(resident: bool):
        runs = {}
        pending = []

        def take(out, stats):
            # completed segments come back in submission order; each is converted to MP4 sample form here, on its own rank
            pos = 0
            while stats and pending:
                k = pending[0]
                cnt = segs[k][1] - segs[k][0]
                if len(stats) < cnt:
                    break
                size = sum(st.bytes for st in stats[:cnt])
                runs[k] = mp4.to_samples(out[pos:pos + size])
                pos += size
                stats = stats[cnt:]
                pending.pop(0)

        for k in mine:
            a, b = segs[k]
            src = dev[k].data_ptr() if resident else host[k].numpy()
            pending.append(k)
            take(*enc.submit(src, b - a, on_device=resident, frame_bytes=fb, src_size=(sw, sh)))
        take(*enc.finish())
        assert not pending
        # host side of the path: gather the per-rank sample runs on rank 0, put the segments back into display order, ONE file
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(runs, gathered, dst=0)
        else:
            gathered = [runs]
        if rank != 0:
            return 0, None
        allruns = {}
        for g_ in gathered:
            allruns.update(g_)
        data = mp4.assemble(track, [allruns[k] for k in range(len(segs))])
        return len(data), data

    for _ in range(max(1, min(args.warmup, 2))):
        one_pass(True)
    res = {}
    for mode in ('resident', 'host'):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            ev0.record()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                size, data = one_pass(mode == 'resident')
            torch.cuda.synchronize()
            ev1.record()
            torch.cuda.synchronize()
            barrier()
            wall = time.perf_counter() - t0
        t = torch.tensor([ev0.elapsed_time(ev1) / 1000.0, wall], device=f'cuda:{local}', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = (float(t[0]), float(t[1]), size, clocks.summary(), data)
    if rank == 0:
        total = n_total * args.steps
        out = ROOT / 'gpurun_out'
        decoded = None
        try:                                  # the muxed result plays as ONE stream: frame count through an independent demuxer + decoder
            import cv2
            out.mkdir(exist_ok=True)
            path = out / f'config4_n{world}.mp4'
            path.write_bytes(res['host'][4])
            cap = cv2.VideoCapture(str(path))
            decoded = 0
            while decoded < n_total + 10:
                ok, _f = cap.read()
                if not ok:
                    break
                decoded += 1
            cap.release()
            path.unlink()
        except Exception as exc:
            decoded = f'not checked: {exc}'
        line = {'metric': METRICS[4], 'value': round(total / res['resident'][0], 3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': 2, 'ms_per_step': round(1000.0 * res['resident'][0] / args.steps, 2), 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'u16', 'data': 'synthetic',
                'config': {'workload': f'one {n_total}-frame {sw}x{sh}@{fps} 8-bit clip -> {tw}x{th} Main10 (SDR signalling kept), fused scale + encode',
                           'baseline_config': 4, 'segments': len(segs), 'segment_frames': params.keyint, 'streams_per_gpu': S,
                           'sharding': 'closed-GOP segment k -> rank k mod N; every rank converts its segments to MP4 sample form, gather_object to rank 0, in-order '
                                       'assembly into ONE hvc1 MP4 (all inside the timed region)',
                           'crf': params.crf, 'vbv_maxrate_kbps': params.vbv_maxrate_kbps, 'level_idc': params.level_idc,
                           'decoded_frames_of_muxed_output': decoded},
                'timing': 'CUDA events on the default stream around K passes (bracketed by device syncs), max over ranks; source segments resident in HBM',
                'e2e': {'value': round(total / res['host'][1], 3), 'unit': UNIT, 'h2d_bytes_per_step': n_total * fb, 'd2h_bytes_per_step': res['host'][2],
                        'timing': 'wall clock, pinned host source frames -> fused scale + encode -> gather -> mux'},
                'gpu_launches': int(enc.launches), 'clocks': res['resident'][3]}
        if cores_bound:
            line['host_cores_per_rank'] = cores_bound
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    enc.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4], help='BASELINE.json configuration (1-based); 2 = the headline metric')
    ap.add_argument('--clip', default=None)
    ap.add_argument('--frames', type=int, default=None, help='frames per step and stream (default: one closed GOP of the configuration)')
    ap.add_argument('--batch', type=int, default=None, help='frames per device batch (default: --frames)')
    ap.add_argument('--streams', type=int, default=3, help='independent encoder streams per GPU (closed-GOP segments round-robin)')
    ap.add_argument('--ref-frames', type=int, default=30, help='--impl reference: frames per step (one closed GOP)')
    ap.add_argument('--files', type=int, default=64, help='config 3: files in the batch')
    ap.add_argument('--clip-frames', type=int, default=None, help='config 3: frames per file (150); config 4: frames of the clip (1800)')
    ap.add_argument('--workers', type=int, default=2, help='config 3: concurrent workers per rank (the GUI default)')
    ap.add_argument('--file-batch', type=int, default=30, help='config 3: frames per device batch of a file worker')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    if args.config in (1, 2):
        from hevc_b200 import derive
        args.clip = args.clip or CONFIG_CLIP[args.config]
        keyint = derive.derive_b200_params(clip_info(args.clip)).keyint
        args.frames = args.frames or keyint
        args.batch = args.batch or args.frames
    if args.impl == 'reference':
        args.clip = args.clip or CONFIG_CLIP[2]
        args.frames = args.frames or 120
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    if args.config == 3:
        args.clip_frames = args.clip_frames or 150
        run_batch(args)
    elif args.config == 4:
        args.clip_frames = args.clip_frames or 1800
        run_upscale(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
