"""Deterministic synthetic clips: the reference's fixtures are `lavfi testsrc` clips
(tests/generate_test_videos.py:10-16) which cannot be generated without ffmpeg, so benches and tests use this
generator: smooth moving gradients + translating textured patches (integer and fractional velocities) +
film-grain-like noise + a hard-edged colour-bar strip.  8-bit limited-range planar 4:2:0 out."""
from __future__ import annotations

import numpy as np

CLIP_TYPES = {  # name -> (width, height, fps, hdr)   -- the reference's five fixture types + BASELINE config 2
    '1080p_sdr': (1920, 1080, 30, False), '720p_sdr': (1280, 720, 30, False), '4k_sdr': (3840, 2160, 30, False),
    '1080p_hdr': (1920, 1080, 30, True), '4k_hdr': (3840, 2160, 30, True), '4k60_hdr': (3840, 2160, 60, True),
}


class SynthClip:
    def __init__(self, width: int, height: int, seed: int = 0, noise: float = 2.0):
        self.w, self.h, self.seed, self.noise = width, height, seed, noise
        rng = np.random.default_rng(seed)
        tw, th = width + 512, height + 512
        # band-limited texture: smoothed noise + a few oriented sinusoids, so that sub-pel motion matters
        base = rng.normal(0, 1, (th // 4 + 2, tw // 4 + 2))
        tex = np.kron(base, np.ones((4, 4)))[:th, :tw]
        k = np.array([1, 4, 6, 4, 1], float) / 16
        for _ in range(2):
            tex = np.apply_along_axis(lambda r: np.convolve(r, k, 'same'), 1, tex)
            tex = np.apply_along_axis(lambda c: np.convolve(c, k, 'same'), 0, tex)
        yy, xx = np.mgrid[0:th, 0:tw]
        tex = tex * 28 + 14 * np.sin(xx / 11.0 + yy / 23.0) + 10 * np.sin(xx / 5.3 - yy / 7.1)
        self.tex = tex
        self.patches = [(rng.integers(0, width - width // 4), rng.integers(0, height - height // 4),
                         width // 6 + int(rng.integers(0, width // 8)), height // 6 + int(rng.integers(0, height // 8)),
                         float(rng.choice([-3, -1.25, 0.5, 2, 3.75])), float(rng.choice([-2, -0.75, 0.25, 1, 2.5])),
                         int(rng.integers(0, 256)), int(rng.integers(0, 256))) for _ in range(5)]

    def frame(self, n: int):
        """(y, u, v) uint8 planes of frame n."""
        w, h = self.w, self.h
        rng = np.random.default_rng(self.seed * 100003 + n)
        yy, xx = np.mgrid[0:h, 0:w]
        # global slow pan of the background texture (0.75 px / frame horizontally, 0.25 vertically)
        ox, oy = 0.75 * n, 0.25 * n
        ix, iy = int(np.floor(ox)), int(np.floor(oy))
        fx, fy = ox - ix, oy - iy
        t = self.tex
        x0, y0 = 128 + ix, 128 + iy
        bg = ((1 - fx) * (1 - fy) * t[y0:y0 + h, x0:x0 + w] + fx * (1 - fy) * t[y0:y0 + h, x0 + 1:x0 + w + 1]
              + (1 - fx) * fy * t[y0 + 1:y0 + h + 1, x0:x0 + w] + fx * fy * t[y0 + 1:y0 + h + 1, x0 + 1:x0 + w + 1])
        luma = 110 + 40 * np.sin((xx + 2 * n) / (w / 6.0)) * np.cos(yy / (h / 4.0)) + bg
        cb = 128 + 30 * np.sin((xx - n) / (w / 3.0)) + 0 * yy
        cr = 128 + 30 * np.cos((yy + n) / (h / 3.0)) + 0 * xx
        for (px, py, pw, ph, vx, vy, pcb, pcr) in self.patches:
            qx, qy = px + vx * n, py + vy * n
            jx, jy = int(np.floor(qx)) % (w - pw), int(np.floor(qy)) % (h - ph)
            sub = t[300 + jy % 64:300 + jy % 64 + ph, 40 + jx % 64:40 + jx % 64 + pw]
            luma[jy:jy + ph, jx:jx + pw] = 128 + 1.6 * sub
            cb[jy:jy + ph, jx:jx + pw] = 0.5 * cb[jy:jy + ph, jx:jx + pw] + 0.5 * pcb
            cr[jy:jy + ph, jx:jx + pw] = 0.5 * cr[jy:jy + ph, jx:jx + pw] + 0.5 * pcr
        bars = h - h // 8                                     # testsrc-like hard-edged bars
        bar_y = np.array([180, 162, 131, 112, 84, 65, 35, 16])[(xx[0] * 8 // w)]
        bar_cb = np.array([128, 44, 156, 72, 184, 100, 212, 128])[(xx[0] * 8 // w)]
        bar_cr = np.array([128, 142, 44, 58, 198, 212, 114, 128])[(xx[0] * 8 // w)]
        luma[bars:, :] = bar_y[None, :]
        cb[bars:, :] = bar_cb[None, :]
        cr[bars:, :] = bar_cr[None, :]
        if self.noise > 0:
            luma = luma + rng.normal(0, self.noise, luma.shape)
        y8 = np.clip(np.rint(luma), 16, 235).astype(np.uint8)
        u8 = np.clip(np.rint(cb[0::2, 0::2]), 16, 240).astype(np.uint8)
        v8 = np.clip(np.rint(cr[0::2, 0::2]), 16, 240).astype(np.uint8)
        return y8, u8, v8
