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
    def __init__(self, width: int, height: int, seed: int = 0, noise: float = 2.0, pan=(0.75, 0.25), patch_speed: float = 1.0):
        self.w, self.h, self.seed, self.noise = width, height, seed, noise
        self.pan, self.patch_speed = pan, patch_speed
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
        # global pan of the background texture (default 0.75 px / frame horizontally, 0.25 vertically)
        ox, oy = self.pan[0] * n, self.pan[1] * n
        ix, iy = int(np.floor(ox)), int(np.floor(oy))
        fx, fy = ox - ix, oy - iy
        t = self.tex
        x0, y0 = 128 + ix % 256, 128 + iy % 256
        bg = ((1 - fx) * (1 - fy) * t[y0:y0 + h, x0:x0 + w] + fx * (1 - fy) * t[y0:y0 + h, x0 + 1:x0 + w + 1]
              + (1 - fx) * fy * t[y0 + 1:y0 + h + 1, x0:x0 + w] + fx * fy * t[y0 + 1:y0 + h + 1, x0 + 1:x0 + w + 1])
        luma = 110 + 40 * np.sin((xx + 2 * n) / (w / 6.0)) * np.cos(yy / (h / 4.0)) + bg
        cb = 128 + 30 * np.sin((xx - n) / (w / 3.0)) + 0 * yy
        cr = 128 + 30 * np.cos((yy + n) / (h / 3.0)) + 0 * xx
        for (px, py, pw, ph, vx, vy, pcb, pcr) in self.patches:
            qx, qy = px + vx * self.patch_speed * n, py + vy * self.patch_speed * n
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


class TorchSynthClip:
    """Same recipe as ``SynthClip`` generated with torch on a device (fast enough for 4K60 benches).  Output layout is
    the encoder's packed planar 4:2:0 8-bit frame: Y plane, U plane, V plane.  Deterministic per (seed, frame)."""

    def __init__(self, width: int, height: int, seed: int = 0, noise: float = 2.0, device='cuda'):
        import torch
        import torch.nn.functional as F
        self.torch, self.w, self.h, self.seed, self.noise, self.device = torch, width, height, seed, noise, device
        g = torch.Generator(device='cpu').manual_seed(seed)
        tw, th = width + 512, height + 512
        low = torch.randn((1, 1, th // 8 + 2, tw // 8 + 2), generator=g).to(device)
        tex = F.interpolate(low, size=(th, tw), mode='bicubic', align_corners=False)[0, 0]
        yy, xx = torch.meshgrid(torch.arange(th, device=device, dtype=torch.float32),
                                torch.arange(tw, device=device, dtype=torch.float32), indexing='ij')
        self.tex = tex * 28 + 14 * torch.sin(xx / 11.0 + yy / 23.0) + 10 * torch.sin(xx / 5.3 - yy / 7.1)
        self.yy, self.xx = yy[:height, :width], xx[:height, :width]
        rng = np.random.default_rng(seed)
        self.patches = [(int(rng.integers(0, width - width // 4)), int(rng.integers(0, height - height // 4)),
                         width // 6 + int(rng.integers(0, width // 8)), height // 6 + int(rng.integers(0, height // 8)),
                         float(rng.choice([-3, -1.25, 0.5, 2, 3.75])), float(rng.choice([-2, -0.75, 0.25, 1, 2.5])),
                         int(rng.integers(0, 256)), int(rng.integers(0, 256))) for _ in range(5)]
        bar = (self.xx[0] * 8 / width).long().clamp(0, 7)
        self.bar_y = torch.tensor([180, 162, 131, 112, 84, 65, 35, 16], device=device, dtype=torch.float32)[bar]
        self.bar_cb = torch.tensor([128, 44, 156, 72, 184, 100, 212, 128], device=device, dtype=torch.float32)[bar]
        self.bar_cr = torch.tensor([128, 142, 44, 58, 198, 212, 114, 128], device=device, dtype=torch.float32)[bar]

    @property
    def frame_bytes(self) -> int:
        return self.w * self.h * 3 // 2

    def frame(self, n: int):
        torch = self.torch
        w, h, t = self.w, self.h, self.tex
        ox, oy = 0.75 * n, 0.25 * n
        ix, iy = int(np.floor(ox)), int(np.floor(oy))
        fx, fy = ox - ix, oy - iy
        x0, y0 = 128 + ix % 256, 128 + iy % 256
        bg = ((1 - fx) * (1 - fy) * t[y0:y0 + h, x0:x0 + w] + fx * (1 - fy) * t[y0:y0 + h, x0 + 1:x0 + w + 1]
              + (1 - fx) * fy * t[y0 + 1:y0 + h + 1, x0:x0 + w] + fx * fy * t[y0 + 1:y0 + h + 1, x0 + 1:x0 + w + 1])
        luma = 110 + 40 * torch.sin((self.xx + 2 * n) / (w / 6.0)) * torch.cos(self.yy / (h / 4.0)) + bg
        cb = 128 + 30 * torch.sin((self.xx - n) / (w / 3.0))
        cr = 128 + 30 * torch.cos((self.yy + n) / (h / 3.0))
        for (px, py, pw, ph, vx, vy, pcb, pcr) in self.patches:
            jx, jy = int(np.floor(px + vx * n)) % (w - pw), int(np.floor(py + vy * n)) % (h - ph)
            sub = t[300 + jy % 64:300 + jy % 64 + ph, 40 + jx % 64:40 + jx % 64 + pw]
            luma[jy:jy + ph, jx:jx + pw] = 128 + 1.6 * sub
            cb[jy:jy + ph, jx:jx + pw] = 0.5 * cb[jy:jy + ph, jx:jx + pw] + 0.5 * pcb
            cr[jy:jy + ph, jx:jx + pw] = 0.5 * cr[jy:jy + ph, jx:jx + pw] + 0.5 * pcr
        bars = h - h // 8
        luma[bars:, :] = self.bar_y[None, :]
        cb[bars:, :] = self.bar_cb[None, :]
        cr[bars:, :] = self.bar_cr[None, :]
        if self.noise > 0:
            g = torch.Generator(device=self.device).manual_seed(self.seed * 100003 + n)
            luma = luma + torch.randn(luma.shape, generator=g, device=self.device) * self.noise
        y8 = luma.round().clamp(16, 235).to(torch.uint8)
        u8 = cb[0::2, 0::2].round().clamp(16, 240).to(torch.uint8)
        v8 = cr[0::2, 0::2].round().clamp(16, 240).to(torch.uint8)
        return torch.cat([y8.reshape(-1), u8.reshape(-1), v8.reshape(-1)])

    def frames(self, start: int, count: int):
        """uint8 tensor [count, frame_bytes] on the device"""
        return self.torch.stack([self.frame(start + i) for i in range(count)])


CONTENT_CLASSES = ('base', 'hardcut', 'pan', 'static', 'grain', 'clean')


def content_clip(kind: str, width: int, height: int, n_frames: int, seed: int = 0):
    """The R-D harness's content classes -> list of (y, u, v) uint8 frames.  ``base``: the default recipe (slow pan, moving
    patches that cover and uncover background, grain sigma 2); ``hardcut``: two unrelated scenes joined in the middle;
    ``pan``: fast global pan (6.5, 2.25 px / frame); ``static``: nothing moves, only the grain changes; ``grain``: sigma 5; ``clean``: no grain."""
    if kind == 'base':
        c = SynthClip(width, height, seed=seed)
        return [c.frame(i) for i in range(n_frames)]
    if kind == 'hardcut':
        a, b = SynthClip(width, height, seed=seed), SynthClip(width, height, seed=seed + 100)
        cut = n_frames // 2 + 1
        return [a.frame(i) if i < cut else tuple(np.ascontiguousarray(p[::-1, ::-1]) for p in b.frame(i + 37)) for i in range(n_frames)]
    if kind == 'pan':
        c = SynthClip(width, height, seed=seed, pan=(6.5, 2.25))
        return [c.frame(i) for i in range(n_frames)]
    if kind == 'static':
        c = SynthClip(width, height, seed=seed, pan=(0.0, 0.0), patch_speed=0.0)
        return [c.frame(i) for i in range(n_frames)]
    if kind == 'clean':
        c = SynthClip(width, height, seed=seed, noise=0.0)
        return [c.frame(i) for i in range(n_frames)]
    if kind == 'grain':
        c = SynthClip(width, height, seed=seed, noise=5.0)
        return [c.frame(i) for i in range(n_frames)]
    raise ValueError(kind)
