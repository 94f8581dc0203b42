"""Synthetic workload of BASELINE.json configs 2-4: 1920x1080 gray frames, moving
filled ellipses over a noisy background.  Pure input generation (numpy); not part of the hot path."""
from __future__ import annotations

import numpy as np


def make_background(h=1080, w=1920, seed=1234):
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(150, 3, (h, w))), 0, 255).astype(np.uint8)


class BlobWorld:
    def __init__(self, h=1080, w=1920, n_blobs=100, seed=1234, semi=(22, 6), margin=40):
        self.h, self.w, self.n = h, w, n_blobs
        self.rng = np.random.default_rng(seed + 1)
        self.bg = make_background(h, w, seed)
        self.pos = np.stack([self.rng.uniform(margin, w - margin, n_blobs), self.rng.uniform(margin, h - margin, n_blobs)], 1)
        self.vel = self.rng.uniform(-4, 4, (n_blobs, 2))
        self.ang = self.rng.uniform(0, np.pi, n_blobs)
        self.grey = self.rng.integers(20, 80, n_blobs)
        self.semi, self.margin = semi, margin

    def step(self):
        self.pos += self.vel
        for d, lim in ((0, self.w), (1, self.h)):
            lo, hi = self.margin, lim - self.margin
            bad = (self.pos[:, d] < lo) | (self.pos[:, d] > hi)
            self.vel[bad, d] *= -1
            self.pos[:, d] = np.clip(self.pos[:, d], lo, hi)

    def frame(self):
        noise = self.rng.integers(-3, 4, (self.h, self.w))
        fr = np.clip(self.bg.astype(np.int16) + noise, 1, 255).astype(np.uint8)
        a, b = self.semi
        r = int(max(a, b)) + 2
        yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
        for i in range(self.n):
            cx, cy = self.pos[i]
            ix, iy = int(round(cx)), int(round(cy))
            c, s = np.cos(self.ang[i]), np.sin(self.ang[i])
            u = (xx + ix - cx) * c + (yy + iy - cy) * s
            v = -(xx + ix - cx) * s + (yy + iy - cy) * c
            m = (u / a) ** 2 + (v / b) ** 2 <= 1.0
            y0, y1, x0, x1 = iy - r, iy + r + 1, ix - r, ix + r + 1
            sy0, sx0 = max(0, -y0), max(0, -x0)
            y0, x0 = max(y0, 0), max(x0, 0)
            y1, x1 = min(y1, self.h), min(x1, self.w)
            sub = m[sy0:sy0 + (y1 - y0), sx0:sx0 + (x1 - x0)]
            fr[y0:y1, x0:x1][sub] = self.grey[i]
        self.step()
        return fr

    def frames(self, n):
        return np.stack([self.frame() for _ in range(n)])


def to_color(gray: np.ndarray, seed=0, channels=3) -> np.ndarray:
    """Grey frames / background (..., H, W) -> interleaved B,G,R(,A) images (..., H, W, channels): every channel is the
    grey value times a fixed gain plus a small per-pixel offset, so that cvtColor of the result stays close to the
    input (the colour variant of the same workload; video sources hand TRex BGR(A) frames)."""
    rng = np.random.default_rng(seed)
    g = gray.astype(np.int16)[..., None]
    gains = np.array([0.9, 1.0, 1.1])
    out = np.clip(np.rint(g * gains) + rng.integers(-1, 2, gray.shape + (3,)), 0, 255).astype(np.uint8)
    if channels == 4:
        out = np.concatenate([out, np.full(gray.shape + (1,), 255, np.uint8)], -1)
    return out
