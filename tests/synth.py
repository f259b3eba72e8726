"""Seeded synthetic 4:2:0 content: low-pass noise background translating by
(3,5) px/frame, three textured objects with their own motion, N(0,2) sensor noise."""
from __future__ import annotations

import numpy as np

from thevc_b200.tlibcuda import HostPic

SEED = 20261018


def _lowpass(a: np.ndarray, k: int) -> np.ndarray:
    out = a.astype(np.float64)
    for axis in (0, 1):
        acc = np.zeros_like(out)
        for s in range(-k, k + 1):
            acc += np.roll(out, s, axis=axis)
        out = acc / (2 * k + 1)
    return out


def make_sequence(w: int, h: int, n: int, seed: int = SEED, bit_depth: int = 8):
    rng = np.random.default_rng(seed)
    big = 256
    bg = _lowpass(rng.uniform(0, 255, (h + big, w + big)), 2)
    bg = (bg - bg.min()) / (bg.max() - bg.min()) * 200 + 20
    tex = [rng.uniform(0, 255, (min(96, h // 2), min(96, w // 2))) for _ in range(3)]
    vel = [(2, -3), (-5, 1), (1, 4)]
    pos0 = [(h // 5, w // 6), (h // 2, w // 2), (h // 3, (2 * w) // 3)]
    frames = []
    maxv = (1 << bit_depth) - 1
    for t in range(n):
        oy, ox = (5 * t) % big, (3 * t) % big
        y = bg[oy:oy + h, ox:ox + w].copy()
        for k in range(3):
            th, tw = tex[k].shape
            py = (pos0[k][0] + vel[k][0] * t) % max(1, h - th)
            px = (pos0[k][1] + vel[k][1] * t) % max(1, w - tw)
            y[py:py + th, px:px + tw] = tex[k]
        y = y + rng.normal(0, 2, y.shape)
        yy = np.clip(np.rint(y), 0, 255)
        u = np.clip(np.rint(128 + 0.3 * (yy[::2, ::2] - 128) + rng.normal(0, 1, (h // 2, w // 2))), 0, 255)
        v = np.clip(np.rint(128 - 0.2 * (yy[::2, ::2] - 128) + rng.normal(0, 1, (h // 2, w // 2))), 0, 255)
        sc = 1 << (bit_depth - 8)
        frames.append(((yy * sc).clip(0, maxv).astype(np.int16), (u * sc).clip(0, maxv).astype(np.int16),
                       (v * sc).clip(0, maxv).astype(np.int16)))
    return frames


def to_hostpic(frame, w: int, h: int, extend: bool = True) -> HostPic:
    p = HostPic(w, h)
    p.y[:] = frame[0]
    p.u[:] = frame[1]
    p.v[:] = frame[2]
    if extend:
        p.extend_border()
    return p


def random_pic(rng, w: int, h: int, bit_depth: int = 8, extend: bool = True) -> HostPic:
    p = HostPic(w, h)
    p.y[:] = rng.integers(0, 1 << bit_depth, (h, w))
    p.u[:] = rng.integers(0, 1 << bit_depth, (h // 2, w // 2))
    p.v[:] = rng.integers(0, 1 << bit_depth, (h // 2, w // 2))
    if extend:
        p.extend_border()
    return p
