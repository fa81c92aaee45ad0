"""Synthetic page photos for benchmarks and tests.

`synth_page_numpy` is the host restatement of the device generator in csrc/synth.cu (same hashes, same
layout; float rounding may differ by +-1 grey level) so that the CPU reference arm of bench.py can build its
inputs without touching the GPU.  `synth_quad` gives the page quad (TL, TR, BR, BL) both generators use.
"""
from __future__ import annotations

import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = np.asarray(x, np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def _seed32(seed: int) -> int:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    hi = int(_mix32(((seed >> 32) + 0x9E3779B9) & 0xFFFFFFFF))
    return int(_mix32((seed & 0xFFFFFFFF) ^ hi))


def synth_quad(seed: int, width: int, height: int) -> np.ndarray:
    W, H = np.float32(width), np.float32(height)
    s32 = _seed32(seed)
    base = np.array([0.10 * W, 0.07 * H, 0.90 * W, 0.09 * H, 0.93 * W, 0.93 * H, 0.07 * W, 0.91 * H], np.float32)
    base = np.array([np.float32(0.10) * W, np.float32(0.07) * H, np.float32(0.90) * W, np.float32(0.09) * H,
                     np.float32(0.93) * W, np.float32(0.93) * H, np.float32(0.07) * W, np.float32(0.91) * H], np.float32)
    jitter = np.float32(0.02) * min(W, H)
    out = np.empty(8, np.float32)
    for i in range(8):
        r = int(_mix32((s32 + 101 * (i + 1)) & 0xFFFFFFFF))
        out[i] = base[i] + jitter * (np.float32(r & 0xFFFF) / np.float32(32767.5) - np.float32(1.0))
    return out.reshape(4, 2)


def _perspective(quad, rect):
    a, b = [], []
    for (sx, sy), (dx, dy) in zip(quad, rect):
        a.append([sx, sy, 1, 0, 0, 0, -sx * dx, -sy * dx]); b.append(dx)
    for (sx, sy), (dx, dy) in zip(quad, rect):
        a.append([0, 0, 0, sx, sy, 1, -sx * dy, -sy * dy]); b.append(dy)
    m = np.linalg.solve(np.array(a, np.float64), np.array(b, np.float64))
    return np.append(m, 1.0).astype(np.float32)


def synth_page_numpy(seed: int, width: int = 3000, height: int = 4000):
    """Returns (photo HxWx3 uint8 BGR, quad (4,2) float32)."""
    f32 = np.float32
    quad = synth_quad(seed, width, height)
    s32 = np.uint64(_seed32(seed))
    ph = f32(0.85) * f32(height)
    pw = ph / f32(1.41421356)
    rect = np.array([[0, 0], [pw - 1, 0], [pw - 1, ph - 1], [0, ph - 1]], np.float32)
    hinv = _perspective(quad.astype(np.float64), rect.astype(np.float64))
    ys, xs = np.mgrid[0:height, 0:width]
    acc = np.zeros((height, width), np.float32)
    line_pitch = ph / f32(75.5)
    top = f32(2.6) * line_pitch
    margin = f32(0.05) * pw
    cell = pw / f32(26.0)
    for s in range(4):
        fx = xs.astype(np.float32) + (f32(0.25) if s & 1 else f32(-0.25))
        fy = ys.astype(np.float32) + (f32(0.25) if s & 2 else f32(-0.25))
        ww = hinv[6] * fx + hinv[7] * fy + hinv[8]
        u = (hinv[0] * fx + hinv[1] * fy + hinv[2]) / ww
        v = (hinv[3] * fx + hinv[4] * fy + hinv[5]) / ww
        off = (u < 0) | (v < 0) | (u >= pw) | (v >= ph)
        lf = (v - top) / line_pitch
        line = np.floor(lf)
        val = np.full((height, width), 235.0, np.float32)
        in_lines = (lf >= 0) & (line < 70)
        line_u = np.where(in_lines, line, 0).astype(np.uint64)
        hl = _mix32((s32 * np.uint64(2654435761) + line_u * np.uint64(97) + np.uint64(13)) & _M32)
        text_h = line_pitch * (f32(0.24) + f32(0.22) * (hl & np.uint64(255)).astype(np.float32) / f32(255.0))
        in_text = in_lines & ((lf - line) * line_pitch <= text_h)
        uf = (u - margin) / cell
        word = np.floor(uf)
        in_cols = (uf >= 0) & (u <= pw - margin)
        word_u = np.where(in_cols, word, 0).astype(np.uint64)
        hw = _mix32((hl + word_u * np.uint64(7919)) & _M32)
        filled = (hw & np.uint64(127)) <= np.uint64(108)
        gap = cell * (f32(0.12) + f32(0.2) * ((hw >> np.uint64(8)) & np.uint64(255)).astype(np.float32) / f32(255.0))
        inside = (uf - word) * cell
        ink = in_text & in_cols & filled & (inside >= gap)
        val = np.where(ink, f32(20.0) + ((hw >> np.uint64(16)) % np.uint64(70)).astype(np.float32), val)
        acc += np.where(off, f32(40.0), val)
    val = acc * f32(0.25)
    val = val * (f32(0.55) + f32(0.45) * (f32(0.6) * xs.astype(np.float32) / f32(width) + f32(0.4) * ys.astype(np.float32) / f32(height)))
    n = _mix32(s32 ^ _mix32((ys.astype(np.uint64) * np.uint64(65537) + xs.astype(np.uint64)) & _M32))
    noise = ((n & np.uint64(255)).astype(np.float32) + ((n >> np.uint64(8)) & np.uint64(255)).astype(np.float32)
             + ((n >> np.uint64(16)) & np.uint64(255)).astype(np.float32) + (n >> np.uint64(24)).astype(np.float32)
             - f32(510.0)) * f32(3.0 / 147.8)
    out = np.empty((height, width, 3), np.uint8)
    for c, g in enumerate((0.97, 1.0, 1.02)):
        out[:, :, c] = np.clip(np.rint(val * f32(g) + noise), 0, 255).astype(np.uint8)
    return out, quad


def synth_angle(seed: int) -> float:
    """Deskew angle the bench supplies for page `seed`: a multiple of 0.5 degrees in [-2, 2]."""
    return (int(_mix32((_seed32(seed) + 77) & 0xFFFFFFFF)) % 9 - 4) * 0.5
