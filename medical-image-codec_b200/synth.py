"""Deterministic synthetic inputs for the BASELINE.json configs.

Language-independent PRNG: splitmix64 finaliser over (seed * 0x9E3779B97F4A7C15 + index),
so the same images can be regenerated from any host language.  Shapes follow the
reference's own generators (makeSmoothFrames multiframe_test.go:115-146,
makeTissueTile / makeWSITestImage wsi_test.go:23-121); only the PRNG differs
(Go's math/rand is not reproducible outside Go).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed: int, n: int, start: int = 0) -> np.ndarray:
    """n 64-bit values: mix(seed*GOLD + start + i)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * _GOLD + np.arange(start, start + n, dtype=np.uint64)
        z = z + _GOLD
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def noise(seed: int, n: int, mod: int, start: int = 0) -> np.ndarray:
    return (splitmix64(seed, n, start) >> np.uint64(33)).astype(np.int64) % mod


def xr_image(seed: int, w: int = 2577, h: int = 2048) -> np.ndarray:
    """Config 2: 12-bit radiograph: ramp + noise[0,64) + ~5 %-area zero collimator border."""
    y = np.arange(h, dtype=np.int64)[:, None]
    x = np.arange(w, dtype=np.int64)[None, :]
    v = (y * 2400) // h + (x * 1200) // w + noise(seed, w * h, 64).reshape(h, w)
    bx, by = int(w * 0.0127), int(h * 0.0127)
    if bx and by:
        v[:by, :] = 0
        v[h - by:, :] = 0
        v[:, :bx] = 0
        v[:, w - bx:] = 0
    return v.astype(np.uint16)


def smooth_image(seed: int, w: int, h: int, noisy_from: int | None = None) -> np.ndarray:
    """12-bit smooth field with changing slopes + noise[0,6): content the gradient-adaptive predictor wins on
    (deltagradcompressu16.go:149-166).  Rows >= noisy_from (if given) are radiograph rows (xr_image) instead, where avg(top,left) wins,
    so a PICA container of the image carries both predictor flags."""
    y = np.arange(h, dtype=np.float64)[:, None]
    x = np.arange(w, dtype=np.float64)[None, :]
    v = (2000.0 + 1500.0 * np.sin(x / 23.0 + seed) * np.cos(y / 31.0)).astype(np.int64)
    v += (np.arange(h, dtype=np.int64)[:, None] * np.arange(w, dtype=np.int64)[None, :]) % 7
    v += noise(seed, w * h, 6).reshape(h, w)
    if noisy_from is not None:
        v[noisy_from:, :] = xr_image(seed + 1, w, 3 * h)[h + noisy_from:2 * h, :]   # middle rows: no top / bottom border
    return v.astype(np.uint16)


def mammo_image(seed: int, rows: int = 4096, cols: int = 3328) -> np.ndarray:
    """Config 3: 14-bit mammogram: breast-shaped mask (~45 % zero background) + smooth field + noise[0,32)."""
    y = np.arange(rows, dtype=np.float64)[:, None] / rows
    x = np.arange(cols, dtype=np.float64)[None, :] / cols
    # half-ellipse attached to the left edge
    inside = (x / 0.82) ** 2 + ((y - 0.5) / 0.47) ** 2 < 1.0
    field = (9000.0 * (1.0 - x) + 3000.0 * np.sin(3.0 * y) ** 2 + 2000.0).astype(np.int64)
    v = field + noise(seed, rows * cols, 32).reshape(rows, cols)
    v = np.where(inside, v, 0)
    return np.clip(v, 0, 16383).astype(np.uint16)


def tomo_stack(seed: int, frames: int = 96, rows: int = 2457, cols: int = 1996) -> np.ndarray:
    """Config 4: 10-bit tomosynthesis stack; frame f = frame f-1 + noise[-5,5] clamped (multiframe_test.go:134-144)."""
    y = np.arange(rows, dtype=np.int64)[:, None]
    x = np.arange(cols, dtype=np.int64)[None, :]
    base = 200 + (y * 500) // rows + (x * 250) // cols + noise(seed, rows * cols, 16).reshape(rows, cols)
    out = np.empty((frames, rows, cols), np.uint16)
    cur = np.clip(base, 0, 1023)
    out[0] = cur
    for f in range(1, frames):
        cur = np.clip(cur + noise(seed + 1000 + f, rows * cols, 11).reshape(rows, cols) - 5, 0, 1023)
        out[f] = cur
    return out


def wsi_region(seed: int, x0: int, y0: int, w: int, h: int, full_w: int, full_h: int) -> np.ndarray:
    """Config 5: RGB8 window [y0:y0+h, x0:x0+w] of a procedurally defined slide (never materialised whole).

    White background; elliptical tissue (~35 % area) with an H&E base colour, per-pixel noise and ~3 % nuclei
    (wsi_test.go:31-45)."""
    yy = (np.arange(y0, y0 + h, dtype=np.int64))[:, None]
    xx = (np.arange(x0, x0 + w, dtype=np.int64))[None, :]
    fy = yy / float(full_h)
    fx = xx / float(full_w)
    tissue = ((fx - 0.5) / 0.36) ** 2 + ((fy - 0.5) / 0.31) ** 2 < 1.0
    idx = (yy * full_w + xx).astype(np.uint64).ravel()
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * _GOLD + idx + _GOLD
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    z = z.reshape(h, w)
    n0 = ((z >> np.uint64(8)) & np.uint64(0xFF)).astype(np.int64) % 15
    n1 = ((z >> np.uint64(20)) & np.uint64(0xFF)).astype(np.int64) % 12
    n2 = ((z >> np.uint64(32)) & np.uint64(0xFF)).astype(np.int64) % 10
    nucleus = ((z >> np.uint64(44)) & np.uint64(0x3FF)).astype(np.int64) < 31
    r = 200 + (30 * yy) // full_h + (10 * xx) // full_w + n0
    g = 140 + (20 * yy) // full_h + n1 + 0 * xx
    b = 170 + (15 * xx) // full_w + n2 + 0 * yy
    r = np.where(nucleus, 80 + n0, r)
    g = np.where(nucleus, 40 + n1, g)
    b = np.where(nucleus, 120 + n2, b)
    out = np.full((h, w, 3), 255, np.uint8)
    out[..., 0] = np.where(tissue, np.clip(r, 0, 255), 255)
    out[..., 1] = np.where(tissue, np.clip(g, 0, 255), 255)
    out[..., 2] = np.where(tissue, np.clip(b, 0, 255), 255)
    return out
