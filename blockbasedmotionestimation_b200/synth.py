"""Seeded synthetic 8-bit luma frame pairs: textured base, global shift, moving patches, noise."""
import numpy as np


def _box5(a):
    """5x5 box blur with edge replication (numpy only)."""
    p = np.pad(a, 2, mode="edge")
    c = np.cumsum(np.cumsum(p, axis=0, dtype=np.float64), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    s = c[5:, 5:] - c[:-5, 5:] - c[5:, :-5] + c[:-5, :-5]
    return (s / 25.0).astype(np.float32)


def make_pair(height, width, seed, shift=(5, -3), patches=6, max_patch_shift=12, noise=2, kind="textured"):
    """Returns (frame1, frame2) uint8.  frame2(x, y) = frame1(x + shift[0], y + shift[1]) away from patches, i.e.
    the true motion vector is (-shift[0], -shift[1]).  kind: textured | noise | constant."""
    rng = np.random.default_rng(seed)
    if kind == "constant":
        v = int(rng.integers(0, 256))
        f = np.full((height, width), v, np.uint8)
        return f, f.copy()
    m = 64
    big = rng.integers(0, 256, (height + 2 * m, width + 2 * m)).astype(np.float32)
    if kind == "textured":
        big = _box5(_box5(big))
        lo, hi = np.percentile(big, [1, 99])
        big = np.clip((big - lo) * (255.0 / max(hi - lo, 1e-6)), 0, 255)
    big = np.rint(big).astype(np.int16)
    dx, dy = shift
    f1 = big[m:m + height, m:m + width].copy()
    f2 = big[m + dy:m + dy + height, m + dx:m + dx + width].copy()
    for _ in range(patches):
        ph = int(rng.integers(max(8, height // 16), max(9, height // 4)))
        pw = int(rng.integers(max(8, width // 16), max(9, width // 4)))
        y = int(rng.integers(0, max(1, height - ph)))
        x = int(rng.integers(0, max(1, width - pw)))
        ox = int(rng.integers(-max_patch_shift, max_patch_shift + 1))
        oy = int(rng.integers(-max_patch_shift, max_patch_shift + 1))
        f2[y:y + ph, x:x + pw] = big[m + y + oy:m + y + oy + ph, m + x + ox:m + x + ox + pw]
    if noise:
        f2 = f2 + rng.integers(-noise, noise + 1, f2.shape, dtype=np.int16)
    return np.clip(f1, 0, 255).astype(np.uint8), np.clip(f2, 0, 255).astype(np.uint8)


def seed_for(config, pair_index):
    return 1000 * int(config) + int(pair_index)
