"""Shared helpers of the parity tests: conversions between the oracle's dense float fields (the reference's
level_flow layout: MV at each block's top-left pixel) and the library's block-granular int16 fields."""
import numpy as np

from blockbasedmotionestimation_b200.synth import make_pair  # noqa: F401


def dense_to_blocks(flow, g):
    return np.ascontiguousarray(np.rint(flow[::g, ::g, :]).astype(np.int16))


def blocks_to_dense(mv, g, h, w):
    out = np.zeros((h, w, 2), np.float32)
    out[::g, ::g, :] = mv.astype(np.float32)
    return out


def mv2_to_dense(mv2):
    """2x2-granular int16 field -> dense float field (what MF::calcMotionBlockMatching returns)."""
    return np.repeat(np.repeat(mv2, 2, axis=0), 2, axis=1).astype(np.float32)


def describe_diff(a, b):
    d = (a != b)
    if d.ndim == 3:
        d = d.any(-1)
    idx = np.argwhere(d)
    head = ", ".join(f"{tuple(i)}: {a[tuple(i)].tolist()} vs {b[tuple(i)].tolist()}" for i in idx[:5])
    return f"{len(idx)} of {d.size} entries differ; first: {head}"


def up4(img):
    """x4 bilinear up-sampling with the pixel-centre convention of cv2.resize(INTER_LINEAR) (main_class.cpp:32-33), in
    22-bit fixed point.  It is the STAND-IN's resampler (deterministic on every box); cv2's own fixed-point variant differs
    from it by +-1 LSB on ~12 % of the pixels.  The up-sampled frames are inputs to the path: oracle and GPU get the
    same bytes."""
    h, w = img.shape

    def coords(n):
        s = (np.arange(4 * n) + 0.5) / 4 - 0.5
        i0 = np.floor(s).astype(np.int64)
        f = s - i0
        return np.clip(i0, 0, n - 1), np.clip(i0 + 1, 0, n - 1), np.rint(f * 2048).astype(np.int64)

    y0, y1, fy = coords(h)
    x0, x1, fx = coords(w)
    a = img.astype(np.int64)
    top = a[y0][:, x0] * (2048 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (2048 - fx) + a[y1][:, x1] * fx
    v = (top * (2048 - fy)[:, None] + bot * fy[:, None] + (1 << 21)) >> 22
    return np.clip(v, 0, 255).astype(np.uint8)
