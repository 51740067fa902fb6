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
