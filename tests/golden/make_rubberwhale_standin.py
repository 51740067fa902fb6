"""Builds the stand-in for BASELINE config 0 (Middlebury RubberWhale, 584x388).  Run HERE (needs /root/reference):

    python tests/golden/make_rubberwhale_standin.py

The real frame10/frame11 PNGs are not in the reference tree (.gitignore:8) and there is no network, so:
frame10 = a synthetic texture, frame11 = frame10 warped by the REAL ground truth gt-flow/RubberWhale/flow10.flo
(unknown pixels -> zero flow).  Output: tests/golden/rubberwhale_standin.npz with frame10, frame11 (uint8) and the
ground-truth flow (float32, bit-identical to the .flo payload) -- a stand-in, clearly not the Middlebury images.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from blockbasedmotionestimation_b200.synth import _box5  # noqa: E402
from oracle import binding as ob  # noqa: E402


def texture(h, w, seed):
    rng = np.random.default_rng(seed)
    fine = _box5(rng.integers(0, 256, (h, w)).astype(np.float32))
    coarse = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
    coarse = np.kron(coarse, np.ones((8, 8), np.float32))[:h, :w]
    coarse = _box5(_box5(coarse))
    t = 0.55 * fine + 0.45 * coarse
    lo, hi = np.percentile(t, [1, 99])
    return np.clip((t - lo) * (255.0 / (hi - lo)), 0, 255)


def bilinear_sample(img, xs, ys):
    h, w = img.shape
    xs = np.clip(xs, 0, w - 1.001)
    ys = np.clip(ys, 0, h - 1.001)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    fx, fy = xs - x0, ys - y0
    a = img[y0, x0] * (1 - fx) + img[y0, x0 + 1] * fx
    b = img[y0 + 1, x0] * (1 - fx) + img[y0 + 1, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def main():
    gt = ob.flo_read("/root/reference/middlebury/gt-flow/RubberWhale/flow10.flo")
    h, w, _ = gt.shape
    f10 = texture(h, w, 10)
    flow = gt.copy()
    unknown = (np.abs(flow[..., 0]) > 1e9) | (np.abs(flow[..., 1]) > 1e9)
    flow[unknown] = 0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    # frame11(p) = frame10(p - f(p)): to first order the block of frame10 at p reappears at p + f(p)
    f11 = bilinear_sample(f10, xx - flow[..., 0], yy - flow[..., 1])
    out = os.path.join(HERE, "rubberwhale_standin.npz")
    np.savez_compressed(out, frame10=np.rint(f10).astype(np.uint8), frame11=np.rint(f11).astype(np.uint8), gt=gt)
    print("wrote", out, os.path.getsize(out), "bytes; unknown gt pixels:", int(unknown.sum()))


if __name__ == "__main__":
    main()
