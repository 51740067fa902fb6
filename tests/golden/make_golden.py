"""Generates the committed golden fixtures.  Run HERE (the container with /root/reference and cv2), not on the GPU box:

    python tests/golden/make_golden.py

 * pyrdown_cv2.npz  : inputs + outputs of the container's cv2 4.13.0 for pyrDown / copyMakeBorder / norm(NORM_L1)
                      (pins the OpenCV arithmetic the reference calls, motion_framework.cpp:60-61,89-90,315).
 * mf_reference.npz : dense CV_32FC2 fields produced by THE REFERENCE ITSELF (oracle/_ref = the reference's
                      motion_framework.cpp compiled from /root/reference against oracle/cvshim) on seeded inputs.
 * flo_gt_digest.json : size + sha256 of the 8 Middlebury gt-flow .flo files and the AEE of a zero field against
                      each, computed by the reference's own Flow::ReadFlowFile / CalculateMSE (rw_flow.cpp).
 * rubberwhale_crop.flo : a 96x64 crop of gt-flow/RubberWhale/flow10.flo that contains unknown-flow pixels
                      (written by the reference's own Flow::WriteFlowFile).
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from blockbasedmotionestimation_b200.synth import make_pair  # noqa: E402
from oracle import binding as ob  # noqa: E402

REF = "/root/reference"

MF_CASES = [
    # name, h, w, search_size, block_size, kind, seed
    ("two_level_bs8", 96, 128, [16, 16], [8, 8], "textured", 11),
    ("padded_three_level", 90, 122, [20, 20, 20], [8, 8, 8], "textured", 12),
    ("constant_all_ties", 96, 128, [16, 16], [8, 8], "constant", 13),
    ("noise_bs4", 120, 200, [14, 12], [4, 4], "noise", 14),
    ("bs2", 64, 96, [6, 6], [2, 2], "noise", 15),
    ("mixed_bs8_bs16", 128, 160, [24, 40], [8, 16], "textured", 16),
    ("bs16_three_level", 192, 256, [48, 48, 48], [16, 16, 16], "textured", 17),
    ("bs32_default_like", 256, 320, [64, 64], [32, 32], "textured", 18),
    ("odd_shift", 100, 140, [17, 9], [8, 8], "textured", 19),
]


def main():
    import cv2
    assert ob.load_ref() is not None, "build oracle/_ref first (make -C oracle ref)"
    # ---- OpenCV arithmetic
    rng = np.random.default_rng(2024)
    out = {}
    for i, (h, w) in enumerate([(64, 64), (96, 128), (50, 70), (51, 71), (34, 38), (272, 480)]):
        a = rng.integers(0, 256, (h, w)).astype(np.uint8)
        out[f"pyr_in_{i}"] = a
        out[f"pyr_out_{i}"] = cv2.pyrDown(a, dstsize=(w // 2, h // 2))
    a = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    out["border_in"] = a
    out["border_out"] = cv2.copyMakeBorder(a, 3, 3, 5, 5, cv2.BORDER_CONSTANT, value=0)
    b = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    out["norm_a"], out["norm_b"] = a, b
    out["norm_l1"] = np.array([cv2.norm(a, b, cv2.NORM_L1)])
    np.savez_compressed(os.path.join(HERE, "pyrdown_cv2.npz"), **out)

    # ---- fields from the reference itself
    ref = {}
    for name, h, w, ss, bs, kind, seed in MF_CASES:
        f1, f2 = make_pair(h, w, seed, shift=(3, -2), max_patch_shift=6, kind=kind)
        flow, dims, _, _ = ob.ref_estimate(f1, f2, ss, bs)
        assert np.all(flow == np.rint(flow)) and np.abs(flow).max() < 32767
        ref[name + "__f1"] = f1
        ref[name + "__f2"] = f2
        ref[name + "__flow"] = flow.astype(np.int16)  # integer-valued: stored compactly
        ref[name + "__dims"] = np.array(dims, np.int32)
        ref[name + "__ss"] = np.array(ss, np.int32)
        ref[name + "__bs"] = np.array(bs, np.int32)
    np.savez_compressed(os.path.join(HERE, "mf_reference.npz"), **ref)

    # ---- .flo known answers, via the reference's own Flow class
    lib = ob.load_ref()
    digest = {}
    for seq in sorted(os.listdir(os.path.join(REF, "middlebury", "gt-flow"))):
        path = os.path.join(REF, "middlebury", "gt-flow", seq, "flow10.flo")
        raw = open(path, "rb").read()
        cap = (len(raw) - 12) // 4
        buf = np.empty(cap, np.float32)
        w, h = C.c_int(0), C.c_int(0)
        assert lib.ref_flow_read(path.encode(), buf.ctypes.data, C.byref(w), C.byref(h), cap) == 0
        gt = buf.reshape(h.value, w.value, 2)
        zero = np.zeros_like(gt)
        unknown = (np.abs(gt[..., 0]) > 1e9) | (np.abs(gt[..., 1]) > 1e9) | np.isnan(gt).any(-1)
        digest[seq] = {"bytes": len(raw), "sha256": hashlib.sha256(raw).hexdigest(), "width": w.value, "height": h.value,
                       "unknown_pixels": int(unknown.sum()),
                       "aee_of_zero_field": lib.ref_flow_mse(gt.ctypes.data, zero.ctypes.data, w.value, h.value)}
        if seq == "RubberWhale":
            ys, xs = np.nonzero(unknown)
            y0 = max(0, min(int(ys[len(ys) // 2]) - 32, h.value - 64))
            x0 = max(0, min(int(xs[len(xs) // 2]) - 48, w.value - 96))
            crop = np.ascontiguousarray(gt[y0:y0 + 64, x0:x0 + 96])
            assert lib.ref_flow_write(os.path.join(HERE, "rubberwhale_crop.flo").encode(), crop.ctypes.data, 96, 64) == 0
            est = np.zeros_like(crop)
            est[..., 0] = 0.25
            digest["rubberwhale_crop"] = {"origin": [x0, y0], "unknown_pixels": int(unknown[y0:y0 + 64, x0:x0 + 96].sum()),
                                          "aee_of_quarter_pixel_field": lib.ref_flow_mse(crop.ctypes.data, est.ctypes.data, 96, 64)}
    json.dump(digest, open(os.path.join(HERE, "flo_gt_digest.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
