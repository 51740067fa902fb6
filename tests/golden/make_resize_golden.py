"""Generates tests/golden/resize_cv2.npz.  Run HERE (the container with cv2 4.13.0), not on the GPU box:

    python tests/golden/make_resize_golden.py

Pins the arithmetic of cv::resize(img, img, Size(), f, f, INTER_LINEAR) on 8-bit images -- the call main() makes before
constructing MF (reference main_class.cpp:32-33) -- as computed by the container's cv2: inputs and cv2's outputs for the
factors the CUDA path supports (2, 4, 8), random and smooth content, odd sizes, a degenerate 2x3 image."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(4242)
    out = {}
    cases = [((37, 53), 4, "noise"), ((33, 47), 2, "noise"), ((20, 30), 8, "noise"), ((48, 64), 4, "smooth"), ((2, 3), 4, "noise"),
             ((97, 146), 4, "smooth")]
    for i, ((h, w), f, kind) in enumerate(cases):
        a = rng.integers(0, 256, (h, w)).astype(np.uint8)
        if kind == "smooth":
            a = cv2.GaussianBlur(cv2.resize(rng.integers(0, 256, (h // 4 + 1, w // 4 + 1)).astype(np.uint8), (w, h),
                                            interpolation=cv2.INTER_CUBIC), (5, 5), 1.2)
        out[f"in_{i}"] = a
        out[f"factor_{i}"] = np.array([f])
        out[f"out_{i}"] = cv2.resize(a, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR)
        assert out[f"out_{i}"].shape == (h * f, w * f)
    out["cv2_version"] = np.array([cv2.__version__])
    np.savez_compressed(os.path.join(HERE, "resize_cv2.npz"), **out)
    print("wrote resize_cv2.npz:", {k: v.shape for k, v in out.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
