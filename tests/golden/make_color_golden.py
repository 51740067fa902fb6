"""Generates tests/golden/flow_color_ref.npz.  Run HERE (the container with /root/reference), not on the GPU box:

    python tests/golden/make_color_golden.py

Input fields (random vectors, unknown-flow pixels, an all-zero field, a crop of a real Middlebury ground-truth file) and
the images THE REFERENCE ITSELF produces for them: Flow::MotionToColor of rw_flow.cpp:202-300, compiled from
/root/reference into oracle/_ref (see oracle/Makefile).  Pins bbme_flow_to_color."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402


def main():
    assert ob.load_ref() is not None, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(99)
    out = {}
    a = (rng.standard_normal((40, 56, 2)) * 5).astype(np.float32)
    a[3:6, 10:20] = 1e10
    a[20, 21, 1] = np.nan
    fields = [(a, -1.0), (a, 3.0), (np.zeros((8, 12, 2), np.float32), -1.0),
              (ob.flo_read("/root/reference/middlebury/gt-flow/RubberWhale/flow10.flo")[100:148, 200:264].copy(), -1.0)]
    for i, (f, mm) in enumerate(fields):
        out[f"flow_{i}"] = f
        out[f"maxmotion_{i}"] = np.array([mm], np.float32)
        out[f"bgr_{i}"] = ob.ref_flow_color(f, mm)
    np.savez_compressed(os.path.join(HERE, "flow_color_ref.npz"), **out)
    print("wrote flow_color_ref.npz", {k: v.shape for k, v in out.items() if k.startswith("bgr_")})


if __name__ == "__main__":
    main()
