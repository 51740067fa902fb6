"""Single-pair latency of BASELINE config 2 (1920x1080, 16x16, +-32, 3 levels) through the host-buffer C ABI call:
wall time per call (H2D + kernels + D2H, numpy buffers) with and without CUDA-graph replay, and the device time."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair
f1, f2 = make_pair(1080, 1920, 2001, patches=12, max_patch_shift=40)
res = {}
for graphs in (0, 1):
    os.environ["BBME_GRAPHS"] = str(graphs)
    with bb.Estimator(1920, 1080, [80] * 3, [16] * 3) as est:
        out = np.empty(est.flow_shape(), np.float32)
        for _ in range(3):
            est.estimate(f1, f2, out)
        t = time.perf_counter()
        for _ in range(20):
            est.estimate(f1, f2, out)
        res[f"wall_ms_graphs_{graphs}"] = (time.perf_counter() - t) / 20 * 1e3
        res[f"checksum_{graphs}"] = float(out.sum())
with bb.Estimator(1920, 1080, [80] * 3, [16] * 3, collect_stats=True) as est:
    est.estimate(f1, f2)
    est.estimate(f1, f2)
    res["device"] = est.stats()
print(json.dumps(res))
