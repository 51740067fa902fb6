"""Times one frame pair of each BASELINE.json configuration through the C ABI (device time by CUDA events, host wall
time of the synchronous call) and checks size-independent properties.  Output: JSON lines (profiles/r01_configs.jsonl)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair

CONFIGS = [
    ("C1 RubberWhale-sized x4 (2336x1552), block 32, search_size 64, 4 levels (repo defaults, main_class.cpp:19-21)", 2336, 1552, [64] * 4, [32] * 4, 2),
    ("C2 1920x1080, block 16, +-32, 3 levels", 1920, 1080, [80] * 3, [16] * 3, 2),
    ("C3 3840x2160, block 8, +-64, 4 levels", 3840, 2160, [136] * 4, [8] * 4, 2),
    ("C5 7680x4320, block 16, +-128, 4 levels, 5 sweeps", 7680, 4320, [272] * 4, [16] * 4, 5),
]
only = sys.argv[1:] or None
for name, w, h, ss, bs, sweeps in CONFIGS:
    if only and not any(name.startswith(o) for o in only):
        continue
    f1, f2 = make_pair(h, w, 99, shift=(7, -5), patches=8, max_patch_shift=24)
    with bb.Estimator(w, h, ss, bs, sweeps=sweeps, collect_stats=True) as est:
        est.estimate(f1, f2)  # warm-up
        reps = 5 if w <= 3840 else 2
        t0 = time.perf_counter()
        for _ in range(reps):
            flow = est.estimate(f1, f2)
        wall = (time.perf_counter() - t0) / reps
        st = est.stats()
        sh = est.shape
    py, px = sh["padding_y"], sh["padding_x"]
    inner = flow[py + 256:py + h - 256, px + 256:px + w - 256]
    frac = float(np.mean((inner[..., 0] == -7) & (inner[..., 1] == 5)))
    print(json.dumps({"config": name, "padded": [sh["padded_width"], sh["padded_height"]], "wall_ms_per_pair_host_buffers": 1e3 * wall,
                      "device_ms": {k: round(st[k], 3) for k in ("ms_total", "ms_pyramid", "ms_search", "ms_regularize", "ms_other")},
                      "search_absdiffs": st["search_absdiffs"], "search_G_absdiff_per_s": st["search_absdiffs"] / (st["ms_search"] * 1e-3) / 1e9,
                      "kernel_launches": st["kernel_launches"], "fix_rounds": st["fix_rounds"],
                      "interior_fraction_with_true_shift": frac, "field_is_2x2_constant": bool(np.array_equal(flow[0::2, 0::2], flow[1::2, 1::2]))}), flush=True)
