"""Device time of one estimate for several chunk sizes under the current BBME_GRID_ROUNDS (env): tuning helper."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair
pairs = [make_pair(1080, 1920, 2001 + i, patches=12, max_patch_shift=40) for i in range(4)]
res = {"gr": os.environ.get("BBME_GRID_ROUNDS", "default")}
for n in (1, 8, 32):
    with bb.Estimator(1920, 1080, [80] * 3, [16] * 3, chunk_pairs=n, collect_stats=True) as est:
        a = [pairs[i % 4][0] for i in range(n)]; b = [pairs[i % 4][1] for i in range(n)]
        est.estimate_batch(a, b)
        est.estimate_batch(a, b)
        st = est.stats()
        res[f"n{n}_reg_ms_per_pair"] = round(st["ms_regularize"] / n, 4)
        res[f"n{n}_total_ms_per_pair"] = round(st["ms_total"] / n, 4)
print(json.dumps(res))
