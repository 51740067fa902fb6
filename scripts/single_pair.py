"""Two single-pair estimates of BASELINE config 2 (for an ncu launch list of the single-pair path)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair
f1, f2 = make_pair(1080, 1920, 2001, patches=12, max_patch_shift=40)
with bb.Estimator(1920, 1080, [80] * 3, [16] * 3, collect_stats=True) as est:
    est.estimate(f1, f2)
    t = time.perf_counter(); est.estimate(f1, f2); dt = time.perf_counter() - t
    print("wall ms", dt * 1e3, est.stats())
