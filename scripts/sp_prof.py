import os, sys
sys.path.insert(0, "/root/repo")
os.environ["BBME_GRAPHS"] = "0"
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair
f1, f2 = make_pair(1080, 1920, 2001, patches=12, max_patch_shift=40)
with bb.Estimator(1920, 1080, [80] * 3, [16] * 3) as est:
    est.estimate(f1, f2); est.estimate(f1, f2)
