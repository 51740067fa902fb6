"""One tiny TMA-kernel search against the oracle (debug helper for compute-sanitizer runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair
from oracle import binding as ob

bs, ss = int(sys.argv[1]), int(sys.argv[2])
h, w = 96, 160
f1, f2 = make_pair(h, w, 1, shift=(3, -2), max_patch_shift=5)
est = bb.Estimator(64, 64, [12], [4], search_kernel=1)
got, st = est.stage_search(f1, f2, bs, ss, None, kernel=2)
want = ob.search_level(f1, f2, bs, ss, np.zeros((h, w, 2), np.float32))[0][::bs, ::bs].astype(np.int16)
print("equal:", np.array_equal(got, want), st)
