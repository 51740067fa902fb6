"""Fuzz: the oracle port (oracle/bbme_oracle.c) against the reference's own sources (oracle/_ref) on random sizes, block
sizes, search sizes, level counts and contents, for a fixed wall time.  CPU only.  Round 1: 16 918 cases, 0 mismatches."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as ob
from blockbasedmotionestimation_b200.synth import make_pair
import blockbasedmotionestimation_b200 as bb
rng = np.random.default_rng(12345)
t0 = time.time(); n = 0; skipped = 0; bad = 0
while time.time() - t0 < 240:
    L = int(rng.integers(1, 4))
    bs = [int(2 ** rng.integers(1, 6)) for _ in range(L)]
    ss = [b + int(rng.integers(0, 13)) for b in bs]
    w, h = int(rng.integers(16, 200)), int(rng.integers(16, 160))
    try:
        sh = bb.plan_shape(w, h, ss, bs)
    except bb.BbmeError:
        skipped += 1; continue
    kind = ["textured", "noise", "constant"][int(rng.integers(0, 3))]
    f1, f2 = make_pair(h, w, int(rng.integers(0, 1 << 30)), shift=(int(rng.integers(-4, 5)), int(rng.integers(-4, 5))), max_patch_shift=4, kind=kind)
    want = ob.ref_estimate(f1, f2, ss, bs)[0]
    got, _ = ob.estimate(f1, f2, ss, bs, 2)
    n += 1
    if not np.array_equal(got, want):
        bad += 1; print("MISMATCH", w, h, ss, bs, kind, int((got != want).any(-1).sum()))
print("cases", n, "skipped", skipped, "mismatches", bad)
