"""BASELINE configs 2, 3 and 5 at FULL size against digests made from the reference itself.

tests/golden/big_digests.json (made by tests/golden/make_big_golden.py in the build container) holds sha256 digests of what
oracle/_ref -- the reference's unmodified motion_framework.cpp -- returns for seeded 1920x1080, 3840x2160 and 7680x4320
pairs (2 sweeps, the reference's hard-coded count; the oracle port agreed bit for bit), of the port's 5-sweep run for
config 5, and of every level's pyramid images and fields after the search / after the regularisation schedule.  The CUDA
path must reproduce every one of them (motion_framework.cpp:113-219 end to end: K64 keys, two-box TMA windows, band
splitting and the 5-sweep schedule at the sizes BASELINE.json names them at)."""
import hashlib
import json
import os

import numpy as np
import pytest

import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_digests.json")
DIGESTS = json.load(open(GOLD)) if os.path.exists(GOLD) else {}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _inputs(case):
    s = dict(case["synth"])
    s["shift"] = tuple(s["shift"])
    f1, f2 = make_pair(case["height"], case["width"], case["seed"], **s)
    assert [sha(f1), sha(f2)] == case["input_sha256"], "the synthetic generator no longer reproduces the golden inputs"
    return f1, f2


def _runs():
    out = []
    for name, case in sorted(DIGESTS.items()):
        out.append((name, "sweeps2", 2))
        if "sweepsN" in case:
            out.append((name, "sweepsN", case["sweepsN"]["sweeps"]))
    return out


def test_digest_file_covers_the_named_configs():
    assert {"c2_1080p", "c3_4k"} <= set(DIGESTS), sorted(DIGESTS)
    for case in DIGESTS.values():
        assert case["ref_equals_port_at_2_sweeps"] is True


@pytest.mark.gpu
@pytest.mark.parametrize("name,key,sweeps", _runs())
def test_full_size_field_equals_reference_digest(name, key, sweeps):
    case = DIGESTS[name]
    want = case[key]
    f1, f2 = _inputs(case)
    ss, bs = case["search_size"], case["block_size"]
    L = len(bs)
    with bb.Estimator(case["width"], case["height"], ss, bs, sweeps=sweeps, collect_stats=True, keep_search_mv=True) as est:
        assert [est.shape["padded_width"], est.shape["padded_height"]] == case["padded"]
        flow = est.estimate(f1, f2)
        st = est.stats()
        assert st["search_absdiffs"] == want["search_absdiffs"]
        for l in range(L):
            assert sha(est.level_image(l, 0)) == want["pyr1"][l], f"{name}: image 1 of level {l}"
            assert sha(est.level_image(l, 1)) == want["pyr2"][l], f"{name}: image 2 of level {l}"
        for l in range(L - 1, -1, -1):  # coarse to fine: the first level that differs is the one to look at
            assert sha(est.level_mv(l, which=1)) == want["level_after_search"][l], f"{name}: field after the search, level {l}"
            assert sha(est.level_mv(l, which=0)) == want["level_after_reg"][l], f"{name}: field after the schedule, level {l}"
    # the dense CV_32FC2 field (motion_framework.cpp:218): integer-valued, 2x2-constant, compact digest equal
    mv2 = flow[::2, ::2]
    assert np.array_equal(flow[1::2, 1::2], mv2) and np.array_equal(flow[0::2, 1::2], mv2) and np.array_equal(flow[1::2, 0::2], mv2)
    c16 = np.rint(mv2).astype(np.int16)
    assert np.array_equal(c16.astype(np.float32), mv2)
    assert sha(c16) == want["field_mv2"], f"{name}: final field"


def test_1080p_digest_reproducible_from_the_oracle(oracle):
    """CPU side of the pin: the committed 1080p digests are what the oracle port (== oracle/_ref) computes here."""
    case = DIGESTS["c2_1080p"]
    f1, f2 = _inputs(case)
    flow, st = oracle.estimate(f1, f2, case["search_size"], case["block_size"], 2)
    assert sha(np.rint(flow[::2, ::2]).astype(np.int16)) == case["sweeps2"]["field_mv2"]
    assert st["search_absdiffs"] == case["sweeps2"]["search_absdiffs"]
    ref = oracle.ref_estimate(f1, f2, case["search_size"], case["block_size"])
    if ref is not None:  # oracle/_ref present (build container, or shipped to the GPU box)
        assert np.array_equal(ref[0], flow)
