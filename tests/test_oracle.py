"""CPU tests of the oracle (test infrastructure) against golden vectors: the container's cv2 for the OpenCV arithmetic,
the compiled reference (oracle/_ref) and its frozen outputs for whole fields, the gt-flow files for the .flo codec."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import make_pair, up4

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF_ROOT = "/root/reference"

# visit order of find_min_block_spiral for shift = 4 (transcribed from motion_framework.cpp:326-411)
SPIRAL_SHIFT4 = [(0, 0), (1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1), (2, -1), (2, 0), (2, 1),
                 (2, 2), (1, 2), (0, 2), (-1, 2), (-2, 2), (-2, 1), (-2, 0), (-2, -1), (-2, -2), (-1, -2), (0, -2),
                 (1, -2), (2, -2)]


def test_spiral_known_answer(oracle):
    assert [tuple(p) for p in oracle.spiral_walk(4).tolist()] == SPIRAL_SHIFT4


@pytest.mark.parametrize("shift", [1, 2, 3, 4, 9, 16, 32, 64, 65, 128, 256])
def test_spiral_covers_square_and_rank_is_closed_form(oracle, shift):
    walk = oracle.spiral_walk(shift)
    R = shift >> 1
    assert len(walk) == (2 * R + 1) ** 2
    assert len({tuple(p) for p in walk.tolist()}) == len(walk)
    assert np.abs(walk).max() == R
    ranks = [oracle.spiral_rank(int(dx), int(dy)) for dx, dy in walk]
    assert ranks == list(range(len(walk)))


def test_pyrdown_border_norm_match_cv2_golden(oracle):
    g = np.load(os.path.join(GOLD, "pyrdown_cv2.npz"))
    i = 0
    while f"pyr_in_{i}" in g:
        assert np.array_equal(oracle.pyrdown(g[f"pyr_in_{i}"]), g[f"pyr_out_{i}"]), i
        i += 1
    assert i >= 6
    assert np.array_equal(oracle.pad_image(g["border_in"], 5, 3), g["border_out"])
    l1 = np.abs(g["norm_a"].astype(np.int64) - g["norm_b"].astype(np.int64)).sum()
    assert l1 == int(g["norm_l1"][0])


def test_pyrdown_matches_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for h, w in [(40, 56), (41, 57), (128, 96), (6, 6)]:
        a = rng.integers(0, 256, (h, w)).astype(np.uint8)
        assert np.array_equal(oracle.pyrdown(a), cv2.pyrDown(a, dstsize=(w // 2, h // 2))), (h, w)


def _golden_cases():
    g = np.load(os.path.join(GOLD, "mf_reference.npz"))
    names = sorted({k.split("__")[0] for k in g.files})
    return g, names


def test_oracle_matches_frozen_reference_fields(oracle):
    """Fields computed by the reference's own motion_framework.cpp (frozen by tests/golden/make_golden.py)."""
    g, names = _golden_cases()
    assert len(names) >= 9
    for n in names:
        ss, bs = g[n + "__ss"].tolist(), g[n + "__bs"].tolist()
        flow, _ = oracle.estimate(g[n + "__f1"], g[n + "__f2"], ss, bs, 2)
        want = g[n + "__flow"].astype(np.float32)
        assert flow.shape == want.shape, n
        assert np.array_equal(flow, want), n
        rc, sh = oracle.plan_shape(g[n + "__f1"].shape[1], g[n + "__f1"].shape[0], bs)
        assert rc == 0
        assert [sh["padded_width"], sh["padded_height"], sh["padding_x"], sh["padding_y"]] == g[n + "__dims"].tolist(), n


def test_oracle_matches_live_reference(oracle):
    """oracle/_ref (the reference's sources compiled against oracle/cvshim) run on fresh seeds, when it is built."""
    if oracle.load_ref() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    for seed, (h, w, ss, bs, kind) in enumerate([(96, 128, [16, 16], [8, 8], "textured"), (90, 122, [20, 20, 20], [8, 8, 8], "noise"),
                                                 (128, 192, [40, 40], [16, 16], "textured"), (64, 64, [10, 10], [4, 4], "constant")]):
        f1, f2 = make_pair(h, w, 900 + seed, shift=(2, 3), max_patch_shift=5, kind=kind)
        flow, _ = oracle.estimate(f1, f2, ss, bs, 2)
        ref_flow, dims, _, _ = oracle.ref_estimate(f1, f2, ss, bs)
        assert np.array_equal(flow, ref_flow), (h, w, ss, bs, kind)


def test_shape_rules(oracle):
    # BASELINE configs
    assert oracle.plan_shape(2336, 1552, [32] * 4)[1]["padded_width"] == 2560
    rc, sh = oracle.plan_shape(2336, 1552, [32] * 4)
    assert (rc, sh["padded_height"], sh["padding_x"], sh["padding_y"]) == (0, 1792, 112, 120)
    rc, sh = oracle.plan_shape(1920, 1080, [16] * 3)
    assert (rc, sh["padded_width"], sh["padded_height"], sh["padding_x"], sh["padding_y"]) == (0, 1920, 1088, 0, 4)
    rc, sh = oracle.plan_shape(3840, 2160, [8] * 4)
    assert (rc, sh["padded_height"], sh["padding_y"]) == (0, 2176, 8)
    rc, sh = oracle.plan_shape(7680, 4320, [16] * 4)
    assert (rc, sh["padded_height"], sh["padding_y"]) == (0, 4352, 16)
    # the reference aborts when an axis reaches twice its size before a multiple is found (motion_framework.cpp:21-26)
    assert oracle.plan_shape(3, 64, [8])[0] == -2
    # odd padding difference and single-block axes are rejected (the reference reads out of bounds there)
    assert oracle.plan_shape(101, 96, [8])[0] == -3
    assert oracle.plan_shape(8, 96, [8])[0] == -4


def test_sign_convention_and_global_shift(oracle):
    # frame2(x, y) = frame1(x + 5, y - 3)  ->  mv = (-5, +3)  
    f1, f2 = make_pair(128, 160, 77, shift=(5, -3), patches=0, noise=0)
    flow, st = oracle.estimate(f1, f2, [24, 24], [8, 8], 2)
    inner = flow[32:-32, 32:-32]
    assert np.mean((inner[..., 0] == -5) & (inner[..., 1] == 3)) > 0.99
    assert st["search_absdiffs"] > 0 and st["reg_absdiffs"] > 0
    assert np.array_equal(flow[0::2, 0::2], flow[1::2, 1::2])  # final 2x2 fill (motion_framework.cpp:205-206)


def test_flo_codec_known_answers(oracle, tmp_path):
    digest = json.load(open(os.path.join(GOLD, "flo_gt_digest.json")))
    # the committed crop, written by the reference's own WriteFlowFile
    crop = oracle.flo_read(os.path.join(GOLD, "rubberwhale_crop.flo"))
    assert crop.shape == (64, 96, 2)
    unknown = (np.abs(crop[..., 0]) > 1e9) | (np.abs(crop[..., 1]) > 1e9)
    assert int(unknown.sum()) == digest["rubberwhale_crop"]["unknown_pixels"] > 0
    est = np.zeros_like(crop)
    est[..., 0] = 0.25
    assert oracle.aee(crop, est) == pytest.approx(digest["rubberwhale_crop"]["aee_of_quarter_pixel_field"], rel=0, abs=1e-12)
    out = tmp_path / "roundtrip.flo"
    assert oracle.flo_write(out, crop) == 0
    assert open(out, "rb").read() == open(os.path.join(GOLD, "rubberwhale_crop.flo"), "rb").read()
    # error paths of ReadFlowFile (rw_flow.cpp:58-133)
    bad = tmp_path / "bad.flo"
    raw = open(out, "rb").read()
    for payload in (raw[:-4], raw + b"\0", b"XXXX" + raw[4:]):
        bad.write_bytes(payload)
        with pytest.raises(ValueError):
            oracle.flo_read(bad)
    with pytest.raises(ValueError):
        oracle.flo_read(tmp_path / "name.txt")


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="reference tree not present (GPU box)")
def test_flo_gt_files_roundtrip_byte_identical(oracle, tmp_path):
    digest = json.load(open(os.path.join(GOLD, "flo_gt_digest.json")))
    for seq, d in digest.items():
        if "sha256" not in d:
            continue
        path = os.path.join(REF_ROOT, "middlebury", "gt-flow", seq, "flow10.flo")
        raw = open(path, "rb").read()
        assert hashlib.sha256(raw).hexdigest() == d["sha256"]
        gt = oracle.flo_read(path)
        assert gt.shape == (d["height"], d["width"], 2)
        assert oracle.aee(gt, np.zeros_like(gt)) == pytest.approx(d["aee_of_zero_field"], rel=0, abs=1e-12)
        out = tmp_path / (seq + ".flo")
        oracle.flo_write(out, gt)
        assert open(out, "rb").read() == raw


def test_config0_rubberwhale_standin_on_the_oracle(oracle):
    """main()'s pipeline (main_class.cpp:19-21,32-33,58-82) on the committed RubberWhale stand-in: pins the oracle's
    work counters (6.36 G search / 0.542 G regularisation abs-diffs) and the sanity AEE."""
    d = np.load(os.path.join(GOLD, "rubberwhale_standin.npz"))
    flow, st = oracle.estimate(up4(d["frame10"]), up4(d["frame11"]), [64] * 4, [32] * 4, 2)
    assert flow.shape == (1792, 2560, 2)
    assert (st["search_absdiffs"], st["reg_absdiffs"]) == (6362474496, 542256000)
    assert (st["search_sad_calls"], st["reg_sad_calls"]) == (6213354, 36426318)
    sub = np.ascontiguousarray(flow[120:1792 - 120:4, 112:2560 - 112:4] / 4.0)
    assert sub.shape == d["gt"].shape
    aee = oracle.aee(d["gt"], sub)
    assert aee == pytest.approx(0.13334927018152679, rel=0, abs=1e-9)


# ------------------------------------------------------------------------------------------ main()'s quarter-pel wrapper
def test_resize_linear_equals_cv2_golden(oracle):
    """cv::resize(INTER_LINEAR) of main_class.cpp:32-33: the oracle's restatement against outputs of the container's cv2
    (tests/golden/make_resize_golden.py), and against cv2 itself where it is importable."""
    d = np.load(os.path.join(GOLD, "resize_cv2.npz"))
    n = len([k for k in d.files if k.startswith("in_")])
    assert n >= 5
    for i in range(n):
        f = int(d[f"factor_{i}"][0])
        got = oracle.resize_linear(d[f"in_{i}"], f)
        assert np.array_equal(got, d[f"out_{i}"]), (i, f, int((got != d[f"out_{i}"]).sum()))
    try:
        import cv2
    except ImportError:
        return
    rng = np.random.default_rng(77)
    for (h, w), f in [((31, 45), 4), ((64, 50), 2), ((17, 23), 8), ((388, 584), 4)]:
        a = rng.integers(0, 256, (h, w)).astype(np.uint8)
        assert np.array_equal(oracle.resize_linear(a, f), cv2.resize(a, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR))


def test_strip_subsample_follows_main(oracle):
    """main_class.cpp:58-70 -- literal loop in numpy against the oracle and against the product's host helper."""
    import blockbasedmotionestimation_b200 as bb
    rng = np.random.default_rng(5)
    shape = bb.plan_shape(4 * 73, 4 * 41, [24, 24], [8, 8])
    ph, pw, px, py = shape["padded_height"], shape["padded_width"], shape["padding_x"], shape["padding_y"]
    flow = rng.integers(-40, 41, (ph, pw, 2)).astype(np.float32)
    want = np.zeros((41, 73, 2), np.float32)
    for i in range(py, ph - py, 4):
        for j in range(px, pw - px, 4):
            want[(i - py) // 4, (j - px) // 4] = flow[i, j] / np.float32(4)
    assert np.array_equal(oracle.strip_subsample(flow, px, py, 4), want)
    assert np.array_equal(bb.Flow().StripAndSubsample(flow, shape, 4), want)


def test_oracle_fuzz_against_live_reference(oracle):
    """Random geometries and contents: oracle port == the reference's own sources (oracle/_ref), bit for bit.  A short slice
    of scripts/fuzz_oracle_vs_reference.py (which ran 16 918 cases without a mismatch in round 1)."""
    if oracle.load_ref() is None:
        pytest.skip("oracle/_ref not built")
    import blockbasedmotionestimation_b200 as bb
    rng = np.random.default_rng(2718)
    done = 0
    while done < 120:
        L = int(rng.integers(1, 4))
        bs = [int(2 ** rng.integers(1, 6)) for _ in range(L)]
        ss = [b + int(rng.integers(0, 13)) for b in bs]
        w, h = int(rng.integers(16, 200)), int(rng.integers(16, 160))
        try:
            bb.plan_shape(w, h, ss, bs)
        except bb.BbmeError:
            continue
        kind = ["textured", "noise", "constant"][int(rng.integers(0, 3))]
        f1, f2 = make_pair(h, w, int(rng.integers(0, 1 << 30)), shift=(int(rng.integers(-4, 5)), int(rng.integers(-4, 5))),
                           max_patch_shift=4, kind=kind)
        want = oracle.ref_estimate(f1, f2, ss, bs)[0]
        got, _ = oracle.estimate(f1, f2, ss, bs, 2)
        assert np.array_equal(got, want), (w, h, ss, bs, kind)
        done += 1


def test_raster_search_and_compensation_equal_the_references_own_functions(oracle):
    """the oracle's find_min_block (:246-294) and draw_MVimage (:887-905) restatements against the reference's
    own (uncalled) member functions, run through oracle/_ref."""
    if oracle.load_ref() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    differs = 0
    for (h, w, bs, ss, kind) in [(64, 96, 8, 24, "textured"), (64, 64, 16, 32, "textured"), (48, 80, 4, 10, "noise"),
                                 (64, 96, 8, 24, "constant"), (96, 128, 32, 48, "textured"), (40, 56, 2, 6, "textured")]:
        f1, f2 = make_pair(h, w, int(rng.integers(1 << 20)), shift=(2, -1), kind=kind)
        pred = np.zeros((h, w, 2), np.float32)
        pred[::bs, ::bs] = rng.integers(-12, 13, (h // bs, w // bs, 2)).astype(np.float32)  # some predictions leave the image
        got = oracle.search_level_raster(f1, f2, bs, ss, pred)
        assert np.array_equal(got, oracle.ref_search_level_raster(f1, f2, bs, ss, pred)), (h, w, bs, ss, kind)
        assert np.array_equal(oracle.compensate(f2, bs, got), oracle.ref_compensate(f1, f2, bs, got))
        spiral, _ = oracle.search_level(f1, f2, bs, ss, pred)
        differs += int((spiral[::bs, ::bs] != got[::bs, ::bs]).any(-1).sum())
    assert differs > 0  # the two searches are different algorithms (tie-break, no centre test)
