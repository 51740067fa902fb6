"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol include/bbme.h declares,
the shape planner equals the oracle's, the .flo codec / AEE equal the golden answers, and without a GPU the library
fails loudly instead of falling back.  No compute kernels are launched here."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bbme.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bbme_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (bbme_[a-z0-9_]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(_lib.SIGNATURES), sorted(set(declared) ^ set(_lib.SIGNATURES))
    assert lib.bbme_version() == 201


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_plan_shape_equals_oracle(oracle):
    rng = np.random.default_rng(3)
    checked = 0
    for _ in range(300):
        w, h = int(rng.integers(4, 700)), int(rng.integers(4, 700))
        L = int(rng.integers(1, 5))
        bs = [int(2 ** rng.integers(1, 6)) for _ in range(L)]
        ss = [b + 8 for b in bs]
        rc_o, sh_o = oracle.plan_shape(w, h, bs)
        try:
            sh = bb.plan_shape(w, h, ss, bs)
            rc = 0
        except bb.BbmeError as e:
            rc = e.status
        assert rc == rc_o, (w, h, bs, rc, rc_o)
        if rc == 0:
            checked += 1
            for k in ("padded_width", "padded_height", "padding_x", "padding_y", "level_width", "level_height"):
                assert sh[k] == sh_o[k], (w, h, bs, k)
    assert checked > 20


def test_baseline_config_shapes():
    sh = bb.plan_shape(2336, 1552, [64] * 4, [32] * 4)  # RubberWhale x4, repo defaults (main_class.cpp:19-21,32-33)
    assert (sh["padded_width"], sh["padded_height"], sh["padding_x"], sh["padding_y"]) == (2560, 1792, 112, 120)
    assert sh["level_width"] == [2560, 1280, 640, 320]
    sh = bb.plan_shape(1920, 1080, [80] * 3, [16] * 3)
    assert (sh["padded_width"], sh["padded_height"], sh["padding_y"]) == (1920, 1088, 4)


def test_invalid_arguments_have_status_codes():
    for args, status in [((100, 100, [12], [6]), -1), ((3, 64, [16], [8]), -2), ((101, 96, [16], [8]), -3),
                         ((8, 96, [16], [8]), -4), ((20000, 64, [16], [8]), -10)]:
        with pytest.raises(bb.BbmeError) as e:
            bb.plan_shape(*args)
        assert e.value.status == status, (args, e.value.status)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bb.BbmeError) as e:
        bb.Estimator(64, 64, [16], [8])
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(bb.BbmeError):
        bb.MF(np.zeros((64, 64), np.uint8), np.zeros((64, 64), np.uint8), [16], [8], 1)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may import, include, link or call it."""
    pat = re.compile(r"(from|import)\s+oracle|oracle/|bbme_oracle|\borc_[a-z]|libbbme_ref|cvshim")
    for top in (os.path.join(ROOT, "blockbasedmotionestimation_b200"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not pat.search(text), os.path.join(dirpath, f)
    mk = open(os.path.join(ROOT, "Makefile")).read()
    link_line = [l for l in mk.splitlines() if "-shared" in l and "$(OBJS)" in l]
    assert link_line and "oracle" not in link_line[0]


def test_flow_class_codec_and_metric(tmp_path):
    digest = json.load(open(os.path.join(GOLD, "flo_gt_digest.json")))
    fl = bb.Flow()
    crop = fl.ReadFlowFile(os.path.join(GOLD, "rubberwhale_crop.flo"))
    assert crop.shape == (64, 96, 2) and crop.dtype == np.float32
    est = np.zeros_like(crop)
    est[..., 0] = 0.25
    assert fl.CalculateMSE(crop, est) == pytest.approx(digest["rubberwhale_crop"]["aee_of_quarter_pixel_field"], rel=0, abs=1e-12)
    out = tmp_path / "w.flo"
    fl.WriteFlowFile(crop, out)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "rubberwhale_crop.flo"), "rb").read()
    raw = open(out, "rb").read()
    for payload, name in ((raw[:-4], "short.flo"), (raw + b"\0", "long.flo"), (b"XXXX" + raw[4:], "tag.flo"), (raw, "ext.txt")):
        p = tmp_path / name
        p.write_bytes(payload)
        with pytest.raises(bb.BbmeError):
            fl.ReadFlowFile(p)
    with pytest.raises(bb.BbmeError):
        fl.WriteFlowFile(crop, tmp_path / "noext")
    with pytest.raises(bb.BbmeError):
        fl.ReadFlowFile(None)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present (GPU box)")
def test_flow_class_roundtrips_gt_files(tmp_path):
    digest = json.load(open(os.path.join(GOLD, "flo_gt_digest.json")))
    fl = bb.Flow()
    for seq, d in digest.items():
        if "sha256" not in d:
            continue
        path = f"/root/reference/middlebury/gt-flow/{seq}/flow10.flo"
        gt = fl.ReadFlowFile(path)
        assert gt.shape == (d["height"], d["width"], 2)
        assert fl.CalculateMSE(gt, np.zeros_like(gt)) == pytest.approx(d["aee_of_zero_field"], rel=0, abs=1e-12)
        out = tmp_path / (seq + ".flo")
        fl.WriteFlowFile(gt, out)
        assert open(out, "rb").read() == open(path, "rb").read()


def test_strip_and_subsample_like_main():
    # main_class.cpp:58-70 on a synthetic padded field
    sh = bb.plan_shape(2336, 1552, [64] * 4, [32] * 4)
    rng = np.random.default_rng(1)
    flow = rng.integers(-40, 40, (sh["padded_height"], sh["padded_width"], 2)).astype(np.float32)
    out = bb.Flow().StripAndSubsample(flow, sh, 4)
    assert out.shape == (388, 584, 2)
    px, py = sh["padding_x"], sh["padding_y"]
    want = flow[py:sh["padded_height"] - py:4, px:sh["padded_width"] - px:4] / 4.0
    assert np.array_equal(out, want)


def test_motion_to_color_equals_the_reference():
    """Flow::MotionToColor (rw_flow.cpp:202-300): bytes identical to what the reference's own function produced
    (tests/golden/flow_color_ref.npz, made by make_color_golden.py from oracle/_ref) and, when oracle/_ref is present,
    to the reference's function run now on fresh random fields."""
    from oracle import binding as ob
    d = np.load(os.path.join(GOLD, "flow_color_ref.npz"))
    fl = bb.Flow()
    n = len([k for k in d.files if k.startswith("flow_")])
    assert n >= 4
    for i in range(n):
        got = fl.MotionToColor(d[f"flow_{i}"], float(d[f"maxmotion_{i}"][0]))
        assert got.dtype == np.uint8 and got.shape == d[f"bgr_{i}"].shape
        assert np.array_equal(got, d[f"bgr_{i}"]), (i, int((got != d[f"bgr_{i}"]).sum()))
    assert np.array_equal(fl.MotionToColor(np.zeros((4, 5, 2), np.float32)), np.full((4, 5, 3), 255, np.uint8))  # zero flow is white
    unknown = np.full((2, 2, 2), 1e10, np.float32)
    assert not fl.MotionToColor(unknown).any()  # unknown flow is black
    if ob.load_ref() is not None:
        rng = np.random.default_rng(8)
        for scale, mm in ((0.3, -1.0), (40.0, -1.0), (7.0, 2.5)):
            f = (rng.standard_normal((33, 47, 2)) * scale).astype(np.float32)
            assert np.array_equal(fl.MotionToColor(f, mm), ob.ref_flow_color(f, mm))


def test_reference_main_compiles_against_dropin_headers(tmp_path):
    """The reference's own main_class.cpp, unmodified, against include/ (MF, Flow incl. MotionToColor, the public ints):
    syntax check only -- OpenCV itself is not installed, so cv::Mat comes from the test oracle's shim headers and
    imread / resize / imwrite are declared by tests/cpp/opencv_io_stubs.hpp.  The reference's own class headers must not
    sit next to main_class.cpp (quote includes look there first), hence the copy."""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "main_class.cpp")):
        pytest.skip("reference tree not present on this box")
    import shutil
    for f in ("main_class.cpp", "standard_headers.h", "opencv_headers.h"):
        shutil.copy(os.path.join(ref, f), tmp_path / f)
    res = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-DBBME_USE_OPENCV", "-I" + os.path.join(ROOT, "include"),
                          "-I" + os.path.join(ROOT, "oracle", "cvshim"), "-include",
                          os.path.join(ROOT, "tests", "cpp", "opencv_io_stubs.hpp"), "main_class.cpp"],
                         cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the reference's CPU path on this box's cores): exactly one line on stdout, JSON, with the
    keys the driver reads.  Also covers the stdout hygiene (libraries must not get to print there)."""
    res = subprocess.run([os.sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p frame-pairs/sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_expand_compact_matches_numpy():
    """The host-side expansion behind the host-buffer entry points (compact int16 D2H + worker threads): every 2x2 pixel
    block of the dense CV_32FC2 field carries its entry's vector (motion_framework.cpp:205-206)."""
    lib = _lib.load()
    rng = np.random.default_rng(7)
    for gh2, gw2, align in [(3, 5, 0), (64, 96, 0), (77, 130, 0), (544, 960, 0), (33, 36, 4)]:
        mv = rng.integers(-2000, 2000, (gh2, gw2, 2)).astype(np.int16)
        buf = np.full(2 * gh2 * 2 * gw2 * 2 + 16, np.nan, np.float32)
        out = buf[align:align + 2 * gh2 * 2 * gw2 * 2].reshape(2 * gh2, 2 * gw2, 2)  # align = 4 floats: not 32-byte aligned
        assert lib.bbme_expand_compact(mv.ctypes.data, gw2, gh2, out.ctypes.data) == 0
        want = np.repeat(np.repeat(mv, 2, axis=0), 2, axis=1).astype(np.float32)
        assert np.array_equal(out, want)
        assert np.isnan(buf[:align]).all() and np.isnan(buf[align + out.size:]).all()


def test_pool_and_mf_cache_fail_loudly_without_a_gpu():
    """No device, no fallback: the multi-GPU pool and the MF context cache report the CUDA error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    p = C.c_void_p()
    assert lib.bbme_pool_create(C.byref(p), 0, None) == -6 and not p.value
    assert b"no CPU fallback" in lib.bbme_last_error(None)
    ctx = C.c_void_p()
    rc = lib.bbme_mf_open(C.byref(ctx), 0, 64, 64, 1, (C.c_int * 1)(16), (C.c_int * 1)(8), 2, None)
    assert rc == -6 and not ctx.value
    lib.bbme_mf_cache_clear()


def test_color_flow_cli(tmp_path):
    """tools/color_flow == the vendored Middlebury tool's command line (middlebury/flow-code/color_flow.cpp:68-98): same console
    line, a PNG whose pixels are Flow::MotionToColor's (itself pinned to the reference's function by flow_color_ref.npz)."""
    cv2 = pytest.importorskip("cv2")
    exe = os.path.join(ROOT, "tools", "color_flow")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ROOT, "tools/color_flow"], stdout=subprocess.DEVNULL)
    fl = bb.Flow()
    rng = np.random.default_rng(5)
    f = (rng.standard_normal((37, 53, 2)) * 6).astype(np.float32)
    f[3:6, 4:9] = 1e10  # unknown flow -> black
    src = tmp_path / "in.flo"
    fl.WriteFlowFile(f, src)
    for extra, maxmotion in ([], -1.0), (["7.5"], 7.5):
        out = tmp_path / "out.png"
        res = subprocess.run([exe, "-quiet", str(src), str(out)] + extra, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        assert res.stdout.startswith("max motion: ") and "motion range: u = " in res.stdout and res.stderr == ""
        assert np.array_equal(cv2.imread(str(out), cv2.IMREAD_COLOR), fl.MotionToColor(f, maxmotion))
    res = subprocess.run([exe, str(src), str(tmp_path / "out.ppm")], capture_output=True, text=True)
    assert res.returncode == 0 and "normalizing by" in res.stderr
    assert subprocess.run([exe, str(src)], capture_output=True).returncode != 0  # usage error
