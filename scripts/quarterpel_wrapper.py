"""main()'s whole flow (main_class.cpp:32-70) on the GPU for BASELINE config 0: RubberWhale stand-in frames at their
native 584x388 -> cv::resize x4 on the device -> MF (search 64 / block 32 / 4 levels) -> strip + every 4th pixel + / 4
-> 584x388x2 sub-pixel field.  Times a batch through the host-buffer C ABI call (copies inside the timed region), next to
(a) the same pairs through the padded-dense path with frames up-sampled on the host beforehand and (b) the CPU chain
(oracle resize + reference MF + strip) on one pair.  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import blockbasedmotionestimation_b200 as bb
from oracle import binding as ob

d = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_standin.npz"))
gt, f10, f11 = d["gt"], d["frame10"], d["frame11"]
ss, bs, F = [64] * 4, [32] * 4, 4
N, CH = 64, 16
h, w = f10.shape
fl = bb.Flow()
with bb.Estimator(F * w, F * h, ss, bs, chunk_pairs=CH, slots=2, collect_stats=True) as est:
    est.estimate_upsampled([f10] * CH, [f11] * CH, F)  # warm-up
    t0 = time.perf_counter()
    out = est.estimate_upsampled([f10] * N, [f11] * N, F)
    t_wrap = time.perf_counter() - t0
    st = est.stats()
    shape = est.shape
    # padded-dense path on frames up-sampled beforehand (what a caller of MF alone pays: 16x the input, 40x the output bytes)
    u10, u11 = ob.resize_linear(f10, F), ob.resize_linear(f11, F)
    est.estimate_batch([u10] * CH, [u11] * CH)
    t0 = time.perf_counter()
    dense = est.estimate_batch([u10] * N, [u11] * N)
    t_dense = time.perf_counter() - t0
px, py = shape["padding_x"], shape["padding_y"]
t0 = time.perf_counter()
c10, c11 = ob.resize_linear(f10, F), ob.resize_linear(f11, F)
t_resize = time.perf_counter() - t0
ref = ob.ref_estimate(c10, c11, ss, bs)
if ref is not None:
    ref_flow, _, t_ctor, t_run = ref
    kind = "reference (oracle/_ref)"
else:
    ref_flow, ost = ob.estimate(c10, c11, ss, bs, 2)
    t_ctor, t_run, kind = ost["t_ctor_s"], ost["t_run_s"], "oracle port"
want = ob.strip_subsample(ref_flow, px, py, F)
print(json.dumps({
    "config": "BASELINE config 0 through main()'s wrapper on the device: 584x388 frames, cv::resize x4, search 64 / block 32 / 4 levels, strip + /4",
    "pairs": N, "chunk_pairs": CH, "slots": 2,
    "wrapper_pairs_per_s_host_buffers": N / t_wrap, "wrapper_ms_per_pair": 1e3 * t_wrap / N,
    "wrapper_bytes_per_pair": {"h2d": 2 * w * h, "d2h": 8 * w * h},
    "device_ms_per_pair": {k: round(st[k] / N, 4) for k in ("ms_total", "ms_pyramid", "ms_search", "ms_regularize", "ms_other")},
    "dense_path_pairs_per_s_host_buffers": N / t_dense,
    "dense_path_bytes_per_pair": {"h2d": 2 * F * F * w * h, "d2h": 8 * shape["padded_width"] * shape["padded_height"]},
    "wrapper_equals_cpu_chain": bool(all(np.array_equal(o, want) for o in out)),
    "dense_equals_cpu_field": bool(np.array_equal(dense[0], ref_flow)),
    "aee_px_vs_gt_flow": fl.CalculateMSE(gt, out[0]),
    "cpu": kind, "cpu_resize_s": t_resize, "cpu_constructor_s": t_ctor, "cpu_calcMotionBlockMatching_s": t_run,
    "note": "stand-in frames (texture warped by the real gt flow); up-sampling here is cv::resize arithmetic (bit-exact vs cv2 4.13)"}))
