"""BASELINE config 0 sanity line: RubberWhale stand-in through main()'s pipeline on the GPU and on the CPU reference
(oracle/_ref when built, else the oracle port).  Prints one JSON object (profiles/r01_middlebury_standin.json)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import blockbasedmotionestimation_b200 as bb
from helpers import up4
from oracle import binding as ob

d = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_standin.npz"))
gt = d["gt"]
im1, im2 = up4(d["frame10"]), up4(d["frame11"])
ss, bs = [64] * 4, [32] * 4
mf = bb.MF(im1, im2, ss, bs, 4, collect_stats=True)
mf.calcMotionBlockMatching()
t0 = time.perf_counter()
flow = mf.calcMotionBlockMatching()
wall = time.perf_counter() - t0
st, shape = mf.stats(), mf._est.shape
mf.close()
fl = bb.Flow()
sub = fl.StripAndSubsample(flow, shape, 4)
aee = fl.CalculateMSE(gt, sub)
ref = ob.ref_estimate(im1, im2, ss, bs)
if ref is not None:
    ref_flow, _, t_ctor, t_run = ref
    kind = "reference (oracle/_ref)"
else:
    ref_flow, ost = ob.estimate(im1, im2, ss, bs, 2)
    t_ctor, t_run, kind = ost["t_ctor_s"], ost["t_run_s"], "oracle port"
print(json.dumps({"config": "BASELINE config 0: RubberWhale stand-in 584x388, x4 bilinear, search 64 / block 32 / 4 levels, strip + /4",
                  "aee_px_vs_gt_flow": aee, "gpu_field_equals_cpu_field": bool(np.array_equal(flow, ref_flow)),
                  "gpu_wall_ms_host_buffers": 1e3 * wall, "gpu_device_ms": round(st["ms_total"], 3),
                  "cpu": kind, "cpu_constructor_s": t_ctor, "cpu_calcMotionBlockMatching_s": t_run,
                  "note": "stand-in frames (texture warped by the real gt flow); reference error.txt logs 0.21-0.43 px on the real Dimetrodon frames"}))
