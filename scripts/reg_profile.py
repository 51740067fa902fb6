"""Regularisation time of a 1080p chunk (device-resident) for several cluster sizes, plus the in-kernel phase profile of pair 0
(BBME_REG_PROFILE=1 -> stderr).  usage: reg_profile.py <pairs> [cluster sizes...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair, seed_for

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
css = [int(v) for v in sys.argv[2:]] or [0]
W, H, SS, BS = 1920, 1080, [80] * 3, [16] * 3
distinct = [make_pair(H, W, seed_for(4, i), shift=(5 - (i % 11), (i % 7) - 3), patches=12, max_patch_shift=40) for i in range(min(P, 16))]
dev = torch.device("cuda", 0)
d1 = torch.empty((P, H, W), dtype=torch.uint8, device=dev)
d2 = torch.empty((P, H, W), dtype=torch.uint8, device=dev)
for i in range(P):
    d1[i].copy_(torch.from_numpy(distinct[i % len(distinct)][0]))
    d2[i].copy_(torch.from_numpy(distinct[i % len(distinct)][1]))
sh = bb.plan_shape(W, H, SS, BS)
Hp, Wp = sh["padded_height"], sh["padded_width"]
out = torch.empty((P, Hp, Wp, 2), dtype=torch.float32, device=dev)
for cs in css:
    if cs:
        os.environ["BBME_REG_CLUSTER"] = str(cs)
    else:
        os.environ.pop("BBME_REG_CLUSTER", None)
    with bb.Estimator(W, H, SS, BS, chunk_pairs=P, collect_stats=True) as est:
        for _ in range(3):
            est.estimate_device(P, d1.data_ptr(), d2.data_ptr(), W, W * H, out.data_ptr(), Hp * Wp * 2)
            est.sync()
        st = est.stats()
    print(json.dumps({"pairs": P, "cluster": cs, "ms_regularize": st["ms_regularize"], "ms_search": st["ms_search"], "ms_total": st["ms_total"],
                      "fix_rounds": st["fix_rounds"], "fix_blocks": st["fix_blocks"], "checksum": float(out.sum())}), flush=True)
