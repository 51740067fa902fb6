"""Runs a batch of one BASELINE geometry through the device path a few times (for ncu captures of its search launches).
usage: search_only.py c1|c3|c5 [pairs]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair
G = {"c1": (2336, 1552, [64] * 4, [32] * 4, 2, 16), "c3": (3840, 2160, [136] * 4, [8] * 4, 2, 8), "c5": (7680, 4320, [272] * 4, [16] * 4, 5, 1),
     "c2": (1920, 1080, [80] * 3, [16] * 3, 2, 128)}
w, h, ss, bs, sweeps, P = G[sys.argv[1]]
if len(sys.argv) > 2:
    P = int(sys.argv[2])
dev = torch.device("cuda", 0)
pairs = [make_pair(h, w, 7000 + i, shift=(7 - 2 * i, i - 5), patches=8, max_patch_shift=24) for i in range(min(P, 2))]
d1 = torch.stack([torch.from_numpy(pairs[i % len(pairs)][0]) for i in range(P)]).to(dev)
d2 = torch.stack([torch.from_numpy(pairs[i % len(pairs)][1]) for i in range(P)]).to(dev)
sh = bb.plan_shape(w, h, ss, bs)
out = torch.empty((P, sh["padded_height"] // 2, sh["padded_width"] // 2, 2), dtype=torch.int16, device=dev)
with bb.Estimator(w, h, ss, bs, sweeps=sweeps, chunk_pairs=P, collect_stats=True) as est:
    for _ in range(2):
        est.estimate_device_compact(P, d1.data_ptr(), d2.data_ptr(), w, w * h, out.data_ptr(), out[0].numel())
        est.sync()
    st = est.stats()
    pk, _ = est.measure_int_peak()
print(json.dumps({"geometry": sys.argv[1], "pairs": P, "ms_search": st["ms_search"], "search_launches": st["search_launches"],
                  "G_absdiff_per_s": st["search_absdiffs"] / (st["ms_search"] * 1e-3) / 1e9, "frac_of_int_peak": st["search_absdiffs"] / (st["ms_search"] * 1e-3) / pk}))
