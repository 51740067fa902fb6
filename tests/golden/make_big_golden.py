"""Full-size golden digests for BASELINE configs 2, 3 and 5, generated from the reference itself.

    python tests/golden/make_big_golden.py c2_1080p c3_4k c5_8k      (run in THIS container: needs oracle/_ref)

For every case the seeded synthetic pair (blockbasedmotionestimation_b200.synth.make_pair, numpy only) is run through
  * oracle/_ref  -- the reference's unmodified motion_framework.cpp compiled against oracle/cvshim (2 sweeps, the
                    reference's hard-coded value, motion_framework.cpp:143,184), and
  * the oracle port (oracle/bbme_oracle.c) with the same 2 sweeps; the two dense fields must be identical.
The port then also yields the per-level state (pyramid images, field after calcLevelBM, field after the regularisation
schedule) and, for config 5, the 5-sweep run the reference cannot do without editing its source (the port's sweep loop
is the reference's loop with the literal 2 replaced by a parameter, pinned at 2 sweeps by the check above).

What is committed (tests/golden/big_digests.json) are sha256 digests: of the input frames (so that a changed generator
is noticed), of the final 2x2-granular int16 field (== the dense CV_32FC2 field, which replicates it 2x2,
motion_framework.cpp:205-206), and per level of the block-granular field after the search and the 2x2-granular field
after the schedule.  The -m gpu tests recompute the same digests from the CUDA path.  CPU cost: c2 seconds, c3 ~10 min,
c5 ~1 h (three runs side by side on three cores).
"""
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blockbasedmotionestimation_b200.synth import make_pair  # noqa: E402
from oracle import binding as ob  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "big_digests.json")

CASES = {
    # BASELINE config 2: 1920x1080, 16x16, +-32, 3 levels
    "c2_1080p": dict(w=1920, h=1080, ss=[80] * 3, bs=[16] * 3, sweeps=2, seed=2001,
                     synth=dict(shift=(5, -3), patches=12, max_patch_shift=40)),
    # BASELINE config 3: 3840x2160, 8x8, +-64, 4 levels
    "c3_4k": dict(w=3840, h=2160, ss=[136] * 4, bs=[8] * 4, sweeps=2, seed=3001,
                  synth=dict(shift=(9, -7), patches=12, max_patch_shift=100)),
    # BASELINE config 5: 7680x4320, +-128, 5 sweeps (16x16 blocks, 4 levels: SURVEY 8, C5)
    "c5_8k": dict(w=7680, h=4320, ss=[272] * 4, bs=[16] * 4, sweeps=5, seed=5001,
                  synth=dict(shift=(-11, 6), patches=12, max_patch_shift=60)),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compact(flow, g=2):
    """Dense float field -> int16 field sampled at every g-th pixel (the block corners)."""
    c = flow[::g, ::g, :]
    r = np.rint(c).astype(np.int16)
    assert np.array_equal(r.astype(np.float32), c), "field is not integer-valued"
    return np.ascontiguousarray(r)


def digests_of(flow, dbg, bs):
    L = len(bs)
    assert np.array_equal(flow[0::2, 0::2], flow[1::2, 1::2]) and np.array_equal(flow[0::2, 0::2], flow[0::2, 1::2])
    return {
        "field_mv2": sha(compact(flow, 2)),
        "level_after_search": [sha(compact(dbg["after_search"][l], bs[l])) for l in range(L)],
        "level_after_reg": [sha(compact(dbg["after_reg"][l], 2)) for l in range(L)],
        "pyr1": [sha(dbg["pyr1"][l]) for l in range(L)],
        "pyr2": [sha(dbg["pyr2"][l]) for l in range(L)],
    }


def run_case(name):
    c = CASES[name]
    f1, f2 = make_pair(c["h"], c["w"], c["seed"], **c["synth"])
    res = {}

    def job(key, fn):
        t0 = time.perf_counter()
        res[key] = fn()
        res[key + "_s"] = time.perf_counter() - t0
        print(f"[{name}] {key} done in {res[key + '_s']:.0f} s", flush=True)

    jobs = [("ref2", lambda: ob.ref_estimate(f1, f2, c["ss"], c["bs"])),
            ("port2", lambda: ob.estimate(f1, f2, c["ss"], c["bs"], 2, debug=True))]
    if c["sweeps"] != 2:
        jobs.append(("portN", lambda: ob.estimate(f1, f2, c["ss"], c["bs"], c["sweeps"], debug=True)))
    th = [threading.Thread(target=job, args=j) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if res.get("ref2") is None:
        raise SystemExit("oracle/_ref is not built: run `make -C oracle` in a container that has /root/reference")
    ref_flow = res["ref2"][0]
    port_flow, port_stats, port_dbg = res["port2"]
    if not np.array_equal(ref_flow, port_flow):
        raise SystemExit(f"{name}: the oracle port differs from oracle/_ref at 2 sweeps -- golden NOT written")
    sh = port_dbg["shape"]
    entry = {
        "width": c["w"], "height": c["h"], "search_size": c["ss"], "block_size": c["bs"], "sweeps": c["sweeps"],
        "seed": c["seed"], "synth": {k: list(v) if isinstance(v, tuple) else v for k, v in c["synth"].items()},
        "input_sha256": [sha(f1), sha(f2)],
        "padded": [sh["padded_width"], sh["padded_height"]], "padding": [sh["padding_x"], sh["padding_y"]],
        "ref_equals_port_at_2_sweeps": True,
        "ref_seconds": {"ctor": res["ref2"][2], "calcMotionBlockMatching": res["ref2"][3]},
        "port_seconds_2_sweeps": res["port2_s"],
        "sweeps2": dict(digests_of(port_flow, port_dbg, c["bs"]),
                        search_absdiffs=int(port_stats["search_absdiffs"]), reg_absdiffs=int(port_stats["reg_absdiffs"]),
                        source="oracle/_ref (reference sources) == oracle port"),
    }
    if c["sweeps"] != 2:
        fl, st, dbg = res["portN"]
        entry["sweepsN"] = dict(digests_of(fl, dbg, c["bs"]), sweeps=c["sweeps"],
                                search_absdiffs=int(st["search_absdiffs"]), reg_absdiffs=int(st["reg_absdiffs"]),
                                source="oracle port (the reference hard-codes 2 sweeps)")
        entry["port_seconds_N_sweeps"] = res["portN_s"]
        # the search does not depend on the sweep count at the coarsest level
        assert entry["sweepsN"]["level_after_search"][-1] == entry["sweeps2"]["level_after_search"][-1]
    return entry


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        entry = run_case(name)
        # merge under a lock-free read-modify-write (cases may be generated by separate invocations)
        tmp = OUT + f".{name}.part"
        json.dump(entry, open(tmp, "w"), indent=1)
        data = {}
        if os.path.exists(OUT):
            data = json.load(open(OUT))
        data[name] = entry
        json.dump(data, open(OUT + ".tmp", "w"), indent=1, sort_keys=True)
        os.replace(OUT + ".tmp", OUT)
        os.unlink(tmp)
        print(f"[{name}] written to {OUT}", flush=True)


if __name__ == "__main__":
    main()
