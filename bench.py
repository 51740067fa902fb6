#!/usr/bin/env python
"""bench.py -- 1080p frame-pairs/s of the block-matching hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one batch of `--pairs` synthetic 1080p frame pairs per GPU through the whole path
(pad -> pyramid -> per-level SAD search -> regularisation with block splitting -> MV propagation -> dense field);
BASELINE.json config 4 (batch of 1080p pairs, 16x16 blocks, +-32 search, 3 levels) sharded over the ranks with no
data-path collective.  `value` is measured with the inputs resident in HBM (C ABI bbme_estimate_device), `e2e`
through the host-buffer C ABI call (bbme_estimate_batch_async: pinned H2D + pipeline + D2H + host-side expansion into
the caller's dense CV_32FC2 buffers inside the timed region).  `other_configs` times the other BASELINE configurations
(single 1080p pair, 4K batch, 8K +-128 / 5 sweeps, RubberWhale x4 geometry) outside the headline.
`--impl reference` times the reference's own CPU implementation (oracle/_ref: the reference's sources compiled against
oracle/cvshim; falls back to the oracle port) on all host cores of rank 0.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
SEARCH_SIZE = [80, 80, 80]  # block 16 + 2 * 32
BLOCK_SIZE = [16, 16, 16]
SWEEPS = 2
METRIC = "1080p frame-pairs/sec"
UNIT = "pairs/s"
WORKLOAD = ("BASELINE config 4: batch of synthetic 1920x1080 8-bit luma frame pairs (config 2 geometry: 16x16 blocks, "
            "+-32 search = search_size 80, 3-level pyramid, 2 regularisation sweeps per block size), sharded by pair")
CONFIG = {"workload": WORKLOAD}  # identical in both arms (the driver compares it); run details go to "run"

_REAL_STDOUT = None


def claim_stdout():
    """Only the JSON line may reach stdout: libraries (NCCL prints its version there) get stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_pairs(n_distinct, rank):
    from blockbasedmotionestimation_b200.synth import make_pair, seed_for
    out = []
    for i in range(n_distinct):
        seed = seed_for(4, rank * 4096 + i)
        out.append(make_pair(HEIGHT, WIDTH, seed, shift=(5 - (i % 11), (i % 7) - 3), patches=12, max_patch_shift=40))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        res = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return res
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)]  # under load = samples above the idle clock floor
            res["sm_mhz"] = float(np.median(busy if busy else sm))
            res["sm_max_mhz"] = float(max(mx))
            res["samples"] = len(sm)
        res["reasons"] = sorted(reasons)
        return res


def cpu_reference_sample(pairs, threads):
    """The reference's CPU path on `threads` host threads over len(pairs) independent pairs.  Returns (pairs/s, kind, s)."""
    from oracle import binding as ob
    use_ref = ob.load_ref() is not None
    t0 = time.perf_counter()
    if use_ref:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL inside ref_mf_run
            list(ex.map(lambda p: ob.ref_estimate(p[0], p[1], SEARCH_SIZE, BLOCK_SIZE), pairs))
    else:
        ob.estimate_many(pairs, SEARCH_SIZE, BLOCK_SIZE, SWEEPS, threads)
    dt = time.perf_counter() - t0
    return len(pairs) / dt, ("reference" if use_ref else "port"), dt


def cpu_single_thread(pair):
    """One pair on ONE thread, split the way main() times it (main_class.cpp:47-55: calcMotionBlockMatching only)."""
    from oracle import binding as ob
    r = ob.ref_estimate(pair[0], pair[1], SEARCH_SIZE, BLOCK_SIZE)
    if r is not None:
        return {"constructor_s": r[2], "calcMotionBlockMatching_s": r[3], "kind": "reference"}
    _, st = ob.estimate(pair[0], pair[1], SEARCH_SIZE, BLOCK_SIZE, SWEEPS)
    return {"constructor_s": st["t_ctor_s"], "calcMotionBlockMatching_s": st["t_run_s"], "kind": "port"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pairs = make_pairs(min(cores, 16), 0)
    sample = [pairs[i % len(pairs)] for i in range(cores)]
    for _ in range(args.warmup):
        cpu_reference_sample(sample[:max(1, cores // 4)], cores)
    dts, kind = [], "port"
    for _ in range(args.steps):
        _, kind, dt = cpu_reference_sample(sample, cores)
        dts.append(dt)
    value = len(sample) * len(dts) / sum(dts)
    desc = (f"{len(sample)} distinct-seed 1080p pairs per step, one per host thread, {cores} threads; "
            + ("reference sources (motion_framework.cpp) compiled -O3 against oracle/cvshim" if kind == "reference"
               else "oracle port (oracle/bbme_oracle.c), -O3"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(dts) / len(dts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": CONFIG, "run": {"pairs_per_step": len(sample)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------- other configurations
OTHER = {
    "single_1080p": dict(name="BASELINE config 2: ONE synthetic 1920x1080 pair, 16x16 blocks, +-32, 3 levels (latency)",
                         w=1920, h=1080, ss=[80] * 3, bs=[16] * 3, sweeps=2, pairs=1, golden="c2_1080p"),
    "4k_batch": dict(name="BASELINE config 3: batch of synthetic 3840x2160 pairs, 8x8 blocks, +-64, 4 levels",
                     w=3840, h=2160, ss=[136] * 4, bs=[8] * 4, sweeps=2, pairs=16, golden="c3_4k"),
    "8k_pair": dict(name="BASELINE config 5: synthetic 7680x4320 pair, 16x16 blocks, +-128, 4 levels, 5 sweeps per block size",
                    w=7680, h=4320, ss=[272] * 4, bs=[16] * 4, sweeps=5, pairs=1, golden="c5_8k"),
    "rubberwhale_x4": dict(name="BASELINE config 0/1 geometry: 2336x1552 (RubberWhale 584x388 x4), repo defaults: 32x32 blocks, "
                                "search_size 64, 4 levels (main_class.cpp:19-21)",
                           w=2336, h=1552, ss=[64] * 4, bs=[32] * 4, sweeps=2, pairs=16, golden=None),
}


def golden_inputs(key):
    """(f1, f2, digest dict) of a full-size golden case (tests/golden/big_digests.json), or None."""
    path = os.path.join(ROOT, "tests", "golden", "big_digests.json")
    try:
        case = json.load(open(path))[key]
    except Exception:
        return None
    from blockbasedmotionestimation_b200.synth import make_pair
    s = dict(case["synth"])
    s["shift"] = tuple(s["shift"])
    f1, f2 = make_pair(case["height"], case["width"], case["seed"], **s)
    if [sha(f1), sha(f2)] != case["input_sha256"]:
        return None
    want = case["sweepsN"] if "sweepsN" in case else case["sweeps2"]
    return f1, f2, want


def run_other(key, dev, local_rank, peak_absdiff, reps, check=True):
    """Device-resident timing of one other configuration: CUDA events around `reps` back-to-back steps (after two warm-up
    steps), stage split and work counters from one extra step with stats on."""
    import torch
    import blockbasedmotionestimation_b200 as bb
    from blockbasedmotionestimation_b200.synth import make_pair
    o = OTHER[key]
    w, h, P = o["w"], o["h"], o["pairs"]
    gold = golden_inputs(o["golden"]) if o["golden"] else None
    distinct = []
    if gold is not None:
        distinct.append((gold[0], gold[1]))
    for i in range(len(distinct), min(P, 4)):
        distinct.append(make_pair(h, w, 7000 + 13 * i + len(key), shift=(7 - 2 * i, i - 5), patches=8, max_patch_shift=24))
    shape = bb.plan_shape(w, h, o["ss"], o["bs"])
    Hp, Wp = shape["padded_height"], shape["padded_width"]
    d1 = torch.empty((P, h, w), dtype=torch.uint8, device=dev)
    d2 = torch.empty((P, h, w), dtype=torch.uint8, device=dev)
    for i in range(P):
        a, b = distinct[i % len(distinct)]
        d1[i].copy_(torch.from_numpy(a))
        d2[i].copy_(torch.from_numpy(b))
    dmv = torch.empty((P, Hp // 2, Wp // 2, 2), dtype=torch.int16, device=dev)
    dflow = torch.empty((P, Hp, Wp, 2), dtype=torch.float32, device=dev)
    res = {"workload": o["name"], "pairs_per_step": P}
    stream = torch.cuda.Stream(device=dev)
    for stats in (False, True):
        est = bb.Estimator(w, h, o["ss"], o["bs"], sweeps=o["sweeps"], device=local_rank, chunk_pairs=P, slots=1, collect_stats=stats)
        est.set_streams([stream.cuda_stream])

        def step():
            est.estimate_device_both(P, d1.data_ptr(), d2.data_ptr(), w, w * h, dflow.data_ptr(), Hp * Wp * 2,
                                     dmv.data_ptr(), (Hp // 2) * (Wp // 2) * 2)
        if not stats:
            for _ in range(2):
                step()
            est.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(reps):
                    step()
                e1.record(stream)
            est.sync()
            ms = e0.elapsed_time(e1) / reps
            res["ms"] = ms
            res["pairs_per_s"] = P / (ms * 1e-3)
        else:
            step()
            est.sync()
            st = est.stats()
            res["stage_ms"] = {k: st[k] for k in ("ms_total", "ms_pyramid", "ms_search", "ms_regularize", "ms_other")}
            res["search_absdiffs"] = st["search_absdiffs"]
            res["search_G_absdiff_per_s"] = st["search_absdiffs"] / (st["ms_search"] * 1e-3) / 1e9 if st["ms_search"] > 0 else None
            res["roofline_frac"] = (st["search_absdiffs"] / (st["ms_search"] * 1e-3) / peak_absdiff) if st["ms_search"] > 0 and peak_absdiff > 0 else None
            res["kernel_launches"] = st["kernel_launches"]
        est.close()
    res["bit_exact"] = None
    if check:
        got = dmv[0].cpu().numpy()
        dense_ok = bool(torch.equal(dflow[0, ::2, ::2].to(torch.int16), dmv[0]) and torch.equal(dflow[0, 1::2, 1::2], dflow[0, ::2, ::2]))
        if gold is not None:
            res["bit_exact"] = bool(sha(got) == gold[2]["field_mv2"]) and dense_ok
            res["bit_exact_against"] = f"tests/golden/big_digests.json:{o['golden']} (reference sources at 2 sweeps / oracle port)"
        else:
            from oracle import binding as ob
            want, _ = ob.estimate(distinct[0][0], distinct[0][1], o["ss"], o["bs"], o["sweeps"])
            res["bit_exact"] = bool(np.array_equal(np.rint(want[::2, ::2]).astype(np.int16), got)) and dense_ok
            res["bit_exact_against"] = "oracle port, pair 0"
    del d1, d2, dmv, dflow
    torch.cuda.empty_cache()
    return res


def dropin_mf_cost(lib, pair, reps=8):
    """What a caller of the reference's API pays per pair: MF::MF + calcMotionBlockMatching + ~MF (main_class.cpp:45-50 builds
    one MF per pair) through the same C-ABI calls include/bbme/dropin.hpp makes, pageable buffers like a cv::Mat."""
    f1, f2 = pair
    L = len(BLOCK_SIZE)
    ss = (C.c_int * L)(*SEARCH_SIZE)
    bs = (C.c_int * L)(*BLOCK_SIZE)
    from blockbasedmotionestimation_b200._lib import BbmeShape
    sh = BbmeShape()
    times = []
    flow = None
    for _ in range(reps + 2):
        t0 = time.perf_counter()
        h = C.c_void_p()
        rc = lib.bbme_mf_open(C.byref(h), 0, WIDTH, HEIGHT, L, ss, bs, SWEEPS, C.byref(sh))
        if rc != 0:
            return None
        flow = np.empty((sh.padded_height, sh.padded_width, 2), np.float32)
        rc = lib.bbme_estimate(h, f1.ctypes.data, f2.ctypes.data, WIDTH, flow.ctypes.data)
        lib.bbme_mf_close(h)
        if rc != 0:
            return None
        times.append(time.perf_counter() - t0)
    lib.bbme_mf_cache_clear()
    return {"ms_per_pair": 1e3 * float(np.median(times[2:])), "first_ms": 1e3 * times[0],
            "what": "bbme_mf_open (geometry-keyed context cache) + bbme_estimate into a fresh pageable buffer + bbme_mf_close, "
                    "host wall clock per 1080p pair, median of %d" % reps}, flow


def run_ours(args, rank, local_rank, world):
    import torch
    import blockbasedmotionestimation_b200 as bb
    from blockbasedmotionestimation_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs
    # ---- synthetic input (distinct seeds per rank), tiled to P pairs
    distinct = make_pairs(min(P, args.distinct), rank)
    shape = bb.plan_shape(WIDTH, HEIGHT, SEARCH_SIZE, BLOCK_SIZE)
    Hp, Wp = shape["padded_height"], shape["padded_width"]
    h1 = torch.empty((P, HEIGHT, WIDTH), dtype=torch.uint8, pin_memory=True)
    h2 = torch.empty((P, HEIGHT, WIDTH), dtype=torch.uint8, pin_memory=True)
    for i in range(P):
        a, b = distinct[i % len(distinct)]
        h1[i].copy_(torch.from_numpy(a))
        h2[i].copy_(torch.from_numpy(b))
    hout = torch.empty((P, Hp, Wp, 2), dtype=torch.float32, pin_memory=True)
    hout_b = torch.empty((P, Hp, Wp, 2), dtype=torch.float32, pin_memory=True)  # steps alternate: no two in-flight steps share a destination
    d1 = h1.to(dev)
    d2 = h2.to(dev)
    dout = torch.empty((P, Hp, Wp, 2), dtype=torch.float32, device=dev)
    # compact (2x2-granular int16) copies of the same fields: what the ranks hand to rank 0 at the end of a step (N > 1)
    dmv = [torch.empty((P, Hp // 2, Wp // 2, 2), dtype=torch.int16, device=dev) for _ in range(2 if world > 1 else 0)]

    # ---- device-resident arm: `value` (stats collection off; the stage split comes from one extra step below)
    est = bb.Estimator(WIDTH, HEIGHT, SEARCH_SIZE, BLOCK_SIZE, sweeps=SWEEPS, device=local_rank, chunk_pairs=args.chunk,
                       slots=args.slots)
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.slots)]
    est.set_streams([s.cuda_stream for s in streams])
    peak_absdiff, peak_mhz = est.measure_int_peak()

    from blockbasedmotionestimation_b200.shard import ResultGather
    main = torch.cuda.current_stream(dev)
    gather = ResultGather(dmv[0], world * P, n_buffers=2) if world > 1 else None
    gather_done = [None, None]

    def step_device(k=0):
        if world == 1:
            est.estimate_device(P, d1.data_ptr(), d2.data_ptr(), WIDTH, WIDTH * HEIGHT, dout.data_ptr(), Hp * Wp * 2)
            return
        # N > 1: pairs are sharded by rank (no data-path collective); the only exchange is the gather of the step's compact
        # fields on rank 0, double-buffered so that it overlaps the next step's kernels
        buf = dmv[k & 1]
        if gather_done[k & 1] is not None:
            for s_ in streams:
                s_.wait_event(gather_done[k & 1])
        est.estimate_device_both(P, d1.data_ptr(), d2.data_ptr(), WIDTH, WIDTH * HEIGHT, dout.data_ptr(), Hp * Wp * 2,
                                 buf.data_ptr(), (Hp // 2) * (Wp // 2) * 2)
        for s_ in streams:
            main.wait_stream(s_)
        if not args.no_gather:
            gather.push(buf, k & 1)  # enqueued on the current stream of this rank
        ev = torch.cuda.Event()
        ev.record(main)
        gather_done[k & 1] = ev

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for k in range(args.warmup):
        step_device(k)
        est.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(main)
    for s in streams:
        s.wait_event(ev0)
    for k in range(args.steps):
        step_device(k)  # no host synchronisation between steps
    for s in streams:
        main.wait_stream(s)
    ev1.record(main)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    est.sync()
    ms_total = ev0.elapsed_time(ev1)
    launches = est.stats()["kernel_launches"] * args.steps

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * P * args.steps / (ms_max * 1e-3)

    # ---- parity of what was timed: EVERY distinct pair of this rank against the oracle (outside the timed region)
    parity = None
    if not args.no_check:
        from oracle import binding as ob
        cores = max(1, (os.cpu_count() or 1) // world)
        wants, _ = ob.estimate_many(distinct, SEARCH_SIZE, BLOCK_SIZE, SWEEPS, cores)
        ok = True
        for i in range(P):
            if i < len(distinct) or i >= P - len(distinct):  # every distinct pair, at its first and last position in the batch
                ok = ok and bool(np.array_equal(dout[i].cpu().numpy(), wants[i % len(distinct)]))
        tt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        parity = bool(tt.item() == 1)

    gather_ok = None
    if world > 1 and not args.no_check and not args.no_gather:
        barrier()
        if rank == 0:
            last = gather.result((args.steps - 1) & 1)
            own = dout[:, ::2, ::2, :].to(torch.int16)
            gather_ok = bool(torch.equal(last[:P], own)) and tuple(last.shape) == (world * P, Hp // 2, Wp // 2, 2)
            # every other rank's slice: 2x2-granular fields of ITS pairs; compare a checksum exchanged through NCCL
        sums = torch.zeros(world, dtype=torch.int64, device=dev)
        sums[rank] = dout[:, ::2, ::2, :].to(torch.int64).sum()
        dist.all_reduce(sums)
        if rank == 0:
            got = torch.stack([last[r * P:(r + 1) * P].to(torch.int64).sum() for r in range(world)])
            gather_ok = gather_ok and bool(torch.equal(got, sums))

    # ---- stage split and roofline of the dominant kernel (the search): one extra, untimed step with stats on
    est.close()
    est_s = bb.Estimator(WIDTH, HEIGHT, SEARCH_SIZE, BLOCK_SIZE, sweeps=SWEEPS, device=local_rank, chunk_pairs=P,
                         slots=1, collect_stats=True)  # the whole step as ONE chunk: the launches the ncu captures under profiles/ show
    for _ in range(2):
        est_s.estimate_device(P, d1.data_ptr(), d2.data_ptr(), WIDTH, WIDTH * HEIGHT, dout.data_ptr(), Hp * Wp * 2)
        est_s.sync()
    st = est_s.stats()
    est_s.close()
    absdiffs = st["search_absdiffs"]
    ms_search = st["ms_search"]
    n_search = max(1, st["search_launches"])
    achieved = absdiffs / (ms_search * 1e-3) if ms_search > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes_pair = 2 * WIDTH * HEIGHT + 8 * Wp * Hp
    roofline = {
        "bound": "int",  # integer SIMD (VABSDIFF4) issue rate; HBM time is ~50x smaller (see roofline_hbm)
        "kernel": "k_search_tma<16,13,24> (all pyramid levels)",
        "achieved": achieved / 1e9, "peak": peak_absdiff / 1e9, "unit": "G absdiff/s",
        "frac": (achieved / peak_absdiff) if peak_absdiff > 0 else None,
        "peak_source": f"live VABSDIFF4.U8.ACC issue-rate micro-benchmark on this GPU at {peak_mhz:.0f} MHz "
                       "(bbme_measure_int_peak; MEASURED_PEAKS.json has no integer entry)",
        "measured": "CUDA events around the search launches of one extra step run as a single chunk on one stream (stats on), not of the timed steps",
        "algorithmic_absdiffs_per_launch": absdiffs / n_search,
        "avg_launch_ms": ms_search / n_search, "launches_per_step": n_search,
        "traffic": None,
        "traffic_from_profile": {"bytes": 1398.4e6, "file": "profiles/r02_search_l0_ncu.txt",
                                 "note": "dram read+write of the level-0 launch of a 128-pair chunk from a committed ncu --set full "
                                         "capture, NOT measured in this run; algorithmic 1 331 MB: image 1 (266 MB), image 2 in its four "
                                         "byte-shifted copies (1 061 MB, written by k_shift4 just before), 4 MB of vectors"},
    }
    roofline_hbm = {
        "bound": "hbm", "algorithmic_bytes_per_pair": alg_bytes_pair,
        "achieved": (alg_bytes_pair * P * args.steps / (ms_max * 1e-3)) / 1e9, "peak": hbm_peak, "unit": "GB/s per GPU",
        "frac": (alg_bytes_pair * P * args.steps / (ms_max * 1e-3)) / 1e9 / hbm_peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
    }
    stage_ms = {k: st[k] for k in ("ms_total", "ms_pyramid", "ms_search", "ms_regularize", "ms_other")}

    # ---- end-to-end arm: host buffers through the C ABI (H2D + pipeline + D2H + expansion inside the timed region)
    est2 = bb.Estimator(WIDTH, HEIGHT, SEARCH_SIZE, BLOCK_SIZE, sweeps=SWEEPS, device=local_rank, chunk_pairs=args.e2e_chunk,
                        slots=args.e2e_slots)
    lib = _lib.load()
    PA = C.c_void_p * P
    in_stride = WIDTH * HEIGHT
    out_stride = Hp * Wp * 2 * 4
    p1 = PA(*[h1.data_ptr() + i * in_stride for i in range(P)])
    p2 = PA(*[h2.data_ptr() + i * in_stride for i in range(P)])
    po = [PA(*[hb.data_ptr() + i * out_stride for i in range(P)]) for hb in (hout, hout_b)]

    def step_e2e(k=0):
        # asynchronous host-buffer call: H2D of this step's frames, the pipeline, D2H of the compact fields and their
        # expansion into the dense host buffers are all enqueued here; successive steps overlap on the slots.
        # est2.sync() below closes the timed region (streams drained AND every field expanded).
        rc = lib.bbme_estimate_batch_async(est2._ctx, P, p1, p2, WIDTH, po[k & 1])
        if rc != 0:
            raise RuntimeError("bbme_estimate_batch_async failed: " + lib.bbme_last_error(est2._ctx).decode())

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.no_e2e:
        e2e_steps = 0
    hout.zero_()
    hout_b.zero_()
    for k in range(min(args.warmup, 2) if e2e_steps else 0):
        step_e2e(k)
    est2.sync()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        step_e2e(k)
    est2.sync()  # every step's fields are in host memory
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()  # sampled over both timed regions (device arm and host-buffer arm)
    e2e_value = world * P * e2e_steps / float(t.item()) if e2e_steps else None
    e2e_ok = None
    if not args.no_check and e2e_steps:
        last = hout if ((e2e_steps - 1) & 1) == 0 else hout_b
        e2e_ok = bool(torch.equal(last, dout.cpu()))
    # ceiling of the host side of the path, measured the same way with the kernels left out: every rank at the same time runs its
    # H2D copies, compact D2H copies and expansions (bbme_debug_skip_compute), i.e. the link, the pinned staging, the worker
    # threads and the host memory under the traffic mix of the real arm
    host_path = None
    if e2e_steps:
        lib.bbme_debug_skip_compute(est2._ctx, 1)
        for k in range(2):
            step_e2e(k)
        est2.sync()
        barrier()
        t0 = time.perf_counter()
        hp_steps = max(2, min(e2e_steps, 10))
        for k in range(hp_steps):
            step_e2e(k)
        est2.sync()
        dt_hp = time.perf_counter() - t0
        lib.bbme_debug_skip_compute(est2._ctx, 0)
        thp = torch.tensor([dt_hp], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(thp, op=dist.ReduceOp.MAX)
        host_path = world * P * hp_steps / float(thp.item())
        step_e2e(0)  # leave real fields in the host buffers again (the check below already ran)
        est2.sync()
    # the individual ceilings (all ranks measure at the same time: the link and the host memory are shared)
    barrier()
    link = est2.measure_host_link(256 << 20) if e2e_steps else None
    if link is not None:
        lt = torch.tensor([link["h2d_gbs"], link["d2h_gbs"], link["duplex_gbs_per_direction"], link["host_stream_write_gbs"]],
                          dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        h2d_b, d2h_b, dense_b = 2 * WIDTH * HEIGHT, Hp * Wp, 8 * Hp * Wp
        tot = [float(x) for x in lt.tolist()]
        ceil_pairs = min(tot[2] * 1e9 / h2d_b, tot[2] * 1e9 / d2h_b, tot[3] * 1e9 / dense_b, value, host_path)
        link = {"h2d_gbs_all_ranks": tot[0], "d2h_gbs_all_ranks": tot[1], "duplex_gbs_per_direction_all_ranks": tot[2],
                "host_stream_write_gbs_all_ranks": tot[3], "host_threads_per_rank": link["host_threads"],
                "host_path_pairs_per_s": host_path,
                "host_path_note": "the same host-buffer call on every rank at once with the kernels left out (copies + expansion only): "
                                  "what the host side of this box sustains for this traffic mix",
                "ceiling_pairs_per_s": ceil_pairs,
                "ceiling_note": "min(host_path_pairs_per_s, device-resident value, and the individually measured link / host-write ceilings)"}
    est2.close()

    # ---- the other BASELINE configurations (outside the headline); under N > 1 only the 4K batch, on every rank
    others = []
    if not args.no_other:
        keys = ["single_1080p", "4k_batch", "8k_pair", "rubberwhale_x4"] if world == 1 else ["4k_batch"]
        for key in keys:
            try:
                r = run_other(key, dev, local_rank, peak_absdiff, reps=args.other_reps, check=not args.no_check)
            except Exception as ex:  # an other-config failure must not take the headline down
                r = {"workload": OTHER[key]["name"], "error": repr(ex)}
            if dist is not None and "ms" in r:
                tm = torch.tensor([r["ms"]], dtype=torch.float64, device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                r["ms"] = float(tm.item())
                r["pairs_per_s"] = world * r["pairs_per_step"] / (r["ms"] * 1e-3)
                r["pairs_per_step"] = world * r["pairs_per_step"]
            others.append(r)

    # ---- the drop-in's per-object cost and the CPU baseline next to it (rank 0, N=1 only)
    cpu = None
    dropin = None
    if rank == 0 and world == 1:
        try:
            d = dropin_mf_cost(lib, distinct[0])
        except AttributeError:
            d = None
        if d is not None:
            dropin = d[0]
            if not args.no_check:
                dropin["bit_exact"] = bool(np.array_equal(d[1], dout[0].cpu().numpy()))
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = [distinct[i % len(distinct)] for i in range(cores)]
        v, kind, dt_cpu = cpu_reference_sample(sample, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{len(sample)} of this run's 1080p pairs, one per host thread ({dt_cpu:.1f} s wall); "
                         + ("reference sources compiled -O3 against oracle/cvshim (oracle/_ref)" if kind == "reference"
                            else "oracle port, -O3"),
               "single_thread_one_pair": cpu_single_thread(distinct[0])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": CONFIG,
            "run": {"pairs_per_gpu_per_step": P, "distinct_pairs_per_gpu": len(distinct),
                    "parallelism": f"pairs sharded over {world} GPU(s), one process per GPU; "
                                   + ("no communication" if world == 1 else
                                      f"per step the compact int16 fields are gathered on rank 0 ({gather.how}), inside the timed region"),
                    "chunk_pairs": args.chunk, "slots": args.slots,
                    "l2": f"inputs larger than L2: {2 * P * WIDTH * HEIGHT / 1e6:.0f} MB of frames and "
                          f"{P * Hp * Wp * 8 / 1e6:.0f} MB of output per step vs 126 MB L2",
                    "timing": "CUDA events on the launching streams, max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * P * WIDTH * HEIGHT,
                    "d2h_bytes_per_step": P * Hp * Wp, "host_expanded_bytes_per_step": P * Hp * Wp * 8,
                    "steps": e2e_steps, "chunk_pairs": args.e2e_chunk, "slots": args.e2e_slots, "matches_device_arm": e2e_ok,
                    "frac_of_value": (e2e_value / value) if e2e_value else None,
                    "frac_of_host_ceiling": (e2e_value / link["ceiling_pairs_per_s"]) if (e2e_value and link) else None,
                    "host_link": link,
                    "note": "D2H moves the 2x2-granular int16 field (1/8 of the dense bytes); worker threads expand it into the "
                            "caller's dense padded CV_32FC2 buffers (motion_framework.cpp:218) inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu,
            "stage_ms_stats_step": stage_ms,
            "fix_rounds_stats_step": st["fix_rounds"], "fix_blocks_stats_step": st["fix_blocks"],
            "bit_exact_vs_oracle": parity,
            "bit_exact_note": "every distinct pair of every rank, whole dense field, against the oracle port (pinned to the reference's sources)",
            "gather_matches_local_fields": gather_ok,
            "other_configs": others,
            "dropin_mf": dropin,
            "wall_s_timed_region": t_wall,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        if gather is not None:
            gather.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=128, help="frame pairs per GPU per step")
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic pairs generated per GPU (tiled to --pairs)")
    # chunks x slots of the two arms (profiles/r02_chunk_slot_sweep.md): several chunks in flight let one chunk's
    # regularisation (one CTA per pair, latency-bound) share the GPU with another chunk's search (every SM, ALU-bound)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--e2e-chunk", type=int, default=64)
    ap.add_argument("--e2e-slots", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=1000, help="cap on the e2e arm's steps (default: same as --steps)")
    ap.add_argument("--other-reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer arm")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="diagnosis only (N > 1): skip the per-step gather of the fields on rank 0")
    args = ap.parse_args()
    claim_stdout()
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
