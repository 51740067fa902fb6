#!/usr/bin/env python
"""bench.py -- 1080p frame-pairs/s of the block-matching hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one batch of `--pairs` synthetic 1080p frame pairs per GPU through the whole path
(pad -> pyramid -> per-level SAD search -> regularisation with block splitting -> MV propagation -> dense field);
BASELINE.json config 4 (batch of 1080p pairs, 16x16 blocks, +-32 search, 3 levels) sharded over the ranks with no
data-path collective.  `value` is measured with the inputs resident in HBM (C ABI bbme_estimate_device), `e2e`
through the host-buffer C ABI call (bbme_estimate_batch: pinned H2D + pipeline + D2H inside the timed region).
`--impl reference` times the reference's own CPU implementation (oracle/_ref: the reference's sources compiled against
oracle/cvshim; falls back to the oracle port) on all host cores of rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
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
            # under load = samples above the idle clock floor
            busy = [v for v in sm if v > 0.5 * max(sm)]
            res["sm_mhz"] = float(np.median(busy if busy else sm))
            res["sm_max_mhz"] = float(max(mx))
            res["samples"] = len(sm)
        res["reasons"] = sorted(reasons)
        return res


def cpu_reference_sample(pairs, threads):
    """The reference's CPU path on `threads` host threads over len(pairs) independent pairs.  Returns (pairs/s, kind)."""
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pairs = make_pairs(min(cores, 16), 0)
    sample = [pairs[i % len(pairs)] for i in range(cores)]
    for _ in range(args.warmup):
        cpu_reference_sample(sample[:max(1, cores // 4)], cores)
    vals, dts, kind = [], [], "port"
    for _ in range(args.steps):
        v, kind, dt = cpu_reference_sample(sample, cores)
        vals.append(v)
        dts.append(dt)
    value = len(sample) * len(dts) / sum(dts)
    desc = (f"{len(sample)} distinct-seed 1080p pairs per step, one per host thread, {cores} threads; "
            + ("reference sources (motion_framework.cpp) compiled -O3 against oracle/cvshim" if kind == "reference"
               else "oracle port (oracle/bbme_oracle.c), -O3"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(dts) / len(dts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": len(sample)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


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
    # compact (2x2-granular int16) copies of the same fields: what the ranks exchange at the end of a step (N > 1)
    dmv = [torch.empty((P, Hp // 2, Wp // 2, 2), dtype=torch.int16, device=dev) for _ in range(2 if world > 1 else 0)]

    # ---- device-resident arm: `value`
    est = bb.Estimator(WIDTH, HEIGHT, SEARCH_SIZE, BLOCK_SIZE, sweeps=SWEEPS, device=local_rank, chunk_pairs=args.chunk,
                       slots=args.slots, collect_stats=True)
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.slots)]
    est.set_streams([s.cuda_stream for s in streams])
    peak_absdiff, peak_mhz = est.measure_int_peak()

    from blockbasedmotionestimation_b200.shard import gather_fields
    main = torch.cuda.current_stream(dev)
    gathered = [None, None]
    gather_done = [None, None]

    def step_device(k=0):
        if world == 1:
            est.estimate_device(P, d1.data_ptr(), d2.data_ptr(), WIDTH, WIDTH * HEIGHT, dout.data_ptr(), Hp * Wp * 2)
            return
        # N > 1: pairs are sharded by rank (no data-path collective); the only exchange is the gather of the step's
        # compact fields over NCCL, double-buffered so that it overlaps the next step's kernels
        buf = dmv[k & 1]
        if gather_done[k & 1] is not None:
            for s_ in streams:
                s_.wait_event(gather_done[k & 1])
        est.estimate_device_both(P, d1.data_ptr(), d2.data_ptr(), WIDTH, WIDTH * HEIGHT, dout.data_ptr(), Hp * Wp * 2,
                                 buf.data_ptr(), (Hp // 2) * (Wp // 2) * 2)
        for s_ in streams:
            main.wait_stream(s_)
        gathered[k & 1] = gather_fields(buf, world * P)
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
    per_step_stats = []
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(main)
    for s in streams:
        s.wait_event(ev0)
    launches = 0
    for k in range(args.steps):
        step_device(k)  # no host synchronisation between steps
    for s in streams:
        main.wait_stream(s)
    ev1.record(main)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    est.sync()
    ms_total = ev0.elapsed_time(ev1)
    st = est.stats()  # of the LAST step (stats reset at each call)
    launches = st["kernel_launches"] * args.steps
    per_step_stats.append(st)

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * P * args.steps / (ms_max * 1e-3)

    # ---- parity spot check of what was timed (rank 0, one pair, against the oracle; not in the timed region)
    parity = None
    if rank == 0 and not args.no_check:
        from oracle import binding as ob
        want, ost = ob.estimate(distinct[0][0], distinct[0][1], SEARCH_SIZE, BLOCK_SIZE, SWEEPS)
        got = dout[0].cpu().numpy()
        parity = bool(np.array_equal(got, want))

    gather_ok = None
    if world > 1 and rank == 0 and not args.no_check:
        last = gathered[(args.steps - 1) & 1]
        own = dout[:, ::2, ::2, :].to(torch.int16)
        gather_ok = bool(torch.equal(last[rank * P:(rank + 1) * P], own)) and tuple(last.shape) == (world * P, Hp // 2, Wp // 2, 2)

    # ---- roofline of the dominant kernel (the search), from the CUDA-event intervals of the last timed step
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
        "algorithmic_absdiffs_per_launch": absdiffs / n_search,
        "avg_launch_ms": ms_search / n_search, "launches_per_step": n_search,
        # dram__bytes_read.sum + dram__bytes_write.sum of the level-0 launch (128 pairs) from the committed ncu capture
        # profiles/r01c_search_l0_ncu.txt (not re-measured here); algorithmic bytes of that launch: 2 frames x 2.09 MB x 128
        "traffic": 566.6e6 if P == 128 else None,
        "traffic_note": "bytes per level-0 launch of 128 pairs (ncu --set full, profiles/r01c_search_l0_ncu.txt); algorithmic 535 MB + 4 MB of vectors",
    }
    roofline_hbm = {
        "bound": "hbm", "algorithmic_bytes_per_pair": alg_bytes_pair,
        "achieved": (alg_bytes_pair * world * P * args.steps / (ms_max * 1e-3)) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": (alg_bytes_pair * P * args.steps / (ms_max * 1e-3)) / 1e9 / hbm_peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
    }
    stage_ms = {k: st[k] for k in ("ms_total", "ms_pyramid", "ms_search", "ms_regularize", "ms_other")}

    # ---- end-to-end arm: host buffers through the C ABI (H2D + pipeline + D2H inside the timed region)
    est.close()
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
        # asynchronous host-buffer call: H2D of this step's frames, the pipeline, D2H of this step's fields are all
        # enqueued here; successive steps overlap on the slots.  est2.sync() below closes the timed region.
        rc = lib.bbme_estimate_batch_async(est2._ctx, P, p1, p2, WIDTH, po[k & 1])
        if rc != 0:
            raise RuntimeError("bbme_estimate_batch_async failed: " + lib.bbme_last_error(est2._ctx).decode())

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.no_e2e:
        e2e_steps = 0
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
    if rank == 0 and not args.no_check and e2e_steps:
        e2e_ok = bool(torch.equal(hout[0], dout[0].cpu()) and torch.equal(hout_b[P - 1], dout[P - 1].cpu()))
    est2.close()

    # ---- CPU baseline next to it (rank 0, N=1 only): bounded sample, all host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = [distinct[i % len(distinct)] for i in range(cores)]
        v, kind, dt_cpu = cpu_reference_sample(sample, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{len(sample)} of this run's 1080p pairs, one per host thread ({dt_cpu:.1f} s wall); "
                         + ("reference sources compiled -O3 against oracle/cvshim (oracle/_ref)" if kind == "reference"
                            else "oracle port, -O3")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": P, "distinct_pairs_per_gpu": len(distinct),
                       "parallelism": f"pairs sharded over {world} GPU(s), one process per GPU; "
                                      + ("no communication" if world == 1 else
                                         "per step one NCCL all-gather of the compact int16 fields (inside the timed region)"),
                       "chunk_pairs": args.chunk, "slots": args.slots,
                       "l2": f"inputs larger than L2: {2 * P * WIDTH * HEIGHT / 1e6:.0f} MB of frames and "
                             f"{P * Hp * Wp * 8 / 1e6:.0f} MB of output per step vs 126 MB L2",
                       "timing": "CUDA events on the launching streams, max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * P * WIDTH * HEIGHT,
                    "d2h_bytes_per_step": P * Hp * Wp * 8, "steps": e2e_steps, "chunk_pairs": args.e2e_chunk,
                    "slots": args.e2e_slots, "matches_device_arm": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu,
            "stage_ms_last_step": stage_ms,
            "fix_rounds_last_step": st["fix_rounds"], "fix_blocks_last_step": st["fix_blocks"], "fix_tail_blocks_last_step": st["reserved"],
            "bit_exact_vs_oracle": parity, "gather_matches_local_fields": gather_ok,
            "wall_s_timed_region": t_wall,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=128, help="frame pairs per GPU per step")
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic pairs generated per GPU (tiled to --pairs)")
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--slots", type=int, default=1)
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--e2e-slots", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=1000, help="cap on the e2e arm's steps (default: same as --steps)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer arm")
    ap.add_argument("--no-check", action="store_true")
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
