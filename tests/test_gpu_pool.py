"""The multi-GPU batch call and the MF-per-pair context cache of the C ABI (csrc/multi.cpp), on whatever GPUs the box has."""
import ctypes as C

import numpy as np
import pytest

import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200 import _lib
from helpers import describe_diff, make_pair

pytestmark = pytest.mark.gpu


def test_pool_batch_equals_single_context(oracle):
    """bbme_pool_estimate_batch (pairs sharded over every visible GPU, ragged last chunk) == the oracle, pair by pair."""
    h, w, ss, bs = 128, 192, [24, 24], [8, 8]
    pairs = [make_pair(h, w, 900 + i, shift=(i % 5 - 2, 1 - i % 3)) for i in range(11)]
    with bb.Pool(w, h, ss, bs, chunk_pairs=4, slots=2) as pool:
        assert pool.device_count >= 1
        got = pool.estimate_batch([p[0] for p in pairs], [p[1] for p in pairs])
    for (f1, f2), g in zip(pairs, got):
        want, _ = oracle.estimate(f1, f2, ss, bs, 2)
        assert np.array_equal(g, want), describe_diff(g, want)


def test_pool_on_named_devices_and_bad_device():
    lib = _lib.load()
    p = C.c_void_p()
    assert lib.bbme_pool_create(C.byref(p), 1, (C.c_int * 1)(0)) == 0
    assert lib.bbme_pool_device_count(p) == 1
    lib.bbme_pool_destroy(p)
    assert lib.bbme_pool_create(C.byref(p), 1, (C.c_int * 1)(999)) == -1  # BBME_E_ARG: device out of range


def test_mf_cache_reuses_contexts(oracle):
    """One MF per pair (main_class.cpp:45-50): bbme_mf_open hands the parked context of the same geometry back."""
    lib = _lib.load()
    h, w, ss, bs = 96, 128, [24, 24], [8, 8]
    L = 2
    ssa, bsa = (C.c_int * L)(*ss), (C.c_int * L)(*bs)
    sh = _lib.BbmeShape()
    seen = []
    for i in range(4):
        f1, f2 = make_pair(h, w, 70 + i)
        ctx = C.c_void_p()
        assert lib.bbme_mf_open(C.byref(ctx), 0, w, h, L, ssa, bsa, 2, C.byref(sh)) == 0
        seen.append(ctx.value)
        flow = np.empty((sh.padded_height, sh.padded_width, 2), np.float32)
        assert lib.bbme_estimate(ctx, f1.ctypes.data, f2.ctypes.data, w, flow.ctypes.data) == 0
        lib.bbme_mf_close(ctx)
        want, _ = oracle.estimate(f1, f2, ss, bs, 2)
        assert np.array_equal(flow, want)
    assert len(set(seen)) == 1  # the same context every time
    # two objects alive at once get two contexts; a different geometry gets its own
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.bbme_mf_open(C.byref(a), 0, w, h, L, ssa, bsa, 2, None) == 0
    assert lib.bbme_mf_open(C.byref(b), 0, w, h, L, ssa, bsa, 2, None) == 0
    assert lib.bbme_mf_open(C.byref(c), 0, w, h, L, ssa, bsa, 3, None) == 0
    assert len({a.value, b.value, c.value}) == 3 and a.value == seen[0]
    for x in (a, b, c):
        lib.bbme_mf_close(x)
    # a geometry the reference cannot pad: status + message, and the context goes away on close
    bad = C.c_void_p()
    rc = lib.bbme_mf_open(C.byref(bad), 0, 3, 64, 1, (C.c_int * 1)(16), (C.c_int * 1)(8), 2, None)
    assert rc == -2 and bad.value
    assert b"multiples of the block size" in lib.bbme_last_error(bad)
    lib.bbme_mf_close(bad)
    lib.bbme_mf_cache_clear()


def test_skip_compute_is_only_a_measurement_aid(oracle):
    """bbme_debug_skip_compute: copies and expansion without kernels (the bench's host-path ceiling); switching it off restores
    the real path."""
    lib = _lib.load()
    h, w, ss, bs = 96, 128, [24, 24], [8, 8]
    f1, f2 = make_pair(h, w, 123)
    want, _ = oracle.estimate(f1, f2, ss, bs, 2)
    with bb.Estimator(w, h, ss, bs) as est:
        assert np.array_equal(est.estimate(f1, f2), want)
        assert lib.bbme_debug_skip_compute(est._ctx, 1) == 0
        g1, g2 = make_pair(h, w, 124, shift=(-4, 3))
        stale = est.estimate(g1, g2)          # no kernel ran: the field of the previous call, expanded again
        assert np.array_equal(stale, want)
        assert lib.bbme_debug_skip_compute(est._ctx, 0) == 0
        want2, _ = oracle.estimate(g1, g2, ss, bs, 2)
        assert np.array_equal(est.estimate(g1, g2), want2)


def test_stamp_epoch_restarts_before_it_wraps(oracle):
    """The de-duplication stamps of the regularisation carry a 32-bit epoch that grows for the lifetime of a plan (ADVICE round
    1): just below the wrap the kernel clears the stamps and restarts the epoch, and the field stays exact."""
    lib = _lib.load()
    h, w, ss, bs = 128, 192, [24, 24], [8, 8]
    pairs = [make_pair(h, w, 300 + i, shift=(i - 1, 2 - i), patches=3) for i in range(3)]
    wants = [oracle.estimate(a, b, ss, bs, 2)[0] for a, b in pairs]
    with bb.Estimator(w, h, ss, bs, chunk_pairs=3) as est:
        got = est.estimate_batch([p[0] for p in pairs], [p[1] for p in pairs])
        assert all(np.array_equal(g, w_) for g, w_ in zip(got, wants))
        for epoch in (0xf0000001, 0xffffff00, 0xefffffff):
            assert lib.bbme_debug_set_stamp_epoch(est._ctx, epoch) == 0
            for _ in range(2):  # the second call runs on the restarted (or still growing) epoch with stale stamps below it
                got = est.estimate_batch([p[0] for p in pairs], [p[1] for p in pairs])
                assert all(np.array_equal(g, w_) for g, w_ in zip(got, wants)), hex(epoch)
