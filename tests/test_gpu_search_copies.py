"""The search kernel over byte-shifted copies of the window image (PRE instantiations of k_search_tma, k_shift4), the block
counter, and many chunks in flight -- all against the oracle (find_min_block_spiral, motion_framework.cpp:296-422), bit-exact.
Geometries are chosen by the window's row pitch in words, ((15 + 2R) >> 2) + bs / 4 + 1 rounded up to a pitch class: 24 and
40 take the copies, the others (and BBME_SEARCH_PRE=0) the funnel-shift kernels."""
import os

import numpy as np
import pytest

import blockbasedmotionestimation_b200 as bb
from helpers import blocks_to_dense, dense_to_blocks, describe_diff, make_pair

pytestmark = pytest.mark.gpu

# (block, search_size, h, w): pitch class, what is special
COPIES = [(16, 80, 192, 288),    # 24 words: BASELINE config 2's kernel, one band
          (32, 64, 256, 384),    # 24 words, three work items per unit: the deep-ring variant
          (8, 64, 128, 192),     # 24 words, 8x8 tile
          (16, 128, 256, 384),   # 40 words, +-56: four copies only fit in bands of candidate rows
          (32, 128, 256, 320)]   # 40 words, 32x32 blocks, 14-bit ranks


def _run(gpu, oracle, bs, ss, h, w, pred_mode, seed):
    f1, f2 = make_pair(h, w, seed, shift=(5, -3), max_patch_shift=12)
    rng = np.random.default_rng(seed)
    gh, gw = h // bs, w // bs
    lim = 6 if pred_mode == "small" else max(h, w) // 3   # "wild": predictions that partly leave the image (:304-310)
    pred = rng.integers(-lim, lim + 1, (gh, gw, 2)).astype(np.int16)
    got, st = gpu.stage_search(f1, f2, bs, ss, pred, kernel=2)
    want_dense, ost = oracle.search_level(f1, f2, bs, ss, blocks_to_dense(pred, bs, h, w))
    return got, dense_to_blocks(want_dense, bs), st, ost


@pytest.mark.parametrize("bs,ss,h,w", COPIES)
@pytest.mark.parametrize("pred_mode", ["small", "wild"])
def test_search_over_shifted_copies_matches_oracle(gpu, oracle, bs, ss, h, w, pred_mode):
    got, want, st, ost = _run(gpu, oracle, bs, ss, h, w, pred_mode, 3000 + bs + ss)
    assert st["search_kernel_used"] == 2
    assert np.array_equal(got, want), describe_diff(got, want)
    assert st["search_absdiffs"] == ost["search_absdiffs"]


@pytest.mark.parametrize("bs,ss,h,w", COPIES[:2] + COPIES[3:4])
def test_funnel_shift_kernels_agree_with_the_copies(gpu, oracle, monkeypatch, bs, ss, h, w):
    """BBME_SEARCH_PRE=0 (read at plan time) keeps the kernels that align bytes in the loop: same field, same work count."""
    a, want, sa, _ = _run(gpu, oracle, bs, ss, h, w, "small", 3100 + bs)
    monkeypatch.setenv("BBME_SEARCH_PRE", "0")
    b, _, sb, _ = _run(gpu, oracle, bs, ss, h, w, "small", 3100 + bs)
    assert np.array_equal(a, want) and np.array_equal(b, want)
    assert sa["search_absdiffs"] == sb["search_absdiffs"]


def test_columns_of_every_byte_phase_at_the_right_image_edge(gpu, oracle):
    """The copies hold src[x + c]; their last columns come from past the row (zero / pitch padding) and may only ever feed
    candidates that leave the image.  A bright right edge and predictions pointing at it would show a leak."""
    h, w, bs, ss = 96, 176, 16, 80
    f1, f2 = make_pair(h, w, 3200, shift=(1, 0))
    f1 = f1.copy(); f2 = f2.copy()
    f1[:, -20:] = 255
    f2[:, -20:] = 255
    for dx in range(8):  # every byte phase of the window origin
        pred = np.zeros((h // bs, w // bs, 2), np.int16)
        pred[..., 0] = 24 + dx
        got, _ = gpu.stage_search(f1, f2, bs, ss, pred, kernel=2)
        want = dense_to_blocks(oracle.search_level(f1, f2, bs, ss, blocks_to_dense(pred, bs, h, w))[0], bs)
        assert np.array_equal(got, want), f"dx {dx}: " + describe_diff(got, want)


def test_eight_slots_and_ragged_chunks(oracle):
    """More chunks in flight than before (slots up to 8), a ragged last chunk, the block counter of every level reused by
    successive chunks of a slot."""
    h, w, ss, bs = 128, 192, [80, 80], [16, 16]   # level 0 runs the copies kernel
    pairs = [make_pair(h, w, 3300 + i, shift=(i % 7 - 3, i % 3 - 1), patches=3) for i in range(21)]
    with bb.Estimator(w, h, ss, bs, chunk_pairs=2, slots=8) as est:
        flows = est.estimate_batch([p[0] for p in pairs], [p[1] for p in pairs])
        again = est.estimate_batch([p[0] for p in pairs[:5]], [p[1] for p in pairs[:5]])
    for i, p in enumerate(pairs):
        want, _ = oracle.estimate(p[0], p[1], ss, bs, 2)
        assert np.array_equal(flows[i], want), f"pair {i}: " + describe_diff(flows[i], want)
    for i in range(5):
        assert np.array_equal(again[i], flows[i])
    with pytest.raises(bb.BbmeError):
        bb.Estimator(w, h, ss, bs, chunk_pairs=2, slots=9)


def test_sequence_mode_uses_the_next_frame_copies(oracle):
    """Sequence mode: pair i's window image is frame i + 1 of the image-1 array; its shifted copies must follow."""
    h, w, ss, bs = 128, 192, [80, 80], [16, 16]
    rng = np.random.default_rng(3400)
    frames = [make_pair(h, w, 3400 + i, shift=(2, 1))[0] for i in range(5)]
    with bb.Estimator(w, h, ss, bs, chunk_pairs=3, slots=2) as est:
        fields = est.estimate_sequence(frames)
    for i in range(4):
        want, _ = oracle.estimate(frames[i], frames[i + 1], ss, bs, 2)
        assert np.array_equal(fields[i], want), f"pair {i}: " + describe_diff(fields[i], want)
    del rng
