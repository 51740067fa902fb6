"""SURVEY 8f rank 4 on the GPU: the reference's other (uncalled) search, MF::find_min_block (motion_framework.cpp:246-294:
raster scan, L1-distance tie-break, clamped window), as a variant of the search kernel, and the motion-compensated frame of
MF::draw_MVimage (:887-905).  The oracle's restatements are pinned to the reference's own functions in tests/test_oracle.py."""
import numpy as np
import pytest

import blockbasedmotionestimation_b200 as bb
from helpers import dense_to_blocks, blocks_to_dense, describe_diff, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,bs,ss,kind", [(64, 96, 8, 24, "textured"), (64, 64, 16, 32, "textured"), (48, 80, 4, 10, "noise"),
                                            (64, 96, 8, 24, "constant"), (96, 128, 32, 48, "textured"), (40, 56, 2, 6, "textured"),
                                            (128, 192, 16, 80, "textured")])
def test_raster_search_stage(gpu, oracle, h, w, bs, ss, kind):
    rng = np.random.default_rng(h * 7 + w + bs)
    f1, f2 = make_pair(h, w, 40 + bs, shift=(2, -1), kind=kind)
    pred = rng.integers(-12, 13, (h // bs, w // bs, 2)).astype(np.int16)  # some predictions leave the image
    got = gpu.stage_search_raster(f1, f2, bs, ss, pred)
    want = dense_to_blocks(oracle.search_level_raster(f1, f2, bs, ss, blocks_to_dense(pred, bs, h, w)), bs)
    assert np.array_equal(got, want), describe_diff(got, want)


@pytest.mark.parametrize("bs", [2, 8, 16])
def test_compensated_frame_stage(gpu, oracle, bs):
    h, w = 96, 128
    rng = np.random.default_rng(bs)
    _, f2 = make_pair(h, w, 77)
    mv = rng.integers(-20, 21, (h // bs, w // bs, 2)).astype(np.int16)  # border blocks point outside: skipped
    got = gpu.stage_compensate(f2, bs, mv)
    want = oracle.compensate(f2, bs, blocks_to_dense(mv, bs, h, w))
    assert np.array_equal(got, want)


def test_whole_path_with_raster_search_and_mc_frame(oracle):
    """search_variant=1 through the whole hierarchy (padding included), then MF::draw_MVimage on the result."""
    h, w, ss, bs = 122, 186, [24, 24, 24], [8, 8, 8]
    f1, f2 = make_pair(h, w, 91, shift=(3, -2), patches=4)
    want = oracle.estimate_raster(f1, f2, ss, bs, 2)
    mf = bb.MF(f1, f2, ss, bs, 3, search_variant=1)
    got = mf.calcMotionBlockMatching()
    assert np.array_equal(got, want), describe_diff(got, want)
    spiral, _ = oracle.estimate(f1, f2, ss, bs, 2)
    assert not np.array_equal(spiral, want)  # it IS a different search
    # the motion-compensated padded frame, block size 2 (what :205 leaves behind), against the oracle's restatement
    pad2 = oracle.pad_image(f2, mf.padding_x, mf.padding_y)
    mc = mf.draw_MVimage()
    assert np.array_equal(mc, oracle.compensate(pad2, 2, want))
    mf.close()
