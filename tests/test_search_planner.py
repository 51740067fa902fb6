"""Host-side logic of the search kernel that needs no GPU: the geometry planner (which kernel family runs a pyramid level, ring
depth, bands, shared memory) and the multiply-high division constants.  The kernels themselves are covered by the -m gpu tests;
this file pins the decisions DESIGN.md 3.1 describes for the BASELINE configurations and the invariants every plan must keep."""
import random

import pytest

from blockbasedmotionestimation_b200 import api

SMEM_CAP = 200 * 1024         # dynamic shared memory of the funnel-shift kernels
SMEM_CAP_COPIES = 224 * 1024  # kernels over byte-shifted copies (stages four times as large)


def test_division_constants_are_exact():
    rng = random.Random(7)
    divisors = list(range(2, 700)) + [rng.randrange(700, 1 << 16) for _ in range(300)] + [(1 << 16) - 1, 1 << 16, (1 << 20) + 3, (1 << 30) - 1]
    for d in divisors:
        m, s = api.div_magic(d)
        assert 0 < m < (1 << 32) and 0 <= s < 32
        xs = {0, 1, d - 1, d, d + 1, (1 << 31) - 1, (1 << 31) - d, ((1 << 31) - 1) // d * d, ((1 << 31) - 1) // d * d - 1}
        xs |= {rng.randrange(0, 1 << 31) for _ in range(40)} | {k * d + r for k in (1, 2, 3, 1000, 65535) for r in (-1, 0, 1)}
        for x in xs:
            if 0 <= x < (1 << 31):
                assert ((x * m) >> 32) >> s == x // d, (d, x)
    with pytest.raises(api.BbmeError):
        api.div_magic(1)


def test_baseline_configurations_are_planned_as_designed():
    c2 = api.search_geometry(1920, 1088, 16, 80)             # 1080p, 16x16, +-32
    assert c2["planned"] and c2["copies"] and not c2["deep_ring"] and not c2["key64"]
    assert (c2["rows_per_lane"], c2["pitch_words"], c2["stages"], c2["bands"]) == (13, 24, 5, 1)
    assert c2["box_w"] == 96 and c2["box_h"] == 65 + 15 and c2["lanes_per_unit"] == 65 * 5
    assert c2 == api.search_geometry(480, 272, 16, 80)       # coarser levels of the same configuration: the same kernel
    c2_funnel = api.search_geometry(1920, 1088, 16, 80, allow_copies=False)
    assert not c2_funnel["copies"] and c2_funnel["stage_bytes"] * 3 < c2["stage_bytes"] and c2_funnel["smem_bytes"] <= SMEM_CAP

    c1 = api.search_geometry(2336, 1568, 32, 64)             # the reference's defaults on RubberWhale x4: 32x32, +-16
    assert c1["copies"] and c1["deep_ring"] and c1["stages"] == 8 and c1["rows_per_lane"] == 11 and c1["lanes_per_unit"] == 99

    c3 = api.search_geometry(3840, 2176, 8, 136)             # 4K, 8x8, +-64: the copies would leave five stages of five items
    assert c3["planned"] and not c3["copies"] and c3["rows_per_lane"] == 43 and c3["pitch_words"] == 40 and c3["bands"] == 1

    c5 = api.search_geometry(7680, 4352, 16, 272)            # 8K, +-128: 64-bit keys, window wider than one TMA box
    assert c5["key64"] and c5["two_boxes"] and not c5["copies"] and c5["bands"] > 1


def test_what_the_tma_kernel_does_not_cover():
    assert not api.search_geometry(640, 480, 4, 12)["planned"]       # block sizes other than 8 / 16 / 32: generic kernel
    assert not api.search_geometry(640, 480, 64, 96)["planned"]
    assert not api.search_geometry(640, 480, 16, 16)["planned"]      # R = 0
    with pytest.raises(api.BbmeError):
        api.search_geometry(640, 480, 12, 40)                        # not a power of two


@pytest.mark.parametrize("allow_copies", [True, False])
def test_invariants_of_every_plan(allow_copies):
    seen_copies = set()
    for bs in (8, 16, 32):
        for R in list(range(1, 40)) + [48, 56, 64, 80, 100, 112, 128, 150, 180]:
            g = api.search_geometry(4096, 2304, bs, bs + 2 * R, allow_copies)
            if not g["planned"]:
                continue
            n = 2 * R + 1
            words = ((15 + 2 * R) >> 2) + bs // 4 + 1
            assert g["pitch_words"] in (16, 24, 32, 40, 48, 64)
            assert g["box_w"] == 4 * g["pitch_words"] <= 256 and 0 < g["box_h"] <= 256
            if g["two_boxes"]:
                assert g["key64"] and 2 * g["pitch_words"] - 4 >= words
            else:
                assert g["pitch_words"] >= words
            assert g["bands"] * g["segments_per_band"] * g["rows_per_lane"] >= n          # every candidate row has a lane
            assert (g["bands"] - 1) * g["segments_per_band"] * g["rows_per_lane"] < n     # and no band is empty
            assert g["box_h"] == g["segments_per_band"] * g["rows_per_lane"] + bs - 1
            assert g["lanes_per_unit"] == max(32, n * g["segments_per_band"])
            assert g["stages"] in (5, 8, 16) and (g["stages"] > 5) == bool(g["deep_ring"])
            cap = SMEM_CAP_COPIES if g["copies"] else SMEM_CAP
            copies = 4 if g["copies"] else 2 if g["two_boxes"] else 1
            assert g["stage_bytes"] >= copies * g["box_h"] * g["box_w"] + bs * max(bs, 16)
            assert g["stages"] * g["stage_bytes"] <= g["smem_bytes"] <= cap
            if not g["key64"]:   # spiral-rank table: (n + rows per lane) rows of 4 * pitch 16-bit entries behind the ring
                assert g["smem_bytes"] - g["stages"] * g["stage_bytes"] >= (n + g["rows_per_lane"]) * 4 * g["pitch_words"] * 2
                assert n * n <= (1 << (14 if bs == 32 else 16)) - 1
            if g["copies"]:
                assert allow_copies and not g["key64"] and g["pitch_words"] in (24, 40)
                seen_copies.add(g["pitch_words"])
    assert seen_copies == ({24, 40} if allow_copies else set())
