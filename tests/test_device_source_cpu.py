"""The product's DEVICE source (mofreak_b200/csrc/*_device.cuh), compiled as plain C++ by tests/host_harness.cpp and
driven with the kernels' phase structure, against the oracle — no GPU needed."""
import numpy as np
import pytest

import mofreak_b200
from tests import util


def test_device_atan2f_is_bit_identical_to_host_libm():
    # FREAK's orientation bin comes from atan2f of two integer sums; the device evaluates glibc's algorithm with
    # IEEE float ops, so the angle is bit-identical (tolerance 0 ulp) and no emitted bit can flip.
    assert util.harness().hh_atan2f_sweep(400, 3000000, 9) == 0


def test_theta_index_matches_libm_expression():
    L = util.harness()
    rng = np.random.default_rng(0)
    for d0, d1 in rng.integers(-4000, 4000, (20000, 2)):
        ang = np.float32(np.float64(np.arctan2(np.float32(d1), np.float32(d0), dtype=np.float32)) * (180.0 / 3.1415926535897932384626433832795))
        want = int(np.float64(np.float32(256.0) * ang) * (1 / 360.0) + 0.5)
        want = want + 256 if want < 0 else want
        want = want - 256 if want >= 256 else want
        # numpy's float32 arctan2 is the same libm atan2f
        assert L.hh_theta_index(int(d0), int(d1)) == want


def test_scale_threshold_table_equals_libm_expression(orc):
    L = util.harness()
    rng = np.random.default_rng(1)
    sizes = np.concatenate([rng.uniform(1.0, 200.0, 200000), np.float32(7.0) * np.float32(2.0) ** (np.arange(0, 64) / 16.0)]).astype(np.float32)
    for s in sizes[:60000]:
        a = L.hh_scale_index_table(float(s))
        assert a == L.hh_scale_index_direct(float(s))
    for s in sizes[::97]:
        assert L.hh_scale_index_table(float(s)) == orc.freak_scale_index(float(s))
    # step edges: one float below every threshold the index is one lower
    for k in range(1, 64):
        lo, hi = np.float32(1.0), np.float32(4096.0)
        while np.nextafter(lo, hi) < hi:
            mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
            if L.hh_scale_index_direct(float(mid)) >= k:
                hi = mid
            else:
                lo = mid
        assert L.hh_scale_index_table(float(hi)) == k and L.hh_scale_index_table(float(lo)) == k - 1


@pytest.mark.parametrize("shape,seed", [((160, 120), 1), ((160, 120), 6), ((320, 240), 2), ((212, 163), 3), ((106, 97), 4)])
def test_device_detector_source_equals_oracle(orc, shape, seed):
    w, h = shape
    clip = mofreak_b200.synth_clip(w, h, 8, seed)
    for _, d in util.diff_frames(clip):
        want = orc.brisk(d)
        got = util.harness_detect(d)
        assert util.kp_equal(got[:, :5].copy(), want)
        assert util.kp_equal(util.harness_detect(d, reverse_rounds=True)[:, :5].copy(), want)
        # the keep flag is FREAK's border filter
        kept, _, _ = orc.freak_compute(d, want)
        assert int(got[:, 5].sum()) == len(kept)


def test_device_detector_source_on_tie_heavy_images(orc):
    rng = np.random.default_rng(12)
    for trial in range(3):
        img = (rng.integers(0, 4, (96, 128)) * 80).astype(np.uint8)
        want = orc.brisk(img)
        assert len(want) > 0
        assert util.kp_equal(util.harness_detect(img)[:, :5].copy(), want)
        assert util.kp_equal(util.harness_detect(img, reverse_rounds=True)[:, :5].copy(), want)


@pytest.mark.parametrize("shape,seed", [((320, 240), 2), ((212, 163), 3), ((131, 70), 8)])
def test_early_removal_of_tile_interior_losers_changes_no_keypoint(orc, shape, seed):
    """k_scores drops a detector hit that has a strictly larger score among its 8 neighbours when the hit lies inside the
    1-pixel frame of its 64 x 32 tile (status 0, never a candidate).  Same keypoints as the reference, also with phase B's
    rounds reversed and on tie-heavy images, with the status map starting from stale bytes."""
    w, h = shape
    clip = mofreak_b200.synth_clip(w, h, 7, seed)
    for _, d in util.diff_frames(clip):
        want = orc.brisk(d)
        for rev in (False, True):
            assert util.kp_equal(util.harness_detect(d, reverse_rounds=rev, prune_tile_losers=True)[:, :5].copy(), want)
    rng = np.random.default_rng(12 + seed)
    img = (rng.integers(0, 4, (96, 160)) * 80).astype(np.uint8)
    want = orc.brisk(img)
    assert len(want) > 0
    for rev in (False, True):
        assert util.kp_equal(util.harness_detect(img, reverse_rounds=rev, prune_tile_losers=True)[:, :5].copy(), want)


def test_device_layers_equal_oracle(orc):
    import ctypes as C
    L = util.harness()
    clip = mofreak_b200.synth_clip(212, 163, 6, 5)
    d = next(util.diff_frames(clip))[1]
    want = [d, orc.twothirdsample(d), orc.halfsample(d)]
    want.append(orc.halfsample(want[1])); want.append(orc.halfsample(want[2])); want.append(orc.halfsample(want[3]))
    for layer in range(6):
        img = np.zeros(d.size, np.uint8); sc = np.zeros(d.size, np.uint8); dims = (C.c_int * 2)()
        L.hh_layer(d.ctypes.data_as(C.c_void_p), 212, 163, layer, img.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), dims)
        n = dims[0] * dims[1]
        assert (dims[1], dims[0]) == want[layer].shape
        assert np.array_equal(img[:n].reshape(want[layer].shape), want[layer])
        assert np.array_equal(sc[:n].reshape(want[layer].shape), orc.score_map(want[layer]))


def test_device_row_source_equals_oracle(orc):
    clip = mofreak_b200.synth_clip(320, 240, 8, 4)
    got = util.harness_extract_clip(clip)
    want = orc.extract_clip(clip)
    assert len(want) > 100
    assert util.rows_equal(got, want), util.describe_diff(got, want)
