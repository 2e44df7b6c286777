"""Pins oracle/brisk_port.cpp (our restatement of the BRISK scale-space detector) against the reference's OWN
object code (oracle/_ref: reference brisk.cpp + oast9_16*.cc + agast5_8*.cc compiled unchanged).
The reference ships no tests or golden vectors; its compiled code is the strongest anchor there is."""
import numpy as np
import pytest

import mofreak_b200
from tests import util


def need_ref(orc):
    if not orc.has_reference:
        pytest.skip("oracle/_ref (reference object code) not built here")


@pytest.mark.parametrize("w", [16, 31, 32, 47, 48, 53, 80, 106, 160, 212, 240, 320, 333, 360, 480, 720])
def test_halfsample_matches_reference(orc, w):
    need_ref(orc)
    rng = np.random.default_rng(w)
    for h in (14, 15):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(orc.halfsample(img), orc.halfsample(img, ref=True))


@pytest.mark.parametrize("w", [15, 16, 29, 30, 45, 47, 48, 160, 161, 212, 320, 321, 322, 700, 720])
def test_twothirdsample_matches_reference(orc, w):
    need_ref(orc)
    rng = np.random.default_rng(w)
    for h in (12, 13, 14):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(orc.twothirdsample(img), orc.twothirdsample(img, ref=True))


def test_oast_score_closed_form_matches_reference_bisection(orc):
    need_ref(orc)
    rng = np.random.default_rng(1)
    # mixtures of flat, noisy and high-contrast patches so that every score range occurs
    for trial in range(6):
        img = rng.integers(0, 256, (40, 40), dtype=np.uint8)
        if trial % 2:
            img = (img // 64 * 64 + rng.integers(0, 8, img.shape)).astype(np.uint8)
        for y in range(3, 37, 3):
            for x in range(3, 37, 2):
                assert orc.oast_score(img, x, y) == orc.ref_oast_score(img, x, y, 0)
                assert orc.agast58_score(img, x, y) == orc.ref_agast58_score(img, x, y, 0)
                # cornerScore started from b0 returns max(b0, closed form)
                assert max(29, orc.oast_score(img, x, y)) == orc.ref_oast_score(img, x, y, 29)


def test_detection_set_is_score_at_least_threshold(orc):
    need_ref(orc)
    clip = mofreak_b200.synth_clip(160, 120, 7, 4)
    for _, d in util.diff_frames(clip):
        hits = orc.ref_oast_detect(d, 30)
        S = orc.score_map(d)
        ys, xs = np.nonzero(S >= 30)
        assert np.array_equal(hits, np.stack([xs, ys], 1).astype(np.int32))     # same set, same raster order


@pytest.mark.parametrize("shape,seeds", [((160, 120), range(1, 9)), ((320, 240), (1, 2)), ((212, 163), (3,)),
                                         ((367, 251), (5,)), ((720, 576), (1,))])
def test_port_equals_reference_on_synthetic_clips(orc, shape, seeds):
    need_ref(orc)
    w, h = shape
    total = 0
    for seed in seeds:
        clip = mofreak_b200.synth_clip(w, h, 9 if w < 700 else 7, seed)
        for _, d in util.diff_frames(clip):
            ref = orc.brisk_ref(d)
            assert util.kp_equal(orc.brisk_port(d), ref)
            # the least favourable schedule of the tie fixed point must not change anything
            assert util.kp_equal(orc.brisk_port(d, jacobi_reverse=True), ref)
            total += len(ref)
    assert total > 0


def test_port_equals_reference_on_tie_heavy_images(orc):
    """Quantised random images produce plateaus of equal scores, i.e. many isMax2D tie-breaks whose outcome
    depends on the reference's memoisation order."""
    need_ref(orc)
    rng = np.random.default_rng(11)
    for trial in range(4):
        img = (rng.integers(0, 4, (96, 128)) * 80).astype(np.uint8)
        if trial >= 2:
            img = np.kron(rng.integers(0, 3, (24, 32)) * 100, np.ones((4, 4))).astype(np.uint8)
        ref = orc.brisk_ref(img)
        assert len(ref) > 0
        assert util.kp_equal(orc.brisk_port(img), ref)
        assert util.kp_equal(orc.brisk_port(img, jacobi_reverse=True), ref)


def test_final_raw_score_map_is_masked_closed_form(orc):
    """Every value the reference leaves in its memoised score maps is either 0 or the closed-form score."""
    need_ref(orc)
    clip = mofreak_b200.synth_clip(160, 120, 6, 9)
    d = next(util.diff_frames(clip))[1]
    for layer in range(6):
        img, raw = orc.ref_dump(d, layer)
        S = orc.score_map(img)
        assert np.all((raw == 0) | (raw == S))
        assert np.array_equal(raw >= 30, S >= 30)


def test_blank_and_tiny_inputs(orc):
    assert len(orc.brisk_port(np.zeros((120, 160), np.uint8))) == 0
    if orc.has_reference:
        assert len(orc.brisk_ref(np.zeros((120, 160), np.uint8))) == 0
    assert len(orc.brisk_port(np.full((120, 160), 255, np.uint8))) == 0
