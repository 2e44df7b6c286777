"""Oracle restatements of the un-vendored OpenCV pieces and of the MIP / BoW / writer code, against independent
implementations (cv2 4.13 where it exists, numpy otherwise) and against hand-derived known answers."""
import numpy as np
import pytest

import mofreak_b200
from tests import util


def test_integral_matches_numpy(orc):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    want = np.zeros((38, 54), np.int64)
    want[1:, 1:] = img.astype(np.int64).cumsum(0).cumsum(1)
    assert np.array_equal(orc.integral(img), want.astype(np.int32))


def test_integral_and_absdiff_match_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (60, 80), dtype=np.uint8); b = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    from tests.oracle_binding import absdiff
    assert np.array_equal(absdiff(a, b), cv2.absdiff(a, b))
    assert np.array_equal(orc.integral(a), cv2.integral(a))


def test_resize_model_matches_cv2(orc):
    """cv::resize(INTER_LINEAR, 8U) is not vendored by the reference; the oracle's fixed-point model is pinned here
    against OpenCV 4.13 for every ROI side MoFREAK can produce (9 .. 110) plus the degenerate small ones."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for s in list(range(2, 112)):
        for _ in range(2):
            roi = rng.integers(0, 256, (s, s), dtype=np.uint8)
            want = cv2.resize(roi, (19, 19), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(orc.resize19(roi), want), "side %d" % s


def test_resize_identity_and_half(orc):
    rng = np.random.default_rng(3)
    roi = rng.integers(0, 256, (19, 19), dtype=np.uint8)
    assert np.array_equal(orc.resize19(roi), roi)
    roi = rng.integers(0, 256, (38, 38), dtype=np.uint8).astype(np.int32)
    want = (roi[0::2, 0::2] + roi[0::2, 1::2] + roi[1::2, 0::2] + roi[1::2, 1::2] + 2) >> 2
    assert np.array_equal(orc.resize19(roi.astype(np.uint8)), want.astype(np.uint8))


def test_mip_byte_known_answers(orc):
    a = np.full((19, 19), 100, np.uint8)
    assert orc.mip_byte(a, a, 9, 9) == 0                       # identical patches: every SSD is 0
    b = np.full((19, 19), 94, np.uint8)
    assert orc.mip_byte(a, b, 9, 9) == 0xFF                    # 9 * 6^2 = 324 > 288 for all eight shifts
    b = np.full((19, 19), 95, np.uint8)
    assert orc.mip_byte(a, b, 9, 9) == 0                       # 9 * 5^2 = 225 <= 288
    # the 3x3 "patch" is 9 consecutive bytes of the flat buffer (the reference's stride quirk):
    # change only the pixel 3 columns to the right of the patch's top-left corner in the previous frame
    b = a.copy()
    x, y = 9, 9
    flat = b.reshape(-1)
    flat[(y - 1) * 19 + (x - 4 - 1) + 3] = 0                   # inside shift 0's (-4, 0) 9-byte run, outside its 3x3 block
    assert orc.mip_byte(a, b, x, y) & 1 == 1                   # 100^2 > 288: bit 0 set although the 3x3 block is equal


def test_mip_extract_uses_truncated_coordinates_and_gray_frames(orc):
    rng = np.random.default_rng(4)
    cur = rng.integers(0, 256, (120, 160), dtype=np.uint8); prev = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    size, x, y = 13.4, 60, 50
    side, half = int(np.ceil(size)), int(size) // 2
    a = orc.resize19(cur[y - half:y - half + side, x - half:x - half + side])
    b = orc.resize19(prev[y - half:y - half + side, x - half:x - half + side])
    centres = [(5, 5), (5, 9), (5, 13), (9, 5), (9, 13), (13, 5), (13, 9), (13, 13)]
    want = [orc.mip_byte(a, b, cx, cy) for cx, cy in centres]
    assert list(orc.mip_extract(cur, prev, size, x, y)) == want


def test_hamming_and_brute_force(orc):
    for a in (0, 1, 37, 128, 255):
        for b in (0, 2, 77, 254, 255):
            assert orc.hamming_u8(a, b) == bin(a ^ b).count("1")
    rng = np.random.default_rng(5)
    cb = rng.integers(0, 256, (50, 16), dtype=np.uint8)
    cb[31] = cb[7]                                            # duplicate codeword: the lower index must win
    desc = rng.integers(0, 256, (300, 16), dtype=np.uint8)
    desc[:50] = cb
    dist = np.unpackbits(desc[:, None, :] ^ cb[None, :, :], axis=2).sum(2)
    assert np.array_equal(orc.bow_assign(desc, cb), dist.argmin(1).astype(np.int32))
    assert orc.bow_assign(cb[31:32], cb)[0] == 7


def test_load_clusters_off_by_one_quirk(orc):
    lines = [" ".join(str((7 * r + c) % 256) for c in range(16)) + " " for r in range(5)]
    got = orc.load_clusters("\n".join(lines) + "\n")
    assert got.shape == (5, 16)
    for k in range(4):                                        # codeword k is file line k+2; line 1 is never used
        assert list(got[k]) == [int(v) for v in lines[k + 1].split()]
    assert list(got[4]) == [int(lines[4].split()[-1])] * 16   # the last one repeats the final parsed value


def test_build_histogram(orc):
    rng = np.random.default_rng(6)
    cb = rng.integers(0, 256, (20, 16), dtype=np.uint8)
    rows = np.zeros(40, util.ROW_DTYPE)
    rows["x"] = rng.random(40) * 100; rows["y"] = rng.random(40) * 100; rows["scale"] = 12; rows["frame"] = 4
    rows["appearance"] = rng.integers(0, 256, (40, 8)); rows["motion"] = rng.integers(0, 256, (40, 8))
    text = orc.format_mofreak(rows)
    ok, hist = orc.build_histogram(text, cb)
    desc = np.concatenate([rows["appearance"], rows["motion"]], 1)
    counts = np.bincount(orc.bow_assign(desc, cb), minlength=20).astype(np.float32)
    assert ok and np.array_equal(hist, counts / np.float32(counts.sum()))
    ok, hist = orc.build_histogram("", cb)
    assert not ok and not hist.any()


def test_mofreak_text_format(orc):
    rows = np.zeros(2, util.ROW_DTYPE)
    rows[0] = (66.47868, 73.52687, 4, 10.404844, [255, 252, 187, 255, 182, 245, 255, 212], [116, 255, 255, 124, 255, 191, 255, 255])
    rows[1] = (74.0, 100000.5, 98, 8.4, [0] * 8, [1] * 8)
    text = orc.format_mofreak(rows)
    lines = text.split("\n")
    assert lines[0] == "66.4787 73.5269 4 10.4048 0 0 255 252 187 255 182 245 255 212 116 255 255 124 255 191 255 255 "
    assert lines[1] == "74 100000 98 8.4 0 0 0 0 0 0 0 0 0 0 1 1 1 1 1 1 1 1 "        # %g: 6 significant digits
    assert lines[2] == "" and len(lines) == 3


def test_freak_tables_structure(orc):
    sizes = [orc.freak_pattern_size(s) for s in range(64)]
    assert sizes[0] == 23 and sizes[4] == 28 and sizes[8] == 33 and sizes[63] == 339      # SURVEY App. A.3 / probe P5
    assert sizes == sorted(sizes)
    assert orc.freak_scale_index(7.0) == 0 and orc.freak_scale_index(1.0) == 0 and orc.freak_scale_index(1e6) == 63
    assert orc.freak_scale_index(14.0) == 16                                             # one octave = 16 scales


def test_freak_descriptor_layout_and_invariants(orc):
    """No FREAK implementation exists in this container (parity unpinned, see oracle/freak_restate.cpp); these are
    the structural properties the published algorithm guarantees."""
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    kps = np.array([[160.3, 120.7, 12.0, 50, 0], [5.0, 5.0, 12.0, 50, 0], [100.0, 100.0, 30.0, 40, 2], [300.0, 200.0, 9.0, 31, 0]], np.float32)
    kept, desc, theta = orc.freak_compute(img, kps)
    assert len(kept) == 2 and kept[0][0] == np.float32(160.3) and kept[1][0] == np.float32(100.0)   # border drops keep order
    assert desc.shape == (2, 64) and ((0 <= theta) & (theta < 256)).all()
    # a constant image: every pair test is val >= val -> all ones; orientation sums are 0 -> atan2(0,0) = 0
    flat = np.full((240, 320), 77, np.uint8)
    kept, desc, theta = orc.freak_compute(flat, kps[:1])
    assert (desc == 255).all() and theta[0] == 0


def test_frame_schedule_and_labels(orc):
    clip = mofreak_b200.synth_clip(160, 120, 12, 1)
    rows = orc.extract_clip(clip)
    assert len(rows) > 0
    assert rows["frame"].min() >= 4 and rows["frame"].max() <= 10          # frame c in 5..11 is labelled c-1
    assert np.all(np.diff(rows["frame"]) >= 0)
    assert len(orc.extract_clip(clip[:5])) == 0 and len(orc.extract_clip(clip[:3])) == 0
    one = orc.extract_clip(clip[:6])
    assert rows_equal_prefix(one, rows)


def rows_equal_prefix(a, b):
    return a.tobytes() == b[:len(a)].tobytes()


def test_as_written_mode_is_identical(orc):
    clip = mofreak_b200.synth_clip(160, 120, 8, 2)
    assert util.rows_equal(orc.extract_clip(clip, as_written=True), orc.extract_clip(clip, as_written=False))
