"""GPU parity proper: everything goes through the C ABI (ctypes) and is compared with the CPU oracle on the same
seeded inputs and with the committed golden fixtures; full-size cases use size-independent properties.
Bar: bit-exact — keypoint coordinates are float32 but must be IDENTICAL (tolerance 0), as are all descriptor
bytes and bag-of-words indices."""
import os
import subprocess
import sys

import numpy as np
import pytest

import mofreak_b200
from mofreak_b200 import sharding
from tests import util

pytestmark = pytest.mark.gpu

CASES = util.golden_cases()


@pytest.fixture(scope="module")
def lib_loaded():
    import os
    L = mofreak_b200.lib()
    assert os.path.samefile(mofreak_b200.LIB_PATH, os.path.join(os.path.dirname(mofreak_b200.__file__), "libmofreak_b200.so"))
    return L


def run_clip(clip, max_pairs=8, store=None):
    n, h, w = clip.shape
    with mofreak_b200.Context(w, h, store or max(16, n), max_pairs) as ctx:
        rows = ctx.extract_clip(clip)
        assert ctx.launch_count() > 0 or n <= 5
    return rows


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_rows_bit_exact(lib_loaded, orc, name):
    c = CASES[name]
    clip = mofreak_b200.synth_clip(c["width"], c["height"], c["frames"], c["seed"])
    got = run_clip(clip, max_pairs=16)
    want = util.golden_rows(name)
    assert util.rows_equal(got, want), util.describe_diff(got, want)
    assert orc.format_mofreak(got) == util.golden_text(name)


@pytest.mark.parametrize("shape,frames,seed", [((160, 120), 14, 11), ((320, 240), 9, 12), ((720, 576), 7, 13),
                                               ((106, 97), 9, 14), ((367, 251), 8, 15), ((48, 48), 8, 16), ((1280, 720), 6, 17),
                                               ((1920, 1080), 6, 18)])
def test_rows_equal_oracle(orc, shape, frames, seed):
    w, h = shape
    clip = mofreak_b200.synth_clip(w, h, frames, seed)
    got = run_clip(clip)
    want = orc.extract_clip(clip)
    assert util.rows_equal(got, want), util.describe_diff(got, want)


def blob_clip(w, h, n, seed, zoom=16, step=3, mix=0.5):
    """Smooth large-scale structure (a random pattern upsampled `zoom` x with cubic interpolation, drifting `step` px per frame)
    mixed with the moving-squares clip: its difference images have corners at EVERY pyramid scale, so keypoints of all six
    layers, large FREAK patterns and MIP regions wider than the staged 32 px occur in numbers."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // zoom + 8, w // zoom + 8)).astype(np.float64)
    big = np.clip(ndimage.zoom(base, zoom, order=3), 0, 255)
    fine = mofreak_b200.synth_clip(w, h, n, seed).astype(np.float64)
    out = np.empty((n, h, w), np.uint8)
    for t in range(n):
        o = step * t
        out[t] = np.clip(mix * big[o:o + h, o:o + w] + (1 - mix) * fine[t], 0, 255).astype(np.uint8)
    return out


def coverage(orc, clip):
    """What a clip exercises, from the oracle's side: keypoints per pyramid layer, FREAK scale indices of the rows, MIP region
    sides on either side of the kernel's staging limit (kernels_desc.cu ROI_MAX = 32), orientation bins."""
    layers = np.zeros(6, int)
    sidx, thetas, staged, unstaged = set(), set(), 0, 0
    for _, d in util.diff_frames(clip):
        kps = orc.brisk(d)
        layers += np.bincount(kps[:, 4].astype(int), minlength=6)
        kept, _, theta = orc.freak_compute(d, kps)
        sidx |= {orc.freak_scale_index(float(sz)) for sz in kept[:, 2]}
        thetas |= set(theta.tolist())
        sides = np.ceil(kept[:, 2])
        staged += int((sides <= 32).sum()); unstaged += int((sides > 32).sum())
    return layers, sidx, thetas, staged, unstaged


@pytest.mark.parametrize("shape,seed,blobs", [((720, 576), 13, False), ((1280, 720), 17, False), ((720, 576), 5, True), ((366, 250), 6, True)])
def test_parity_cases_cover_every_layer_scale_and_mip_path(orc, shape, seed, blobs):
    """Rows equal the oracle's AND the inputs are known to exercise the paths: every pyramid layer emits keypoints, the
    FREAK scale index takes >= 20 distinct values, MIP regions are staged (side <= 32) and unstaged (> 32), and for widths
    that are not a multiple of 4 the synchronous (non-cp.async) staging path runs."""
    w, h = shape
    clip = blob_clip(w, h, 7, seed) if blobs else mofreak_b200.synth_clip(w, h, 7, seed)
    layers, sidx, thetas, staged, unstaged = coverage(orc, clip)
    assert (layers > 0).all(), "keypoints per layer %s" % layers
    assert len(sidx) >= 20 and len(thetas) >= 64
    assert staged > 0 and unstaged > 0
    if blobs:
        assert layers[3:].sum() > layers.sum() // 4          # the upper layers carry a large share here
        assert unstaged >= (100 if w >= 720 else 5)
    got = run_clip(clip)
    want = orc.extract_clip(clip)
    assert len(want) > 100 and util.rows_equal(got, want), util.describe_diff(got, want)


def test_stages_equal_oracle(orc):
    w, h, n = 212, 163, 8
    clip = mofreak_b200.synth_clip(w, h, n, 21)
    with mofreak_b200.Context(w, h, 16, 8) as ctx:
        ctx.upload_frames(clip, 0)
        cur = np.arange(5, n)
        ctx.extract_pairs(cur, cur - 5, cur - 1)
        for pair, (c, d) in enumerate(util.diff_frames(clip)):
            layers = [d, orc.twothirdsample(d), orc.halfsample(d)]
            layers.append(orc.halfsample(layers[1])); layers.append(orc.halfsample(layers[2])); layers.append(orc.halfsample(layers[3]))
            for li, want in enumerate(layers):
                assert np.array_equal(ctx.debug_layer(0, li, pair), want), "pyramid layer %d" % li
                assert np.array_equal(ctx.debug_layer(1, li, pair), orc.score_map(want)), "score map layer %d" % li
            assert util.kp_equal(ctx.debug_keypoints(pair), orc.brisk(d))
        rows, per_pair = ctx.fetch_rows(100000, n - 5)
        assert per_pair.sum() == len(rows)
        want = orc.extract_clip(clip)
        assert util.rows_equal(rows, want)
        assert [int((want["frame"] == c - 1).sum()) for c in range(5, n)] == list(per_pair)


def test_tie_heavy_frames(orc):
    """Quantised noise: thousands of equal-score neighbours, the order-dependent part of the reference's NMS."""
    rng = np.random.default_rng(31)
    w, h = 128, 96
    frames = np.zeros((7, h, w), np.uint8)
    frames[5] = (rng.integers(0, 4, (h, w)) * 80).astype(np.uint8)
    frames[6] = np.kron(rng.integers(0, 3, (h // 4, w // 4)) * 100, np.ones((4, 4))).astype(np.uint8)
    with mofreak_b200.Context(w, h, 16, 4, cand_cap_per_pair=w * h * 2, row_cap_per_pair=w * h) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.extract_pairs([5, 6], [0, 1], [4, 5])
        for pair, c in enumerate((5, 6)):
            want = orc.brisk(frames[c])          # prev frame is zero, so the diff image is the frame itself
            assert len(want) > 50
            assert util.kp_equal(ctx.debug_keypoints(pair), want)
        rows, _ = ctx.fetch_rows(200000)
    assert util.rows_equal(rows, orc.extract_clip(frames))


def test_empty_ragged_and_blank_inputs(orc):
    clip = mofreak_b200.synth_clip(160, 120, 9, 41)
    with mofreak_b200.Context(160, 120, 16, 8) as ctx:
        for n in (0, 1, 5):
            assert len(ctx.extract_clip(clip[:n])) == 0                      # no frame pair exists yet
        assert util.rows_equal(ctx.extract_clip(clip[:6]), orc.extract_clip(clip[:6]))
        blank = np.zeros((8, 120, 160), np.uint8)
        assert len(ctx.extract_clip(blank)) == 0
        still = np.repeat(clip[:1], 8, axis=0)                               # no motion -> zero difference image
        assert len(ctx.extract_clip(still)) == 0
        ctx.extract_pairs([], [], [])
        rows, per = ctx.fetch_rows(10, 0)
        assert len(rows) == 0


def test_batching_and_windowing_do_not_change_rows(orc):
    clip = mofreak_b200.synth_clip(160, 120, 40, 51)
    want = orc.extract_clip(clip)
    for max_pairs, store in ((1, 6), (3, 7), (7, 13), (64, 64)):
        got = run_clip(clip, max_pairs=max_pairs, store=store)      # store < clip length: ring reuse with the 5-frame look-back
        assert util.rows_equal(got, want), "max_pairs=%d store=%d: %s" % (max_pairs, store, util.describe_diff(got, want))


def test_ring_wraparound_and_arbitrary_pair_order(orc):
    clip = mofreak_b200.synth_clip(160, 120, 12, 61)
    with mofreak_b200.Context(160, 120, 16, 5) as ctx:
        ctx.upload_frames(clip, 10)                                  # slots 10..15, 0..5
        cur = np.array([11, 5, 8, 8], np.int32)
        ctx.extract_pairs((cur + 10) % 16, (cur - 5 + 10) % 16, cur - 1)
        rows, per = ctx.fetch_rows(100000, 4)
    want = orc.extract_clip(clip)
    parts = [want[want["frame"] == c - 1] for c in cur]
    assert list(per) == [len(p) for p in parts]
    assert util.rows_equal(rows, np.concatenate(parts))


def test_uploads_queued_ahead_of_their_extractions(orc):
    """A launch sequence waits only for the uploads that carry its frames (per-upload events); records that fall out of the
    event ring (more than 256 uploads in flight) fold into an in-order event.  All frames are uploaded one by one FIRST - more
    uploads than the ring holds, the ring wrapping twice over its slots - then every pair is extracted."""
    clip = mofreak_b200.synth_clip(160, 120, 40, 91)
    want = orc.extract_clip(clip)
    with mofreak_b200.Context(160, 120, 40, 8) as ctx:
        for rep in range(8):                                         # 320 single-frame uploads, the last pass holds the clip
            src = clip if rep == 7 else np.roll(clip, rep + 1, axis=0)
            for k in range(40):
                ctx.upload_frames(src[k:k + 1], k)
        cur = np.arange(5, 40, dtype=np.int32)
        for a in range(0, len(cur), 5):                              # several extraction calls after all the uploads
            ctx.extract_pairs(cur[a:a + 5], cur[a:a + 5] - 5, cur[a:a + 5] - 1)
        rows, _ = ctx.fetch_rows(200000)
    assert util.rows_equal(rows, want), util.describe_diff(rows, want)


def test_capacity_overflow_is_reported_not_silent():
    clip = mofreak_b200.synth_clip(320, 240, 7, 71)
    with mofreak_b200.Context(320, 240, 16, 2, cand_cap_per_pair=64, row_cap_per_pair=8) as ctx:
        with pytest.raises(mofreak_b200.MofreakError) as e:
            ctx.extract_clip(clip)
        assert e.value.code == mofreak_b200.capi.MFK_ERR_BUDGET


def test_small_caller_buffers_and_budget_overflow_are_told_apart(orc):
    """MFK_ERR_CAPACITY = the caller's buffer is too small: sizes are reported and the accumulation is kept, so a re-read
    with larger buffers gives the full result.  MFK_ERR_BUDGET = a device budget overflowed: no sizes, nothing to re-read."""
    import ctypes as C
    clip = mofreak_b200.synth_clip(320, 240, 9, 72)
    want = orc.extract_clip(clip)
    assert len(want) > 50
    L = mofreak_b200.lib()
    with mofreak_b200.Context(320, 240, 16, 4) as ctx:
        rows = np.empty(8, mofreak_b200.ROW_DTYPE)
        text = np.empty(64, np.uint8)
        n, nb = C.c_size_t(), C.c_size_t()
        rc = L.mfk_extract_clip_text(ctx._h, clip.ctypes.data_as(C.c_void_p), 320, 9, text.ctypes.data_as(C.c_void_p), text.size, C.byref(nb),
                                     rows.ctypes.data_as(C.c_void_p), rows.size, C.byref(n))
        assert rc == mofreak_b200.capi.MFK_ERR_CAPACITY and n.value == len(want) and nb.value > 64
        text_full = ctx.rows_text()
        got, _ = ctx.fetch_rows(n.value)
        assert util.rows_equal(got, want)
        assert text_full.decode() == orc.format_mofreak(want) and len(text_full) == nb.value
    with mofreak_b200.Context(320, 240, 16, 4, cand_cap_per_pair=64, row_cap_per_pair=8) as ctx:
        rows = np.empty(100000, mofreak_b200.ROW_DTYPE)
        text = np.empty(1 << 22, np.uint8)
        n, nb = C.c_size_t(7), C.c_size_t(7)
        rc = L.mfk_extract_clip_text(ctx._h, clip.ctypes.data_as(C.c_void_p), 320, 9, text.ctypes.data_as(C.c_void_p), text.size, C.byref(nb),
                                     rows.ctypes.data_as(C.c_void_p), rows.size, C.byref(n))
        assert rc == mofreak_b200.capi.MFK_ERR_BUDGET and n.value == 0 and nb.value == 0


def test_negative_slots_are_rejected():
    with mofreak_b200.Context(160, 120, 16, 4) as ctx:
        ctx.upload_frames(mofreak_b200.synth_clip(160, 120, 8, 3), 0)
        with pytest.raises(mofreak_b200.MofreakError) as e:
            ctx.extract_pairs(np.array([5, -1], np.int32), np.array([0, 1], np.int32), np.array([4, 5], np.int32))
        assert e.value.code == mofreak_b200.capi.MFK_ERR_ARG
        ctx.extract_pairs(np.array([5], np.int32), np.array([0], np.int32), np.array([4], np.int32))     # the context still works
        rows, _ = ctx.fetch_rows(10000)
        assert len(rows) >= 0


def test_mip_facade(orc):
    clip = mofreak_b200.synth_clip(320, 240, 6, 81)
    rng = np.random.default_rng(82)
    kps = np.stack([rng.uniform(60, 260, 200), rng.uniform(60, 180, 200), rng.uniform(8.4, 100, 200)], 1).astype(np.float32)
    with mofreak_b200.Context(320, 240, 8, 1) as ctx:
        ctx.upload_frames(clip, 0)
        got = ctx.mip_describe(5, 0, kps)
        want = np.stack([orc.mip_extract(clip[5], clip[0], float(s), int(x), int(y)) for x, y, s in kps])
        assert np.array_equal(got, want)
        with pytest.raises(mofreak_b200.MofreakError):
            ctx.mip_describe(5, 0, np.array([[3.0, 3.0, 40.0]], np.float32))     # ROI leaves the frame


@pytest.mark.parametrize("K", [1, 2, 600, 1000, 10100, 20000])
def test_bow_assignment_and_histogram(orc, K):
    rng = np.random.default_rng(K)
    cb = rng.integers(0, 256, (K, 16), dtype=np.uint8)
    if K > 20:
        cb[K // 2] = cb[3]; cb[K - 1] = cb[3]                      # duplicates: the lowest index must win
    desc = rng.integers(0, 256, (5000, 16), dtype=np.uint8)
    desc[:min(K, 600)] = cb[:min(K, 600)]
    desc[700] = cb[K - 1]
    with mofreak_b200.Context(160, 120, 8, 1) as ctx:
        ctx.bow_set_codebook(cb)
        idx = ctx.bow_assign(desc)
        want = orc.bow_assign(desc[:800], cb) if K > 5000 else orc.bow_assign(desc, cb)
        assert np.array_equal(idx[:len(want)], want)
        dist = lambda d, k: int(np.unpackbits(d ^ cb[k]).sum())    # property check on the rest: it is an arg-min
        for i in range(len(want), len(desc), 97):
            best = dist(desc[i], idx[i])
            sample = rng.integers(0, K, 200)
            assert all(dist(desc[i], int(k)) > best or (dist(desc[i], int(k)) == best and k >= idx[i]) for k in sample)
        assert np.array_equal(ctx.bow_histogram(desc), np.bincount(idx, minlength=K))
        assert not ctx.bow_histogram(desc[:0]).any()
        assert len(ctx.bow_assign(desc[:0])) == 0


def test_bow_on_device_rows(orc):
    clip = mofreak_b200.synth_clip(320, 240, 9, 91)
    rng = np.random.default_rng(92)
    cb = rng.integers(0, 256, (600, 16), dtype=np.uint8)
    with mofreak_b200.Context(320, 240, 16, 4) as ctx:
        ctx.bow_set_codebook(cb)
        ctx.upload_frames(clip, 0)
        cur = np.arange(5, 9)
        ctx.extract_pairs(cur, cur - 5, cur - 1)
        counts = ctx.bow_histogram_rows()
        rows, _ = ctx.fetch_rows(100000)
    desc = np.concatenate([rows["appearance"], rows["motion"]], 1)
    assert np.array_equal(counts, np.bincount(orc.bow_assign(desc, cb), minlength=600))
    # the text path of the reference (buildHistogram on the .mofreak file) gives the same normalised histogram
    ok, hist = orc.build_histogram(orc.format_mofreak(rows), cb)
    assert ok and np.array_equal(hist, counts.astype(np.float32) / np.float32(counts.sum()))


def test_full_size_properties_trecvid_shape():
    """BASELINE configs[3] at bench size is too big for the CPU oracle; check what must hold at any size:
    determinism, independence from batch composition, frame-range sharding with the 5-frame halo, and the
    per-pair accounting."""
    w, h, n = 720, 576, 37
    clip = mofreak_b200.synth_clip(w, h, n, 101)
    with mofreak_b200.Context(w, h, n, 32) as ctx:
        a = ctx.extract_clip(clip)
        b = ctx.extract_clip(clip)
        assert util.rows_equal(a, b) and len(a) > 20000
        ctx.upload_frames(clip, 0)
        order = np.arange(5, n)[::-1].copy()                        # reversed pair order, different batch composition
        ctx.extract_pairs(order, order - 5, order - 1)
        rows, per = ctx.fetch_rows(len(a) + 10, len(order))
        assert per.sum() == len(a)
        back = np.concatenate([rows[per[:i].sum():per[:i + 1].sum()] for i in range(len(order))][::-1])
        assert util.rows_equal(back, a)
    parts = []
    for first, lo, hi in sharding.shard_stream(n, 4):               # what 4 GPUs would each do
        with mofreak_b200.Context(w, h, 16, 8) as ctx:
            r = ctx.extract_clip(clip[first:hi])
            r["frame"] += first
            parts.append(r)
    assert util.rows_equal(np.concatenate(parts), a)
    assert np.all(np.diff(a["frame"]) >= 0) and a["frame"].min() == 4 and a["frame"].max() == n - 2


def test_full_size_properties_kth_batch():
    """BASELINE configs[1] shape: many small clips in one batch; rows of a clip must not depend on its neighbours."""
    clips = [mofreak_b200.synth_clip(160, 120, 30, 200 + i) for i in range(24)]
    with mofreak_b200.Context(160, 120, 24 * 30, 512) as ctx:
        alone = [ctx.extract_clip(c) for c in clips[:3]]
        ctx.upload_frames(np.concatenate(clips), 0)
        cur = np.concatenate([np.arange(5, 30) + 30 * i for i in range(24)])
        lab = np.concatenate([np.arange(5, 30) - 1 for _ in range(24)])
        ctx.extract_pairs(cur, cur - 5, lab)
        rows, per = ctx.fetch_rows(1 << 20, len(cur))
    per_clip = per.reshape(24, 25).sum(1)
    offs = np.concatenate([[0], np.cumsum(per_clip)]).astype(np.int64)
    for i in range(3):
        assert util.rows_equal(rows[offs[i]:offs[i + 1]], alone[i])
    assert per.sum() == len(rows)


def test_serial_and_pipelined_launch_sequences_agree(orc):
    """Consecutive launch sequences rotate through four buffer sets on their own streams; the rows must not depend on it."""
    clip = mofreak_b200.synth_clip(320, 240, 30, 77)
    want = orc.extract_clip(clip)
    with mofreak_b200.Context(320, 240, 30, 4) as ctx:          # 25 pairs -> 7 launch sequences
        a = ctx.extract_clip(clip)
        ctx.set_serial(True)
        b = ctx.extract_clip(clip)
        ctx.set_serial(False)
        c = ctx.extract_clip(clip)
    assert util.rows_equal(a, want) and util.rows_equal(b, want) and util.rows_equal(c, want)


def test_wide_frames_take_the_separate_dense_kernels(orc):
    """Frames wider than the fused front kernel's band buffer (2048 px) run absdiff / pyramid / row sums as separate kernels."""
    clip = mofreak_b200.synth_clip(2064, 96, 7, 9)
    want = orc.extract_clip(clip)
    with mofreak_b200.Context(2064, 96, 8, 2) as ctx:
        ctx.profile_enable(True)
        got = ctx.extract_clip(clip)
        prof = ctx.profile_read()
    assert len(want) > 0 and util.rows_equal(got, want), util.describe_diff(got, want)
    assert prof["front"][0] == 0 and prof["pyramid"][0] == 5 and prof["absdiff"][0] == 1


@pytest.mark.parametrize("shape", [(720, 576), (212, 163), (100, 58)])
def test_fused_front_equals_separate_dense_kernels(orc, shape, monkeypatch):
    """The fused front (TMA-staged bands when the width is a multiple of 16, word reads otherwise; partial last band) and
    the separate kernels produce the same pyramid, hence the same rows."""
    w, h = shape
    clip = mofreak_b200.synth_clip(w, h, 8, 17)
    layers = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("MFK_FUSED_FRONT", fused)
        with mofreak_b200.Context(w, h, 8, 3) as ctx:
            ctx.upload_frames(clip, 0)
            ctx.profile_enable(True)
            ctx.extract_pairs(np.array([5, 6, 7], np.int32), np.array([0, 1, 2], np.int32), np.array([4, 5, 6], np.int32))
            rows, _ = ctx.fetch_rows(100000, 3)
            prof = ctx.profile_read()
            assert (prof["front"][0] == 1) == (fused == "1")
            layers[fused] = ([ctx.debug_layer(0, l, p) for l in range(6) for p in range(3)], rows.copy())
    for a, b in zip(layers["1"][0], layers["0"][0]):
        assert np.array_equal(a, b)
    assert util.rows_equal(layers["1"][1], layers["0"][1])
    want = orc.extract_clip(clip)
    assert util.rows_equal(layers["1"][1], want)


def test_timer_and_profile_hooks():
    clip = mofreak_b200.synth_clip(160, 120, 9, 301)
    with mofreak_b200.Context(160, 120, 16, 4) as ctx:
        ctx.upload_frames(clip, 0)
        ctx.profile_enable(True)
        before = ctx.launch_count()
        ctx.timer_start()
        cur = np.arange(5, 9)
        ctx.extract_pairs(cur, cur - 5, cur - 1)
        ms = ctx.timer_stop()
        prof = ctx.profile_read()
        assert ms > 0 and ctx.launch_count() - before == 20      # one launch sequence: 20 kernels with the fused front
        assert prof["scores"][0] == 1 and prof["scores"][1] > 0 and prof["describe"][1] > 0 and prof["front"][0] == 1


@pytest.mark.gpu
def test_profile_timeline_lists_every_launcher_call_in_stream_order():
    clip = mofreak_b200.synth_clip(320, 240, 12, 5)
    with mofreak_b200.Context(320, 240, 16, 2) as ctx:          # 7 pairs in batches of 2 -> 4 launch sequences on 4 lanes
        ctx.upload_frames(clip)
        ctx.profile_enable(True)
        ctx.timer_start()
        cur = np.arange(5, 12, dtype=np.int32)
        ctx.extract_pairs(cur, cur - 5, cur - 1)
        ms = ctx.timer_stop()
        tl = ctx.profile_timeline()
        names = [t[0] for t in tl]
        assert names.count("scores") == 4 and names.count("describe") == 4 and names.count("compact") == 8
        assert all(0.0 <= a <= b <= ms + 1e-3 for _, a, b in tl)
        # within one launch sequence the families follow the dependency order
        first = names[:names.index("describe") + 1]
        assert first.index("front") < first.index("scores") < first.index("phase_a") < first.index("phase_b") < first.index("describe")
        prof = ctx.profile_read()
        assert ctx.profile_timeline() == []                        # consumed by profile_read
        assert prof["scores"][0] == 4
        ctx.profile_enable(False)
        ctx.fetch_rows(1 << 16)


def _gray_fixed_point(bgr, bits):
    cb, cg, cr = (1868, 9617, 4899) if bits == 14 else (3735, 19235, 9798)
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * cb + g * cg + r * cr + (1 << (bits - 1))) >> bits).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(160, 120), (53, 49)])
@pytest.mark.parametrize("bits", [14, 15])
def test_bgr_ingest_matches_cvtcolor_arithmetic(shape, bits):
    w, h = shape
    rng = np.random.default_rng(21)
    n = 5
    bgr = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    want = _gray_fixed_point(bgr, bits)
    if bits == 15:
        cv2 = pytest.importorskip("cv2")                     # OpenCV >= 3 arithmetic is what this container's cv2 computes
        assert np.array_equal(want[0], cv2.cvtColor(bgr[0], cv2.COLOR_BGR2GRAY))
    with mofreak_b200.Context(w, h, 2 * n, n) as ctx:
        ctx.upload_frames_bgr(bgr, slot0=0, coeff_bits=bits)      # device conversion into slots 0..n-1
        ctx.upload_frames(want, slot0=n)                           # host conversion into slots n..2n-1
        cur = np.arange(n, dtype=np.int32)
        ctx.extract_pairs(cur, cur + n, cur)
        for pair in range(n):                                       # layer 0 is |slot cur - slot prev|: zero iff identical
            assert not ctx.debug_layer(0, 0, pair).any()
        ctx.fetch_rows(1 << 16)
        with pytest.raises(mofreak_b200.MofreakError):
            ctx.upload_frames_bgr(bgr, coeff_bits=16)


@pytest.mark.gpu
def test_back_to_back_steps_and_interleaved_contexts_agree(orc):
    """The bench's shapes: steps queued back to back without a sync (the row reset is applied in-line, the front of step
    k + 1 runs under the tail of step k), and two contexts alternating (the end-to-end double buffer)."""
    clip = mofreak_b200.synth_clip(320, 240, 21, 31)
    want = orc.extract_clip(clip)
    cur = np.arange(5, 21, dtype=np.int32)
    with mofreak_b200.Context(320, 240, 21, 3) as a, mofreak_b200.Context(320, 240, 21, 5) as b:
        a.upload_frames(clip); b.upload_frames(clip)
        for _ in range(4):                                   # only the last step's rows may remain
            a.reset_rows()
            a.extract_pairs(cur, cur - 5, cur - 1)
        rows, per = a.fetch_rows(len(want) + 8, len(cur))
        assert util.rows_equal(rows, want) and per.sum() == len(want)
        got = []
        for k in range(4):                                   # submit on one context, then fetch the other one's previous step
            c = (a, b)[k % 2]
            c.reset_rows()
            c.extract_pairs(cur, cur - 5, cur - 1)
            if k:
                got.append((a, b)[(k - 1) % 2].fetch_rows(len(want) + 8)[0])
        got.append(b.fetch_rows(len(want) + 8)[0])
        assert all(util.rows_equal(g, want) for g in got) and len(got) == 4
        # a reset that no launch sequence consumed is applied by the next reader
        a.extract_pairs(cur[:2], cur[:2] - 5, cur[:2] - 1)
        a.reset_rows()
        assert a.rows_device()[1] == 0 and a.rows_text() == b""


@pytest.mark.gpu
@pytest.mark.parametrize("wc", [False, True])
def test_host_buffers_from_the_library(wc):
    clip = mofreak_b200.synth_clip(160, 120, 9, 4)
    buf = mofreak_b200.HostBuffer(clip.size, write_combined=wc)
    buf.array[:] = clip.reshape(-1)
    with mofreak_b200.Context(160, 120, 16, 4) as ctx:
        want = ctx.extract_clip(clip)
        ctx.upload_raw(buf.ptr, 9, 0)
        cur = np.arange(5, 9, dtype=np.int32)
        ctx.extract_pairs(cur, cur - 5, cur - 1)
        got, _ = ctx.fetch_rows(len(want) + 8)
        assert util.rows_equal(got, want) and len(want) > 0
    buf.close()


_SWITCH_CACHE = {}
_SWITCH_SCRIPT = r"""
import hashlib, sys
sys.path.insert(0, %r)
import mofreak_b200
out = []
for (w, h, n, seed) in [(720, 576, 7, 13), (367, 251, 8, 15), (160, 120, 14, 11)]:
    clip = mofreak_b200.synth_clip(w, h, n, seed)
    with mofreak_b200.Context(w, h, max(16, n), 3) as ctx:
        rows = ctx.extract_clip(clip)
    out.append("%%d:%%s" %% (len(rows), hashlib.sha256(rows.tobytes()).hexdigest()))
print(" ".join(out))
"""


@pytest.mark.parametrize("env", [{"MFK_SCORES_NO_TENSOR_MAP": "1"}, {"MFK_EARLY_LOSERS": "0"}, {"MFK_PRIORITY": "0", "MFK_EARLY_COMPACT": "0"},
                                 {"MFK_TAIL_OVERLAP": "1", "MFK_LANES": "2"}, {"MFK_FUSED_FRONT": "0", "MFK_PB_GROUP": "256"}])
def test_measurement_switches_do_not_change_the_rows(env):
    """The environment switches choose between implementations of the same step (tile staging by the TMA tensor copy or by
    thread loads, early removal of 8-neighbour losers, stream priorities, lane count, fused or separate dense front, phase B
    group size): whatever they are set to, the rows are the same bytes.  Each setting runs in a process of its own because the
    library reads the switches once."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = _SWITCH_SCRIPT % root

    def run(extra):
        e = dict(os.environ)
        for k in ("MFK_SCORES_NO_TENSOR_MAP", "MFK_EARLY_LOSERS", "MFK_PRIORITY", "MFK_EARLY_COMPACT", "MFK_TAIL_OVERLAP", "MFK_LANES",
                  "MFK_FUSED_FRONT", "MFK_PB_GROUP", "MFK_PA_BLOCKS", "MFK_DESC_BLOCKS"):
            e.pop(k, None)
        e.update(extra)
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600, env=e)
        assert r.returncode == 0, r.stderr[-2000:]
        return r.stdout.strip().splitlines()[-1]

    if "base" not in _SWITCH_CACHE:
        _SWITCH_CACHE["base"] = run({})
    base = _SWITCH_CACHE["base"]
    assert base.count(":") == 3 and not base.startswith("0:")
    assert run(env) == base
