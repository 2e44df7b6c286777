"""The C++ host classes that keep the reference's call surface (mofreak_b200/host/): MoFREAKUtilities,
MotionInterchangePatterns, BagOfWordsRepresentation — driven through their extern "C" shim."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mofreak_b200
from tests import util

HOST_DIR = os.path.join(os.path.dirname(mofreak_b200.__file__), "host")


@pytest.fixture(scope="module")
def host():
    path = os.path.join(HOST_DIR, "libmofreak_host.so")
    if not os.path.exists(path):
        mofreak_b200.build.build()
        subprocess.run(["make", "-s", "-C", HOST_DIR], check=True)
    L = C.CDLL(path)
    L.mfkh_compute_mofreak_from_file.restype = C.c_long
    L.mfkh_rewrite_mofreak.restype = C.c_long
    L.mfkh_compute_mofreak_clear.restype = C.c_long
    L.mfkh_compute_mofreak_clear.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_long]
    return L


def clusters_text(cb):
    return "".join(" ".join(str(int(v)) for v in row) + " \n" for row in cb)


def libsvm_line(label, hist):
    return "%d " % label + "".join("%d:%g " % (i + 1, float(v)) for i, v in enumerate(hist))


# ---------------------------------------------------------------------------------------------- CPU side
def test_reader_writer_round_trip_is_byte_identical(host, tmp_path):
    name = "kth_160x120x100_seed1"
    src = tmp_path / "in.mofreak"; dst = tmp_path / "out.mofreak"
    src.write_text(util.golden_text(name))
    n = host.mfkh_rewrite_mofreak(str(src).encode(), str(dst).encode())
    assert n == len(util.golden_rows(name))
    # %g keeps 6 significant digits, so text -> float -> text is the identity on the reference's own output
    assert dst.read_text() == util.golden_text(name)


def test_load_clusters_quirk_matches_oracle(host, orc, tmp_path):
    rng = np.random.default_rng(3)
    cb = rng.integers(0, 256, (9, 16), dtype=np.uint8)
    (tmp_path / "clusters.txt").write_text(clusters_text(cb))
    out = np.zeros((16, 16), np.uint8)
    K = host.mfkh_bow_codebook(9, str(tmp_path).encode(), out.ctypes.data_as(C.c_void_p), 16)
    want = orc.load_clusters(clusters_text(cb))
    assert K == len(want) == 9
    assert np.array_equal(out[:K], want)
    assert np.array_equal(out[:8], cb[1:]) and (out[8] == cb[8, 15]).all()      # line k+2; last word repeats the last value


def test_gray_video_container(host, tmp_path):
    clip = mofreak_b200.synth_clip(160, 120, 3, 5)
    p = tmp_path / "clip.y8"
    assert host.mfkh_write_gray_video(str(p).encode(), clip.ctypes.data_as(C.c_void_p), 160, 120, 3) == 0
    raw = p.read_bytes()
    assert raw[:8] == b"MFKGRAY1" and np.frombuffer(raw[8:20], "<i4").tolist() == [160, 120, 3]
    assert raw[20:] == clip.tobytes()


def _cluster(host, tmp_path, files, classes, K, n_classes, seed):
    arr = (C.c_char_p * len(files))(*[str(f).encode() for f in files])
    cls = (C.c_int * len(files))(*classes)
    host.mfkh_cluster.argtypes = [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_uint]
    filled = host.mfkh_cluster(K, n_classes, str(tmp_path).encode(), arr, cls, len(files), seed)
    text = (tmp_path / "clusters.txt").read_text()
    cb = np.array([[int(t) for t in line.split()] for line in text.split("\n") if line], np.uint8)
    return filled, text, cb


def test_codebook_selection_is_seeded_balanced_and_drawn_from_the_class(host, tmp_path):
    """Clustering::randomClusters (Clustering.cpp:165-213) with an explicit seed: reproducible, an equal quota per class,
    every codeword one of that class's descriptors, clusters.txt in the format loadClusters reads."""
    names = ["kth_160x120x100_seed1", "ucf_320x240x12_seed3", "trecvid_720x576x7_seed5"]
    files = []
    for i, n in enumerate(names):
        f = tmp_path / ("class%d.mofreak" % (i + 1)); f.write_text(util.golden_text(n)); files.append(f)
    K, n_classes = 12, 3
    filled, text, cb = _cluster(host, tmp_path, files, [1, 2, 3], K, n_classes, 7)
    assert filled == K and cb.shape == (K, 16)
    assert text == clusters_text(cb)                                   # K lines of 16 integers, each followed by a space
    for c, n in enumerate(names):
        rows = util.golden_rows(n)
        pool = {bytes(r["appearance"]) + bytes(r["motion"]) for r in rows}
        for k in range(4 * c, 4 * c + 4):                               # K / classes consecutive codewords per class
            assert bytes(cb[k]) in pool
    assert _cluster(host, tmp_path, files, [1, 2, 3], K, n_classes, 7)[1] == text        # same seed, same file
    assert _cluster(host, tmp_path, files, [1, 2, 3], K, n_classes, 8)[1] != text        # another seed, another draw


# ---------------------------------------------------------------------------------------------- GPU side
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kth_160x120x100_seed1", "odd_212x163x9_seed7"])
def test_compute_mofreak_from_file_writes_the_golden_file(host, tmp_path, name):
    c = util.golden_cases()[name]
    clip = mofreak_b200.synth_clip(c["width"], c["height"], c["frames"], c["seed"])
    video = tmp_path / "person01_boxing_d1_uncomp.y8"; out = tmp_path / "person01_boxing_d1_uncomp.mofreak"
    host.mfkh_write_gray_video(str(video).encode(), clip.ctypes.data_as(C.c_void_p), c["width"], c["height"], c["frames"])
    n = host.mfkh_compute_mofreak_from_file(0, str(video).encode(), str(out).encode())
    assert n == c["rows"]
    assert out.read_text() == util.golden_text(name)          # the .mofreak file is byte-identical


@pytest.mark.gpu
@pytest.mark.parametrize("rows_hint", [0, 1])
def test_clear_after_computation_writes_device_text_without_building_features(host, tmp_path, rows_hint):
    name = "ucf_320x240x12_seed3"
    c = util.golden_cases()[name]
    clip = mofreak_b200.synth_clip(c["width"], c["height"], c["frames"], c["seed"])
    video = tmp_path / "v.y8"; out = tmp_path / "v.mofreak"
    host.mfkh_write_gray_video(str(video).encode(), clip.ctypes.data_as(C.c_void_p), c["width"], c["height"], c["frames"])
    # rows_hint = 1: the first host buffers are too small, the device keeps the rows and they are read again
    n = host.mfkh_compute_mofreak_clear(3, str(video).encode(), str(out).encode(), rows_hint)
    assert n == c["rows"]
    assert out.read_text() == util.golden_text(name)


@pytest.mark.gpu
def test_command_line_driver(tmp_path):
    exe = os.path.join(HOST_DIR, "mofreak_main")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", HOST_DIR], check=True)
    c = util.golden_cases()["ucf_320x240x12_seed3"]
    clip = mofreak_b200.synth_clip(c["width"], c["height"], c["frames"], c["seed"])
    video = tmp_path / "v_Archery_g01_c01.y8"
    video.write_bytes(b"MFKGRAY1" + np.array([c["width"], c["height"], c["frames"]], "<i4").tobytes() + clip.tobytes())
    r = subprocess.run([exe, "detect", "UCF101", str(tmp_path), str(video)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "v_Archery_g01_c01.mofreak").read_text() == util.golden_text("ucf_320x240x12_seed3")


@pytest.mark.gpu
def test_bag_of_words_cross_validation_files(host, orc, tmp_path):
    rng = np.random.default_rng(5)
    K, groups = 40, 3
    cb = rng.integers(0, 256, (K + 1, 16), dtype=np.uint8)         # first line is skipped by the loader
    cb[-1] = cb[-1, 15]                                            # keep the last (repeated-value) codeword well defined
    (tmp_path / "clusters.txt").write_text(clusters_text(cb))
    loaded = orc.load_clusters(clusters_text(cb))
    names = ["person01_boxing_d1_uncomp.mofreak", "person02_boxing_d2_uncomp.mofreak", "person02_walking_d1_uncomp.mofreak",
             "person03_running_d4_uncomp.mofreak"]
    files, expect = [], []
    actions = {}
    for i, nm in enumerate(names):
        clip = mofreak_b200.synth_clip(160, 120, 14, 30 + i)
        rows = orc.extract_clip(clip)
        assert len(rows) > 0
        p = tmp_path / nm
        p.write_text(orc.format_mofreak(rows))
        files.append(str(p).encode())
        ok, hist = orc.build_histogram(orc.format_mofreak(rows), loaded)
        act = nm.split("_")[1]
        actions.setdefault(act, len(actions))
        group = int(nm.split("_")[0][-2:]) - 1
        expect.append((group, libsvm_line(actions[act] + 1, hist)))
    arr = (C.c_char_p * len(files))(*files)
    assert host.mfkh_bow_run(K + 1, str(tmp_path).encode(), groups, 0, arr, len(files)) == K + 1
    for g in range(groups):
        train = (tmp_path / ("%d.train" % (g + 1))).read_text().splitlines()
        test = (tmp_path / ("%d.test" % (g + 1))).read_text().splitlines()
        assert test == [l for (grp, l) in expect if grp == g]
        assert train == [l for (grp, l) in expect if grp != g]


@pytest.mark.gpu
def test_sliding_bag_of_words(host, orc, tmp_path):
    rng = np.random.default_rng(6)
    cb = rng.integers(0, 256, (21, 16), dtype=np.uint8); cb[-1] = 9
    (tmp_path / "clusters.txt").write_text(clusters_text(cb))
    loaded = orc.load_clusters(clusters_text(cb))
    clip = mofreak_b200.synth_clip(160, 120, 30, 44)
    rows = orc.extract_clip(clip)
    mof = tmp_path / "stream.mofreak"; mof.write_text(orc.format_mofreak(rows))
    alpha = 4
    outp = tmp_path / "out.txt"
    host.mfkh_bow_sliding(21, str(tmp_path).encode(), str(mof).encode(), alpha, 7, str(outp).encode())
    desc = np.concatenate([rows["appearance"], rows["motion"]], 1)
    idx = orc.bow_assign(desc, loaded)
    frames = rows["frame"]
    uniq = sorted(set(frames.tolist()))[:-1]                      # the last frame's rows never form a histogram
    hists = [np.bincount(idx[frames == f], minlength=21).astype(np.float32) for f in uniq]
    want = []
    for i in range(alpha - 1, len(hists)):
        h = np.sum(hists[i - alpha + 1:i + 1], axis=0, dtype=np.float32)
        want.append(libsvm_line(7, h / np.float32(h.sum())))
    assert len(want) > 3
    assert outp.read_text().splitlines() == want


@pytest.mark.gpu
def test_mip_facade_class(host, orc):
    clip = mofreak_b200.synth_clip(320, 240, 6, 81)
    out = np.zeros(8, np.uint8)
    host.mfkh_mip(clip[5].ctypes.data_as(C.c_void_p), clip[0].ctypes.data_as(C.c_void_p), 320, 240, C.c_float(17.3), 150, 100,
                  out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, orc.mip_extract(clip[5], clip[0], 17.3, 150, 100))


@pytest.mark.gpu
def test_host_class_grows_device_budgets_for_busy_frames(host, orc, tmp_path):
    """Quantised noise has far more detector candidates than the default per-pair budget; the reference has no such
    limit, so the host class must grow the budgets and still write the reference's file."""
    rng = np.random.default_rng(9)
    w, h = 160, 120
    frames = np.zeros((7, h, w), np.uint8)
    frames[5] = (rng.integers(0, 4, (h, w)) * 80).astype(np.uint8)
    frames[6] = (rng.integers(0, 3, (h, w)) * 100).astype(np.uint8)
    video = tmp_path / "noise.y8"; out = tmp_path / "noise.mofreak"
    host.mfkh_write_gray_video(str(video).encode(), frames.ctypes.data_as(C.c_void_p), w, h, 7)
    n = host.mfkh_compute_mofreak_from_file(0, str(video).encode(), str(out).encode())
    want = orc.extract_clip(frames)
    assert n == len(want)
    assert out.read_text() == orc.format_mofreak(want)


def test_command_line_cluster_stage(tmp_path):
    exe = os.path.join(HOST_DIR, "mofreak_main")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", HOST_DIR], check=True)
    a = tmp_path / "a.mofreak"; a.write_text(util.golden_text("kth_160x120x100_seed1"))
    b = tmp_path / "b.mofreak"; b.write_text(util.golden_text("ucf_320x240x12_seed3"))
    args = [exe, "cluster", "KTH", "10", "2", str(tmp_path), "5", "1:%s" % a, "2:%s" % b]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    first = (tmp_path / "clusters.txt").read_text()
    lines = [l for l in first.split("\n") if l]
    assert len(lines) == 10 and all(len(l.split()) == 16 and l.endswith(" ") for l in lines)
    subprocess.run(args, capture_output=True, text=True, check=True)
    assert (tmp_path / "clusters.txt").read_text() == first                      # seeded: reproducible


@pytest.mark.gpu
def test_detection_to_classification_pipeline_files(orc, tmp_path):
    """BASELINE configs[4] up to the SVM (main.cpp:944-962: computeMoFREAKFiles -> cluster -> computeBOWRepresentation) on a
    6-clip 320x240 synthetic set, through the command-line driver; every file it writes is compared byte for byte with what
    the oracle pipeline produces: `.mofreak` (extraction), clusters.txt (seeded codebook selection; cluster_restate.cpp fed
    with the generator's numbers), <k>.train / <k>.test (loadClusters quirk, histograms, leave-one-group-out fan-out)."""
    from tests.test_oracle_next_rows import lcg15
    exe = os.path.join(HOST_DIR, "mofreak_main")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", HOST_DIR], check=True)
    names = [("person01_boxing_d1_uncomp", 1), ("person02_boxing_d2_uncomp", 1), ("person03_boxing_d1_uncomp", 1),
             ("person01_walking_d1_uncomp", 2), ("person02_walking_d3_uncomp", 2), ("person03_walking_d2_uncomp", 2)]
    K, n_classes, groups, seed = 24, 2, 3, 11
    args, rows_of = [], {}
    for i, (nm, cls) in enumerate(names):
        clip = mofreak_b200.synth_clip(320, 240, 11, 500 + i)
        video = tmp_path / (nm + ".y8")
        video.write_bytes(b"MFKGRAY1" + np.array([320, 240, 11], "<i4").tobytes() + clip.tobytes())
        args.append("%d:%s" % (cls, video))
        rows_of[nm] = orc.extract_clip(clip)
        assert len(rows_of[nm]) > K
    r = subprocess.run([exe, "pipeline", "KTH", str(K), str(n_classes), str(groups), str(tmp_path), str(seed)] + args,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # 1. extraction
    for nm, _ in names:
        assert (tmp_path / (nm + ".mofreak")).read_text() == orc.format_mofreak(rows_of[nm])
    # 2. codebook: per class, files are read in order, each file's rows reversed (the reader), three seeded shuffles
    centers, center_row, shuffles = np.zeros((K, 16), np.uint8), 0, 0
    for cls in (1, 2):
        rows = np.concatenate([rows_of[nm][::-1] for nm, c in names if c == cls])
        data = np.zeros((len(rows), 19), np.uint8)
        data[:, 3:11] = rows["appearance"]; data[:, 11:] = rows["motion"]
        rands = []
        for _ in range(3):
            rands += lcg15(seed + 0x9E3779B9 * shuffles, len(rows) - 1); shuffles += 1
        centers, center_row, _, _ = orc.random_clusters(data, K, n_classes, [1, 2], True, rands, centers, center_row)
    assert center_row == K
    text = orc.clusters_text(centers)
    assert (tmp_path / "clusters.txt").read_text() == text
    # 3. bag of words with the loader's off-by-one (first line unused, last codeword repeats the last value)
    loaded = orc.load_clusters(text)
    assert len(loaded) == K
    expect, actions = [], {}
    for nm, _ in names:
        ok, hist = orc.build_histogram(orc.format_mofreak(rows_of[nm]), loaded)
        assert ok
        act = nm.split("_")[1]
        actions.setdefault(act, len(actions))
        expect.append((int(nm.split("_")[0][-2:]) - 1, libsvm_line(actions[act] + 1, hist)))
    for g in range(groups):
        assert (tmp_path / ("%d.test" % (g + 1))).read_text().splitlines() == [l for (grp, l) in expect if grp == g]
        assert (tmp_path / ("%d.train" % (g + 1))).read_text().splitlines() == [l for (grp, l) in expect if grp != g]
