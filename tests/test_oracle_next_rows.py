"""SURVEY 8(f) "next" rows N2 / N3: the oracle restatements of computeSlidingBagOfWords
(BagOfWordsRepresentation.cpp:176-378) and of the codebook selection (Clustering.cpp:6-25, 165-213, 233-257), checked
(a) against independent re-derivations in numpy / plain Python, and (b) the product's host classes against them."""
import ctypes as C
import os

import numpy as np
import pytest

import mofreak_b200
from tests import util
from tests.test_host_layer import clusters_text, host, libsvm_line  # noqa: F401  (host is a fixture)


def lcg15(seed, n):
    """rand() of the reference's runtime (MSVC): 15-bit LCG; what the host class draws when seeded."""
    out, state = [], seed & 0xffffffff
    for _ in range(n):
        state = (state * 214013 + 2531011) & 0xffffffff
        out.append((state >> 16) & 0x7fff)
    return out


# ------------------------------------------------------------------------------------------------ N2: sliding window
def sliding_expected(rows, idx, K, alpha, label):
    frames = rows["frame"]
    uniq = sorted(set(frames.tolist()))[:-1]                      # the last frame's rows never form a histogram
    hists = [np.bincount(idx[frames == f], minlength=K).astype(np.float32) for f in uniq]
    lines = []
    for i in range(alpha - 1, len(hists)):
        h = np.sum(hists[i - alpha + 1:i + 1], axis=0, dtype=np.float32)
        lines.append(libsvm_line(label, h / np.float32(h.sum())))
    if not lines:                                                 # no window ever filled: one line from what there is
        h = np.sum(hists, axis=0, dtype=np.float32) if hists else np.zeros(K, np.float32)
        with np.errstate(invalid="ignore"):
            lines.append(libsvm_line(label, h / np.float32(h.sum())))
    return lines


@pytest.mark.parametrize("alpha", [1, 4, 12, 500])
def test_sliding_restatement_equals_numpy_derivation(orc, alpha):
    rows = util.golden_rows("kth_160x120x100_seed1")
    text = util.golden_text("kth_160x120x100_seed1")
    rng = np.random.default_rng(11)
    cb = rng.integers(0, 256, (37, 16), dtype=np.uint8)
    desc = np.concatenate([rows["appearance"], rows["motion"]], 1)
    idx = orc.bow_assign(desc, cb)
    got = orc.sliding_bow(text, cb, alpha, 3).splitlines()
    want = sliding_expected(rows, idx, 37, alpha, 3)
    assert len(got) == len(want) and (alpha == 500) == (len(want) == 1)
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("alpha", [3, 12, 500])
def test_host_sliding_bag_of_words_equals_oracle_restatement(host, orc, tmp_path, alpha):
    rng = np.random.default_rng(6)
    cb = rng.integers(0, 256, (21, 16), dtype=np.uint8); cb[-1] = 9
    (tmp_path / "clusters.txt").write_text(clusters_text(cb))
    loaded = orc.load_clusters(clusters_text(cb))
    text = util.golden_text("ucf_320x240x12_seed3")
    mof = tmp_path / "stream.mofreak"; mof.write_text(text)
    outp = tmp_path / "out.txt"
    host.mfkh_bow_sliding(21, str(tmp_path).encode(), str(mof).encode(), alpha, 7, str(outp).encode())
    assert outp.read_text() == orc.sliding_bow(text, loaded, alpha, 7)


# ------------------------------------------------------------------------------------------------ N3: codebook selection
def py_shuffle(data, rands):
    d = [list(r) for r in data]
    used, row0 = 0, len(d) - 1
    while row0 > 0:
        row1 = rands[used] % row0; used += 1
        d[row1], d[row0] = d[row0], d[row1]
        row0 -= 1
    return d, used


def test_cluster_restatement_equals_plain_python(orc):
    rng = np.random.default_rng(2)
    rows, dim = 57, 16
    data = rng.integers(0, 256, (rows, dim + 3), dtype=np.uint8)
    data[:, 0] = rng.integers(1, 4, rows)                          # classes 1..3
    rands = lcg15(99, 3 * (rows - 1))
    centers, center_row, used, shuffled = orc.random_clusters(data, 12, 3, [1, 2, 3], False, rands)
    d, at = data.tolist(), 0
    for _ in range(3):
        d, u = py_shuffle(d, rands[at:]); at += u
    assert used == at == 3 * (rows - 1) and shuffled.tolist() == d
    want = []
    for c in (1, 2, 3):
        want += [r[3:] for r in d if r[0] == c][:4]                # 12 / 3 per class, from the top
    assert center_row == len(want) == 12 and centers.tolist() == want
    assert orc.clusters_text(centers) == "".join("".join("%d " % v for v in row) + "\n" for row in want)
    # the shuffle never leaves a row in place as its own partner: rand() % row0 < row0
    _, _, _, once = orc.random_clusters(np.arange(5 * 19, dtype=np.uint8).reshape(5, 19), 1, 1, [0], True, [0] * 12)
    assert once.shape == (5, 19)


def test_host_clustering_equals_oracle_restatement(host, orc, tmp_path):
    """Clustering (host class, seeded) over two classes of `.mofreak` files, the shape of cluster() (main.cpp:160-213),
    against the restatement fed with the numbers the seeded generator draws."""
    names = ["kth_160x120x100_seed1", "ucf_320x240x12_seed3"]
    files, classes = [], []
    for k, name in enumerate(names):
        p = tmp_path / ("c%d.mofreak" % k); p.write_text(util.golden_text(name))
        files.append(str(p).encode()); classes.append(k + 1)
    seed, K = 5, 20
    arr = (C.c_char_p * len(files))(*files)
    cls = (C.c_int * len(files))(*classes)
    filled = host.mfkh_cluster(K, 2, str(tmp_path).encode(), arr, cls, len(files), seed)
    got = (tmp_path / "clusters.txt").read_text()
    centers, center_row, shuffles = np.zeros((K, 16), np.uint8), 0, 0
    for k, name in enumerate(names):
        rows = util.golden_rows(name)[::-1]                         # the reader stores a file's rows in reverse order
        data = np.zeros((len(rows), 19), np.uint8)                  # action / person / video are 0 for KTH-style names
        data[:, 3:11] = rows["appearance"]; data[:, 11:] = rows["motion"]
        rands = []
        for _ in range(3):                                          # the host class reseeds per shuffle: seed + golden ratio * count
            rands += lcg15(seed + 0x9E3779B9 * shuffles, len(rows) - 1); shuffles += 1
        centers, center_row, used, _ = orc.random_clusters(data, K, 2, [1, 2], True, rands, centers, center_row)
        assert used == 3 * (len(rows) - 1)
    assert filled == center_row == K
    assert got == orc.clusters_text(centers)
