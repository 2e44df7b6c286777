"""ctypes binding of the CPU oracle (oracle/).  TEST INFRASTRUCTURE: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libmofreak_ref.so")
PORT_LIB = os.path.join(ORACLE_DIR, "libmofreak_oracle.so")

ROW_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("frame", "<i4"), ("scale", "<f4"),
                      ("appearance", "u1", (8,)), ("motion", "u1", (8,))])


def build(target="all"):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, path):
        self.path = path
        L = self.L = C.CDLL(path)
        self.has_reference = bool(L.orc_has_reference())
        L.orc_time_extract.restype = C.c_double
        L.orc_time_bow_assign.restype = C.c_double
        L.orc_extract_clip.restype = C.c_long
        L.orc_format_mofreak.restype = C.c_long
        L.orc_hamming_u8.restype = C.c_uint
        L.orc_mip_byte.restype = C.c_uint

    # ---- detector ----
    def _detect(self, fn, img, *extra):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = 1 << 16
        out = np.zeros((cap, 5), np.float32)
        n = fn(_p(img), w, h, *extra, _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def brisk_port(self, img, thresh=30, jacobi_reverse=False):
        return self._detect(self.L.orc_brisk_detect, img, thresh, 1 if jacobi_reverse else 0)

    def brisk_ref(self, img, thresh=30, octaves=3):
        assert self.has_reference
        return self._detect(self.L.ref_brisk_detect, img, thresh, octaves)

    def brisk(self, img):
        """Reference object code when available, else the (pinned) port."""
        return self.brisk_ref(img) if self.has_reference else self.brisk_port(img)

    def halfsample(self, img, ref=False):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        out = np.zeros((h // 2, w // 2), np.uint8)
        (self.L.ref_halfsample if ref else self.L.orc_halfsample)(_p(img), w, h, _p(out))
        return out

    def twothirdsample(self, img, ref=False):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        out = np.zeros((2 * (h // 3), 2 * (w // 3)), np.uint8)
        (self.L.ref_twothirdsample if ref else self.L.orc_twothirdsample)(_p(img), w, h, _p(out))
        return out

    def score_map(self, img):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        out = np.zeros((h, w), np.uint8)
        self.L.orc_score_map(_p(img), w, h, _p(out))
        return out

    def oast_score(self, img, x, y):
        return self.L.orc_oast_score(_p(img), img.shape[1], x, y)

    def agast58_score(self, img, x, y):
        return self.L.orc_agast58_score(_p(img), img.shape[1], x, y)

    def ref_oast_score(self, img, x, y, b):
        return self.L.ref_oast_score(_p(img), img.shape[1], img.shape[0], x, y, b)

    def ref_agast58_score(self, img, x, y, b):
        return self.L.ref_agast58_score(_p(img), img.shape[1], img.shape[0], x, y, b)

    def ref_oast_detect(self, img, b):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        xy = np.zeros((w * h, 2), np.int32)
        n = self.L.ref_oast_detect(_p(img), w, h, b, _p(xy), w * h)
        return xy[:n].copy()

    def ref_dump(self, img, layer):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        li = np.zeros(w * h, np.uint8); ls = np.zeros(w * h, np.uint8); dims = (C.c_int * 2)()
        self.L.ref_brisk_dump(_p(img), w, h, 30, 3, layer, _p(li), _p(ls), dims)
        n = dims[0] * dims[1]
        return li[:n].reshape(dims[1], dims[0]).copy(), ls[:n].reshape(dims[1], dims[0]).copy()

    # ---- FREAK / MIP ----
    def integral(self, img):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        out = np.zeros((h + 1, w + 1), np.int32)
        self.L.orc_integral(_p(img), w, h, _p(out))
        return out

    def freak_pattern_size(self, s):
        return self.L.orc_freak_pattern_size(s)

    def freak_scale_index(self, size):
        return self.L.orc_freak_scale_index(C.c_float(size))

    def freak_compute(self, img, kps):
        img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
        k = np.ascontiguousarray(kps, np.float32).reshape(-1, 5).copy()
        desc = np.zeros((max(1, k.shape[0]), 64), np.uint8)
        theta = np.zeros(max(1, k.shape[0]), np.int32)
        n = self.L.orc_freak_compute(_p(img), w, h, _p(k), k.shape[0], _p(desc), _p(theta))
        return k[:n].copy(), desc[:n].copy(), theta[:n].copy()

    def resize19(self, roi):
        roi = np.ascontiguousarray(roi, np.uint8); s = roi.shape[0]
        assert roi.shape == (s, s)
        out = np.zeros((19, 19), np.uint8)
        self.L.orc_resize19(_p(roi), s, s, _p(out))
        return out

    def mip_byte(self, cur19, prev19, x, y):
        return self.L.orc_mip_byte(_p(np.ascontiguousarray(cur19, np.uint8)), _p(np.ascontiguousarray(prev19, np.uint8)), x, y)

    def mip_extract(self, cur, prev, size, x, y):
        cur = np.ascontiguousarray(cur, np.uint8); prev = np.ascontiguousarray(prev, np.uint8)
        h, w = cur.shape
        out = np.zeros(8, np.uint8)
        self.L.orc_mip_extract(_p(cur), _p(prev), w, h, C.c_float(size), int(x), int(y), _p(out))
        return out

    # ---- extraction ----
    def extract_clip(self, frames, use_ref=None, as_written=False):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        if use_ref is None:
            use_ref = self.has_reference
        cap = max(1024, n * 4096)
        rows = np.zeros(cap, ROW_DTYPE)
        got = self.L.orc_extract_clip(_p(frames), w, h, n, 1 if use_ref else 0, 1 if as_written else 0, _p(rows), C.c_long(cap))
        assert got <= cap
        return rows[:got].copy()

    def time_extract(self, clips, use_ref=None, as_written=False, threads=1):
        clips = np.ascontiguousarray(clips, np.uint8)
        nc, n, h, w = clips.shape
        if use_ref is None:
            use_ref = self.has_reference
        totals = (C.c_long * 2)()
        sec = self.L.orc_time_extract(_p(clips), w, h, n, nc, 1 if use_ref else 0, 1 if as_written else 0, threads, totals)
        return sec, totals[0], totals[1]

    def format_mofreak(self, rows):
        rows = np.ascontiguousarray(rows)
        assert rows.dtype.itemsize == 32
        n = self.L.orc_format_mofreak(_p(rows), C.c_long(rows.shape[0]), None, C.c_long(0))
        buf = C.create_string_buffer(n + 1)
        self.L.orc_format_mofreak(_p(rows), C.c_long(rows.shape[0]), buf, C.c_long(n))
        return buf.raw[:n].decode()

    # ---- bag of words ----
    def hamming_u8(self, a, b):
        return self.L.orc_hamming_u8(int(a), int(b))

    def bow_assign(self, desc, codebook):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 16)
        cb = np.ascontiguousarray(codebook, np.uint8).reshape(-1, 16)
        idx = np.zeros(d.shape[0], np.int32)
        self.L.orc_bow_assign(_p(d), C.c_long(d.shape[0]), _p(cb), cb.shape[0], _p(idx))
        return idx

    def time_bow_assign(self, desc, codebook, threads=1):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 16)
        cb = np.ascontiguousarray(codebook, np.uint8).reshape(-1, 16)
        idx = np.zeros(d.shape[0], np.int32)
        return self.L.orc_time_bow_assign(_p(d), C.c_long(d.shape[0]), _p(cb), cb.shape[0], threads, _p(idx)), idx

    def load_clusters(self, text):
        b = text.encode()
        K = self.L.orc_load_clusters(b, C.c_long(len(b)), None, 0)
        out = np.zeros((max(K, 1), 16), np.uint8)
        self.L.orc_load_clusters(b, C.c_long(len(b)), _p(out), K)
        return out[:K].copy()

    def build_histogram(self, text, codebook):
        cb = np.ascontiguousarray(codebook, np.uint8).reshape(-1, 16)
        b = text.encode()
        hist = np.zeros(cb.shape[0], np.float32)
        ok = self.L.orc_build_histogram(b, C.c_long(len(b)), _p(cb), cb.shape[0], _p(hist))
        return bool(ok), hist

    # ---- SURVEY 8(f) "next" rows ----
    def sliding_bow(self, text, codebook, alpha, label):
        cb = np.ascontiguousarray(codebook, np.uint8).reshape(-1, 16)
        b = text.encode()
        self.L.orc_sliding_bow.restype = C.c_long
        n = self.L.orc_sliding_bow(b, C.c_long(len(b)), _p(cb), cb.shape[0], alpha, label, None, C.c_long(0))
        buf = C.create_string_buffer(n + 1)
        self.L.orc_sliding_bow(b, C.c_long(len(b)), _p(cb), cb.shape[0], alpha, label, buf, C.c_long(n))
        return buf.raw[:n].decode()

    def random_clusters(self, data, num_clusters, num_classes, classes, only_one_class, rands, centers=None, center_row=0):
        """data: [rows, dim + 3] uint8 (copied, shuffled copy returned); rands: the numbers rand() would return."""
        d = np.ascontiguousarray(data, np.uint8).copy()
        rows, cols = d.shape
        cl = np.ascontiguousarray(classes, np.int32)
        r = np.ascontiguousarray(rands, np.int32)
        assert r.size >= 3 * max(rows - 1, 0)
        if centers is None:
            centers = np.zeros((num_clusters, cols - 3), np.uint8)
        cr = C.c_int(center_row)
        self.L.orc_random_clusters.restype = C.c_long
        used = self.L.orc_random_clusters(_p(d), rows, cols - 3, num_clusters, num_classes, _p(cl), cl.size, 1 if only_one_class else 0,
                                          _p(r), _p(centers), C.byref(cr))
        return centers, cr.value, int(used), d

    def clusters_text(self, centers):
        c = np.ascontiguousarray(centers, np.uint8)
        self.L.orc_clusters_text.restype = C.c_long
        n = self.L.orc_clusters_text(_p(c), c.shape[0], c.shape[1], None, C.c_long(0))
        buf = C.create_string_buffer(n + 1)
        self.L.orc_clusters_text(_p(c), c.shape[0], c.shape[1], buf, C.c_long(n))
        return buf.raw[:n].decode()


_cached = None


def load(prefer_ref=True):
    global _cached
    if _cached is not None:
        return _cached
    if not os.path.exists(PORT_LIB) or (os.path.isdir("/root/reference") and not os.path.exists(REF_LIB)):
        build("all")
    path = REF_LIB if (prefer_ref and os.path.exists(REF_LIB)) else PORT_LIB
    _cached = Oracle(path)
    return _cached


def absdiff(a, b):
    return np.abs(a.astype(np.int16) - b.astype(np.int16)).astype(np.uint8)
