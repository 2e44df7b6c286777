import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ROW_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("frame", "<i4"), ("scale", "<f4"),
                      ("appearance", "u1", (8,)), ("motion", "u1", (8,))])


def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def golden_rows(name):
    return np.fromfile(os.path.join(GOLDEN, name + ".rows.bin"), ROW_DTYPE)


def golden_text(name):
    return open(os.path.join(GOLDEN, name + ".mofreak")).read()


def rows_equal(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def describe_diff(a, b):
    if a.shape != b.shape:
        return "row counts differ: %d vs %d" % (len(a), len(b))
    out = []
    for f in ROW_DTYPE.names:
        bad = (a[f] != b[f]).reshape(len(a), -1).any(1).sum()
        if bad:
            out.append("%s: %d rows" % (f, bad))
    return "; ".join(out) or "identical"


def diff_frames(clip):
    """(c, |I_c - I_{c-5}|) for every processed frame of a clip."""
    for c in range(5, clip.shape[0]):
        yield c, np.abs(clip[c].astype(np.int16) - clip[c - 5].astype(np.int16)).astype(np.uint8)


def kp_equal(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


_harness = None


def harness():
    """CPU build of the product's device source (tests/host_harness.cpp)."""
    global _harness
    if _harness is None:
        path = os.path.join(HERE, "libhost_harness.so")
        import subprocess
        subprocess.run(["make", "-s", "-C", HERE], check=True)      # always: a no-op when the library is newer than the device sources
        L = C.CDLL(path)
        L.hh_atan2f_sweep.restype = C.c_long
        L.hh_atan2f.restype = C.c_float
        L.hh_atan2f.argtypes = [C.c_float, C.c_float]
        L.hh_scale_index_table.argtypes = [C.c_float]
        L.hh_scale_index_direct.argtypes = [C.c_float]
        _harness = L
    return _harness


def harness_detect(diff, reverse_rounds=False, prune_tile_losers=False):
    L = harness()
    diff = np.ascontiguousarray(diff, np.uint8)
    h, w = diff.shape
    out = np.zeros((1 << 16, 6), np.float32)
    n = L.hh_detect(diff.ctypes.data_as(C.c_void_p), w, h, (1 if reverse_rounds else 0) | (2 if prune_tile_losers else 0), out.ctypes.data_as(C.c_void_p), 1 << 16)
    return out[:n].copy()


def harness_extract_clip(clip):
    L = harness()
    n, h, w = clip.shape
    rows = []
    buf = np.zeros(1 << 16, ROW_DTYPE)
    for c in range(5, n):
        cur = np.ascontiguousarray(clip[c]); prev = np.ascontiguousarray(clip[c - 5])
        k = L.hh_extract_pair(cur.ctypes.data_as(C.c_void_p), prev.ctypes.data_as(C.c_void_p), w, h, c - 1,
                              buf.ctypes.data_as(C.c_void_p), 1 << 16)
        rows.append(buf[:k].copy())
    return np.concatenate(rows) if rows else np.zeros(0, ROW_DTYPE)
