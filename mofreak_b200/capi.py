"""ctypes binding of include/mofreak_b200.h — harness glue for tests/ and bench.py.

The product is the C-ABI library (CUDA, sm_100a) and the C++ host classes in mofreak_b200/host/; this module
only loads the library and marshals numpy buffers.  There is no fallback of any kind: if the library is
missing or no CUDA device is usable, loading / context creation raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmofreak_b200.so")

MFK_OK, MFK_ERR_ARG, MFK_ERR_CUDA, MFK_ERR_CAPACITY, MFK_ERR_NOMEM, MFK_ERR_STATE, MFK_ERR_RANGE, MFK_ERR_BUDGET = 0, -1, -2, -3, -4, -5, -6, -7
FRAME_GAP = 5

ROW_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("frame", "<i4"), ("scale", "<f4"),
                      ("appearance", "u1", (8,)), ("motion", "u1", (8,))])
assert ROW_DTYPE.itemsize == 32

SYNTH_LIB_PATH = os.path.join(HERE, "synth", "libmofreak_synth.so")

# every symbol include/mofreak_b200.h (the drop-in surface) and include/mofreak_b200_debug.h (test and measurement
# instruments) declare; tests check that the library exports all of them
SYMBOLS = [
    "mfk_create", "mfk_create_ex", "mfk_destroy", "mfk_last_error", "mfk_sync", "mfk_host_alloc", "mfk_host_free", "mfk_upload_frames", "mfk_upload_frames_bgr", "mfk_frame_store",
    "mfk_extract_pairs", "mfk_fetch_rows", "mfk_rows_device", "mfk_reset_rows", "mfk_extract_clip", "mfk_rows_text", "mfk_format_rows", "mfk_extract_clip_text", "mfk_mip_describe",
    "mfk_bow_set_codebook", "mfk_bow_assign", "mfk_bow_assign_device", "mfk_bow_histogram_rows", "mfk_bow_histogram", "mfk_bow_histogram_videos",
    "mfk_allreduce_counts", "mfk_nccl_unique_id", "mfk_nccl_comm_init", "mfk_nccl_comm_destroy", "mfk_timer_start", "mfk_timer_stop", "mfk_profile_enable", "mfk_profile_read", "mfk_profile_timeline",
    "mfk_set_serial", "mfk_launch_count", "mfk_debug_layer", "mfk_debug_keypoints", "mfk_debug_counters", "mfk_freak_scale_index", "mfk_version",
]


class _Span(C.Structure):
    _fields_ = [("name", C.c_char_p), ("start_ms", C.c_float), ("end_ms", C.c_float)]


class MofreakError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("mofreak_b200 error %d: %s" % (code, message))
        self.code = code


class _ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_uint32), ("total_ms", C.c_float)]


_lib = None


def lib():
    """Loads mofreak_b200/libmofreak_b200.so (built by mofreak_b200/build.py).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m mofreak_b200.build` (or __graft_entry__.build()). "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    p, u8p, i32p, u32p, f32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    L.mfk_create.argtypes = [C.POINTER(p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mfk_create_ex.argtypes = [C.POINTER(p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mfk_destroy.argtypes = [p]; L.mfk_destroy.restype = None
    L.mfk_last_error.argtypes = [p]; L.mfk_last_error.restype = C.c_char_p
    L.mfk_sync.argtypes = [p]
    L.mfk_upload_frames.argtypes = [p, u8p, C.c_size_t, C.c_int, C.c_int]
    L.mfk_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]
    L.mfk_host_free.argtypes = [C.c_void_p]
    L.mfk_upload_frames_bgr.argtypes = [p, u8p, C.c_size_t, C.c_int, C.c_int, C.c_int]
    L.mfk_frame_store.argtypes = [p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.mfk_extract_pairs.argtypes = [p, i32p, i32p, i32p, C.c_int]
    L.mfk_fetch_rows.argtypes = [p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), u32p, C.c_size_t]
    L.mfk_rows_device.argtypes = [p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.mfk_reset_rows.argtypes = [p]
    L.mfk_extract_clip.argtypes = [p, u8p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.mfk_rows_text.argtypes = [p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.mfk_format_rows.argtypes = [p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.mfk_extract_clip_text.argtypes = [p, u8p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                        C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.mfk_mip_describe.argtypes = [p, C.c_int, C.c_int, f32p, C.c_int, u8p]
    L.mfk_bow_set_codebook.argtypes = [p, u8p, C.c_int]
    L.mfk_bow_assign.argtypes = [p, u8p, C.c_size_t, i32p]
    L.mfk_bow_assign_device.argtypes = [p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.mfk_bow_histogram_rows.argtypes = [p, u32p]
    L.mfk_bow_histogram.argtypes = [p, u8p, C.c_size_t, u32p]
    L.mfk_bow_histogram_videos.argtypes = [p, u32p, C.c_int, u32p, C.c_void_p]
    L.mfk_allreduce_counts.argtypes = [p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.mfk_nccl_unique_id.argtypes = [C.c_void_p]
    L.mfk_nccl_comm_init.argtypes = [p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    L.mfk_nccl_comm_destroy.argtypes = [C.c_void_p]
    L.mfk_timer_start.argtypes = [p]
    L.mfk_timer_stop.argtypes = [p, C.POINTER(C.c_float)]
    L.mfk_profile_enable.argtypes = [p, C.c_int]
    L.mfk_profile_read.argtypes = [p, C.POINTER(_ProfEntry), C.c_int, C.POINTER(C.c_int)]
    L.mfk_profile_timeline.argtypes = [p, C.POINTER(_Span), C.c_int, C.POINTER(C.c_int)]
    L.mfk_set_serial.argtypes = [p, C.c_int]
    L.mfk_launch_count.argtypes = [p]; L.mfk_launch_count.restype = C.c_uint64
    L.mfk_debug_layer.argtypes = [p, C.c_int, C.c_int, C.c_int, u8p, C.POINTER(C.c_int)]
    L.mfk_debug_keypoints.argtypes = [p, C.c_int, f32p, C.c_int, C.POINTER(C.c_int)]
    L.mfk_debug_counters.argtypes = [p, C.c_void_p]
    L.mfk_freak_scale_index.argtypes = [C.c_float]
    L.mfk_version.restype = C.c_char_p
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostBuffer:
    """Page-locked host memory from mfk_host_alloc as a uint8 numpy array (`.array`); freed on close()/GC."""

    def __init__(self, nbytes, write_combined=False):
        L = lib()
        p = C.c_void_p()
        rc = L.mfk_host_alloc(C.byref(p), nbytes, 1 if write_combined else 0)
        if rc != MFK_OK:
            raise MofreakError(rc, (L.mfk_last_error(None) or b"").decode())
        self._L, self.ptr, self.nbytes = L, p.value, nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            self._L.mfk_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_synth = None


def synth_clip(w, h, n_frames, seed, t0=0):
    """Deterministic "moving squares" clip, uint8 [n_frames, h, w] (mofreak_b200/synth/synth.cpp).  Lives in its own
    library (no CUDA), so generating frames does not map the product library into the calling process."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError("%s is missing: run `python -m mofreak_b200.build`" % SYNTH_LIB_PATH)
        _synth = C.CDLL(SYNTH_LIB_PATH)
        _synth.mfk_synth_clip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int]
        _synth.mfk_synth_clip.restype = None
    out = np.empty((n_frames, h, w), np.uint8)
    _synth.mfk_synth_clip(_ptr(out), w, h, n_frames, C.c_uint32(seed), t0)
    return out


def freak_scale_index(size):
    return lib().mfk_freak_scale_index(C.c_float(size))


class Context:
    """One mfk_ctx.  Thin: every method is one C-ABI call."""

    def __init__(self, width, height, store_frames, max_pairs, device=0, cand_cap_per_pair=0, row_cap_per_pair=0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.mfk_create_ex(C.byref(self._h), device, width, height, store_frames, max_pairs,
                                   cand_cap_per_pair, row_cap_per_pair)
        if rc != MFK_OK:
            raise MofreakError(rc, (self._L.mfk_last_error(None) or b"").decode())
        self.width, self.height, self.store_frames, self.max_pairs = width, height, store_frames, max_pairs

    def close(self):
        if getattr(self, "_h", None):
            self._L.mfk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != MFK_OK:
            raise MofreakError(rc, (self._L.mfk_last_error(self._h) or b"").decode())

    # --- frames ---
    def upload_frames(self, frames, slot0=0):
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 3 and frames.shape[1:] == (self.height, self.width)
        self._check(self._L.mfk_upload_frames(self._h, _ptr(frames), self.width, frames.shape[0], slot0))

    def upload_frames_bgr(self, frames_bgr, slot0=0, coeff_bits=14):
        """frames_bgr: uint8 [n, H, W, 3]; converted to gray on the device (cv::cvtColor(BGR2GRAY) arithmetic)."""
        f = np.ascontiguousarray(frames_bgr, np.uint8)
        assert f.ndim == 4 and f.shape[1:] == (self.height, self.width, 3)
        self._check(self._L.mfk_upload_frames_bgr(self._h, _ptr(f), self.width * 3, f.shape[0], slot0, coeff_bits))

    def upload_raw(self, host_ptr, n_frames, slot0=0):
        """host_ptr: integer address of (ideally pinned) dense frames."""
        self._check(self._L.mfk_upload_frames(self._h, C.c_void_p(host_ptr), self.width, n_frames, slot0))

    def frame_store(self):
        base, n = C.c_void_p(), C.c_int()
        self._check(self._L.mfk_frame_store(self._h, C.byref(base), C.byref(n)))
        return base.value, n.value

    # --- extraction ---
    def extract_pairs(self, cur, prev, label):
        cur = np.ascontiguousarray(cur, np.int32); prev = np.ascontiguousarray(prev, np.int32)
        label = np.ascontiguousarray(label, np.int32)
        assert cur.shape == prev.shape == label.shape
        self._check(self._L.mfk_extract_pairs(self._h, _ptr(cur), _ptr(prev), _ptr(label), cur.size))

    def fetch_rows(self, cap, n_pairs=0, out=None):
        rows = out if out is not None else np.empty(cap, ROW_DTYPE)
        per_pair = np.zeros(n_pairs, np.uint32)
        n = C.c_size_t()
        self._check(self._L.mfk_fetch_rows(self._h, _ptr(rows), cap, C.byref(n), _ptr(per_pair) if n_pairs else None, n_pairs))
        return rows[:n.value], per_pair

    def fetch_rows_raw(self, host_ptr, cap):
        n = C.c_size_t()
        self._check(self._L.mfk_fetch_rows(self._h, C.c_void_p(host_ptr), cap, C.byref(n), None, 0))
        return n.value

    def rows_device(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.mfk_rows_device(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def reset_rows(self):
        self._check(self._L.mfk_reset_rows(self._h))

    def extract_clip(self, frames, cap=None):
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 3 and frames.shape[1:] == (self.height, self.width)
        if cap is None:
            cap = max(1024, frames.shape[0] * 4096)
        rows = np.empty(cap, ROW_DTYPE)
        n = C.c_size_t()
        self._check(self._L.mfk_extract_clip(self._h, _ptr(frames), self.width, frames.shape[0], _ptr(rows), cap, C.byref(n)))
        return rows[:n.value].copy()

    # --- `.mofreak` text, formatted on the device ---
    def rows_text(self):
        """Text of the accumulated rows (bytes); the accumulation is kept."""
        n = C.c_size_t()
        rc = self._L.mfk_rows_text(self._h, None, 0, C.byref(n))
        if rc != 0 and not (rc == -3 and n.value > 0):
            self._check(rc)
        buf = np.empty(max(1, n.value), np.uint8)
        self._check(self._L.mfk_rows_text(self._h, _ptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].tobytes()

    def rows_text_raw(self, host_ptr, cap):
        n = C.c_size_t()
        self._check(self._L.mfk_rows_text(self._h, C.c_void_p(host_ptr), cap, C.byref(n)))
        return n.value

    def format_rows(self, rows):
        rows = np.ascontiguousarray(rows, ROW_DTYPE)
        buf = np.empty(max(1, rows.size * 128), np.uint8)
        n = C.c_size_t()
        self._check(self._L.mfk_format_rows(self._h, _ptr(rows), rows.size, _ptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].tobytes()

    def extract_clip_text(self, frames, want_rows=True, cap=None):
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 3 and frames.shape[1:] == (self.height, self.width)
        if cap is None:
            cap = max(1024, frames.shape[0] * 4096)
        rows = np.empty(cap, ROW_DTYPE) if want_rows else None
        text = np.empty(cap * 128, np.uint8)
        n, nb = C.c_size_t(), C.c_size_t()
        self._check(self._L.mfk_extract_clip_text(self._h, _ptr(frames), self.width, frames.shape[0], _ptr(text), text.size, C.byref(nb),
                                                  _ptr(rows) if want_rows else None, cap if want_rows else 0, C.byref(n)))
        return text[:nb.value].tobytes(), (rows[:n.value].copy() if want_rows else n.value)

    def mip_describe(self, cur_slot, prev_slot, kp_xys):
        kp = np.ascontiguousarray(kp_xys, np.float32).reshape(-1, 3)
        out = np.zeros((kp.shape[0], 8), np.uint8)
        self._check(self._L.mfk_mip_describe(self._h, cur_slot, prev_slot, _ptr(kp), kp.shape[0], _ptr(out)))
        return out

    # --- bag of words ---
    def bow_set_codebook(self, codewords):
        cw = np.ascontiguousarray(codewords, np.uint8).reshape(-1, 16)
        self._check(self._L.mfk_bow_set_codebook(self._h, _ptr(cw), cw.shape[0]))
        self.K = cw.shape[0]

    def bow_assign(self, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 16)
        idx = np.empty(d.shape[0], np.int32)
        self._check(self._L.mfk_bow_assign(self._h, _ptr(d), d.shape[0], _ptr(idx)))
        return idx

    def bow_assign_device(self, d_desc_ptr, n, d_idx_ptr):
        self._check(self._L.mfk_bow_assign_device(self._h, C.c_void_p(d_desc_ptr), n, C.c_void_p(d_idx_ptr)))

    def bow_histogram(self, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 16)
        counts = np.zeros(self.K, np.uint32)
        self._check(self._L.mfk_bow_histogram(self._h, _ptr(d), d.shape[0], _ptr(counts)))
        return counts

    def bow_histogram_rows(self):
        counts = np.zeros(self.K, np.uint32)
        self._check(self._L.mfk_bow_histogram_rows(self._h, _ptr(counts)))
        return counts

    def bow_histogram_videos(self, pairs_per_video=None, comm=None):
        """[n_videos, K] word counts of the accumulated rows, one vector per group of consecutive queued pairs; with `comm`
        the vectors are summed over the ranks (collective)."""
        if pairs_per_video is None:
            ppv, n = None, 1
        else:
            ppv = np.ascontiguousarray(pairs_per_video, np.uint32); n = ppv.size
        counts = np.zeros((n, self.K), np.uint32)
        self._check(self._L.mfk_bow_histogram_videos(self._h, _ptr(ppv) if ppv is not None else None, n, _ptr(counts), comm))
        return counts

    # --- multi-GPU exchange of word counts ---
    @staticmethod
    def nccl_unique_id():
        buf = np.zeros(128, np.uint8)
        rc = lib().mfk_nccl_unique_id(_ptr(buf))
        if rc != MFK_OK:
            raise MofreakError(rc, (lib().mfk_last_error(None) or b"").decode())
        return buf

    def nccl_comm_init(self, n_ranks, rank, unique_id):
        comm = C.c_void_p()
        uid = np.ascontiguousarray(unique_id, np.uint8)
        self._check(self._L.mfk_nccl_comm_init(self._h, C.byref(comm), n_ranks, rank, _ptr(uid)))
        return comm

    def allreduce_counts(self, d_counts_ptr, n, comm):
        self._check(self._L.mfk_allreduce_counts(self._h, C.c_void_p(d_counts_ptr), n, comm))

    # --- measurement ---
    def sync(self):
        self._check(self._L.mfk_sync(self._h))

    def timer_start(self):
        self._check(self._L.mfk_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self._L.mfk_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._check(self._L.mfk_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        ent = (_ProfEntry * 16)()
        n = C.c_int()
        self._check(self._L.mfk_profile_read(self._h, ent, 16, C.byref(n)))
        return {ent[i].name.decode(): (ent[i].launches, ent[i].total_ms) for i in range(n.value)}

    def profile_timeline(self, cap=65536):
        """[(family, start_ms, end_ms)] of the recorded spans, relative to the last timer_start; call before profile_read."""
        ent = (_Span * cap)()
        n = C.c_int()
        self._check(self._L.mfk_profile_timeline(self._h, ent, cap, C.byref(n)))
        return [(ent[i].name.decode(), ent[i].start_ms, ent[i].end_ms) for i in range(n.value)]

    def set_serial(self, on=True):
        self._check(self._L.mfk_set_serial(self._h, 1 if on else 0))

    def launch_count(self):
        return int(self._L.mfk_launch_count(self._h))

    # --- stage access ---
    def debug_layer(self, what, layer, pair):
        dims = (C.c_int * 2)()
        buf = np.zeros(self.width * self.height, np.uint8)
        self._check(self._L.mfk_debug_layer(self._h, what, layer, pair, _ptr(buf), dims))
        return buf[:dims[0] * dims[1]].reshape(dims[1], dims[0]).copy()

    def debug_counters(self):
        out = np.zeros(8, np.uint32)
        self._check(self._L.mfk_debug_counters(self._h, _ptr(out)))
        return out

    def debug_keypoints(self, pair, cap=65536):
        out = np.zeros((cap, 5), np.float32)
        n = C.c_int()
        self._check(self._L.mfk_debug_keypoints(self._h, pair, _ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()
