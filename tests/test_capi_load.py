"""The C-ABI library: loads without a GPU, exports every symbol include/mofreak_b200.h declares, and refuses to
work (loudly, no fallback) when no CUDA device exists."""
import ctypes
import os
import re

import pytest

import mofreak_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(names=("mofreak_b200.h", "mofreak_b200_debug.h")):
    out = set()
    for name in names:
        text = open(os.path.join(ROOT, "include", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        out |= set(re.findall(r"\b(mfk_[a-z0-9_]+)\s*\(", text))
    return sorted(out)


def test_library_exports_every_declared_symbol():
    L = mofreak_b200.lib()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libmofreak_b200.so does not export %s" % n
    assert sorted(mofreak_b200.SYMBOLS) == names


def test_public_header_is_the_drop_in_surface_only():
    """Debug/profile/synthetic-data instruments live in mofreak_b200_debug.h / mofreak_synth.h, not in the public header."""
    public = header_symbols(("mofreak_b200.h",))
    assert not [n for n in public if n.startswith(("mfk_debug_", "mfk_profile_")) or n in ("mfk_set_serial", "mfk_synth_clip")]
    assert len(public) <= 33
    synth = ctypes.CDLL(mofreak_b200.capi.SYNTH_LIB_PATH)
    assert hasattr(synth, "mfk_synth_clip") and not hasattr(mofreak_b200.lib(), "mfk_synth_clip")


def test_row_layout_is_32_bytes():
    assert mofreak_b200.ROW_DTYPE.itemsize == 32
    assert mofreak_b200.ROW_DTYPE.fields["appearance"][1] == 16 and mofreak_b200.ROW_DTYPE.fields["motion"][1] == 24


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mofreak_b200.MofreakError) as e:
        mofreak_b200.Context(160, 120, 16, 8)
    assert e.value.code == mofreak_b200.capi.MFK_ERR_CUDA and "no CPU path" in str(e.value)


def test_argument_validation_happens_before_cuda():
    L = mofreak_b200.lib()
    h = ctypes.c_void_p()
    assert L.mfk_create(ctypes.byref(h), 0, 10, 10, 16, 8) == mofreak_b200.capi.MFK_ERR_ARG
    assert L.mfk_create(ctypes.byref(h), 0, 160, 120, 2, 8) == mofreak_b200.capi.MFK_ERR_ARG
    assert b"outside" in L.mfk_last_error(None) or b"required" in L.mfk_last_error(None)


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load, link or include anything of oracle/."""
    bad = re.compile(r'oracle_binding|libmofreak_oracle|libmofreak_ref|#\s*include\s*[<"][^>"]*oracle|import\s+oracle|from\s+oracle|dlopen\([^)]*oracle')
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mofreak_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(text), "%s reaches into oracle/" % os.path.join(dirpath, f)
