"""`.mofreak` text formatter (mofreak_b200/csrc/text_device.cuh).

CPU side: the device source compiled as plain C++ (tests/host_harness.cpp) against glibc/Python "%g" and against
the golden `.mofreak` files.  GPU side: the kernels through the C ABI (mfk_format_rows, mfk_rows_text,
mfk_extract_clip_text) against the same."""
import ctypes as C

import numpy as np
import pytest

import mofreak_b200
from tests import util

CASES = util.golden_cases()


def python_text(rows):
    out = []
    for r in rows:
        tok = ["%g" % float(r["x"]), "%g" % float(r["y"]), "%d" % int(r["frame"]), "%g" % float(r["scale"]), "0", "0"]
        tok += ["%d" % v for v in r["appearance"]] + ["%d" % v for v in r["motion"]]
        out.append(" ".join(tok) + " \n")
    return "".join(out)


def random_rows(n, seed):
    rng = np.random.default_rng(seed)
    rows = np.zeros(n, util.ROW_DTYPE)
    # bit patterns uniform over [1e-4, 1e6): every binade, i.e. every digit layout "%g" has in fixed notation
    for f in ("x", "y", "scale"):
        rows[f] = rng.integers(0x38D1B718, 0x49742400, n, dtype=np.uint32).view(np.float32)
    rows["x"][::7] = np.round(rows["x"][::7] * 4) / 4             # short decimals: trailing-zero stripping
    rows["y"][::5] = np.floor(rows["y"][::5])
    rows["scale"][::3] = np.float32(12.0) * rng.choice([1, 1.5, 2, 3, 4, 6], len(rows["scale"][::3])).astype(np.float32)
    rows["x"][::11] *= -1
    rows["frame"] = rng.integers(-2**31, 2**31 - 1, n)
    rows["appearance"] = rng.integers(0, 256, (n, 8))
    rows["motion"] = rng.integers(0, 256, (n, 8))
    return rows


EDGE_FLOATS = [0.0, -0.0, 1.0, 0.5, 0.1, 0.3, 12.0, 18.0, 719.5, 123.456, 99999.95, 999999.4, 123456.5, 123457.5, 100000.0,
               0.00010000001, 0.000123456789, 0.001, 0.01, 9.9999, 9.99999, 9.999995, 99.99996, 0.99999994, 0.0009999999,
               1.5, 2.5, 999999.4375, 8191.9995, 16.000002, 4096.0, 65536.0, 524288.0, 2.0 ** -13, 2.0 ** -10 * 1.5]


def harness_float(v):
    buf = C.create_string_buffer(64)
    L = util.harness()
    L.hh_format_float.argtypes = [C.c_float, C.c_char_p]
    n = L.hh_format_float(C.c_float(v), buf)
    return None if n < 0 else buf.raw[:n].decode()


def harness_text(rows, count_only=False):
    L = util.harness()
    L.hh_format_rows.argtypes = [C.c_void_p, C.c_long, C.c_char_p, C.c_int]
    L.hh_format_rows.restype = C.c_long
    rows = np.ascontiguousarray(rows)
    buf = C.create_string_buffer(max(1, len(rows) * 128))
    n = L.hh_format_rows(rows.ctypes.data_as(C.c_void_p), len(rows), buf, int(count_only))
    assert n >= 0, "row %d not formattable" % (-n - 1)
    return n if count_only else buf.raw[:n].decode()


def test_percent_g_edge_cases():
    for v in EDGE_FLOATS:
        v = float(np.float32(v))
        assert harness_float(v) == "%g" % v, v


def test_percent_g_out_of_range_is_reported_not_misprinted():
    for v in [1e6, 999999.97, 999999.9375, 999999.5, 1e7, 3e38, float("inf"), float("nan"), 9.9e-5, 1e-10, 1e-40, float(np.float32(1e-4))]:
        assert harness_float(v) is None, v


def test_percent_g_random_bit_patterns():
    rng = np.random.default_rng(5)
    vals = rng.integers(0x38D1B718, 0x49742400, 200000, dtype=np.uint32).view(np.float32)
    for v in vals:
        v = float(v)
        assert harness_float(v) == "%g" % v, v


def test_exact_ties_round_half_even():
    # floats that sit exactly between two 6-digit decimals
    for v, want in [(123456.5, "123456"), (123457.5, "123458"), (0.5, "0.5"), (1024.0625, "1024.06"), (2048.1875, "2048.19"),
                    (512.03125, "512.031"), (65536.25, "65536.2"), (65536.75, "65536.8")]:
        assert harness_float(v) == want == "%g" % v


def test_rows_text_matches_python_and_counts():
    rows = random_rows(20000, 3)
    text = harness_text(rows)
    assert text == python_text(rows)
    assert harness_text(rows, count_only=True) == len(text)
    assert max(len(l) for l in text.split("\n")) + 1 <= 128


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_files_from_golden_rows(name):
    assert harness_text(util.golden_rows(name)) == util.golden_text(name)


# ---- GPU ----------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_format_rows_random():
    rows = random_rows(300001, 9)
    with mofreak_b200.Context(160, 120, 8, 2) as ctx:
        before = ctx.launch_count()
        text = ctx.format_rows(rows)
        assert ctx.launch_count() - before == 5
        assert text.decode() == python_text(rows)
        assert ctx.format_rows(rows[:1]).decode() == python_text(rows[:1])
        assert ctx.format_rows(rows[:0]) == b""
        assert ctx.format_rows(rows[:33]).decode() == python_text(rows[:33])


@pytest.mark.gpu
def test_gpu_format_rows_reports_exponent_notation():
    rows = random_rows(100, 1)
    rows["scale"][57] = 1e7
    with mofreak_b200.Context(160, 120, 8, 2) as ctx:
        with pytest.raises(mofreak_b200.MofreakError) as e:
            ctx.format_rows(rows)
        assert e.value.code == -6


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_clip_text_is_the_golden_file(name):
    c = CASES[name]
    clip = mofreak_b200.synth_clip(c["width"], c["height"], c["frames"], c["seed"])
    with mofreak_b200.Context(c["width"], c["height"], 16, 8) as ctx:
        text, rows = ctx.extract_clip_text(clip)
        assert text.decode() == util.golden_text(name)
        assert util.rows_equal(rows, util.golden_rows(name))
        text2, n = ctx.extract_clip_text(clip, want_rows=False)          # text only; the accumulation is reset either way
        assert text2 == text and n == len(rows)
        # rows_text keeps the accumulation: fetch afterwards still sees the rows
        ctx.upload_frames(clip[:7])
        ctx.extract_pairs([5, 6], [0, 1], [4, 5])
        t = ctx.rows_text()
        got, _ = ctx.fetch_rows(1 << 16)
        assert t.decode() == python_text(got) and len(got) > 0
        assert ctx.rows_text() == b""


def test_percent_g_whole_float_space_is_exact_or_refused():
    """Any bit pattern: either the formatter refuses (exponent notation / non-finite -> host fallback) or its text is glibc's."""
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2**32, 100000, dtype=np.uint64).astype(np.uint32)
    refused = 0
    for v in bits.view(np.float32):
        got = harness_float(float(v))
        if got is None:
            refused += 1
            want = "%g" % float(v)
            assert ("e" in want) or ("n" in want) or float(v) == np.float32(1e-4), (v, want)      # inf/nan, exponent form, or the one float just below 1e-4
        else:
            assert got == "%g" % float(v), v
    assert 0 < refused < len(bits)
