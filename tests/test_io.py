"""CPU tests of the sample-file I/O (include/pcq_io.h, pytuq_b200/utils/io.py): the oracle is
numpy itself — np.savetxt / np.loadtxt with default options are what the reference's applications call
(apps/uqpc/uq_pc.py:202-271).  Bars: written files byte-identical, read values bit-identical, same shapes and the
same exception types."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import ROOT
from pytuq_b200.utils import io


def _same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(np.atleast_1d(a).view(np.uint64) | 0, np.atleast_1d(b).view(np.uint64) | 0)


def _nan_payload_free(a):
    # numpy parses 'nan' to the default quiet nan; so do we — compare nans by position only
    a = np.array(a, dtype=np.float64, copy=True)
    a[np.isnan(a)] = np.nan
    return a


def test_io_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pcq_io.h")).read()
    declared = set(re.findall(r"\b(pcqio_[a-z0-9_]+)\s*\(", header))
    assert declared == set(io.SIGNATURES), declared ^ set(io.SIGNATURES)
    handle = ctypes.CDLL(io.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert io.lib().pcqio_version() == 100


def test_conversions_agree_with_libc_on_random_values():
    slow_f, slow_p = ctypes.c_int64(), ctypes.c_int64()
    n = 400000
    for seed in (1, 2):
        bad = io.lib().pcqio_selftest(seed, n, ctypes.byref(slow_f), ctypes.byref(slow_p))
        assert bad == 0
        # the libc route is the exception, not the rule (ties, subnormal results, inf/nan)
        assert slow_f.value < 0.02 * n and slow_p.value < 0.1 * 3 * n


CASES = {
    "wide_range": lambda rng: rng.standard_normal((257, 7)) * 10.0 ** rng.integers(-300, 300, (257, 7)),
    "vector": lambda rng: rng.standard_normal(33),
    "one_row": lambda rng: rng.standard_normal((1, 5)),
    "one_value": lambda rng: rng.standard_normal((1, 1)),
    "one_column": lambda rng: rng.standard_normal((9, 1)),
    "specials": lambda rng: np.array([[0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1.7976931348623157e308,
                                       2.2250738585072014e-308, 0.1, 1e22, 1e23, 9.999999999999999e22, 0.5,
                                       2.5, 1.0, 1e-5, 123456789.0, 9007199254740993.0]]),
    "integers": lambda rng: np.arange(-6, 6).reshape(3, 4),
    "float32": lambda rng: rng.standard_normal((4, 3)).astype(np.float32),
    "fortran_order": lambda rng: np.asfortranarray(rng.standard_normal((11, 6))),
    "strided": lambda rng: rng.standard_normal((40, 12))[::3, 1::2],
    "no_rows": lambda rng: np.zeros((0, 3)),
    "ties_at_digit_20": lambda rng: (2.0 * rng.integers(10 ** 8, 10 ** 9, 64) + 1) * 5.0 ** 10 / 2.0 ** 11 / 10.0 ** 10,
    "many_blocks": lambda rng: rng.uniform(-1, 1, (30000, 10)),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads", [1, 3])
def test_savetxt_bytes_and_loadtxt_bits_match_numpy(tmp_path, name, threads):
    X = CASES[name](np.random.default_rng(sorted(CASES).index(name)))
    f_np, f_us = str(tmp_path / "numpy.txt"), str(tmp_path / "native.txt")
    np.savetxt(f_np, X)
    io.savetxt(f_us, X, threads=threads)
    assert open(f_np, "rb").read() == open(f_us, "rb").read()
    for ndmin in (0, 1, 2):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # "input contained no data" for the empty case
            want = np.loadtxt(f_np, ndmin=ndmin)
            got = io.loadtxt(f_np, ndmin=ndmin, threads=threads)
        assert got.dtype == np.float64 and got.flags.c_contiguous or got.ndim == 0
        assert _same_bits(_nan_payload_free(want), _nan_payload_free(got)), (ndmin, want.shape, got.shape)
    if X.size and name != "specials":
        back = io.loadtxt(f_us, ndmin=2)
        assert np.array_equal(back, np.asarray(X, dtype=np.float64).reshape(back.shape))   # 19 digits round-trip


def test_loadtxt_accepts_what_numpy_accepts(tmp_path):
    text = ("# header line\n"
            "1 2\t3.5   # trailing comment\n"
            "\n"
            "   \t \n"
            "-4e0 +5E-1 .25\r\n"
            "1. inf -Infinity\n"
            "NaN -0 1e400\n"
            "1e-400 123456789012345678901234567890 0.000000000000000000000000000001\n"
            "7 8 9")                                   # no newline at the end
    f = tmp_path / "mixed.txt"
    f.write_bytes(text.encode())
    want = np.loadtxt(str(f))
    got = io.loadtxt(str(f))
    assert want.shape == got.shape == (6, 3)
    assert _same_bits(_nan_payload_free(want), _nan_payload_free(got))
    # the reshape idiom of the applications (uq_pc.py:232-233)
    assert np.array_equal(io.loadtxt(str(f)).reshape(-1, 3), want, equal_nan=True)


@pytest.mark.parametrize("text, exc", [
    ("1 2 3\n4 5\n", ValueError),                 # ragged
    ("1 2\n3 abc\n", ValueError),                 # not a number
    ("1 2\n3 0x10\n", ValueError),                # hex is not accepted by numpy's float converter
    ("1 2\n3 1_0\n", ValueError),
    ("1 2\n3 1e\n", ValueError),
    ("1,2\n3,4\n", ValueError),                   # a delimiter other than whitespace
])
def test_loadtxt_rejects_what_numpy_rejects(tmp_path, text, exc):
    f = tmp_path / "bad.txt"
    f.write_text(text)
    with pytest.raises(exc):
        np.loadtxt(str(f))
    with pytest.raises(exc):
        io.loadtxt(str(f))


def test_ragged_rows_are_found_in_any_segment(tmp_path):
    rows = ["%d %d %d" % (i, i + 1, i + 2) for i in range(200000)]     # > 1 MB: several segments
    rows[150001] = "1 2"
    f = tmp_path / "ragged.txt"
    f.write_text("\n".join(rows) + "\n")
    with pytest.raises(ValueError, match="150002"):
        io.loadtxt(str(f), threads=4)
    rows[150001] = "1 2 x"
    f.write_text("\n".join(rows) + "\n")
    with pytest.raises(ValueError, match="'x'.*150002"):
        io.loadtxt(str(f), threads=4)


def test_errors_and_unsupported_inputs(tmp_path):
    with pytest.raises(FileNotFoundError):
        np.loadtxt(str(tmp_path / "missing.txt"))
    with pytest.raises(FileNotFoundError):
        io.loadtxt(str(tmp_path / "missing.txt"))
    with pytest.raises(OSError):
        io.savetxt(str(tmp_path / "no_such_dir" / "x.txt"), np.zeros(3))
    with pytest.raises(ValueError):
        np.savetxt(str(tmp_path / "x.txt"), np.zeros((2, 2, 2)))
    with pytest.raises(ValueError):
        io.savetxt(str(tmp_path / "x.txt"), np.zeros((2, 2, 2)))
    with pytest.raises(ValueError):
        io.savetxt(str(tmp_path / "x.txt"), np.float64(1.0))
    with pytest.raises(TypeError):
        io.savetxt(str(tmp_path / "x.txt"), np.zeros(2, dtype=complex))
    empty = tmp_path / "empty.txt"
    empty.write_text("# nothing here\n\n")
    with pytest.warns(UserWarning, match="no data"):
        assert io.loadtxt(str(empty)).shape == (0,)
    # C-ABI argument checks: status + message, no exception across the boundary
    lib = io.lib()
    assert lib.pcqio_savetxt(None, None, 1, 1, 1, 0) == io.ERR_ARG
    assert lib.pcqio_savetxt(b"/tmp/x", None, 2, 2, 1, 0) == io.ERR_ARG
    assert b"bad argument" in lib.pcqio_last_error()
    assert lib.pcqio_read(None, None, 0) == io.ERR_ARG
    assert lib.pcqio_close(None) == 0


def test_savetxt_takes_torch_tensors(tmp_path):
    import torch
    X = np.random.default_rng(6).standard_normal((17, 3))
    io.savetxt(str(tmp_path / "t.txt"), torch.from_numpy(X))
    np.savetxt(str(tmp_path / "n.txt"), X)
    assert open(tmp_path / "t.txt", "rb").read() == open(tmp_path / "n.txt", "rb").read()


def test_save_load_pick_the_format_from_the_name(tmp_path):
    X = np.random.default_rng(5).standard_normal((50, 4))
    for name in ("x.npy", "x.txt", "x.dat"):
        io.save(str(tmp_path / name), X)
        assert np.array_equal(io.load(str(tmp_path / name)), X)
    assert open(tmp_path / "x.txt", "rb").read() == open(tmp_path / "x.dat", "rb").read()
    assert np.array_equal(np.load(str(tmp_path / "x.npy")), X)
    assert isinstance(io.load(str(tmp_path / "x.npy"), mmap_mode="r"), np.memmap)
    io.save(str(tmp_path / "v.npy"), X[:, 0])
    assert io.load(str(tmp_path / "v.npy"), ndmin=2).shape == (50, 1)
    io.save(str(tmp_path / "v.txt"), X[:, 0])
    assert io.load(str(tmp_path / "v.txt"), ndmin=2).shape == (50, 1)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(allow_nan=False, allow_infinity=True, width=64), min_size=1, max_size=24))
def test_round_trip_of_arbitrary_doubles(values):
    import tempfile
    X = np.array(values, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        f_np, f_us = os.path.join(d, "a"), os.path.join(d, "b")
        np.savetxt(f_np, X)
        io.savetxt(f_us, X)
        assert open(f_np, "rb").read() == open(f_us, "rb").read()
        assert _same_bits(io.loadtxt(f_us, ndmin=1), X)
        assert _same_bits(np.loadtxt(f_us, ndmin=1), X)


def test_germ_sample_files_of_the_uqpc_app_are_reproduced_byte_for_byte(tmp_path):
    """apps/uqpc/uq_pc.py:206-207,224-226: seeded germ samples written with savetxt.  The germ stream is the host's
    numpy stream (no GPU), so this part of the app's file exchange is checked here; the parameter files, which go
    through evalPC, are checked on the GPU (tests/test_gpu_parity.py::test_uqpc_app_replay_through_sample_files)."""
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.utils.mindex import get_mi
    gold = os.path.join(ROOT, "tests", "golden", "uqpc")
    pdom = io.loadtxt(os.path.join(gold, "pdomain.txt")).reshape(-1, 2)
    assert np.array_equal(pdom, np.loadtxt(os.path.join(gold, "pdomain.txt")).reshape(-1, 2))
    pcf_all = np.vstack((0.5 * (pdom[:, 1] + pdom[:, 0]), np.diag(0.5 * (pdom[:, 1] - pdom[:, 0]))))
    pc = PCRV(3, 3, "LU", mi=get_mi(1, 3), cfs=pcf_all.T)
    for name, n in (("qtrain.txt", 240), ("qtest.txt", 50)):
        io.savetxt(str(tmp_path / name), pc.sampleGerm(n, seed=7))
        assert open(tmp_path / name, "rb").read() == open(os.path.join(gold, name), "rb").read()
    for name in ("ptrain.txt", "ytrain.txt", "qtrain.txt"):
        want = np.loadtxt(os.path.join(gold, name))
        assert np.array_equal(io.loadtxt(os.path.join(gold, name)), want)


_sep = st.sampled_from([" ", "  ", "\t", " \t "])
_eol = st.sampled_from(["\n", "\r\n", "\n\n", " \n", "\n \t\n"])
_num = st.one_of(
    st.floats(allow_nan=False, allow_infinity=False, width=64).map(repr),
    st.floats(allow_nan=False, allow_infinity=False, width=64).map(lambda v: "%.18e" % v),
    st.integers(-10 ** 25, 10 ** 25).map(str),
    st.sampled_from(["inf", "-inf", "+Infinity", "nan", "NaN", ".5", "5.", "+0", "-0.0", "1E5", "1e-320", "00012"]),
)


@settings(max_examples=150, deadline=None)
@given(st.integers(1, 5), st.integers(1, 6), st.data())
def test_arbitrary_layouts_read_like_numpy(ncols, nrows, data):
    import tempfile
    lines = []
    for _ in range(nrows):
        toks = [data.draw(_num) for _ in range(ncols)]
        seps = [data.draw(_sep) for _ in range(ncols - 1)]
        line = data.draw(st.sampled_from(["", " ", "\t"])) + "".join(t + s for t, s in zip(toks, seps + [""]))
        line += data.draw(st.sampled_from(["", "  # note 1 2 3", "#x"]))
        lines.append(line + data.draw(_eol))
    text = data.draw(st.sampled_from(["", "# head\n", "\n"])) + "".join(lines)
    if data.draw(st.booleans()):
        text = text.rstrip("\r\n \t")                      # no newline at the end
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "t.txt")
        with open(f, "wb") as fh:
            fh.write(text.encode())
        want = np.loadtxt(f, ndmin=2)
        got = io.loadtxt(f, ndmin=2)
    assert _same_bits(_nan_payload_free(want), _nan_payload_free(got)), text
