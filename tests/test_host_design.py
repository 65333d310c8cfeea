"""Host side of libllzfilter_cuda (llz_design.c): tap design, estimators, polyphase plans.

These functions run on the CPU in the product too (they are the reference's init-time code,
SURVEY.md section 8a rows a1-a11), so they are tested without a GPU, bit for bit against the oracle.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import KIND, WIN


class Plan(C.Structure):
    """Mirror of llz_plan_t (llzlab_b200/csrc/llz_internal.h)."""
    _dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    _fields_ = [("kind", C.c_int), ("L", C.c_int), ("M", C.c_int), ("n", C.c_int), ("rows", C.c_int),
                ("cols", C.c_int), ("num_in", C.c_int), ("num_out", C.c_int), ("proto", _dp), ("bank", _dp),
                ("crows", C.c_int), ("ctaps", C.c_int), ("shift", C.c_int), ("frame_len", C.c_int),
                ("hist_len", C.c_int), ("cbank", _dp), ("order", _ip), ("single_tap", _ip),
                ("abs_row_sum", C.c_double)]


def build_plan(z, kind, L, M, win, k_override=0):
    lib = z.lib()
    lib.llz_plan_build.argtypes = [C.POINTER(Plan), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.llz_plan_free.argtypes = [C.POINTER(Plan)]
    lib.llz_plan_free.restype = None
    p = Plan()
    rc = lib.llz_plan_build(C.byref(p), kind, L, M, win, k_override)
    return (p if rc == 0 else None), lib


def test_windows_and_estimators(zlib, port, kat):
    for c in kat["windows"]:
        if c["win"] == "KAISER_BETA":
            w = zlib.kaiser_beta(c["N"], c["beta"])
            assert w.tobytes() == port.kaiser_beta(c["N"], c["beta"]).tobytes()
        else:
            w = zlib.window(c["N"], WIN[c["win"]])
            assert w.tobytes() == port.window(c["N"], WIN[c["win"]]).tobytes()
        assert f"{port.fnv64(w):016x}" == c["fnv"]
    for c in kat["cof_num"]:
        assert zlib.cof_num(WIN[c["win"]], c["ftrans"], c["atten"]) == c["n"]
    for c in kat["atten2beta"]:
        assert zlib.lib().llz_kaiser_atten2beta(c["atten"]) == c["beta"]


def test_designs_bit_identical(zlib, port, kat):
    for c in kat["designs"]:
        h = zlib.fir_design(KIND[c["kind"]], c["N"], c["fc1"], c["fc2"], WIN[c["win"]])
        assert len(h) == c["n_used"]
        assert f"{port.fnv64(h):016x}" == c["fnv"], c
    for win in range(3):
        for kind in range(4):
            for N in (2, 3, 8, 63, 64, 255, 1000):
                for f1, f2 in ((0.05, 0.2), (0.3, 0.7), (0.5, 0.99)):
                    a = zlib.fir_design(kind, N, f1, f2, win)
                    b = port.fir_design(kind, N, f1, f2, win)
                    assert a.tobytes() == b.tobytes(), (kind, N, f1, f2, win)


def test_design_rejects_bad_arguments(zlib):
    hp = C.POINTER(C.c_double)()
    assert zlib.lib().llz_fir_lpf_cof(C.byref(hp), 1, 0.2, 0) == -1
    assert zlib.lib().llz_fir_lpf_cof(C.byref(hp), 16, 0.2, 7) == -1
    assert "window" in zlib.last_error() or "bad" in zlib.last_error()


def test_llz_conv_matches_reference_order(zlib, port):
    x = port.lcg_f64(300, 3)
    h = port.fir_design(0, 127, 0.23, 0.0, 0)
    want = port.fir_run(h, x)
    for t in (126, 200, 299):
        assert zlib.conv(x, t, h) == want[t]


@pytest.mark.parametrize("L,M,win,k", [(160, 147, 1, 0), (1, 3, 1, 0), (320, 147, 1, 0), (147, 160, 1, 0),
                                       (3, 2, 2, 0), (2, 1, 0, 0), (320, 147, 1, 128), (7, 5, 0, 3)])
def test_resample_plan(zlib, port, L, M, win, k):
    p, lib = build_plan(zlib, 2, L, M, win, k)
    o = port.resample_plan(L, M, win, k)
    assert (p.n, p.rows, p.cols, p.num_in, p.num_out) == (o.n, o.rows, o.cols, o.num_in, o.num_out)
    proto = np.ctypeslib.as_array(p.proto, shape=(p.n,))
    bank = np.ctypeslib.as_array(p.bank, shape=(p.rows, p.cols))
    assert proto.tobytes() == o.h.tobytes()
    assert bank.tobytes() == o.bank.tobytes()
    # canonical shape of an L/M bank is the reference bank itself
    assert (p.crows, p.ctaps, p.shift, p.frame_len, p.hist_len) == (L, o.cols, 0, 0, o.cols - 1)
    cb = np.ctypeslib.as_array(p.cbank, shape=(p.crows, p.ctaps))
    assert cb.tobytes() == o.bank.tobytes()
    single = np.ctypeslib.as_array(p.single_tap, shape=(p.crows,))
    for r in range(L):
        nz = np.flatnonzero(o.bank[r])
        assert single[r] == (nz[0] if len(nz) == 1 else -1)
    if L > 1 and L >= M:      # fc = 1/L: L-th band filter
        assert single[0] >= 0          # the Nyquist phase: one non-zero tap (knife-edge row)
    assert p.abs_row_sum == pytest.approx(np.abs(o.bank).sum(axis=1).max(), rel=1e-12)
    lib.llz_plan_free(C.byref(p))


def test_c4_extension_shape(zlib):
    """BASELINE config 4: '256-tap bank' = k_override 128 -> n = 81921, Q = 257."""
    p, lib = build_plan(zlib, 2, 320, 147, 1, 128)
    assert (p.n, p.cols, p.num_in, p.num_out) == (81921, 257, 47040, 102400)
    lib.llz_plan_free(C.byref(p))


@pytest.mark.parametrize("M,win", [(1, 0), (2, 0), (3, 1), (16, 2)])
def test_decimate_plan_canonical_form(zlib, port, M, win):
    """canonical taps reproduce the reference's double loop: y[i] = sum_k c[k] x[i*M - k]."""
    p, lib = build_plan(zlib, 0, 1, M, win)
    o = port.decimate_plan(M, win)
    assert (p.n, p.rows, p.cols, p.num_in, p.num_out) == (o.n, o.rows, o.cols, o.num_in, o.num_out)
    assert np.ctypeslib.as_array(p.bank, shape=(p.rows, p.cols)).tobytes() == o.bank.tobytes()
    assert (p.ctaps, p.shift, p.hist_len) == (o.n + 1, 0, o.n)
    c = np.ctypeslib.as_array(p.cbank, shape=(p.ctaps,)).copy()
    order = np.ctypeslib.as_array(p.order, shape=(p.ctaps,)).copy()
    x = port.lcg_s16(o.num_in * 2, 17).astype(np.float64)
    want = port.decimate_run(o, 1.0, x.astype(np.int16), 2 * o.num_out)
    xp = np.concatenate([np.zeros(p.ctaps), x])
    for i in (0, 1, 57, 2 * o.num_out - 1):
        acc = 0.0
        for k in order:                      # reference accumulation order
            acc = acc + xp[p.ctaps + i * M - k] * c[k]
        acc = min(max(acc, -32768.0), 32767.0)
        assert int(acc) == want[i]
    lib.llz_plan_free(C.byref(p))


@pytest.mark.parametrize("L,win", [(2, 0), (3, 1), (16, 2)])
def test_interp_plan_canonical_form(zlib, port, L, win):
    p, lib = build_plan(zlib, 1, L, 1, win)
    o = port.interp_plan(L, win)
    assert (p.n, p.rows, p.cols, p.num_in, p.num_out) == (o.n, o.rows, o.cols, o.num_in, o.num_out)
    assert (p.ctaps, p.shift, p.frame_len, p.hist_len) == (o.cols, o.cols - 1, 1024, 0)
    cb = np.ctypeslib.as_array(p.cbank, shape=(L, p.ctaps))
    assert np.array_equal(cb, o.bank[::-1, ::-1])
    lib.llz_plan_free(C.byref(p))


def test_plan_range_checks(zlib):
    for kind, L, M in ((2, 17, 1), (2, 1, 17), (0, 1, 17), (1, 17, 1), (2, 0, 3), (2, 3, -1)):
        p, _ = build_plan(zlib, kind, L, M, 1)
        assert p is None, (kind, L, M)
    p, _ = build_plan(zlib, 2, 3, 2, 9)
    assert p is None
