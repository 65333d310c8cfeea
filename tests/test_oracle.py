"""The oracle (oracle/llz_oracle.c) against the reference's own outputs.

Two pins: (1) tests/golden/kat.json + vectors.npz, produced by the unmodified reference
(tests/golden/make_golden.py) and available everywhere; (2) the compiled reference itself
(oracle/_ref) when present, over wider grids.  Everything is compared bit for bit.
"""
import numpy as np
import pytest

from conftest import KIND, WIN


def fnv(port, a):
    return f"{port.fnv64(a):016x}"


# ---- golden pins ----------------------------------------------------------------------------------------
def test_windows_golden(port, kat):
    for c in kat["windows"]:
        if c["win"] == "KAISER_BETA":
            w = port.kaiser_beta(c["N"], c["beta"])
        else:
            w = port.window(c["N"], WIN[c["win"]])
        assert fnv(port, w) == c["fnv"], c
        assert w[0] == c["w0"] and w[c["N"] // 2] == c["wmid"]


def test_estimators_golden(port, kat):
    for c in kat["cof_num"]:
        assert port.cof_num(WIN[c["win"]], c["ftrans"], c["atten"]) == c["n"], c
    for c in kat["atten2beta"]:
        assert port.atten2beta(c["atten"]) == c["beta"]


def test_designs_golden(port, kat, vectors):
    for c in kat["designs"]:
        h = port.fir_design(KIND[c["kind"]], c["N"], c["fc1"], c["fc2"], WIN[c["win"]])
        assert len(h) == c["n_used"]
        assert fnv(port, h) == c["fnv"], c
        assert int((h == 0).sum()) == c["zeros"]
    h = port.fir_design(0, 127, 0.23, 0.0, 0)
    assert h.tobytes() == vectors["design_lpf_127_0.23_hamming"].tobytes()


def test_survey_known_answers(port):
    """Plain values recorded in SURVEY.md section 9 from the survey's own build of the reference."""
    h = port.fir_design(0, 127, 0.23, 0.0, 0)
    assert h[0] == 0.00040400358027932475 and int((h == 0).sum()) == 0
    h = port.fir_design(0, 127, 0.25, 0.0, 0)
    assert h[0] == -0.00028581470354193853 and h[63] == 0.25 and int((h == 0).sum()) == 30
    h = port.fir_design(0, 4095, 0.11, 0.0, 2)
    assert h[0] == -7.516322910774977e-08 and int((h == 0).sum()) == 40
    assert port.fir_design(0, 64, 0.3, 0.0, 1)[0] == 1.3850955800294518e-19
    p = port.resample_plan(160, 147, 1)
    assert (p.n, p.cols, p.num_in, p.num_out) == (7041, 45, 23520, 25600)
    y = port.resample_run(p, 1.0, port.lcg_s16(94080, 777), 102400)
    assert y[160:164].tolist() == [14823, 9577, 5007, -15400]
    p = port.resample_plan(1, 3, 1)
    assert (p.n, p.cols, p.num_in, p.num_out) == (133, 134, 1536, 512)
    y = port.resample_run(p, 1.0, port.lcg_s16(98304, 777), 32768)
    assert y[:8].tolist() == [0, 0, 0, -1, 2, -5, 9, -14]
    p = port.resample_plan(320, 147, 1)
    assert (p.n, p.cols, p.num_in, p.num_out) == (14081, 45, 47040, 102400)
    y = port.resample_run(p, 1.0, port.lcg_s16(94080, 777), 204800)
    assert y[320:324].tolist() == [14823, 11690, 9577, 9209]
    p = port.resample_plan(147, 160, 1)
    assert (p.n, p.cols, p.num_in, p.num_out) == (6763, 47, 23520, 21609)
    p = port.decimate_plan(3, 1)
    assert (p.n, p.cols, p.num_in, p.num_out) == (133, 45, 1023, 341)
    x = port.lcg_f64(4096 * 8, 12345)
    y = port.fir_run(port.fir_design(0, 127, 0.23, 0.0, 0), x, n_out=len(x) + 126)
    assert y[0] == -0.00038751807690026437 and y[len(x) - 1] == 0.23309247051030579
    x = port.lcg_f64(8192 * 4, 12345)
    y = port.fir_run(port.fir_design(0, 4095, 0.11, 0.0, 2), x, n_out=len(x) + 4094)
    assert y[len(x) - 1] == -0.25930909330770052


def test_fir_streams_golden(port, kat, vectors):
    for c in kat["fir_streams"]:
        h = port.fir_design(KIND[c["kind"]], c["N"], c["fc1"], c["fc2"], WIN[c["win"]])
        x = port.lcg_f64(c["frame"] * c["frames"], c["seed"])
        y = port.fir_run(h, x, n_out=len(x) + len(h) - 1)
        assert len(y) == c["n_out"]
        assert fnv(port, y) == c["fnv"], c
        assert y[0] == c["y0"] and y[-1] == c["ylast"]
    x = port.lcg_f64(4096 * 2, 12345)
    y = port.fir_run(port.fir_design(0, 127, 0.23, 0.0, 0), x, n_out=len(x) + 126)
    assert y.tobytes() == vectors["fir_lpf_127_y"].tobytes()


def test_resample_golden(port, kat, vectors):
    for c in kat["resample"]:
        p = port.resample_plan(c["L"], c["M"], WIN[c["win"]])
        assert p.num_in == c["num_in"]
        x = port.lcg_s16(c["num_in"] * c["frames"], c["seed"])
        y = port.resample_run(p, c["gain"], x, c["n_out"])
        assert fnv(port, y) == c["fnv"], c
        assert y[:8].tolist() == c["head"]
    for c in kat["framelen"]:
        assert port.resample_plan(c["L"], c["M"], WIN[c["win"]]).num_in == c["num_in"]
    p = port.resample_plan(160, 147, 1)
    y = port.resample_run(p, 1.0, port.lcg_s16(p.num_in * 2, 777), p.num_out * 2)
    assert y.tobytes() == vectors["resample_160_147_y"].tobytes()
    p = port.resample_plan(1, 3, 1)
    y = port.resample_run(p, 1.0, port.lcg_s16(p.num_in * 8, 777), p.num_out * 8)
    assert y.tobytes() == vectors["resample_1_3_y"].tobytes()


def test_resample_saturates_in_golden(kat):
    assert any(c["sat"] > 0 for c in kat["resample"]), "golden set must cover the clamp"


def test_decimate_interp_golden(port, kat):
    for c in kat["decimate"]:
        p = port.decimate_plan(c["M"], WIN[c["win"]])
        x = port.lcg_s16(c["num_in"] * c["frames"], c["seed"])
        y = port.decimate_run(p, c["gain"], x, c["n_out"])
        assert fnv(port, y) == c["fnv"], c
    for c in kat["interp"]:
        p = port.interp_plan(c["L"], WIN[c["win"]])
        x = port.lcg_s16(1024 * c["frames"], c["seed"])
        y = port.interp_run(p, c["gain"], x)
        assert len(y) == c["n_out"]
        assert fnv(port, y) == c["fnv"], c


# ---- wider grids against the compiled reference (build container, and the GPU box's prebuilt copy) ---------
@pytest.mark.parametrize("win", [0, 1, 2])
def test_design_grid_vs_reference(port, ref, win):
    for N in (2, 3, 5, 16, 31, 32, 127, 500, 1023):
        assert port.window(N, win).tobytes() == ref.window(N, win).tobytes()
        for kind in range(4):
            for f1, f2 in ((0.1, 0.3), (0.25, 0.5), (0.5, 0.9), (1.0 / 3, 2.0 / 3)):
                a = port.fir_design(kind, N, f1, f2, win)
                b = ref.fir_design(kind, N, f1, f2, win)
                assert a.tobytes() == b.tobytes(), (kind, N, f1, f2, win)
    for ft in np.linspace(0.0007, 0.2, 57):
        assert port.cof_num(win, float(ft)) == ref.cof_num(win, float(ft))


def test_fir_stream_vs_reference(port, ref):
    rng = np.random.default_rng(1)
    for kind, N, frame in ((0, 5, 16), (0, 64, 64), (1, 33, 40), (2, 77, 100), (3, 129, 256), (0, 513, 1024)):
        x = rng.standard_normal(frame * 5)
        want = ref.fir_stream(kind, N, 0.3, 0.6, 1, x, frame, flush=frame >= N)
        h = port.fir_design(kind, N, 0.3, 0.6, 1)
        got = port.fir_run(h, x, n_out=len(want))
        assert got.tobytes() == want.tobytes(), (kind, N)


@pytest.mark.parametrize("L,M,win", [(160, 147, 1), (147, 160, 0), (1, 3, 1), (3, 1, 2), (5, 7, 0), (16, 1, 1),
                                     (1, 16, 1), (320, 147, 1), (2, 3, 2), (1, 1, 0), (48, 125, 1)])
def test_resample_vs_reference(port, ref, L, M, win):
    p = port.resample_plan(L, M, win)
    assert p.num_in == ref.resample_framelen(L, M, win)
    x = port.lcg_s16(p.num_in * 3, 4242 + L)
    x[100:140] = 32767            # full-scale burst: exercises the clamp
    x[200:240] = -32768
    for gain in (1.0, 2.5):
        want = ref.resample_stream(L, M, gain, win, x)
        got = port.resample_run(p, gain, x, len(want))
        assert np.array_equal(got, want), (L, M, win, gain, int((got != want).sum()))


def test_resample_rejects_ratio(port, ref):
    assert port.resample_plan(17, 1, 1) is None and port.resample_plan(1, 17, 1) is None
    assert ref.resample_stream(17, 1, 1.0, 1, np.zeros(0, np.int16)) is None


@pytest.mark.parametrize("M,win", [(1, 0), (2, 0), (3, 1), (5, 2), (16, 1)])
def test_decimate_vs_reference(port, ref, M, win):
    p = port.decimate_plan(M, win)
    x = port.lcg_s16(p.num_in * 6, 99 + M)
    want = ref.decimate_stream(M, 1.5, win, x)
    got = port.decimate_run(p, 1.5, x, len(want))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("L,win", [(1, 0), (2, 0), (3, 1), (7, 2), (16, 1)])
def test_interp_vs_reference(port, ref, L, win):
    p = port.interp_plan(L, win)
    x = port.lcg_s16(1024 * 3, 5 + L)
    want = ref.interp_stream(L, 0.9, win, x)
    got = port.interp_run(p, 0.9, x)
    assert np.array_equal(got, want)


def test_phase0_knife_edge(port):
    """phase-0 outputs of an L/M bank are x*(1-2^-53) truncated toward zero."""
    p = port.resample_plan(160, 147, 1)
    x = port.lcg_s16(p.num_in, 777)
    y = port.resample_run(p, 1.0, x, p.num_out)
    m = np.arange(0, p.num_out, 160)
    idx = (m * 147) // 160 - (p.n // 2) // 160        # the centre tap sits in column (n/2)/L = 22
    src = np.where(idx >= 0, x[np.maximum(idx, 0)], 0).astype(np.int32)
    expect = np.where(src > 0, src - 1, np.where(src < 0, src + 1, 0))
    assert np.array_equal(y[m].astype(np.int32), expect)


@pytest.mark.parametrize("a,b", [([1.0, -1.8, 0.81], [0.2, 0.3, 0.2]), ([1.0, -0.5], [1.0]), ([1.0], [0.25, 0.25, 0.25, 0.25]),
                                 ([1.0, -2.369513, 2.313988, -1.054665, 0.187379], [0.004824, 0.019297, 0.028946, 0.019297, 0.004824])])
def test_iir_restatement_vs_reference(port, ref, a, b):
    """llz_iir.c:103-156 frame by frame (+ flush) against the whole-signal restatement, bit for bit"""
    a, b = np.array(a), np.array(b)
    x = port.lcg_f64(6000, 77)
    want = ref.iir_stream(a, b, x, frame=1024)
    st = (np.zeros(len(b)), np.zeros(len(a)))
    got = [port.iir_run(a, b, x, state=st)]
    if len(b) > 1:
        got.append(port.iir_run(a, b, None, len(b) - 1, state=st))
    assert np.concatenate(got).tobytes() == want.tobytes()
