"""The exactness guard of the resampler kernels, exercised on purpose.

The exact mode (ACC_F64) is "fast evaluation + near-integer guard + reference-order recompute"
(llz_poly_device.cuh: poly_reference_order_sum; the reference's loop is libllzfilter/llz_resample.c:590-601).
With the production band the recompute runs for about 1e-8 of the outputs, so ordinary parity tests never enter it.
Two families of tests do:
  * a widened band (llz_cuda_resample_bank_set_guard_scale): ~1 % of the outputs take the recompute in every kernel that
    has the branch, and the int16 output must still be the reference's, bit for bit;
  * adversarial inputs (tests/adversarial.py): outputs whose reference sum is a few ulps above / below a non-zero integer,
    with the PRODUCTION band -- the guard has to catch every one of them.
"""
import numpy as np
import pytest

from adversarial import reference_sum, terms_decimate, terms_resample, tune_output

pytestmark = pytest.mark.gpu

# (kind, L, M, k_override, tiles, kernel the call lands on)
CASES = [
    ("resample", 160, 147, 0, 1, "poly_bank_imma_kernel: second look -> third level"),
    ("resample", 160, 147, 0, 4, "poly_bank_umma_kernel (tcgen05): second look -> third level"),
    ("resample", 160, 147, 0, 2, "poly_bank_dmma_kernel"),
    ("resample", 160, 147, 0, 3, "poly_bank_kernel (DFMA register tile)"),
    ("resample", 320, 147, 128, 4, "poly_bank_umma_kernel (tcgen05), Q = 257: 80 phase tiles of the 16-fold replicated bank"),
    ("resample", 1, 3, 0, 4, "poly_bank_umma_kernel (tcgen05): decimating bank as 64 phases of a 192-sample cycle"),
    ("decimate", 1, 4, 0, 4, "poly_bank_umma_kernel (tcgen05): llz_decimate's tap order"),
    ("resample", 320, 147, 128, 1, "poly_bank_imma_kernel, Q = 257"),
    ("resample", 320, 147, 128, 2, "poly_bank_dmma_kernel, Q = 257"),
    ("resample", 3, 2, 0, 0, "few phases: repeated rows on the phase-bank tiles"),
    ("resample", 1, 3, 0, 0, "poly_slide_kernel"),
    ("decimate", 1, 3, 0, 0, "poly_slide_kernel, decimator order"),
    ("decimate", 1, 4, 0, 0, "poly_slide_kernel, M = 4"),
    ("resample", 147, 160, 0, 0, "down-sampling phase bank"),
]


def make_bank(zlib, port, kind, L_, M, k, tiles, C_):
    if kind == "decimate":
        bank = zlib.ResampleBank(zlib.KIND_DECIMATE, 1, M, C_)
        plan = port.decimate_plan(M, 1)
    else:
        bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L_, M, C_, k_override=k)
        plan = port.resample_plan(L_, M, 1, k)
    if tiles:
        bank.set_tiles(tiles)
    return bank, plan


def oracle_run(port, kind, plan, x, n_out):
    if kind == "decimate":
        return port.decimate_run(plan, 1.0, x, n_out)
    return port.resample_run(plan, 1.0, x, n_out)


def run_bank(torch, bank, x, n_out):
    C_, n_in = x.shape
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, n_out, dtype=torch.int16, device="cuda")
    got_out = bank.run(dx, n_in, n_in, dy, n_out)
    torch.cuda.synchronize()
    assert got_out == n_out
    return dy.cpu().numpy()


@pytest.mark.parametrize("kind,L_,M,k,tiles,what", CASES)
def test_widened_guard_band_takes_the_recompute_and_stays_bit_identical(zlib, port, cuda, kind, L_, M, k, tiles, what):
    torch = cuda
    C_ = 2
    bank, plan = make_bank(zlib, port, kind, L_, M, k, tiles, C_)
    n_in = plan.num_in * 3
    x = np.stack([port.lcg_s16(n_in, 9100 + c) for c in range(C_)])
    n_out = bank.out_len(n_in)
    bank.set_guard_scale(3e6)                                  # production band ~2e-9 -> ~6e-3: about 1 % of the outputs
    got = run_bank(torch, bank, x, n_out)
    hits = bank.guard_count()
    for c in range(C_):
        want = oracle_run(port, kind, plan, x[c], n_out)
        assert np.array_equal(got[c], want), (what, c)
    # outputs of single-tap (knife-edge) rows never take the guard; everything else hits with probability ~2*band
    assert hits >= n_out * C_ // 1000, (what, hits, n_out)
    assert hits <= n_out * C_ // 2, (what, hits, n_out)         # the band grows with Q: ~17 % at Q = 257
    bank.close()


def test_widened_guard_band_interp_general_kernel(zlib, port, cuda):
    """llz_interp runs on poly_general_kernel, whose guard is poly_emit's"""
    torch = cuda
    L_ = 2
    bank = zlib.ResampleBank(zlib.KIND_INTERP, L_, 1, 1)
    plan = port.interp_plan(L_, 1)
    n_in = plan.num_in * 4
    x = port.lcg_s16(n_in, 515).reshape(1, -1)
    bank.set_guard_scale(3e6)
    got = run_bank(torch, bank, x, n_in * L_)
    assert np.array_equal(got[0], port.interp_run(plan, 1.0, x[0]))
    assert bank.guard_count() >= n_in * L_ // 2000
    bank.close()


def test_widened_guard_band_interp_on_tcgen05(zlib, port, cuda):
    """llz_interp on the tcgen05 kernel: first-level hits take the warp's FP64 second look with the frame-local window,
    what is still near an integer goes to the reference-order sum"""
    torch = cuda
    L_, C_ = 4, 2
    bank = zlib.ResampleBank(zlib.KIND_INTERP, L_, 1, C_)
    bank.set_tiles(zlib.TILES_INT8_TCGEN05)
    plan = port.interp_plan(L_, 1)
    n_in = plan.num_in * 6
    x = np.stack([port.lcg_s16(n_in, 616 + c) for c in range(C_)])
    bank.set_guard_scale(3e6)
    got = run_bank(torch, bank, x, n_in * L_)
    assert bank.last_run()[1].startswith("poly_bank_umma_kernel")
    for c in range(C_):
        assert np.array_equal(got[c], port.interp_run(plan, 1.0, x[c]))
    assert bank.guard_count() >= n_in * L_ * C_ // 2000
    bank.close()


@pytest.mark.parametrize("kind,L_,M,k,tiles,what", CASES[:6] + CASES[7:9])
def test_adversarial_near_integer_sums_with_the_production_band(zlib, port, cuda, kind, L_, M, k, tiles, what):
    """16 outputs per channel tuned to within a few ulps of a non-zero integer, half of them at or just above it and half
    just below: the fast evaluation may land on either side, the guard has to send every one of them to the
    reference-order sum"""
    torch = cuda
    C_ = 2
    bank, plan = make_bank(zlib, port, kind, L_, M, k, tiles, C_)
    n_in = plan.num_in * (2 if L_ > 1 else 10)                # windows of neighbouring targets must not overlap
    n_out = bank.out_len(n_in)
    rng = np.random.default_rng(L_ * 1000 + M + tiles)
    x = np.stack([port.lcg_s16(n_in, 3300 + c) for c in range(C_)]).copy()
    targets = []
    n_targets = 16
    step = (n_out - 2 * plan.cols * max(L_, 1)) // n_targets
    for c in range(C_):
        for j in range(n_targets):
            o = plan.cols * max(L_, 1) + j * step + int(rng.integers(0, max(1, step // 4)))
            r = None
            for attempt in range(4):                           # a phase whose taps are near small rationals cannot be tuned: next
                if kind == "decimate":
                    terms = terms_decimate(plan, o)
                else:
                    if np.count_nonzero(plan.bank[o % plan.L]) < 3:
                        o += 1                                 # the knife-edge phase has a single tap: no guard needed
                    terms = terms_resample(plan, o)
                r = tune_output(rng, terms, x[c], 1.0, side=1 if j % 2 == 0 else -1, tries=24)
                if r is not None:
                    break
                o += 1
            assert r is not None, (what, c, j)
            targets.append((c, o, terms, r))
    # later targets never touch earlier windows' samples: re-evaluate every target on the final input
    near = 0
    sides = set()
    for c, o, terms, _ in targets:
        v = reference_sum(terms, x[c], 1.0)
        d = v - round(v)
        if abs(d) < 1e-10:
            near += 1
            sides.add(d >= 0)
    assert near >= len(targets) * 3 // 4 and sides == {True, False}, (what, near)
    got = run_bank(torch, bank, x, n_out)
    hits = bank.guard_count()
    for c in range(C_):
        want = oracle_run(port, kind, plan, x[c], n_out)
        assert np.array_equal(got[c], want), (what, c, np.flatnonzero(got[c] != want)[:8])
    assert hits >= near, (what, hits, near)
    bank.close()
