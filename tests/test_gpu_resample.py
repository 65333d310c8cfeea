"""Polyphase resampler parity on the GPU through the C-ABI.

Bar: int16 output bit-exact (value, length, phase/index sequence) in the FP64 modes; the FP32 mode
may differ by at most 1 LSB on a small fraction of samples (SURVEY.md 8d note 4)."""
import numpy as np
import pytest

from conftest import WIN

pytestmark = pytest.mark.gpu


def test_dropin_resample_matches_golden_hashes(zlib, port, kat, cuda):
    for c in kat["resample"]:
        r = zlib.Resampler(zlib.KIND_RESAMPLE, c["L"], c["M"], c["gain"], WIN[c["win"]])
        assert r.bytes_in == 2 * c["num_in"]
        x = port.lcg_s16(c["num_in"] * c["frames"], c["seed"])
        y = r.stream(x)
        r.close()
        assert len(y) == c["n_out"]
        assert f"{port.fnv64(y):016x}" == c["fnv"], c
        assert y[:8].tolist() == c["head"]


def test_dropin_decimate_interp_match_golden_hashes(zlib, port, kat, cuda):
    for c in kat["decimate"]:
        r = zlib.Resampler(zlib.KIND_DECIMATE, 1, c["M"], c["gain"], WIN[c["win"]])
        y = r.stream(port.lcg_s16(c["num_in"] * c["frames"], c["seed"]))
        r.close()
        assert len(y) == c["n_out"] and f"{port.fnv64(y):016x}" == c["fnv"], c
    for c in kat["interp"]:
        r = zlib.Resampler(zlib.KIND_INTERP, c["L"], 1, c["gain"], WIN[c["win"]])
        y = r.stream(port.lcg_s16(1024 * c["frames"], c["seed"]))
        r.close()
        assert len(y) == c["n_out"] and f"{port.fnv64(y):016x}" == c["fnv"], c


def test_dropin_error_conventions(zlib, cuda):
    L = zlib.lib()
    assert L.llz_resample_filter_init(17, 1, 1.0, 1) == zlib.FAIL          # llz_resample.c:375-378
    assert L.llz_decimate_init(17, 1.0, 1) == zlib.FAIL
    assert L.llz_interp_init(17, 1.0, 1) == zlib.FAIL
    r = zlib.Resampler(zlib.KIND_RESAMPLE, 160, 147)
    import ctypes as C
    osz = C.c_int(0)
    buf = np.zeros(100, np.int16)
    assert L.llz_resample(r.handle, buf.ctypes.data, 200, buf.ctypes.data, C.byref(osz)) == -1
    assert L.llz_decimate(r.handle, buf.ctypes.data, r.bytes_in, buf.ctypes.data, C.byref(osz)) == -1
    L.llz_decimate_uninit(r.handle)           # any uninit takes any kind (main.c:125)
    r.handle = 0


def oracle_resample(port, L, M, win, k, gain, x, n_out):
    p = port.resample_plan(L, M, win, k)
    return np.stack([port.resample_run(p, gain, x[c], n_out) for c in range(x.shape[0])])


CASES = [(160, 147, 1, 0), (147, 160, 0, 0), (1, 3, 1, 0), (1, 2, 0, 0), (1, 16, 2, 0), (3, 1, 1, 0), (320, 147, 1, 0),
         (320, 147, 1, 8), (5, 7, 2, 0), (1, 1, 1, 0), (16, 1, 0, 0)]


@pytest.mark.parametrize("L,M,win,k", CASES)
def test_bank_exact_modes(zlib, port, cuda, L, M, win, k):
    torch = cuda
    C_ = 3
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_, win=win, k_override=k, acc=zlib.ACC_F64)
    strict = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_, win=win, k_override=k, acc=zlib.ACC_F64_STRICT)
    info = bank.info
    p = port.resample_plan(L, M, win, k)
    assert (info.n, info.taps_per_phase, info.num_in, info.num_out) == (p.n, p.cols, p.num_in, p.num_out)
    assert bank.proto().tobytes() == p.h.tobytes() and bank.bank().tobytes() == p.bank.tobytes()
    n_in = p.num_in * 2 + 37                          # not a whole number of frames
    x = np.stack([port.lcg_s16(n_in, 777 + c) for c in range(C_)])
    x[1, 50:90] = 32767
    x[1, 300:340] = -32768
    x[2, 1000:3000] = 0                                # a silent stretch: exact-zero sums
    n_out = bank.out_len(n_in)
    assert n_out == -((-n_in * L) // M)
    want = oracle_resample(port, L, M, win, k, 1.0, x, n_out)
    dx = torch.from_numpy(x).cuda()
    for b in (bank, strict):
        dy = torch.zeros(C_, n_out + 8, dtype=torch.int16, device="cuda")
        assert b.run(dx, n_in, n_in, dy, n_out + 8) == n_out
        torch.cuda.synchronize()
        got = dy.cpu().numpy()
        assert np.array_equal(got[:, :n_out], want), (L, M, int((got[:, :n_out] != want).sum()))
        assert not got[:, n_out:].any()                # nothing written past the end
    assert bank.guard_count() < max(64, n_out * C_ // 4)
    bank.close(); strict.close()


@pytest.mark.parametrize("L,M,win,k", CASES + [(320, 147, 1, 128), (441, 320, 1, 0)])
def test_bank_f32_within_one_lsb(zlib, port, cuda, L, M, win, k):
    torch = cuda
    C_ = 2
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_, win=win, k_override=k, acc=zlib.ACC_F32)
    n_in = bank.info.num_in * 3
    x = np.stack([port.lcg_s16(n_in, 31 + c) for c in range(C_)])
    n_out = bank.out_len(n_in)
    want = oracle_resample(port, L, M, win, k, 1.0, x, n_out).astype(np.int32)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, n_out, dtype=torch.int16, device="cuda")
    bank.run(dx, n_in, n_in, dy, n_out)
    torch.cuda.synchronize()
    diff = np.abs(dy.cpu().numpy().astype(np.int32) - want)
    assert diff.max() <= 1
    assert (diff != 0).mean() <= 0.02, (diff != 0).mean()
    print(f"f32 mismatch rate L={L} M={M} Q={bank.info.taps_per_phase}: {(diff != 0).mean():.5f}")
    if L > 1 and L >= M:                               # knife-edge phase stays exact in the fast mode
        assert not diff[:, ::L].any()
    bank.close()


@pytest.mark.parametrize("gain", [1.0, -2.5])
@pytest.mark.parametrize("L,M,k", [(160, 147, 0), (320, 147, 128), (147, 160, 0), (3, 2, 0), (1, 3, 0)])
def test_bank_fast_mode_on_tcgen05(zlib, port, cuda, L, M, k, gain):
    """ACC_F32 on the tcgen05 kernel (llz_cuda_polybank_umma.cu with three int8 digit planes of the taps: exact products
    and sums, 22-bit taps, no guard): within 1 LSB of the reference, knife-edge phase exact, ragged two-call stream."""
    torch = cuda
    C_ = 3
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_, k_override=k, gain=gain, acc=zlib.ACC_F32)
    bank.set_tiles(zlib.TILES_INT8_TCGEN05)
    plan = port.resample_plan(L, M, 1, k)
    n_in = plan.num_in * 2 + 777
    x = np.stack([port.lcg_s16(n_in, 5000 + c) for c in range(C_)])
    x[0, :5000] = 32767
    x[1, 100:4000] = -32768
    n_out = bank.out_len(n_in)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, n_out + 8, dtype=torch.int16, device="cuda")
    cut = plan.num_in + 123
    o1 = bank.run(dx, n_in, cut, dy, n_out + 8)
    o2 = bank.run(dx.data_ptr() + 2 * cut, n_in, n_in - cut, dy.data_ptr() + 2 * o1, n_out + 8)
    torch.cuda.synchronize()
    assert o1 + o2 == n_out
    got = dy.cpu().numpy()
    assert not got[:, n_out:].any()
    want = np.stack([port.resample_run(plan, gain, x[c], n_out) for c in range(C_)]).astype(np.int32)
    diff = np.abs(got[:, :n_out].astype(np.int32) - want)
    assert diff.max() <= 1
    # 22-bit fixed-point taps: the rounding errors of Q taps add up to ~0.02 LSB rms on full-scale noise (Q = 257), and an
    # output is off by one when that error crosses an integer
    assert (diff != 0).mean() <= 0.03, (diff != 0).mean()
    if L >= M:
        assert not diff[:, ::L].any()
    bank.close()


@pytest.mark.parametrize("gain", [0.5, 1.0, 3.0, -2.0, 0.0])
def test_gain_and_saturation(zlib, port, cuda, gain):
    torch = cuda
    L, M = 3, 2
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, 1, gain=gain, win=2)
    n_in = bank.info.num_in * 2
    x = port.lcg_s16(n_in, 9)[None, :] * 2             # wraps: large swings
    n_out = bank.out_len(n_in)
    want = oracle_resample(port, L, M, 2, 0, gain, x, n_out)
    dy = torch.zeros(1, n_out, dtype=torch.int16, device="cuda")
    bank.run(torch.from_numpy(x).cuda(), n_in, n_in, dy, n_out)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), want)
    bank.close()


@pytest.mark.parametrize("L,M,win", [(160, 147, 1), (1, 3, 1), (7, 5, 0)])
def test_streaming_chunks_equal_one_shot(zlib, port, cuda, L, M, win):
    torch = cuda
    C_ = 2
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_, win=win)
    n_in = bank.info.num_in * 2 + 11
    x = np.stack([port.lcg_s16(n_in, 1 + c) for c in range(C_)])
    total = -((-n_in * L) // M)
    want = oracle_resample(port, L, M, win, 0, 1.0, x, total)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, total, dtype=torch.int16, device="cuda")
    pos_in = pos_out = 0
    for step in (1, 2, 5, 40, 1000, 3, 10 ** 9):
        m = min(step, n_in - pos_in)
        if m <= 0:
            break
        pos_out += bank.run(dx.data_ptr() + 2 * pos_in, n_in, m, dy.data_ptr() + 2 * pos_out, total)
        pos_in += m
    torch.cuda.synchronize()
    assert pos_out == total
    assert np.array_equal(dy.cpu().numpy(), want)
    bank.close()


def test_time_segments_with_halo_are_byte_identical(zlib, port, cuda):
    """SURVEY.md 8e, config C4 shape at reduced length: segments start at phase 0 with a Q-1 halo"""
    torch = cuda
    L, M, C_ = 320, 147, 2
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_, k_override=16)
    info = bank.info
    frames = 11
    n_in = info.num_in * frames
    x = np.stack([port.lcg_s16(n_in, 777 + c) for c in range(C_)])
    dx = torch.from_numpy(x).cuda()
    total = bank.out_len(n_in)
    one = torch.zeros(C_, total, dtype=torch.int16, device="cuda")
    bank.run(dx, n_in, n_in, one, total)
    want = oracle_resample(port, L, M, 1, 16, 1.0, x, total)
    torch.cuda.synchronize()
    assert np.array_equal(one.cpu().numpy(), want)
    for world in (2, 4, 8):
        out = torch.zeros_like(one)
        covered = 0
        for rank in range(world):
            s = zlib.shard_resample_segments(n_in, L, M, info.taps_per_phase, info.num_in, world, rank)
            assert s.out_start == covered and s.out_start % L == 0 and s.in_start * L == s.out_start * M
            covered += s.out_count
            bank.set_history((dx.data_ptr() + 2 * (s.in_start - s.halo)) if s.halo else None, n_in)
            got = bank.run(dx.data_ptr() + 2 * s.in_start, n_in, s.in_count, out.data_ptr() + 2 * s.out_start, total)
            assert got == s.out_count
        assert covered == total
        torch.cuda.synchronize()
        assert torch.equal(out, one), world
    bank.close()


def test_decimate_and_interp_banks(zlib, port, cuda):
    torch = cuda
    for M, win in ((2, 0), (3, 1), (16, 2)):
        p = port.decimate_plan(M, win)
        for acc in (zlib.ACC_F64, zlib.ACC_F64_STRICT):
            bank = zlib.ResampleBank(zlib.KIND_DECIMATE, 1, M, 2, gain=1.5, win=win, acc=acc)
            n_in = p.num_in * 5
            x = np.stack([port.lcg_s16(n_in, 3 + c) for c in range(2)])
            n_out = bank.out_len(n_in)
            assert n_out == n_in // M
            want = np.stack([port.decimate_run(p, 1.5, x[c], n_out) for c in range(2)])
            dy = torch.zeros(2, n_out, dtype=torch.int16, device="cuda")
            bank.run(torch.from_numpy(x).cuda(), n_in, n_in, dy, n_out)
            torch.cuda.synchronize()
            assert np.array_equal(dy.cpu().numpy(), want), (M, acc)
            bank.close()
    for L_, win in ((2, 0), (3, 1), (16, 2)):
        p = port.interp_plan(L_, win)
        for acc in (zlib.ACC_F64, zlib.ACC_F64_STRICT):
            bank = zlib.ResampleBank(zlib.KIND_INTERP, L_, 1, 2, gain=0.8, win=win, acc=acc)
            n_in = 1024 * 3
            x = np.stack([port.lcg_s16(n_in, 8 + c) for c in range(2)])
            want = np.stack([port.interp_run(p, 0.8, x[c]) for c in range(2)])
            dy = torch.zeros(2, n_in * L_, dtype=torch.int16, device="cuda")
            assert bank.run(torch.from_numpy(x).cuda(), n_in, n_in, dy, n_in * L_) == n_in * L_
            torch.cuda.synchronize()
            assert np.array_equal(dy.cpu().numpy(), want), (L_, acc)
            bank.close()


@pytest.mark.parametrize("acc", ["exact", "fast"])
@pytest.mark.parametrize("L_,win,gain", [(2, 0, 1.0), (3, 1, 0.8), (4, 1, -1.5), (16, 2, 1.0), (5, 1, 1.0)])
def test_interp_bank_on_tcgen05(zlib, port, cuda, L_, win, gain, acc):
    """llz_interp on the tcgen05 phase-bank kernel: the frames lie one after the other in the sample planes, each followed
    by the zeros the reference's frame-local window reads (quirk R4), a tile's 128 rows are whole frames (5-D tensor map);
    exact mode bit-identical to the reference frame loop, fast mode within 1 LSB.  Two calls of whole frames."""
    torch = cuda
    C_ = 3
    p = port.interp_plan(L_, win)
    bank = zlib.ResampleBank(zlib.KIND_INTERP, L_, 1, C_, gain=gain, win=win, acc=zlib.ACC_F64 if acc == "exact" else zlib.ACC_F32)
    bank.set_tiles(zlib.TILES_INT8_TCGEN05)
    frames = 7
    n_in = p.num_in * frames
    x = np.stack([port.lcg_s16(n_in, 80 + c) for c in range(C_)])
    x[0, 1000:1100] = 32767                              # full scale across a frame boundary
    want = np.stack([port.interp_run(p, gain, x[c]) for c in range(C_)]).astype(np.int32)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, n_in * L_ + 8, dtype=torch.int16, device="cuda")
    cut = p.num_in * 3
    o1 = bank.run(dx, n_in, cut, dy, n_in * L_ + 8)
    o2 = bank.run(dx.data_ptr() + 2 * cut, n_in, n_in - cut, dy.data_ptr() + 2 * o1, n_in * L_ + 8)
    torch.cuda.synchronize()
    assert (o1, o2) == (cut * L_, (n_in - cut) * L_)
    launches, kernel = bank.last_run()
    assert kernel.startswith("poly_bank_umma_kernel"), kernel
    got = dy.cpu().numpy()
    assert not got[:, n_in * L_:].any()
    diff = np.abs(got[:, :n_in * L_].astype(np.int32) - want)
    if acc == "exact":
        assert not diff.any(), (L_, int(diff.max()), int((diff != 0).sum()))
    else:
        assert diff.max() <= 1 and (diff != 0).mean() <= 0.03
    bank.close()


def test_run_host_pipeline(zlib, port, cuda):
    torch = cuda
    L, M, C_ = 1, 3, 8
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, C_)
    n_in = bank.info.num_in * 4000                     # 6.1 M samples x 8 ch x 2 B = 98 MB: two chunks
    x = zlib.host_alloc(C_ * n_in * 2, np.int16).reshape(C_, n_in)
    for c in range(C_):
        x[c] = port.lcg_s16(n_in, 777 + c)
    total = n_in // 3
    y = zlib.host_alloc(C_ * total * 2, np.int16).reshape(C_, total)
    assert bank.run_host(x, n_in, n_in, y, total) == total
    dx = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    dy = torch.zeros(C_, total, dtype=torch.int16, device="cuda")
    bank.reset()
    bank.run(dx, n_in, n_in, dy, total)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), y)
    p = port.resample_plan(1, 3, 1)
    want = port.resample_run(p, 1.0, x[5, :300000], 100000)
    assert np.array_equal(y[5, :100000], want)
    bank.close()
    zlib.host_free(x.reshape(-1)); zlib.host_free(y.reshape(-1))


def test_c3_shape_spot_checks(zlib, port, cuda):
    """BASELINE config 3 (48 k -> 16 k, 64 channels) at 1/10 length, device LCG input; windows
    checked against the oracle at random positions, and a 64-bit-index check far into the stream."""
    torch = cuda
    C_, n_in = 64, 2_880_000
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, 1, 3, C_)
    dx = torch.empty(C_, n_in, dtype=torch.int16, device="cuda")
    zlib.synth_lcg(dx, n_in, C_, n_in, 2, 777)
    total = bank.out_len(n_in)
    dy = torch.zeros(C_, total, dtype=torch.int16, device="cuda")
    bank.run(dx, n_in, n_in, dy, total)
    torch.cuda.synchronize()
    p = port.resample_plan(1, 3, 1)
    rng = np.random.default_rng(3)
    for c in (0, 31, 63):
        xc = port.lcg_s16(n_in, 777 + c)
        assert np.array_equal(dx[c, :5000].cpu().numpy(), xc[:5000])
        yc = dy[c].cpu().numpy()
        for m0 in [0, total - 2000] + rng.integers(0, total - 2000, 6).tolist():
            want = port.resample_run(p, 1.0, xc, 2000, m0=m0)
            assert np.array_equal(yc[m0:m0 + 2000], want), (c, m0)
    bank.close()


@pytest.mark.parametrize("tiles", [1, 2, 3, 4])
@pytest.mark.parametrize("L_,M,k", [(160, 147, 0), (320, 147, 128), (147, 160, 0), (3, 2, 0), (1, 3, 0), (2, 1, 0), (1, 4, 0),
                                    (513, 512, 0)])          # 513 x 32 phases: the last tcgen05 phase tile is half empty
def test_bank_exact_mode_integer_and_fp64_tensor_tiles(zlib, port, cuda, L_, M, k, tiles):
    """The exact mode's four tile kernels: INT8 tensor cores through mma.sync (1) and through tcgen05 with TMEM
    accumulators (4) -- taps as five int8 digit planes, samples as two byte planes, exact s32 accumulation, a two-level
    near-integer guard --, FP64 tensor cores (2) and the DFMA register tile (3), llz_cuda_resample_bank_set_tiles: the
    int16 output must be the reference's, bit for bit, from all of them."""
    torch = cuda
    C_ = 3
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L_, M, C_, k_override=k)
    bank.set_tiles(tiles)
    plan = port.resample_plan(L_, M, 1, k)
    n_in = plan.num_in * 2 + 777
    x = np.stack([port.lcg_s16(n_in, 4000 + c) for c in range(C_)])
    x[0, :5000] = 32767                                  # full scale, saturating outputs
    x[1, 100:4000] = -32768
    n_out = bank.out_len(n_in)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, n_out + 8, dtype=torch.int16, device="cuda")
    # in two ragged calls: the second one starts inside the history and off the tile grid
    cut = plan.num_in + 123
    o1 = bank.run(dx, n_in, cut, dy, n_out + 8)
    o2 = bank.run(dx.data_ptr() + 2 * cut, n_in, n_in - cut, dy.data_ptr() + 2 * o1, n_out + 8)
    torch.cuda.synchronize()
    assert o1 + o2 == n_out
    got = dy.cpu().numpy()[:, :n_out]
    for c in range(C_):
        want = port.resample_run(plan, 1.0, x[c], n_out)
        assert np.array_equal(got[c], want), (L_, M, c, int(np.abs(got[c].astype(np.int32) - want).max()))
    bank.close()
