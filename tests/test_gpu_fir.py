"""FIR parity on the GPU, through the C-ABI of libllzfilter_cuda.so, against the oracle.

Tolerances (SURVEY.md 8d note 4 / BASELINE.json north_star):
  drop-in llz_fir_filter and F64_STRICT banks : bit-identical doubles
  F64 banks (FMA)                             : max |err| <= 1e-12 relative to full scale
  F32 banks                                   : SNR >= 120 dB and max |err| <= 1e-5 of full scale
"""
import numpy as np
import pytest

from conftest import KIND, WIN

pytestmark = pytest.mark.gpu

TOL_F64 = 1e-12
TOL_F32_ABS = 1e-5
SNR_F32_DB = 120.0


def snr_db(want, got):
    err = (got.astype(np.float64) - want).ravel()
    return 10 * np.log10((want.ravel() ** 2).sum() / max((err ** 2).sum(), 1e-300))


def oracle_bank(port, h, x, n_out=None):
    return np.stack([port.fir_run(h, x[c], n_out=n_out) for c in range(x.shape[0])])


def test_dropin_streams_match_golden_hashes(zlib, port, kat, cuda):
    """llz_fir_filter frame by frame + flush == the reference's bytes (hash from the unmodified reference)."""
    for c in kat["fir_streams"]:
        f = zlib.FirFilter(KIND[c["kind"]], c["frame"], c["N"], c["fc1"], c["fc2"], WIN[c["win"]])
        x = port.lcg_f64(c["frame"] * c["frames"], c["seed"])
        parts = [f.filter(x[i:i + c["frame"]]) for i in range(0, len(x), c["frame"])]
        parts.append(f.flush())
        y = np.concatenate(parts)
        f.close()
        assert len(y) == c["n_out"]
        assert f"{port.fnv64(y):016x}" == c["fnv"], c
        assert y[0] == c["y0"] and y[-1] == c["ylast"]


def test_dropin_short_last_frame_and_aliasing(zlib, port, cuda):
    h = port.fir_design(0, 63, 0.3, 0.0, 1)
    x = port.lcg_f64(256 * 3 + 100, 5)
    want = port.fir_run(h, x)
    f = zlib.FirFilter(0, 256, 63, 0.3, 0.0, 1)
    got = np.concatenate([f.filter(x[i:i + 256]) for i in range(0, len(x), 256)])
    assert got.tobytes() == want.tobytes()
    # in-place call (the reference copies the input first, llz_fir.c:565-566)
    f2 = zlib.FirFilter(0, 256, 63, 0.3, 0.0, 1)
    buf = x[:256].copy()
    rc = zlib.lib().llz_fir_filter(f2.handle, buf.ctypes.data, buf.ctypes.data, 256)
    assert rc == 256 and buf.tobytes() == want[:256].tobytes()
    # a frame longer than the handle's is rejected, not asserted
    big = np.zeros(300)
    assert zlib.lib().llz_fir_filter(f2.handle, big.ctypes.data, big.ctypes.data, 300) == -1
    f.close(); f2.close()


@pytest.mark.parametrize("frame,N,tail", [(256, 200, 100), (32, 63, 32), (512, 127, 1), (256, 257, 256)])
def test_dropin_history_between_frames(zlib, port, cuda, frame, N, tail):
    """Full frames leave their history in the previous frame's device buffer (no history kernel in between); frames
    shorter than flt_len-1, a short last frame and the flush must still see the right flt_len-1 predecessors."""
    h = port.fir_design(0, N, 0.2, 0.0, 0)
    x = port.lcg_f64(frame * 4 + tail, 11)
    want = port.fir_run(h, np.concatenate([x, np.zeros(N - 1)]))
    f = zlib.FirFilter(0, frame, N, 0.2, 0.0, 0)
    parts = [f.filter(x[i:i + frame]) for i in range(0, len(x), frame)]
    if frame >= N - 1:                                     # the reference's flush needs frame_len >= flt_len-1 (quirk F2)
        parts.append(f.flush())
    got = np.concatenate(parts)
    assert got.tobytes() == want[:len(got)].tobytes()
    # the bank entry points continue a handle's stream from the deferred history, and a reset forgets it
    torch = cuda
    f2 = zlib.FirFilter(0, frame, N, 0.2, 0.0, 0)
    head = np.concatenate([f2.filter(x[i:i + frame]) for i in range(0, 2 * frame, frame)])
    dx = torch.from_numpy(x[2 * frame:]).cuda()
    dy = torch.empty_like(dx)
    zlib.FirBank.run(f2, dx, len(dx), dy, len(dx), len(dx))
    torch.cuda.synchronize()
    assert np.concatenate([head, dy.cpu().numpy()]).tobytes() == want[:len(x)].tobytes()
    zlib.FirBank.reset(f2)
    assert f2.filter(x[:frame]).tobytes() == want[:frame].tobytes()
    f.close(); f2.close()


@pytest.mark.parametrize("N", [1, 2, 15, 16, 17, 33, 64, 127, 128, 255, 1000, 4095])
def test_bank_f64_all_tap_counts(zlib, port, cuda, N):
    torch = cuda
    rng = np.random.default_rng(N)
    h = rng.standard_normal(N) / max(N, 1) ** 0.5
    C_, n = 3, 9000 + N
    x = rng.uniform(-1, 1, (C_, n))
    want = oracle_bank(port, h, x)
    dx = torch.from_numpy(x).cuda()
    cases = [(zlib.F64_STRICT, zlib.FIR_AUTO, True), (zlib.F64, zlib.FIR_DIRECT, False)]
    if N <= 12289:
        cases.append((zlib.F64, zlib.FIR_FFT, False))
    for dtype, algo, exact in cases:
        bank = zlib.FirBank(C_, dtype, taps=h, algo=algo)
        assert bank.algo == (zlib.FIR_FFT if algo == zlib.FIR_FFT else zlib.FIR_DIRECT)
        dy = torch.empty_like(dx)
        bank.run(dx, n, dy, n, n)
        torch.cuda.synchronize()
        got = dy.cpu().numpy()
        if exact:
            assert got.tobytes() == want.tobytes(), N
        else:
            scale = np.abs(h).sum()
            assert np.abs(got - want).max() <= TOL_F64 * max(scale, 1.0), N
        bank.close()


@pytest.mark.parametrize("N,fc,win,algo", [(127, 0.23, 0, 1), (4095, 0.11, 2, 1), (48, 0.4, 1, 1),
                                          (127, 0.23, 0, 2), (48, 0.4, 1, 2), (513, 0.11, 2, 2), (897, 0.05, 1, 2),
                                          (4095, 0.11, 2, 2), (6145, 0.02, 1, 2)])
def test_bank_f32_meets_snr(zlib, port, cuda, N, fc, win, algo):
    torch = cuda
    h = port.fir_design(0, N, fc, 0.0, win)
    C_, n = 4, 50000
    x = np.stack([port.lcg_f64(n, 12345 + c) for c in range(C_)])
    want = oracle_bank(port, h, x)
    bank = zlib.FirBank(C_, zlib.F32, kind=zlib.LPF, flt_len=N, fc1=fc, win=win, algo=algo)
    assert bank.taps().tobytes() == h.tobytes()
    dx = torch.from_numpy(x.astype(np.float32)).cuda()
    dy = torch.empty_like(dx)
    bank.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    got = dy.cpu().numpy()
    assert snr_db(want, got) >= SNR_F32_DB, snr_db(want, got)
    assert np.abs(got - want).max() <= TOL_F32_ABS
    bank.close()


def test_bank_streaming_equals_one_shot(zlib, port, cuda):
    """ragged chunk sizes (incl. shorter than the history) and unaligned strides keep the stream exact"""
    torch = cuda
    h = port.fir_design(2, 129, 0.2, 0.6, 2)
    C_, n = 5, 40001
    x = np.stack([port.lcg_f64(n, 100 + c) for c in range(C_)])
    want = oracle_bank(port, h, x, n_out=n + 128)
    stride = n + 3                                       # odd stride: the non-vector path
    dx = torch.zeros(C_, stride, dtype=torch.float64, device="cuda")
    dx[:, :n] = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, stride + 200, dtype=torch.float64, device="cuda")
    bank = zlib.FirBank(C_, zlib.F64_STRICT, taps=h)
    pos = 0
    for step in (1, 7, 100, 128, 129, 5000, 3, 20000, 10 ** 9):
        m = min(step, n - pos)
        if m <= 0:
            break
        bank.run(dx.data_ptr() + 8 * pos, stride, dy.data_ptr() + 8 * pos, stride + 200, m)
        pos += m
    assert pos == n
    assert bank.flush(dy.data_ptr() + 8 * n, stride + 200) == 128
    torch.cuda.synchronize()
    got = dy.cpu().numpy()[:, :n + 128]
    assert got.tobytes() == want.tobytes()
    # after a flush the stream restarts from silence
    bank.run(dx, stride, dy, stride + 200, 1000)
    torch.cuda.synchronize()
    assert dy.cpu().numpy()[:, :1000].tobytes() == want[:, :1000].tobytes()
    bank.close()


def test_time_segments_with_halo_are_byte_identical(zlib, port, cuda):
    """the multi-GPU time-segment plan, all segments run on one GPU"""
    torch = cuda
    N, C_, n = 255, 2, 100003
    h = port.fir_design(0, N, 0.11, 0.0, 2)
    x = np.stack([port.lcg_f64(n, 7 + c) for c in range(C_)])
    dx = torch.from_numpy(x).cuda()
    one = torch.empty_like(dx)
    bank = zlib.FirBank(C_, zlib.F64, taps=h, algo=zlib.FIR_DIRECT)
    bank.run(dx, n, one, n, n)
    for world in (2, 4, 8):
        seg_out = torch.zeros_like(dx)
        covered = 0
        for rank in range(world):
            s = zlib.shard_fir_segments(n, N, world, rank)
            assert s.out_start == covered and s.in_start == s.out_start
            covered += s.out_count
            bank.reset()
            if s.halo:
                assert s.halo == N - 1
                bank.set_history(dx.data_ptr() + 8 * (s.in_start - s.halo), n)
            bank.run(dx.data_ptr() + 8 * s.in_start, n, seg_out.data_ptr() + 8 * s.out_start, n, s.in_count)
        assert covered == n
        torch.cuda.synchronize()
        assert torch.equal(seg_out, one), world
    want = oracle_bank(port, h, x)
    assert np.abs(one.cpu().numpy() - want).max() <= TOL_F64
    bank.close()
    # overlap-save kernels: arbitrary cuts agree with the one-shot run to rounding (the block grid moves with the
    # segment start); cuts at multiples of the work-item length are byte-identical (llz_cuda_fir_bank_block_len)
    for taps, n2 in ((h, n), (port.fir_design(0, 1025, 0.11, 0.0, 2), 300_001)):
        N2 = len(taps)
        x2 = np.stack([port.lcg_f64(n2, 7 + c) for c in range(C_)])
        dx2 = torch.from_numpy(x2).cuda()
        fbank = zlib.FirBank(C_, zlib.F64, taps=taps, algo=zlib.FIR_FFT)
        blk = fbank.block_len
        assert blk == (2 * (1024 - 256) if N2 == 255 else 2 * (8192 - 1024))
        assert blk == fbank.block_len
        one2 = torch.empty_like(dx2)
        fbank.run(dx2, n2, one2, n2, n2)
        want2 = oracle_bank(port, taps, x2)
        assert np.abs(one2.cpu().numpy() - want2).max() <= TOL_F64
        for world in (2, 8):
            for aligned in (False, True):
                seg_out = torch.zeros_like(dx2)
                for rank in range(world):
                    s = (zlib.shard_fir_segments_aligned(n2, N2, blk, world, rank) if aligned
                         else zlib.shard_fir_segments(n2, N2, world, rank))
                    fbank.reset()
                    if s.halo:
                        fbank.set_history(dx2.data_ptr() + 8 * (s.in_start - s.halo), n2)
                    fbank.run(dx2.data_ptr() + 8 * s.in_start, n2, seg_out.data_ptr() + 8 * s.out_start, n2, s.in_count)
                torch.cuda.synchronize()
                if aligned:
                    assert torch.equal(seg_out, one2), (N2, world)
                else:
                    assert np.abs(seg_out.cpu().numpy() - want2).max() <= TOL_F64, (N2, world)
        fbank.close()


@pytest.mark.parametrize("N,size", [(898, 8192), (2049, 8192), (4095, 8192), (6145, 8192),
                                    (2305, 16384), (3700, 16384), (4095, 16384), (8191, 16384), (12289, 16384)])
def test_fft8k_16k_bank_interior_and_edge_items(zlib, port, cuda, N, size):
    """8192-point (one CTA per item) and 16384-point (one CTA per item in two rounds, half of the item in an L2-resident scratch)
    overlap-save kernels: several interior items, history splice, ragged end"""
    torch = cuda
    rng = np.random.default_rng(N)
    h = rng.standard_normal(N) / N ** 0.5
    C_, n = 3, 70001 if size == 8192 else 150001
    x = rng.uniform(-1, 1, (C_, n))
    want = oracle_bank(port, h, x)
    scale = max(np.abs(h).sum(), 1.0)
    for dtype, tdt, npdt, tol in ((zlib.F64, torch.float64, np.float64, TOL_F64 * scale),
                                  (zlib.F32, torch.float32, np.float32, TOL_F32_ABS * scale)):
        dx = torch.from_numpy(x.astype(npdt)).cuda()
        bank = zlib.FirBank(C_, dtype, taps=h, algo=zlib.FIR_FFT)
        bank.set_fft_size(size)
        assert bank.block_len == 2 * (size - (N - 1 + size // 32 - 1) // (size // 32) * (size // 32))
        dy = torch.full((C_, n + 64), 7.0, dtype=tdt, device="cuda")
        bank.run(dx, n, dy, n + 64, n)
        # second call continues the stream from the stored history: same samples again -> history splice
        dy2 = torch.empty(C_, 20000, dtype=tdt, device="cuda")
        bank.run(dx, n, dy2, 20000, 20000)
        torch.cuda.synchronize()
        got = dy.cpu().numpy()
        assert np.abs(got[:, :n] - want).max() <= tol, (N, dtype)
        assert (got[:, n:] == 7.0).all()
        if dtype == zlib.F32:
            assert snr_db(want, got[:, :n]) >= SNR_F32_DB
        want2 = oracle_bank(port, h, np.concatenate([x, x[:, :20000]], axis=1))[:, n:]
        assert np.abs(dy2.cpu().numpy() - want2).max() <= tol, (N, dtype)
        bank.close()


@pytest.mark.parametrize("dtype,N,size", [("f64", 544, 1024), ("f64", 545, 8192), ("f64", 3328, 8192), ("f64", 3329, 16384),
                                          ("f64", 4095, 16384), ("f32", 2304, 8192), ("f32", 2305, 16384), ("f32", 4095, 16384)])
def test_auto_transform_length_by_tap_count(zlib, cuda, dtype, N, size):
    """LLZ_CUDA_FIR_ALGO_AUTO: the transform length follows the measured crossovers (profiles/r02_crossover_fft16k.txt)"""
    bank = zlib.FirBank(2, zlib.F64 if dtype == "f64" else zlib.F32, kind=zlib.LPF, flt_len=N, fc1=0.2)
    pad = {1024: 32, 8192: 256, 16384: 512}[size]
    assert bank.algo == zlib.FIR_FFT
    assert bank.block_len == 2 * (size - (N - 1 + pad - 1) // pad * pad)
    bank.close()


@pytest.mark.parametrize("dtype,N", [("f64", 129), ("f32", 129), ("f64", 2049), ("f64", 4095), ("f32", 2561)])
def test_fft_bank_streaming_ragged_chunks(zlib, port, cuda, dtype, N):
    """overlap-save kernels: chunk sizes below / around the history and the block size, odd strides, flush, restart"""
    torch = cuda
    h = port.fir_design(2, N, 0.2, 0.6, 2)
    C_, n = 5, 40001
    x = np.stack([port.lcg_f64(n, 100 + c) for c in range(C_)])
    want = oracle_bank(port, h, x, n_out=n + N - 1)
    tdt, npdt, es = (torch.float64, np.float64, 8) if dtype == "f64" else (torch.float32, np.float32, 4)
    tol = TOL_F64 if dtype == "f64" else TOL_F32_ABS
    stride = n + 3
    dx = torch.zeros(C_, stride, dtype=tdt, device="cuda")
    dx[:, :n] = torch.from_numpy(x.astype(npdt)).cuda()
    ostride = stride + 200 + N
    dy = torch.full((C_, ostride), 7.0, dtype=tdt, device="cuda")
    bank = zlib.FirBank(C_, zlib.F64 if dtype == "f64" else zlib.F32, taps=h, algo=zlib.FIR_FFT)
    pos = 0
    for step in (1, 7, 100, 128, 129, 895, 896, 1792, 1793, 5000, 3, 20000, 10 ** 9):
        m = min(step, n - pos)
        if m <= 0:
            break
        bank.run(dx.data_ptr() + es * pos, stride, dy.data_ptr() + es * pos, ostride, m)
        pos += m
    assert pos == n
    assert bank.flush(dy.data_ptr() + es * n, ostride) == N - 1
    torch.cuda.synchronize()
    got = dy.cpu().numpy()
    assert np.abs(got[:, :n + N - 1] - want).max() <= tol
    assert (got[:, n + N - 1:] == 7.0).all()                 # nothing written past the end
    if dtype == "f32":
        assert snr_db(want, got[:, :n + N - 1]) >= SNR_F32_DB
    bank.run(dx, stride, dy, ostride, 1000)
    torch.cuda.synchronize()
    assert np.abs(dy.cpu().numpy()[:, :1000] - want[:, :1000]).max() <= tol
    bank.close()


def test_fft_algo_selection(zlib, cuda):
    """AUTO picks overlap-save for tolerance-mode banks of 48..12289 taps; STRICT and longer filters stay direct"""
    h = np.ones(127) / 127
    for dtype, N, want in ((zlib.F64, 127, zlib.FIR_FFT), (zlib.F32, 127, zlib.FIR_FFT), (zlib.F64, 47, zlib.FIR_DIRECT),
                           (zlib.F64, 898, zlib.FIR_FFT), (zlib.F64, 4095, zlib.FIR_FFT), (zlib.F64, 6146, zlib.FIR_FFT),
                           (zlib.F64, 12290, zlib.FIR_DIRECT),
                           (zlib.F64_STRICT, 127, zlib.FIR_DIRECT)):
        b = zlib.FirBank(2, dtype, taps=np.ones(N) / N)
        assert b.algo == want, (dtype, N)
        b.close()
    b = zlib.FirBank(2, zlib.F64_STRICT, taps=h)
    assert zlib.lib().llz_cuda_fir_bank_set_algo(b.handle, zlib.FIR_FFT) == -1
    assert zlib.lib().llz_cuda_fir_bank_set_algo(b.handle, 9) == -1
    assert b.algo == zlib.FIR_DIRECT
    b.close()
    b = zlib.FirBank(2, zlib.F64, taps=np.ones(12290) / 12290)
    assert zlib.lib().llz_cuda_fir_bank_set_algo(b.handle, zlib.FIR_FFT) == -1
    b.close()


def test_run_host_pipeline(zlib, port, cuda):
    """host buffers -> chunked H2D / kernel / D2H; several chunks; pinned and pageable"""
    torch = cuda
    N, C_, n = 127, 16, 1_500_000                   # 192 MB in: three pipeline chunks
    h = port.fir_design(0, N, 0.23, 0.0, 0)
    x = zlib.host_alloc(C_ * n * 8, np.float64).reshape(C_, n)
    for c in range(C_):
        x[c] = port.lcg_f64(n, 12345 + c)
    y = zlib.host_alloc(C_ * n * 8, np.float64).reshape(C_, n)
    bank = zlib.FirBank(C_, zlib.F64, taps=h, algo=zlib.FIR_DIRECT)
    bank.run_host(x, n, y, n, n)
    dx = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    dy = torch.empty_like(dx)
    bank.reset()
    bank.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), y)
    for c in (0, 7, 15):
        for t in (0, 126, 524288, 524289, n - 1):
            lo = max(0, t - N + 1)
            want = port.fir_run(h, x[c, lo:t + 1])[-1]
            assert abs(y[c, t] - want) <= TOL_F64
    # pageable memory takes the same path
    xp, yp = np.array(x[:, :300000]), np.empty((C_, 300000))
    bank.reset()
    bank.run_host(xp, 300000, yp, 300000, 300000)
    assert np.array_equal(yp, y[:, :300000])
    bank.close()
    # the overlap-save kernel through the same pipeline: chunks are whole work items -> the one-shot bytes
    fbank = zlib.FirBank(C_, zlib.F64, taps=h)
    assert fbank.algo == zlib.FIR_FFT
    y2 = np.empty_like(np.asarray(y))
    fbank.run_host(x, n, y2, n, n)
    assert np.abs(y2 - y).max() <= TOL_F64
    fbank.reset()
    fbank.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), y2)
    fbank.close()
    zlib.host_free(x.reshape(-1)); zlib.host_free(y.reshape(-1))


@pytest.mark.parametrize("algo", [1, 2])
def test_run_host_channel_groups_with_padded_rows(zlib, port, cuda, small_pipe_slots, algo):
    """many short channels -> the host pipeline runs groups of whole channels; padded (non-dense) rows take the
    row-wise copy inside a group; the stream state carries over to the next call for every group"""
    torch = cuda
    # small_pipe_slots: 4 MiB slots -> groups of 2 channels, 21 groups
    N, C_, n = 127, 41, 100_003
    h = port.fir_design(0, N, 0.23, 0.0, 0)
    xs, ys = n + 5, n + 9
    x = np.zeros((C_, xs))
    for c in range(C_):
        x[c, :n] = port.lcg_f64(n, 500 + c)
    y = np.full((C_, ys), 7.0)
    bank = zlib.FirBank(C_, zlib.F64, taps=h, algo=algo)
    half = 60_000
    bank.run_host(x, xs, y, ys, half)                           # two calls: history of every channel group carries over
    bank.run_host(x[:, half:], xs, y[:, half:], ys, n - half)
    want = oracle_bank(port, h, x[:, :n])
    assert np.abs(y[:, :n] - want).max() <= TOL_F64
    assert (y[:, n:] == 7.0).all()
    # dense rows, same bank: one contiguous copy per group; equals the device-resident run bit for bit
    xd, yd = np.ascontiguousarray(x[:, :n]), np.empty((C_, n))
    bank.reset()
    bank.run_host(xd, n, yd, n, n)
    dx = torch.from_numpy(xd).cuda()
    dy = torch.empty_like(dx)
    bank.reset()
    bank.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), yd)
    bank.close()


def test_c2_full_size_properties(zlib, port, cuda):
    """BASELINE config 2 at full size: 1024 channels x 480,000 samples, 127-tap LPF (f64).
    Spot checks against the oracle plus linearity and a DC-gain check (size-independent properties)."""
    torch = cuda
    C_, n, N = 1024, 480_000, 127
    h = port.fir_design(0, N, 0.23, 0.0, 0)
    dx = torch.empty(C_, n, dtype=torch.float64, device="cuda")
    zlib.synth_lcg(dx, n, C_, n, 0, 12345)
    bank = zlib.FirBank(C_, zlib.F64, kind=zlib.LPF, flt_len=N, fc1=0.23, win=zlib.HAMMING)
    dy = torch.empty_like(dx)
    bank.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    for c in [0, 1, 511, 1023] + rng.integers(0, C_, 12).tolist():
        xc = port.lcg_f64(n, 12345 + c)
        assert np.array_equal(dx[c, :4096].cpu().numpy(), xc[:4096])        # device LCG == host LCG
        yc = dy[c].cpu().numpy()
        for t in [0, 1, 125, 126, 127, 3583, 3584, n - 1] + rng.integers(0, n, 40).tolist():
            lo = max(0, t - N + 1)
            want = port.fir_run(h, xc[lo:t + 1])[-1]
            assert abs(yc[t] - want) <= TOL_F64, (c, t)
    # linearity: F(2x) == 2 F(x) exactly (power-of-two scaling commutes with every rounding)
    bank.reset()
    dx.mul_(2.0)
    dy2 = torch.empty_like(dy)
    bank.run(dx, n, dy2, n, n)
    torch.cuda.synchronize()
    assert torch.equal(dy2, dy * 2.0)
    # DC: a constant input settles to sum(h)
    bank.reset()
    dx.fill_(1.0)
    bank.run(dx, n, dy2, n, n)
    torch.cuda.synchronize()
    assert torch.allclose(dy2[:, N:], torch.full_like(dy2[:, N:], float(h.sum())), rtol=0, atol=1e-13)
    bank.close()


def test_bad_handles_and_arguments(zlib, cuda):
    L = zlib.lib()
    assert L.llz_cuda_fir_bank_run(0, None, 0, None, 0, 10, None) == -1
    assert L.llz_cuda_fir_bank_init(0, 127, 0.2, 0.0, 0, 0, 0) == zlib.FAIL
    assert L.llz_cuda_fir_bank_init(0, 127, 0.2, 0.0, 0, 4, 9) == zlib.FAIL
    r = zlib.ResampleBank(zlib.KIND_RESAMPLE, 2, 1, 1)
    assert L.llz_cuda_fir_bank_run(r.handle, None, 0, None, 0, 10, None) == -1
    assert "not a FIR handle" in zlib.last_error()
    r.close()
