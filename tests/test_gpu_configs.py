"""BASELINE configs 4 and 5 at full per-channel length (one or two channels), checked through windows against the
oracle: exercises 64-bit index / phase arithmetic (m*M exceeds 2^32 for C4) and long-stream tiling."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c4_full_length_one_channel(zlib, port, cuda):
    """44.1 kHz -> 96 kHz, 1 h: 158,760,000 in -> 345,600,000 out per channel, L=320 M=147, k=128 (Q=257)"""
    torch = cuda
    L, M, n_in = 320, 147, 158_760_000
    for acc, tol in ((zlib.ACC_F64, 0), (zlib.ACC_F32, 1)):
        bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L, M, 1, k_override=128, acc=acc)
        assert bank.info.taps_per_phase == 257 and n_in % bank.info.num_in == 0
        dx = torch.empty(n_in, dtype=torch.int16, device="cuda")
        zlib.synth_lcg(dx, n_in, 1, n_in, 2, 777)
        n_out = bank.out_len(n_in)
        assert n_out == 345_600_000
        dy = torch.empty(n_out, dtype=torch.int16, device="cuda")
        assert bank.run(dx, n_in, n_in, dy, n_out) == n_out
        torch.cuda.synchronize()
        plan = port.resample_plan(L, M, 1, 128)
        rng = np.random.default_rng(4)
        for m0 in [0, n_out - 5000, 2 ** 31 // 147 * 320 // 320 - 100] + rng.integers(0, n_out - 5000, 5).tolist():
            # the oracle needs x only around floor(m*M/L): rebuild that window from the LCG (device copy)
            lo = max(0, (m0 * M) // L - 300)
            hi = min(n_in, ((m0 + 5000) * M) // L + 2)
            xw = np.zeros(hi, np.int16)                      # zeros before `lo` are never touched by these outputs
            xw[lo:hi] = dx[lo:hi].cpu().numpy()
            want = port.resample_run(plan, 1.0, xw, 5000, m0=m0)
            got = dy[m0:m0 + 5000].cpu().numpy()
            assert np.abs(got.astype(np.int32) - want).max() <= tol, (acc, m0)
        if acc == zlib.ACC_F64:
            assert bank.guard_count() < 1000
        bank.close()
        del dx, dy
        torch.cuda.empty_cache()


def test_c5_full_length_one_channel(zlib, port, cuda):
    """4095-tap KAISER low-pass over 1 h @ 192 kHz = 691,200,000 samples (5.5 GB in, 5.5 GB out as double)"""
    torch = cuda
    n, N = 691_200_000, 4095
    h = port.fir_design(0, N, 0.11, 0.0, 2)
    dx = torch.empty(n, dtype=torch.float64, device="cuda")
    zlib.synth_lcg(dx, n, 1, n, 0, 12345)
    dy = torch.empty_like(dx)
    bank = zlib.FirBank(1, zlib.F64, kind=zlib.LPF, flt_len=N, fc1=0.11, win=zlib.KAISER)
    bank.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    rng = np.random.default_rng(5)
    for t0 in [0, 4094, n - 200, 2 ** 29 - 100, 2 ** 32 // 8 - 50] + rng.integers(0, n - 200, 4).tolist():
        lo = max(0, t0 - (N - 1))
        xw = dx[lo:t0 + 200].cpu().numpy()
        want = port.fir_run(h, xw)[t0 - lo:]
        got = dy[t0:t0 + 200].cpu().numpy()
        assert np.abs(got - want).max() <= 1e-12, t0
    # the same stream in two time segments with halo (C5's multi-GPU plan): the overlap-save kernel (AUTO picks the
    # 16384-point one at 4095 taps, profiles/r02_crossover_fft16k.txt) agrees to rounding -- its block grid starts at
    # the segment -- and the direct kernel bit for bit
    blk = bank.block_len
    assert bank.algo == zlib.FIR_FFT and blk == 2 * (16384 - 4096)      # the 16384-point kernel
    seg = zlib.shard_fir_segments(n, N, 2, 1)
    bank.reset()
    bank.set_history(dx.data_ptr() + 8 * (seg.in_start - seg.halo), n)
    tail = torch.empty(1_000_000, dtype=torch.float64, device="cuda")
    bank.run(dx.data_ptr() + 8 * seg.in_start, n, tail, 1_000_000, 1_000_000)
    torch.cuda.synchronize()
    assert (tail - dy[seg.out_start:seg.out_start + 1_000_000]).abs().max().item() <= 1e-12
    # ... and bit for bit when the cut is a multiple of the kernel's work-item length (bench.py's plan for C5)
    aseg = zlib.shard_fir_segments_aligned(n, N, blk, 8, 5)
    assert aseg.in_start % blk == 0 and aseg.halo == N - 1
    bank.reset()
    bank.set_history(dx.data_ptr() + 8 * (aseg.in_start - aseg.halo), n)
    m = blk * 40
    bank.run(dx.data_ptr() + 8 * aseg.in_start, n, tail, m, m)
    torch.cuda.synchronize()
    assert torch.equal(tail[:m], dy[aseg.out_start:aseg.out_start + m])
    bank.close()
    direct = zlib.FirBank(1, zlib.F64, kind=zlib.LPF, flt_len=N, fc1=0.11, win=zlib.KAISER, algo=zlib.FIR_DIRECT)
    lead = 8192                                            # one-shot run over a window that starts before the segment
    whole = torch.empty(lead + 1_000_000, dtype=torch.float64, device="cuda")
    direct.run(dx.data_ptr() + 8 * (seg.in_start - lead), n, whole, whole.numel(), whole.numel())
    direct.reset()
    direct.set_history(dx.data_ptr() + 8 * (seg.in_start - seg.halo), n)
    direct.run(dx.data_ptr() + 8 * seg.in_start, n, tail, 1_000_000, 1_000_000)
    torch.cuda.synchronize()
    assert torch.equal(tail, whole[lead:])
    assert (tail - dy[seg.out_start:seg.out_start + 1_000_000]).abs().max().item() <= 1e-12
    direct.close()
