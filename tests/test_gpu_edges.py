"""Edge cases of the C-ABI on the GPU: empty and tiny calls, limits, repeated flush, handle misuse."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fir_empty_and_tiny_calls(zlib, port, cuda):
    torch = cuda
    h = port.fir_design(0, 33, 0.3, 0.0, 0)
    bank = zlib.FirBank(2, zlib.F64_STRICT, taps=h)
    x = np.stack([port.lcg_f64(100, 1), port.lcg_f64(100, 2)])
    want = np.stack([port.fir_run(h, x[c]) for c in range(2)])
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros_like(dx)
    bank.run(dx, 100, dy, 100, 0)                      # n = 0: nothing happens, state untouched
    pos = 0
    for m in (1, 1, 2, 31, 32, 33):                    # pieces shorter than, equal to and longer than the history
        bank.run(dx.data_ptr() + 8 * pos, 100, dy.data_ptr() + 8 * pos, 100, m)
        pos += m
    bank.run(dx.data_ptr() + 8 * pos, 100, dy.data_ptr() + 8 * pos, 100, 100 - pos)
    torch.cuda.synchronize()
    assert dy.cpu().numpy().tobytes() == want.tobytes()
    assert zlib.lib().llz_cuda_fir_bank_run(bank.handle, dx.data_ptr(), 100, dy.data_ptr(), 100, -1, None) == -1
    # flushing twice: the second flush sees silence
    tail = torch.zeros(2, 32, dtype=torch.float64, device="cuda")
    assert bank.flush(tail, 32) == 32
    assert bank.flush(tail, 32) == 32
    torch.cuda.synchronize()
    assert not tail.cpu().numpy().any()
    bank.close()


def test_fir_single_tap_and_limits(zlib, cuda):
    torch = cuda
    bank = zlib.FirBank(1, zlib.F64, taps=np.array([2.5]))
    x = torch.arange(1000, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(x)
    bank.run(x, 1000, y, 1000, 1000)
    torch.cuda.synchronize()
    assert torch.equal(y, x * 2.5)
    assert bank.flush(y, 1000) == 0                     # no history to flush
    bank.close()
    # a filter whose tile cannot fit shared memory is refused at run time with a message, not a crash
    big = zlib.FirBank(1, zlib.F64, taps=np.ones(20000) / 20000)
    rc = zlib.lib().llz_cuda_fir_bank_run(big.handle, x.data_ptr(), 1000, y.data_ptr(), 1000, 1000, None)
    assert rc == -1 and "shared memory" in zlib.last_error()
    big.close()
    L = zlib.lib()
    assert L.llz_cuda_fir_bank_init_taps(None, 0, 1, 0) == zlib.FAIL
    assert L.llz_cuda_fir_bank_init(0, 127, 0.2, 0.0, 0, 70000, 0) == zlib.FAIL   # more channels than one launch takes


def test_resample_empty_tiny_and_out_len(zlib, port, cuda):
    torch = cuda
    for L_, M in ((160, 147), (1, 3), (3, 2)):
        bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L_, M, 1)
        assert bank.out_len(0) == 0
        n_in = 5000
        x = port.lcg_s16(n_in, 3)
        total = -((-n_in * L_) // M)
        p = port.resample_plan(L_, M, 1)
        want = port.resample_run(p, 1.0, x, total)
        dx = torch.from_numpy(x).cuda()
        dy = torch.zeros(total + 4, dtype=torch.int16, device="cuda")
        assert bank.run(dx, n_in, 0, dy, total) == 0
        pos_in = pos_out = 0
        for m in [1] * 7 + [2, 3, 5, n_in]:             # one sample at a time: some calls produce no output
            m = min(m, n_in - pos_in)
            expect = -((-(pos_in + m) * L_) // M) - pos_out
            assert bank.out_len(m) == expect
            got = bank.run(dx.data_ptr() + 2 * pos_in, n_in, m, dy.data_ptr() + 2 * pos_out, total)
            assert got == expect
            pos_in += m
            pos_out += got
        torch.cuda.synchronize()
        assert pos_out == total
        assert np.array_equal(dy.cpu().numpy()[:total], want), (L_, M)
        bank.close()


def test_interp_requires_whole_frames(zlib, cuda):
    bank = zlib.ResampleBank(zlib.KIND_INTERP, 2, 1, 1)
    t = cuda.zeros(4096, dtype=cuda.int16, device="cuda")
    n_out = C.c_longlong(0)
    rc = zlib.lib().llz_cuda_resample_bank_run(bank.handle, t.data_ptr(), 100, 100, t.data_ptr(), 4096, C.byref(n_out), None)
    assert rc == -1 and "whole frames" in zlib.last_error()
    bank.close()


def test_handles_are_tagged(zlib, cuda):
    L = zlib.lib()
    f = zlib.FirFilter(0, 64, 15, 0.3)
    r = zlib.Resampler(zlib.KIND_RESAMPLE, 3, 2)
    buf = np.zeros(4096, np.int16)
    osz = C.c_int(0)
    assert L.llz_resample(f.handle, buf.ctypes.data, r.bytes_in, buf.ctypes.data, C.byref(osz)) == -1
    assert L.llz_fir_filter(r.handle, buf.ctypes.data, buf.ctypes.data, 16) == -1
    assert L.llz_get_resample_framelen_bytes(f.handle) == -1
    L.llz_fir_filter_uninit(r.handle)                  # wrong uninit is a no-op, not a crash
    assert L.llz_get_resample_framelen_bytes(r.handle) == r.bytes_in
    f.close(); r.close()


def test_two_handles_on_two_host_threads(zlib, port, cuda):
    """no globals in the library: distinct handles may run from distinct threads (SURVEY.md 8b threading)"""
    import threading
    out = {}

    def work(seed):
        r = zlib.Resampler(zlib.KIND_RESAMPLE, 160, 147)
        x = port.lcg_s16(r.bytes_in // 2 * 3, seed)
        out[seed] = (x, r.stream(x))
        r.close()

    ths = [threading.Thread(target=work, args=(s,)) for s in (11, 22, 33)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    p = port.resample_plan(160, 147, 1)
    for seed, (x, y) in out.items():
        assert np.array_equal(y, port.resample_run(p, 1.0, x, len(y)))


@pytest.mark.parametrize("kind,L_,M", [(2, 160, 147), (2, 1, 3), (0, 1, 3), (2, 3, 2)])
def test_dropin_frames_then_bank_run_on_the_same_handle(zlib, port, cuda, kind, L_, M):
    """Drop-in frames leave their history in the previous frame's device buffer (no history kernel between frames);
    a bank run on the same handle must pick the stream up from there, and a reset must forget it."""
    torch = cuda
    r = zlib.Resampler(kind, L_, M, 1.0, zlib.BLACKMAN)
    nf = r.bytes_in // 2
    plan = port.decimate_plan(M, 1) if kind == zlib.KIND_DECIMATE else port.resample_plan(L_, M, 1)
    run = port.decimate_run if kind == zlib.KIND_DECIMATE else port.resample_run
    x = port.lcg_s16(nf * 5, 99)
    total = plan.num_out * 5
    want = run(plan, 1.0, x, total)
    got = [r.frame(x[i * nf:(i + 1) * nf]) for i in range(3)]
    dx = torch.from_numpy(x[3 * nf:]).cuda()
    dy = torch.zeros(2 * plan.num_out + 8, dtype=torch.int16, device="cuda")
    n_out = zlib.ResampleBank.run(r, dx, 2 * nf, 2 * nf, dy, dy.numel())      # same handle through the bank entry point
    torch.cuda.synchronize()
    assert n_out == 2 * plan.num_out
    got.append(dy.cpu().numpy()[:n_out])
    assert np.array_equal(np.concatenate(got), want)
    # ... and frames again after the bank run continue the same stream state machine from a reset
    zlib.ResampleBank.reset(r)
    again = np.concatenate([r.frame(x[i * nf:(i + 1) * nf]) for i in range(2)])
    assert np.array_equal(again, want[:2 * plan.num_out])
    r.close()
