"""llz_iir on the GPU: the drop-in handle of llz_iir.h and the batched banks must give the
reference's doubles bit for bit (oracle: the restatement of libllzfilter/llz_iir.c:103-156, pinned to the compiled
reference by tests/test_oracle.py), frame by frame and across calls, including the flush."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FILTERS = [
    ([1.0, -1.8, 0.81], [0.2, 0.3, 0.2]),                                   # biquad, complex poles near the unit circle
    ([1.0, -0.5], [1.0]),                                                   # one pole, one tap
    ([1.0], [0.25, 0.25, 0.25, 0.25]),                                      # no feedback: an FIR
    ([1.0, -2.369513, 2.313988, -1.054665, 0.187379], [0.004824, 0.019297, 0.028946, 0.019297, 0.004824]),   # 4th-order Butterworth
    ([1.0, 0.1, -0.2, 0.05, 0.01, -0.02, 0.003, 0.004, -0.001], None),      # b == NULL: all-zero numerator (llz_iir.c:56-60)
]


@pytest.mark.parametrize("a,b", FILTERS)
def test_dropin_frames_and_flush_are_bit_identical(zlib, port, cuda, a, b):
    a = np.array(a)
    bb = np.array(b) if b is not None else np.zeros(1)
    x = port.lcg_f64(5000, 41)
    f = zlib.IirFilter(a, b)
    got = [f.filter(x[t0:t0 + 777]) for t0 in range(0, len(x), 777)]
    got.append(f.flush())
    f.close()
    st = (np.zeros(len(bb)), np.zeros(len(a)))
    want = [port.iir_run(a, bb, x, state=st)]
    if len(bb) > 1:
        want.append(port.iir_run(a, bb, None, len(bb) - 1, state=st))
    assert np.concatenate(got).tobytes() == np.concatenate(want).tobytes()


@pytest.mark.parametrize("C_,n", [(1, 100), (33, 4097), (70, 1000)])
def test_bank_channels_and_streaming(zlib, port, cuda, C_, n):
    torch = cuda
    a, b = np.array(FILTERS[3][0]), np.array(FILTERS[3][1])
    x = np.stack([port.lcg_f64(n, 900 + c) for c in range(C_)])
    bank = zlib.IirBank(a, b, C_)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(C_, n + 5, dtype=torch.float64, device="cuda")
    cut = n // 3 + 1
    bank.run(dx, n, dy, n + 5, cut)                                          # two calls: the state carries over
    bank.run(dx.data_ptr() + 8 * cut, n, dy.data_ptr() + 8 * cut, n + 5, n - cut)
    torch.cuda.synchronize()
    got = dy.cpu().numpy()
    assert not got[:, n:].any()
    for c in range(C_):
        assert got[c, :n].tobytes() == port.iir_run(a, b, x[c]).tobytes(), c
    bank.reset()
    bank.run(dx, n, dx, n, n)                                                # in place, after a reset
    torch.cuda.synchronize()
    assert dx.cpu().numpy().tobytes() == got[:, :n].tobytes()
    bank.close()
