"""Multi-GPU contexts and sharded jobs (include/llz_cuda.h, "Multi-GPU"): one process driving every visible GPU (up to 4).
SURVEY.md section 8(e)'s invariant: the concatenated result of an N-way sharded job is byte-identical to the one-GPU
call, whatever the shard mode and the gather mode.  On a one-GPU box the context has world 1 and the same code paths
run degenerate (chunked compute, result buffer, the rank's own peer "mapping"); tools/mgpu_check.py repeats the checks
with one process per GPU under torchrun (profiles/r02_mgpu_check_*.txt)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mgpu(zlib, cuda):
    n = min(cuda.cuda.device_count(), 4)
    ctx = zlib.Mgpu(n)
    assert ctx.world == n and ctx.nlocal == n and ctx.ranks == list(range(n))
    yield ctx
    ctx.close()
    cuda.cuda.set_device(0)


def run_job(zlib, torch, ctx, job, n_total, x, np_dt, tdt, gather, chunks=3):
    """scatter the input shards (with their halo), run, return the assembled planar result"""
    C_ = x.shape[0]
    n_out = job.out_len(n_total)
    d_in, in_stride, d_out, out_stride, shards = [], [], [], [], []
    for i, r in enumerate(ctx.ranks):
        sh = job.plan(n_total, r)
        shards.append(sh)
        dev = torch.device("cuda", ctx.devices[i])
        lo, hi = sh.seg.in_start - sh.seg.halo, sh.seg.in_start + sh.seg.in_count
        xi = np.ascontiguousarray(x[sh.first_channel:sh.first_channel + sh.n_channels, lo:hi])
        d_in.append(torch.from_numpy(xi).to(dev))
        in_stride.append(hi - lo)
        d_out.append(torch.full((sh.n_channels, sh.seg.out_count), 77, dtype=tdt, device=dev))
        out_stride.append(sh.seg.out_count)
    es = np.dtype(np_dt).itemsize
    if gather != zlib.GATHER_NONE:
        ctx.result_alloc(0, C_ * n_out * es)
    job.run(n_total, d_in, in_stride, d_out, out_stride, n_out, gather, chunks, [0] * ctx.nlocal)
    for d in ctx.devices:
        torch.cuda.synchronize(d)
    if gather == zlib.GATHER_NONE:
        full = np.zeros((C_, n_out), np_dt)
        for sh, o in zip(shards, d_out):
            full[sh.first_channel:sh.first_channel + sh.n_channels,
                 sh.seg.out_start:sh.seg.out_start + sh.seg.out_count] = o.cpu().numpy()
    else:
        import ctypes
        root_dev = ctx.devices[0]
        torch.cuda.set_device(root_dev)
        buf = torch.empty(C_ * n_out, dtype=tdt, device=torch.device("cuda", root_dev))
        zlib_cuda_memcpy(torch, buf, ctx.result_ptr(0), C_ * n_out * es)
        full = buf.cpu().numpy().reshape(C_, n_out)
        ctx.result_free()
    torch.cuda.set_device(0)
    return full


def zlib_cuda_memcpy(torch, dst, src_ptr, nbytes):
    """device-to-device copy out of the context's result buffer (a raw pointer) into a torch tensor"""
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rc = rt.cudaMemcpy(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src_ptr), ctypes.c_size_t(nbytes), 3)
    assert rc == 0, rc
    torch.cuda.synchronize()


@pytest.mark.parametrize("gather", [0, 1, 2, 3])
@pytest.mark.parametrize("mode,taps,C_,n,dtype", [(0, 127, 37, 60_000, "f64"), (1, 127, 3, 400_000, "f64"),
                                                   (1, 2049, 2, 300_000, "f32"), (0, 31, 9, 20_000, "f64")])
def test_fir_job_is_byte_identical_to_the_one_gpu_call(zlib, port, cuda, mgpu, gather, mode, taps, C_, n, dtype):
    torch = cuda
    if mode == 0 and C_ < mgpu.world:
        pytest.skip("fewer channels than ranks")
    f32 = dtype == "f32"
    np_dt, tdt = (np.float32, torch.float32) if f32 else (np.float64, torch.float64)
    x = np.stack([port.lcg_f64(n, 50 + c) for c in range(C_)]).astype(np_dt)
    torch.cuda.set_device(0)
    one = zlib.FirBank(C_, zlib.F32 if f32 else zlib.F64, kind=zlib.LPF, flt_len=taps, fc1=0.2, win=zlib.HAMMING)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty_like(dx)
    one.run(dx, n, dy, n, n)
    torch.cuda.synchronize()
    want = dy.cpu().numpy()
    one.close()
    job = zlib.MgpuJob.fir(mgpu, C_, zlib.F32 if f32 else zlib.F64, mode, flt_len=taps, fc1=0.2, win=zlib.HAMMING)
    got = run_job(zlib, torch, mgpu, job, n, x, np_dt, tdt, gather)
    job.close()
    assert got.tobytes() == want.tobytes(), (gather, mode, taps, float(np.abs(got - want).max()))


@pytest.mark.parametrize("gather", [0, 1, 2, 3])
@pytest.mark.parametrize("mode,L_,M,k,C_,frames", [(0, 1, 3, 0, 9, 12), (1, 320, 147, 128, 2, 9), (1, 160, 147, 0, 3, 7),
                                                    (0, 160, 147, 0, 5, 3)])
def test_resample_job_is_byte_identical_to_the_one_gpu_call_and_the_oracle(zlib, port, cuda, mgpu, gather, mode, L_, M, k,
                                                                             C_, frames):
    torch = cuda
    if mode == 0 and C_ < mgpu.world:
        pytest.skip("fewer channels than ranks")
    plan = port.resample_plan(L_, M, 1, k)
    n = plan.num_in * frames
    x = np.stack([port.lcg_s16(n, 900 + c) for c in range(C_)])
    job = zlib.MgpuJob.resample(mgpu, L_, M, C_, mode, k_override=k)
    n_out = job.out_len(n)
    got = run_job(zlib, torch, mgpu, job, n, x, np.int16, torch.int16, gather)
    job.close()
    for c in range(C_):
        want = port.resample_run(plan, 1.0, x[c], n_out)
        assert np.array_equal(got[c], want), (gather, mode, c)


def test_job_plans_tile_the_job(zlib, cuda, mgpu):
    job = zlib.MgpuJob.fir(mgpu, 16, zlib.F64, zlib.SHARD_TIME, flt_len=4095, fc1=0.11, win=zlib.KAISER)
    n = 3_000_000
    pos = 0
    for r in range(mgpu.world):
        sh = job.plan(n, r)
        assert sh.first_channel == 0 and sh.n_channels == 16
        assert sh.seg.in_start == pos and sh.seg.out_start == pos and sh.seg.halo == (4094 if r else 0)
        pos += sh.seg.in_count
    assert pos == n
    job.close()
