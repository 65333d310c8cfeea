"""One small launch of every kernel family of libllzfilter_cuda: interior AND edge tiles / items, history splice,
widened guard band.  Written for compute-sanitizer (tools/run_sanitizers.sh); the pool this repo is measured on refuses
compute-sanitizer (profiles/r02_sanitizer_closed.txt), so the script carries its own checks: every output buffer sits
between canary margins that must come back untouched, and every result is compared with the oracle (bit-exact or
within tolerance).  Exit code 0 = all of that held."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llzlab_b200 as z  # noqa: E402
import oracle  # noqa: E402

P = oracle.port()
only = set(sys.argv[1:])
MARGIN, CANARY = 512, 7.25


def want(name):
    return not only or name in only


def fir_case(name, taps, dtype, algo, n, fft_size=0, C_=2):
    if not want(name):
        return
    rng = np.random.default_rng(taps)
    h = rng.standard_normal(taps) / taps ** 0.5
    x = rng.uniform(-1, 1, (C_, n))
    ref = np.stack([P.fir_run(h, x[c]) for c in range(C_)])
    f32 = dtype == z.F32
    dx = torch.from_numpy(x.astype(np.float32 if f32 else np.float64)).cuda()
    ys = n + 2 * MARGIN                                   # every output row sits between two canary margins
    dy = torch.full((C_, ys), CANARY, dtype=dx.dtype, device="cuda")
    bank = z.FirBank(C_, dtype, taps=h, algo=algo)
    if fft_size:
        bank.set_fft_size(fft_size)
    half = n // 2 + 3
    es = dx.element_size()
    bank.run(dx, n, dy.data_ptr() + MARGIN * es, ys, half)   # second call: history splice on the edge items
    bank.run(dx.data_ptr() + half * es, n, dy.data_ptr() + (MARGIN + half) * es, ys, n - half)
    torch.cuda.synchronize()
    full = dy.cpu().numpy()
    assert (full[:, :MARGIN] == CANARY).all() and (full[:, MARGIN + n:] == CANARY).all(), (name, "canary overwritten")
    got = full[:, MARGIN:MARGIN + n].astype(np.float64)
    scale = max(np.abs(h).sum(), 1.0)
    if dtype == z.F64_STRICT:
        assert got.tobytes() == ref.tobytes(), name
    else:
        assert np.abs(got - ref).max() <= (2e-5 if f32 else 1e-12) * scale, (name, np.abs(got - ref).max())
    bank.close()
    print("ok", name, flush=True)


def poly_case(name, kind, L_, M, k, acc, tiles, frames, C_=2, scale=1.0):
    if not want(name):
        return
    if kind == z.KIND_DECIMATE:
        bank, plan = z.ResampleBank(kind, 1, M, C_, acc=acc), P.decimate_plan(M, 1)
    elif kind == z.KIND_INTERP:
        bank, plan = z.ResampleBank(kind, L_, 1, C_, acc=acc), P.interp_plan(L_, 1)
    else:
        bank, plan = z.ResampleBank(kind, L_, M, C_, k_override=k, acc=acc), P.resample_plan(L_, M, 1, k)
    if tiles:
        bank.set_tiles(tiles)
    if scale > 1.0:
        bank.set_guard_scale(scale)
    n_in = plan.num_in * frames
    x = np.stack([P.lcg_s16(n_in, 600 + c) for c in range(C_)])
    n_out = bank.out_len(n_in)
    dx = torch.from_numpy(x).cuda()
    ys = n_out + 2 * MARGIN
    dy = torch.full((C_, ys), 12345, dtype=torch.int16, device="cuda")
    cut = plan.num_in * (frames // 2)
    o1 = bank.run(dx, n_in, cut, dy.data_ptr() + 2 * MARGIN, ys)
    bank.run(dx.data_ptr() + 2 * cut, n_in, n_in - cut, dy.data_ptr() + 2 * (MARGIN + o1), ys)
    torch.cuda.synchronize()
    full = dy.cpu().numpy()
    assert (full[:, :MARGIN] == 12345).all() and (full[:, MARGIN + n_out:] == 12345).all(), (name, "canary overwritten")
    got = full[:, MARGIN:MARGIN + n_out]
    for c in range(C_):
        if kind == z.KIND_DECIMATE:
            ref = P.decimate_run(plan, 1.0, x[c], n_out)
        elif kind == z.KIND_INTERP:
            ref = P.interp_run(plan, 1.0, x[c])
        else:
            ref = P.resample_run(plan, 1.0, x[c], n_out)
        d = np.abs(got[c].astype(np.int32) - ref.astype(np.int32)).max()
        assert d <= (1 if acc == z.ACC_F32 else 0), (name, c, int(d))
    if scale > 1.0:
        assert bank.guard_count() > 0, name
    bank.close()
    print("ok", name, flush=True)


torch.cuda.set_device(0)
# FIR: direct (strict / fma / f32 blocked), overlap-save 1024 (staged f64, gather f32), 8192, 16384 (two rounds, L2 scratch)
fir_case("fir_strict", 127, z.F64_STRICT, z.FIR_DIRECT, 9000)
fir_case("fir_direct_f64", 127, z.F64, z.FIR_DIRECT, 9000)
fir_case("fir_direct_f32", 513, z.F32, z.FIR_DIRECT, 9000)
fir_case("fir_fft1k_f64", 127, z.F64, z.FIR_FFT, 40_000)
fir_case("fir_fft1k_f32", 127, z.F32, z.FIR_FFT, 40_000)
fir_case("fir_fft8k_f64", 2049, z.F64, z.FIR_FFT, 90_000, 8192)
fir_case("fir_fft8k_f32", 2049, z.F32, z.FIR_FFT, 90_000, 8192)
fir_case("fir_fft16k_f64", 4095, z.F64, z.FIR_FFT, 140_000, 16384, C_=1)
fir_case("fir_fft16k_f32", 4095, z.F32, z.FIR_FFT, 140_000, 16384, C_=1)
# resampler: INT8 tensor tiles (producer / consumer warpgroups), DMMA, DFMA tile, HMMA fast mode, FFMA tile,
# sliding (exact + fast), general (interp), all with a widened guard band so the recompute branch runs too
poly_case("bank_imma", z.KIND_RESAMPLE, 160, 147, 0, z.ACC_F64, z.TILES_INT8, 2, scale=3e6)
poly_case("bank_imma_q257", z.KIND_RESAMPLE, 320, 147, 128, z.ACC_F64, z.TILES_INT8, 2, C_=1, scale=3e6)
poly_case("bank_dmma", z.KIND_RESAMPLE, 160, 147, 0, z.ACC_F64, z.TILES_FP64_TENSOR, 2, scale=3e6)
poly_case("bank_dfma", z.KIND_RESAMPLE, 160, 147, 0, z.ACC_F64, z.TILES_CUDA_CORE, 2, scale=3e6)
poly_case("bank_hmma", z.KIND_RESAMPLE, 160, 147, 0, z.ACC_F32, z.TILES_AUTO, 2)
poly_case("bank_ffma", z.KIND_RESAMPLE, 160, 147, 0, z.ACC_F32, z.TILES_CUDA_CORE, 2)
poly_case("slide_f64", z.KIND_RESAMPLE, 1, 3, 0, z.ACC_F64, 0, 8, scale=3e6)
poly_case("slide_f32", z.KIND_RESAMPLE, 1, 3, 0, z.ACC_F32, 0, 8)
poly_case("slide_decimate", z.KIND_DECIMATE, 1, 4, 0, z.ACC_F64, 0, 8, scale=3e6)
poly_case("general_interp", z.KIND_INTERP, 3, 1, 0, z.ACC_F64, 0, 4, scale=3e6)
poly_case("general_strict", z.KIND_RESAMPLE, 160, 147, 0, z.ACC_F64_STRICT, 0, 2, C_=1)
if want("pcm"):
    fr = torch.randint(-32768, 32767, (5000, 6), dtype=torch.int16, device="cuda")
    pl = torch.zeros(6, 5000, dtype=torch.int16, device="cuda")
    z.pcm_deinterleave(fr, z.PCM_S16, 6, 5000, pl, z.PLANAR_S16, 5000)
    back = torch.zeros_like(fr)
    z.pcm_interleave(pl, z.PLANAR_S16, 5000, 6, 5000, back, z.PCM_S16)
    torch.cuda.synchronize()
    assert torch.equal(fr, back) and torch.equal(pl, fr.t().contiguous())
    print("ok pcm", flush=True)
print("SANITIZE_KERNELS_OK")
