"""Interleaved PCM <-> planar channels.  The reference has no counterpart (it treats every
file as mono, quirk R7), so the checker is the numpy statement of the conversions documented in llz_cuda.h."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_frames(rng, fmt, n_frames, C):
    if fmt == 0:
        return rng.integers(-32768, 32768, (n_frames, C), dtype=np.int16)
    if fmt == 1:
        v = rng.integers(-(1 << 23), 1 << 23, (n_frames, C), dtype=np.int32)
        b = np.empty((n_frames, C, 3), np.uint8)
        b[..., 0], b[..., 1], b[..., 2] = v & 255, (v >> 8) & 255, (v >> 16) & 255
        return b, v
    return (rng.standard_normal((n_frames, C)) * 0.6).astype(np.float32)


def expect_planar(fmt, vals, ptype):
    """vals: integer / float sample values [n_frames][C] -> planar [C][n_frames] of the requested type"""
    v = vals.T
    if ptype == 0:
        if fmt == 0:
            return v.astype(np.int16)
        if fmt == 1:
            return (v >> 8).astype(np.int16)
        s = np.clip(v.astype(np.float32) * np.float32(32768.0), -32768.0, 32767.0)
        return np.trunc(s).astype(np.int16)
    dt = np.float32 if ptype == 1 else np.float64
    if fmt == 2:
        return v.astype(dt)
    return v.astype(dt) * dt(1.0 / 32768.0 if fmt == 0 else 1.0 / 8388608.0)


@pytest.mark.parametrize("C,n_frames", [(1, 1000), (2, 100_003), (6, 48_000), (64, 20_001), (3, 63), (128, 5000)])
@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("ptype", [0, 1, 2])
def test_deinterleave_matches_numpy(zlib, cuda, C, n_frames, fmt, ptype):
    torch = cuda
    rng = np.random.default_rng(C * 131 + fmt * 7 + ptype)
    made = make_frames(rng, fmt, n_frames, C)
    raw, vals = (made if fmt == 1 else (made, made))
    want = expect_planar(fmt, vals, ptype)
    d_frames = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint8).reshape(-1)).cuda()
    tdt = (torch.int16, torch.float32, torch.float64)[ptype]
    stride = n_frames + 5                                   # odd pitch: exercises the unaligned row path
    d_planar = torch.zeros(C, stride, dtype=tdt, device="cuda")
    zlib.pcm_deinterleave(d_frames, fmt, C, n_frames, d_planar, ptype, stride)
    torch.cuda.synchronize()
    got = d_planar.cpu().numpy()
    assert np.array_equal(got[:, :n_frames], want)
    assert not got[:, n_frames:].any()


@pytest.mark.parametrize("C,n_frames", [(2, 100_003), (8, 65_536), (5, 777)])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_round_trip_and_saturation(zlib, cuda, C, n_frames, fmt):
    torch = cuda
    rng = np.random.default_rng(5 + C + fmt)
    # planar f64 in [-1.2, 1.2): interleave saturates, deinterleave brings the clipped values back exactly
    x = rng.uniform(-1.2, 1.2, (C, n_frames))
    dx = torch.from_numpy(x).cuda()
    bps = (2, 3, 4)[fmt]
    d_frames = torch.zeros(n_frames * C * bps, dtype=torch.uint8, device="cuda")
    zlib.pcm_interleave(dx, zlib.PLANAR_F64, n_frames, C, n_frames, d_frames, fmt)
    back = torch.zeros_like(dx)
    zlib.pcm_deinterleave(d_frames, fmt, C, n_frames, back, zlib.PLANAR_F64, n_frames)
    torch.cuda.synchronize()
    if fmt == 2:
        want = x.astype(np.float32).astype(np.float64)
    else:
        full = 32768.0 if fmt == 0 else 8388608.0
        want = np.trunc(np.clip(x * full, -full, full - 1)) / full
    assert np.array_equal(back.cpu().numpy(), want)
    # the interleaved bytes themselves: frame f, channel c at (f*C + c)*bps
    fr = d_frames.cpu().numpy()
    if fmt == 0:
        assert np.array_equal(fr.view(np.int16).reshape(n_frames, C), (want.T * 32768.0).astype(np.int16))


def test_deinterleave_feeds_a_bank(zlib, port, cuda):
    """stereo s16 frames -> planar -> 2-channel resampler bank == the oracle per channel (what R7 gets wrong)"""
    torch = cuda
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, 160, 147, 2)
    n = bank.info.num_in * 2
    left, right = port.lcg_s16(n, 1), port.lcg_s16(n, 2)
    frames = np.stack([left, right], axis=1)
    d_frames = torch.from_numpy(frames.reshape(-1)).cuda()
    d_planar = torch.zeros(2, n, dtype=torch.int16, device="cuda")
    zlib.pcm_deinterleave(d_frames, zlib.PCM_S16, 2, n, d_planar, zlib.PLANAR_S16, n)
    n_out = bank.out_len(n)
    dy = torch.zeros(2, n_out, dtype=torch.int16, device="cuda")
    bank.run(d_planar, n, n, dy, n_out)
    d_out = torch.zeros(n_out * 2, dtype=torch.int16, device="cuda")
    zlib.pcm_interleave(dy, zlib.PLANAR_S16, n_out, 2, n_out, d_out, zlib.PCM_S16)
    torch.cuda.synchronize()
    p = port.resample_plan(160, 147, 1)
    want = np.stack([port.resample_run(p, 1.0, left, n_out), port.resample_run(p, 1.0, right, n_out)], axis=1)
    assert np.array_equal(d_out.cpu().numpy().reshape(n_out, 2), want)
    bank.close()


def test_bad_arguments(zlib, cuda):
    L = zlib.lib()
    assert L.llz_cuda_pcm_deinterleave(None, 0, 2, 10, None, 0, 10, None) == -1
    t = cuda.zeros(64, dtype=cuda.int16, device="cuda")
    assert L.llz_cuda_pcm_deinterleave(t.data_ptr(), 7, 2, 10, t.data_ptr(), 0, 10, None) == -1
    assert L.llz_cuda_pcm_deinterleave(t.data_ptr(), 0, 0, 10, t.data_ptr(), 0, 10, None) == -1


@pytest.mark.parametrize("fmt", ["s16", "s24", "f32"])
@pytest.mark.parametrize("kind,L_,M,C_", [("resample", 160, 147, 2), ("resample", 1, 3, 6), ("interp", 4, 1, 2)])
def test_resampler_reads_interleaved_frames_directly(zlib, port, cuda, fmt, kind, L_, M, C_):
    """llz_cuda_resample_bank_run_pcm: the de-interleave (and the s24 / f32 conversion) is fused into the load stage of the
    tcgen05 kernel.  The result must be what de-interleaving first and running the planar entry point gives -- which for
    s16 is the reference's output per channel, bit for bit -- across two calls (the history comes from interleaved input)."""
    torch = cuda
    if kind == "interp":
        bank = zlib.ResampleBank(zlib.KIND_INTERP, L_, 1, C_)
        ref = zlib.ResampleBank(zlib.KIND_INTERP, L_, 1, C_)
        plan = port.interp_plan(L_, 1)
    else:
        bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L_, M, C_)
        ref = zlib.ResampleBank(zlib.KIND_RESAMPLE, L_, M, C_)
        plan = port.resample_plan(L_, M, 1)
    bank.set_tiles(zlib.TILES_INT8_TCGEN05)
    n = plan.num_in * 6
    rng = np.random.default_rng(L_ * 100 + M + C_)
    if fmt == "s16":
        frames = rng.integers(-32768, 32768, size=(n, C_), dtype=np.int16)
        planar = np.ascontiguousarray(frames.T)
        raw, code = frames, zlib.PCM_S16
    elif fmt == "s24":
        v = rng.integers(-(1 << 23), 1 << 23, size=(n, C_), dtype=np.int32)
        raw = np.zeros((n, C_, 3), np.uint8)
        raw[..., 0], raw[..., 1], raw[..., 2] = v & 255, (v >> 8) & 255, (v >> 16) & 255
        planar = np.ascontiguousarray((v >> 8).astype(np.int16).T)
        code = zlib.PCM_S24
    else:
        f = (rng.random((n, C_), dtype=np.float32) * 2.2 - 1.1).astype(np.float32)        # some samples clip
        raw, code = f, zlib.PCM_F32
        planar = np.ascontiguousarray(np.trunc(np.clip(f * np.float32(32768.0), -32768.0, 32767.0)).astype(np.int16).T)
    d_raw = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy()).cuda()
    bps = {"s16": 2, "s24": 3, "f32": 4}[fmt]
    n_out = ref.out_len(n)
    want = torch.zeros(C_, n_out, dtype=torch.int16, device="cuda")
    ref.run(torch.from_numpy(planar).cuda(), n, n, want, n_out)
    got = torch.zeros(C_, n_out, dtype=torch.int16, device="cuda")
    cut = plan.num_in * 2
    o1 = bank.run_pcm(d_raw, code, cut, got, n_out)
    assert bank.last_run()[1].startswith("poly_bank_umma_kernel")
    o2 = bank.run_pcm(d_raw.data_ptr() + cut * C_ * bps, code, n - cut, got.data_ptr() + 2 * o1, n_out)
    torch.cuda.synchronize()
    assert o1 + o2 == n_out
    assert torch.equal(got, want)
    if fmt == "s16" and kind == "resample":
        for c in range(C_):
            assert np.array_equal(got[c].cpu().numpy(), port.resample_run(plan, 1.0, planar[c], n_out))
    bank.close(); ref.close()


def test_run_pcm_host_stereo_file_sized_job(zlib, port, cuda):
    """host frames in, host frames out (what the CLI's --channels mode calls): every channel of the interleaved result is
    the reference's output for that channel"""
    L_, M, C_ = 160, 147, 2
    bank = zlib.ResampleBank(zlib.KIND_RESAMPLE, L_, M, C_)
    plan = port.resample_plan(L_, M, 1)
    n = plan.num_in * 40
    frames = np.stack([port.lcg_s16(n, 900 + c) for c in range(C_)], axis=1).copy()
    n_out = bank.out_len(n)
    out = np.zeros((n_out, C_), np.int16)
    assert bank.run_pcm_host(frames, zlib.PCM_S16, n, out, zlib.PCM_S16, n_out) == n_out
    for c in range(C_):
        assert np.array_equal(out[:, c], port.resample_run(plan, 1.0, np.ascontiguousarray(frames[:, c]), n_out))
    bank.close()
