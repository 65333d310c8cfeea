"""One process per GPU (torchrun): the multi-process model of the multi-GPU context (llz_cuda_mgpu_init_rank with a
shared NCCL id).  Every rank computes its shard of four jobs under the three gather modes and compares it -- and rank 0
the gathered result -- with the one-GPU call, byte for byte.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_check.py
"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llzlab_b200 as z  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [z.mgpu_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx = z.Mgpu(unique_id=ids[0], world=world, rank=rank)
rt = ctypes.CDLL("libcudart.so.12")
ok_all = True


def lcg(kind, C_, n):
    t = torch.empty(C_, n, dtype={0: torch.float64, 1: torch.float32, 2: torch.int16}[kind], device="cuda")
    z.synth_lcg(t, n, C_, n, kind, 4242)
    return t


def check(name, job, one_gpu_result, x, n):
    global ok_all
    C_, n_out = one_gpu_result.shape
    sh = job.plan(n, rank)
    lo, hi = sh.seg.in_start - sh.seg.halo, sh.seg.in_start + sh.seg.in_count
    xin = x[sh.first_channel:sh.first_channel + sh.n_channels, lo:hi].contiguous()
    want_shard = one_gpu_result[sh.first_channel:sh.first_channel + sh.n_channels,
                                sh.seg.out_start:sh.seg.out_start + sh.seg.out_count]
    es = x.element_size()
    for gather, gname in ((z.GATHER_NONE, "none"), (z.GATHER_NCCL, "nccl"), (z.GATHER_PEER, "peer"), (z.GATHER_COPY, "copy")):
        out = torch.zeros(sh.n_channels, sh.seg.out_count, dtype=x.dtype, device="cuda")
        if gather != z.GATHER_NONE:
            ctx.result_alloc(0, C_ * n_out * es)
        job.run(n, [xin], [hi - lo], [out], [sh.seg.out_count], n_out, gather, 4, [torch.cuda.current_stream().cuda_stream])
        torch.cuda.synchronize()
        good = True
        if gather == z.GATHER_NONE or (gather in (z.GATHER_NCCL, z.GATHER_COPY) and rank != 0):
            good = bool(torch.equal(out, want_shard))
        if gather != z.GATHER_NONE and rank == 0:
            full = torch.empty(C_, n_out, dtype=x.dtype, device="cuda")
            assert rt.cudaMemcpy(ctypes.c_void_p(full.data_ptr()), ctypes.c_void_p(ctx.result_ptr(0)),
                                 ctypes.c_size_t(C_ * n_out * es), 3) == 0
            torch.cuda.synchronize()
            good = good and bool(torch.equal(full, one_gpu_result))
        if gather != z.GATHER_NONE:
            dist.barrier()
            ctx.result_free()
        flags = [None] * world
        dist.all_gather_object(flags, good)
        if rank == 0:
            print(f"{name:28s} gather={gname:5s} world={world} bit_identical per rank: {flags}", flush=True)
        ok_all = ok_all and all(flags)


# FIR, channel shards (C2's filter) and time segments (C5's filter, 8192-point overlap-save)
for name, mode, taps, fc, win, C_, n, kind in (("fir 127 taps by channel", z.SHARD_CHANNEL, 127, 0.23, z.HAMMING, 64, 480_000, 0),
                                               ("fir 4095 taps by time", z.SHARD_TIME, 4095, 0.11, z.KAISER, 4, 4_000_000, 0),
                                               ("fir 127 taps f32 by time", z.SHARD_TIME, 127, 0.23, z.HAMMING, 8, 1_000_000, 1)):
    x = lcg(kind, C_, n)
    dtype = z.F32 if kind == 1 else z.F64
    one = z.FirBank(C_, dtype, kind=z.LPF, flt_len=taps, fc1=fc, win=win)
    y = torch.empty_like(x)
    one.run(x, n, y, n, n)
    torch.cuda.synchronize()
    one.close()
    job = z.MgpuJob.fir(ctx, C_, dtype, mode, flt_len=taps, fc1=fc, win=win)
    check(name, job, y, x, n)
    job.close()
# resampler, channel shards (C3's ratio) and time segments (C4's bank)
for name, mode, L_, M, k, C_, frames in (("resample 1/3 by channel", z.SHARD_CHANNEL, 1, 3, 0, 16, 200),
                                         ("resample 320/147 Q257 by time", z.SHARD_TIME, 320, 147, 128, 4, 24)):
    one = z.ResampleBank(z.KIND_RESAMPLE, L_, M, C_, k_override=k)
    n = one.info.num_in * frames
    x = lcg(2, C_, n)
    n_out = one.out_len(n)
    y = torch.zeros(C_, n_out, dtype=torch.int16, device="cuda")
    one.run(x, n, n, y, n_out)
    torch.cuda.synchronize()
    one.close()
    job = z.MgpuJob.resample(ctx, L_, M, C_, mode, k_override=k)
    check(name, job, y, x, n)
    job.close()
ctx.close()
if rank == 0:
    print("MGPU_CHECK_OK" if ok_all else "MGPU_CHECK_FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
