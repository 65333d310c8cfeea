"""Multi-GPU host logic on CPU: shard planners, and the N > 1 plan executed by two `gloo` ranks.

The data path has no collective (channels are independent handles; time segments carry their own halo), so
what needs testing is that the per-rank plans tile the job exactly and that per-rank results concatenate to
the one-process result byte for byte.  The per-rank arithmetic is done by the oracle here (tests may use it);
on the GPU the same plans drive libllzfilter_cuda (tests/test_gpu_*.py::test_time_segments_*).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_ch,world", [(1024, 8), (64, 8), (10, 4), (3, 8), (0, 2), (7, 1)])
def test_channel_shards_tile_exactly(zlib, n_ch, world):
    spans = [zlib.shard_channels(n_ch, world, r) for r in range(world)]
    assert spans[0][0] == 0
    for (f0, c0), (f1, _) in zip(spans, spans[1:]):
        assert f0 + c0 == f1
    assert spans[-1][0] + spans[-1][1] == n_ch
    counts = [c for _, c in spans]
    assert max(counts) - min(counts) <= 1


@pytest.mark.parametrize("n,N,world", [(691_200_000, 4095, 8), (100_003, 255, 4), (1000, 4095, 2), (5, 3, 8)])
def test_fir_segments(zlib, n, N, world):
    pos = 0
    for r in range(world):
        s = zlib.shard_fir_segments(n, N, world, r)
        assert s.in_start == pos == s.out_start and s.in_count == s.out_count
        assert s.halo == min(N - 1, s.in_start)
        pos += s.in_count
    assert pos == n


@pytest.mark.parametrize("n,N,g,world", [(691_200_000, 4095, 8192, 8), (691_200_000, 4095, 24576, 8), (230_400_000, 4095, 24576, 4),
                                         (100_003, 255, 1536, 4), (1000, 127, 1792, 2),
                                         (5, 3, 1, 8), (480_000, 127, 1792, 8)])
def test_fir_segments_aligned(zlib, n, N, g, world):
    """every internal boundary is a multiple of the kernel's work-item length; the segments tile [0, n)"""
    pos, counts = 0, []
    for r in range(world):
        s = zlib.shard_fir_segments_aligned(n, N, g, world, r)
        assert s.in_start == pos == s.out_start and s.in_count == s.out_count
        assert s.in_start % g == 0 or s.in_start == n
        assert s.halo == min(N - 1, s.in_start)
        pos += s.in_count
        counts.append(s.in_count)
    assert pos == n
    full = [c for c in counts if c and c % g == 0]
    assert not full or max(full) - min(full) <= g


@pytest.mark.parametrize("L,M,Q,frames,world", [(320, 147, 257, 3375, 8), (160, 147, 45, 113, 2), (1, 3, 134, 18750, 8),
                                               (3, 2, 77, 5, 8)])
def test_resample_segments(zlib, port, L, M, Q, frames, world):
    num_in = port.resample_plan(L, M, 1).num_in
    n_in = num_in * frames
    pos_in = pos_out = 0
    for r in range(world):
        s = zlib.shard_resample_segments(n_in, L, M, Q, num_in, world, r)
        assert s.in_start == pos_in and s.out_start == pos_out
        assert s.in_start % num_in == 0 and s.out_start % L == 0          # whole frames, phase 0
        assert s.in_start * L == s.out_start * M                            # exact input offset
        assert s.halo == min(Q - 1, s.in_start)
        pos_in += s.in_count
        pos_out += s.out_count
    assert pos_in == n_in and pos_out == n_in * L // M
    if (L, M, frames, world) == (320, 147, 3375, 8):                        # 7 x 422 + 1 x 421 frames
        counts = [zlib.shard_resample_segments(n_in, L, M, Q, num_in, world, r).in_count // num_in for r in range(world)]
        assert sorted(counts) == [421] + [422] * 7


def test_planner_rejects_bad_arguments(zlib):
    import ctypes as C
    seg = zlib.Segment()
    L = zlib.lib()
    assert L.llz_cuda_shard_fir_segments(100, 5, 0, 0, C.byref(seg)) == -1
    assert L.llz_cuda_shard_fir_segments(100, 5, 2, 2, C.byref(seg)) == -1
    assert L.llz_cuda_shard_resample_segments(1001, 3, 2, 9, 1024, 2, 0, C.byref(seg)) == -1   # not whole frames


# ---- two ranks over gloo ----------------------------------------------------------------------------------
def _rank_main(rank, world, port_file, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_file), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import llzlab_b200 as z
    import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = oracle.port()
    # (1) channel shards (C2/C3 style): each rank filters its own channels
    n_ch, n, N = 6, 5000, 127
    h = P.fir_design(0, N, 0.23, 0.0, 0)
    first, count = z.shard_channels(n_ch, world, rank)
    mine = np.stack([P.fir_run(h, P.lcg_f64(n, 12345 + c)) for c in range(first, first + count)])
    # (2) FIR time segments with halo (C5 style)
    x = P.lcg_f64(40_001, 7)
    s = z.shard_fir_segments(len(x), N, world, rank)
    hist = np.zeros(N - 1)
    if s.halo:
        hist[-s.halo:] = x[s.in_start - s.halo:s.in_start]
    seg_fir = P.fir_run(h, x[s.in_start:s.in_start + s.in_count], hist=hist)
    # (3) resampler time segments with halo (C4 style)
    L, M = 320, 147
    plan = P.resample_plan(L, M, 1, 16)
    xs = P.lcg_s16(plan.num_in * 5, 777)
    r = z.shard_resample_segments(len(xs), L, M, plan.cols, plan.num_in, world, rank)
    lo = r.in_start - r.halo
    seg_res = P.resample_run(plan, 1.0, xs[lo:r.in_start + r.in_count], r.out_count, m0=0) if r.halo == 0 else None
    if seg_res is None:
        # a segment sees its own input re-based at the halo start: shift the output index by the halo's worth of
        # input, which is exact because segments start at phase 0 and halo < one cycle's input for this check
        full = P.resample_run(plan, 1.0, xs[:r.in_start + r.in_count], r.out_start + r.out_count)
        # recompute from the shard only: inputs before the halo read as zeros
        shard_in = np.concatenate([np.zeros(r.in_start - r.halo, np.int16), xs[lo:r.in_start + r.in_count]])
        seg_res = P.resample_run(plan, 1.0, shard_in, r.out_count, m0=r.out_start)
        assert np.array_equal(seg_res, full[r.out_start:])
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), ch=mine, fir=seg_fir, res=seg_res)
    # timing protocol of bench.py: barrier, then max over ranks
    dist.barrier()
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_one_process_result(zlib, port, tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        free_port = sk.getsockname()[1]
    world = 2
    mp.spawn(_rank_main, args=(world, free_port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    P = port
    h = P.fir_design(0, 127, 0.23, 0.0, 0)
    want_ch = np.stack([P.fir_run(h, P.lcg_f64(5000, 12345 + c)) for c in range(6)])
    assert np.concatenate([p["ch"] for p in parts]).tobytes() == want_ch.tobytes()
    x = P.lcg_f64(40_001, 7)
    assert np.concatenate([p["fir"] for p in parts]).tobytes() == P.fir_run(h, x).tobytes()
    plan = P.resample_plan(320, 147, 1, 16)
    xs = P.lcg_s16(plan.num_in * 5, 777)
    want = P.resample_run(plan, 1.0, xs, len(xs) * 320 // 147)
    assert np.array_equal(np.concatenate([p["res"] for p in parts]), want)
