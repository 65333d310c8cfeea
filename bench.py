#!/usr/bin/env python
"""bench.py -- throughput of the llzlab FIR / resampling hot path on B200 (libllzfilter_cuda).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--dtype f64|f32]
    python bench.py --impl reference ...        # the reference's own CPU code, all host threads

One "step" = one pass of the hot path over one batch of synthetic input.  The default workload is
BASELINE.json configs[1] (C2): llz_fir 127-tap low-pass on 1024 independent channels x 10 s @ 48 kHz,
in the reference's sample type (double).  With N > 1 (torchrun, one rank per GPU) the channels are
independent units: every rank runs its own 1024-channel batch, no data-path collective (weak scaling).

The JSON line carries: value (device-timed, inputs resident in HBM), e2e (host buffers through the
C-ABI, H2D + D2H inside the timed region), roofline (dominant kernel, algorithmic bytes and flops
against measured / nominal peaks), cpu_baseline (the unmodified reference on one host core, bounded
sample), clocks and gpu_launches.  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "output Msamples/s (device-timed)"
UNIT = "Msamples/s"
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 37.2
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.4

# name -> description of the synthetic workload
WORKLOADS = {
    "c2": dict(kind="fir", desc="llz_fir 127-tap lowpass (fc 0.23, HAMMING), 1024 channels x 480000 samples (10 s @ 48 kHz)",
               channels=1024, n=480_000, taps=127, fc=0.23, win=0, seed=12345, shard="channel"),
    "c5": dict(kind="fir", desc="llz_fir 4095-tap lowpass (fc 0.11, KAISER), 16 channels x 57.6 M samples (5 min @ 192 kHz slice of the 1 h stream)",
               channels=16, n=57_600_000, taps=4095, fc=0.11, win=2, seed=12345, shard="time"),
    "c3": dict(kind="resample", desc="llz_resample 48 kHz -> 16 kHz (L=1, M=3, BLACKMAN, Q=134), 64 channels x 28.8 M samples (10 min)",
               channels=64, n=28_800_000, L=1, M=3, k=0, win=1, seed=777, shard="channel"),
    "c4": dict(kind="resample", desc="llz_resample 44.1 kHz -> 96 kHz (L=320, M=147, BLACKMAN, 256-tap bank: k=128, Q=257), 8 channels x 15.876 M samples (6 min slice of the 1 h stream)",
               channels=8, n=47_040 * 338, L=320, M=147, k=128, win=1, seed=777, shard="time"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- clocks --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons through NVML while the timed region runs"""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.samples, self.reasons, self.max_mhz, self.error = [], set(), None, None
        self.nv, self.h = None, None
        try:                                        # NVML is initialised before the timed region, not inside it
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                      # noqa: BLE001
            self.error = repr(e)

    def sample_now(self):
        """one sample from the calling thread (the main thread calls it while the timed launches are in flight)"""
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception as e:                      # noqa: BLE001
            self.error = repr(e)

    def run(self):
        while not self.stop_flag.is_set():
            self.sample_now()
            time.sleep(0.002)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.error:
            out["error"] = self.error
        return out


# ---- the reference on host cores ---------------------------------------------------------------------------------
def cpu_reference_rate(wl: dict, threads: int, channels_per_thread: int, n: int, repeats: int = 1):
    """Msamples/s of the reference's own C code (oracle/_ref when present, else the oracle port) on
    `threads` host threads, each filtering `channels_per_thread` channels of n samples frame by frame."""
    import oracle
    R, P = oracle.ref(), oracle.port()
    kind = "reference" if R is not None else "port"
    frame = 4096 if wl["kind"] == "fir" else None
    outs = [0] * threads

    def work(t):
        total = 0
        for c in range(channels_per_thread):
            seed = wl["seed"] + t * channels_per_thread + c
            if wl["kind"] == "fir":
                nn = n // frame * frame
                x = P.lcg_f64(nn, seed)
                if R is not None:
                    y = R.fir_stream(0, wl["taps"], wl["fc"], 0.0, wl["win"], x, frame)
                else:
                    y = P.fir_run(P.fir_design(0, wl["taps"], wl["fc"], 0.0, wl["win"]), x)
                total += len(y)
            else:
                plan = P.resample_plan(wl["L"], wl["M"], wl["win"], wl["k"])
                nn = n // plan.num_in * plan.num_in
                x = P.lcg_s16(nn, seed)
                if R is not None and wl["k"] == 0:
                    y = R.resample_stream(wl["L"], wl["M"], 1.0, wl["win"], x)
                else:                               # k_override is not expressible through the reference API
                    y = P.resample_run(plan, 1.0, x, nn * wl["L"] // wl["M"])
                total += len(y)
        outs[t] = total

    # inputs are generated inside the threads (LCG cost is ~1 % of the filtering cost at 127 taps)
    best = None
    for _ in range(repeats):
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sum(outs) / best / 1e6, kind, best


def reference_sample_shape(wl: dict):
    """(channels per thread, samples per channel) of one bounded reference step: ~0.5 s per thread"""
    if wl["kind"] == "fir":
        per_thread = 4.0e6 * 127 / wl["taps"]              # ~8 Msamples/s at 127 taps per core
        n = min(wl["n"], 480_000)
        return max(1, int(per_thread // n)), n
    q = 134 if wl["L"] == 1 else (2 * wl["k"] + 1 if wl["k"] else 45)
    per_thread_out = 2.5e6 * 134 / q
    n_in = int(per_thread_out * wl["M"] / wl["L"])
    return 1, max(50_000, n_in)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    cpt, n = reference_sample_shape(wl)
    for _ in range(args.warmup):
        cpu_reference_rate(wl, threads, 1, min(n, 65536))
    rates, kind, total_t = [], "port", 0.0
    for _ in range(args.steps):
        r, kind, dt = cpu_reference_rate(wl, threads, cpt, n)
        rates.append(r)
        total_t += dt
    value = statistics.mean(rates)
    sample = f"per step: {threads} threads x {cpt} channel(s) x {n} input samples of the workload, one reference handle per channel"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if wl["kind"] == "fir" else "s16 io / f64 acc", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "timing": "host wall clock (CPU implementation, no device)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(index: int):
    """Run this rank on the CPUs NVML reports as local to its GPU, so that the page-locked host buffers of the
    end-to-end leg are first-touched on the GPU's own NUMA node (with N ranks on a two-socket host, buffers on the
    wrong socket push every DMA across the inter-socket link)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:                               # noqa: BLE001  (affinity is an optimisation, never a requirement)
        pass


# ---- the CUDA arm ------------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import llzlab_b200 as z

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libllzfilter_cuda has no CPU path (use --impl reference for the CPU arm)")
    bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    z.lib()
    wl = WORKLOADS[args.workload]
    C_, n = wl["channels"], wl["n"]
    stream = torch.cuda.current_stream().cuda_stream
    hbm_peak, peak_src = peaks()

    # ---- this rank's share of the job --------------------------------------------------------------
    # channel-sharded workloads (C2, C3): every rank runs its own full batch of channels (weak scaling, no collective).
    # time-sharded workloads (C4, C5) at N > 1: the stated stream is cut into N segments; a rank owns
    # [seg.in_start, +in_count) and reads `halo` samples before it (strong scaling, no collective: the halo comes
    # with the rank's own slice of the input).
    time_sharded = wl["shard"] == "time" and world > 1
    scaling = "strong" if time_sharded else "weak"
    seg_first, halo = 0, 0
    if wl["kind"] == "fir":
        f32 = args.dtype == "f32"
        tdt, es = (torch.float32, 4) if f32 else (torch.float64, 8)
        bank = z.FirBank(C_, z.F32 if f32 else z.F64, kind=z.LPF, flt_len=wl["taps"], fc1=wl["fc"], win=wl["win"],
                         algo={"auto": z.FIR_AUTO, "direct": z.FIR_DIRECT, "fft": z.FIR_FFT}[args.algo])
        if time_sharded:
            # boundaries at multiples of the kernel's work-item length: the concatenated output is then byte-identical
            # to the one-GPU run for the overlap-save kernels as well (tests/test_gpu_fir.py)
            seg = z.shard_fir_segments_aligned(n, wl["taps"], bank.block_len, world, rank)
            seg_first, halo, n = seg.in_start - seg.halo, seg.halo, seg.in_count
        dx_all = torch.empty(C_, halo + n, dtype=tdt, device="cuda")
        z.synth_lcg_at(dx_all, halo + n, C_, seg_first, halo + n, 1 if f32 else 0, wl["seed"], stream)
        x_stride = halo + n
        dx = dx_all[:, halo:]
        dy = torch.empty(C_, n, dtype=tdt, device="cuda")
        n_out = n
        flop_per_out, bytes_per_out = 2.0 * wl["taps"], 2.0 * es
        dtype_name = "f32" if f32 else "f64"
        fma_peak_nominal = FP32_NOMINAL_TFLOPS if f32 else FP64_NOMINAL_TFLOPS
        fma_dtype = z.F32 if f32 else z.F64
        fir_fft = bank.algo == z.FIR_FFT
        if fir_fft and wl["taps"] >= 545:
            # 8192-point overlap-save kernel (llz_cuda_fir_fft8k.cu): one CTA of 256 threads turns 2*B outputs out of
            # 2580 FMA-pipe instructions per thread
            halo_pad = (wl["taps"] - 1 + 255) // 256 * 256
            fft_instr_per_out = 2580.0 * 256 / (2 * (8192 - halo_pad))
            fft_desc = "8192-point FFT per CTA"
            kernel = f"fir_fft8k_kernel<{'float' if f32 else 'double'}>"
        elif fir_fft:
            # overlap-save kernel (llz_cuda_fir_fft.cu): one warp turns 2*B outputs out of 1928 FMA-pipe instructions per lane
            halo_pad = (wl["taps"] - 1 + 31) // 32 * 32
            fft_instr_per_out = 1928.0 * 32 / (2 * (1024 - halo_pad))
            fft_desc = "1024-point FFT per warp"
            kernel = f"fir_fft_kernel<{'float' if f32 else 'double'}>"
        else:
            kernel = f"fir_tile_kernel<{'float' if f32 else 'double'}>"

        def step():
            bank.reset()
            if halo:
                bank.set_history(dx_all, x_stride, stream)
            bank.run(dx_all.data_ptr() + halo * es, x_stride, dy, n, n, stream)
        # fir_tile_kernel (or fir_fft_kernel interior + edge instantiations) + fir_history_kernel
        launches_per_step = 3 if fir_fft else 2
    else:
        acc = z.ACC_F32 if args.dtype == "f32" else z.ACC_F64
        bank = z.ResampleBank(z.KIND_RESAMPLE, wl["L"], wl["M"], C_, win=wl["win"], k_override=wl["k"], acc=acc)
        q = bank.info.taps_per_phase
        if time_sharded:
            seg = z.shard_resample_segments(n, wl["L"], wl["M"], q, bank.info.num_in, world, rank)
            seg_first, halo, n = seg.in_start - seg.halo, seg.halo, seg.in_count
        dx_all = torch.empty(C_, halo + n, dtype=torch.int16, device="cuda")
        z.synth_lcg_at(dx_all, halo + n, C_, seg_first, halo + n, 2, wl["seed"], stream)
        x_stride = halo + n
        dx = dx_all[:, halo:]
        n_out = bank.out_len(n)
        dy = torch.empty(C_, n_out, dtype=torch.int16, device="cuda")
        flop_per_out, bytes_per_out = 2.0 * q, 2.0 * (1.0 + wl["M"] / wl["L"])
        dtype_name = "s16 io / f32 acc" if args.dtype == "f32" else "s16 io / f64 acc"
        fma_peak_nominal = FP32_NOMINAL_TFLOPS if args.dtype == "f32" else FP64_NOMINAL_TFLOPS
        fma_dtype = z.F32 if args.dtype == "f32" else z.F64
        exact_tiles = "poly_bank_dmma_kernel" if os.environ.get("LLZ_BANK_NO_IMMA", "0") not in ("", "0") else "poly_bank_imma_kernel"
        kernel = "poly_slide_kernel" if wl["L"] == 1 else ("poly_bank_hmma_kernel" if args.dtype == "f32" else exact_tiles)

        def step():
            if halo:
                bank.set_history(dx_all, x_stride, stream)        # also rewinds the phase to output index 0
            else:
                bank.reset()
            bank.run(dx_all.data_ptr() + halo * 2, x_stride, n, dy, n_out, stream)
        launches_per_step = 2
        fir_fft = False
    outs_per_step = C_ * n_out

    fma_peak_measured = z.probe_fma(fma_dtype)

    # ---- device-timed region ----
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)                   # NVML init happens here, outside the bracket
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sampler.sample_now()                            # the launches above are still running: a sample under load
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.result()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    outs_all = outs_per_step
    if world > 1:
        t = torch.tensor([float(outs_per_step)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        outs_all = int(t.item())
    value = outs_all / (ms_step * 1e-3) / 1e6

    # per-rank kernel figures (rank 0's own time for the roofline of the kernel)
    ms_local = e0.elapsed_time(e1) / args.steps
    ach_gbs = outs_per_step * bytes_per_out / (ms_local * 1e-3) / 1e9
    ach_tf = outs_per_step * flop_per_out / (ms_local * 1e-3) / 1e12

    # ---- optional exchange step: every rank's output to every rank with NCCL over NVLink ----
    # The path itself needs no collective (independent channels / segments that carry their own halo); a job that
    # wants the whole result on one device adds this gather.  It is reported separately, never inside the timed step.
    gather = None
    if world > 1 and not args.no_gather:
        try:
            cnt = torch.tensor([dy.numel()], device="cuda", dtype=torch.int64)
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
            m = int(cnt.item())                                      # time segments may differ by one work item
            es_out = dy.element_size()
            if world * m * es_out > 48e9:
                gather = {"skipped": f"gathered result would be {world * m * es_out / 1e9:.1f} GB per rank"}
            else:
                flat = dy.reshape(-1)
                if flat.numel() < m:
                    flat = torch.cat([flat, flat.new_zeros(m - flat.numel())])
                full = torch.empty(world * m, dtype=dy.dtype, device="cuda")
                dist.all_gather_into_tensor(full, flat)              # warm-up: communicator channels, buffers
                torch.cuda.synchronize()
                dist.barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                dist.all_gather_into_tensor(full, flat)
                g1.record()
                torch.cuda.synchronize()
                t = torch.tensor([g0.elapsed_time(g1)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                gms = float(t.item())
                own_ok = bool(torch.equal(full[rank * m: rank * m + dy.numel()], dy.reshape(-1)))
                gather = {"collective": "NCCL all-gather of the planar outputs (torch.distributed.all_gather_into_tensor)",
                          "ms": gms, "bytes_per_rank": m * es_out, "gathered_bytes": world * m * es_out,
                          "algbw_gbs": world * m * es_out / (gms * 1e-3) / 1e9,
                          "busbw_gbs": (world - 1) * m * es_out / (gms * 1e-3) / 1e9,
                          "own_slice_intact": own_ok, "compute_ms_per_step": ms_step,
                          "note": "not part of the timed step; the filter itself needs no collective; shards land rank-major"}
                del full
        except Exception as ex:                     # noqa: BLE001
            gather = {"error": repr(ex)}

    # ---- end to end through the C-ABI with host buffers ----
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        in_bytes, out_bytes = dx.numel() * dx.element_size(), dy.numel() * dy.element_size()
        np_dt = {torch.float64: np.float64, torch.float32: np.float32, torch.int16: np.int16}[dx.dtype]
        hx = z.host_alloc(in_bytes, np_dt).reshape(C_, n)
        hy = z.host_alloc(out_bytes, np_dt).reshape(C_, n_out)
        torch.from_numpy(hx).copy_(dx)                # this rank's own samples (the halo stays in the bank's history)
        torch.cuda.synchronize()
        e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            if halo:
                bank.set_history(dx_all, x_stride, stream)
                torch.cuda.synchronize()
            else:
                bank.reset()
            if wl["kind"] == "fir":
                bank.run_host(hx, n, hy, n, n)
            else:
                bank.run_host(hx, n, n, hy, n_out)
        e2e_step()                                  # warm-up: staging buffers, streams
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()                              # synchronous: result is in hy on return
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the result read back is the full output; verify it is the device-resident result
        dev = dy[0, :4096].cpu().numpy()
        same = bool(np.array_equal(hy[0, :4096], dev))
        max_diff = float(np.abs(hy[0, :4096].astype(np.float64) - dev.astype(np.float64)).max())
        e2e = {"value": outs_all * e_steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes, "steps": e_steps, "matches_device_result": same,
               "max_abs_diff_vs_device": max_diff,
               "api": "llz_cuda_fir_bank_run_host" if wl["kind"] == "fir" else "llz_cuda_resample_bank_run_host",
               "host_memory": "page-locked (llz_cuda_host_alloc)"}
        z.host_free(hx.reshape(-1))
        z.host_free(hy.reshape(-1))
    except Exception as ex:                         # noqa: BLE001
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline: the unmodified reference, one thread, bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu:
        if wl["kind"] == "fir":
            n_cpu = 480_000 if wl["taps"] < 1000 else 40_960
            ch_cpu = max(1, int(1.0e8 * 127 / wl["taps"] / n_cpu))      # ~12 s at the probed 8.4 Msamples/s
        else:
            n_cpu, ch_cpu = reference_sample_shape(wl)[1] * 4, 6
        rate, kind, dt = cpu_reference_rate(wl, 1, ch_cpu, n_cpu)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
               "sample": f"{ch_cpu} channel(s) x {n_cpu} input samples of the workload, frame by frame through the reference API, 1 thread (the reference is single-threaded); host has {os.cpu_count()} logical cores"}

    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(f"{args.workload}_{args.dtype}")

    if fir_fft:
        # FMA-pipe instructions (DFMA/DADD/DMUL or FFMA/FADD/FMUL, one lane) the kernel executes, against the measured
        # FMA issue rate (probe TFLOP/s / 2 flop per FMA)
        ginstr = outs_per_step * fft_instr_per_out / (ms_local * 1e-3) / 1e12
        fma_pipe = {"achieved": ginstr, "unit": "T lane-instr/s", "instr_per_output": fft_instr_per_out,
                    "peak_measured": fma_peak_measured / 2 if fma_peak_measured else None,
                    "peak_nominal": fma_peak_nominal / 2,
                    "frac_of_measured": ginstr / (fma_peak_measured / 2) if fma_peak_measured else None,
                    "frac_of_nominal": ginstr / (fma_peak_nominal / 2),
                    "direct_form_equivalent_tflops": ach_tf,
                    "note": f"overlap-save ({fft_desc}, arithmetic in the bank's type): "
                            f"{fft_instr_per_out:.1f} FMA-pipe instructions per output instead of {wl['taps']} FMAs; "
                            "the FMA pipe and the shared-memory/LSU pipe are co-limiters below the HBM roof (DESIGN.md 4.1b)"}
    else:
        fma_pipe = {"achieved": ach_tf, "unit": "TFLOP/s", "peak_measured": fma_peak_measured,
                    "peak_nominal": fma_peak_nominal,
                    "frac_of_measured": ach_tf / fma_peak_measured if fma_peak_measured else None,
                    "frac_of_nominal": ach_tf / fma_peak_nominal,
                    "note": "direct-form FIR on CUDA cores: the FMA pipe, not HBM, is the binding roof "
                            f"(ceiling of the HBM fraction = {fma_peak_nominal * 1e12 / flop_per_out * bytes_per_out / 1e9 / hbm_peak:.3f})"}
        if kernel == "poly_bank_imma_kernel":
            fma_pipe["note"] = ("exact integer evaluation on the INT8 tensor cores (ten IMMA digit products per multiply-add, "
                                "1144 TOP/s measured = 114 TFLOP/s of exact multiply-adds): the fraction above is against the "
                                "FP64 FMA pipe this kernel no longer uses, i.e. the speed relative to the FP64 roof")
            fma_pipe["frac_of_imma_roof"] = ach_tf / 114.4

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "channels_per_gpu": C_, "samples_per_channel": n,
                   "outputs_per_step_per_gpu": outs_per_step,
                   "sharding": ("time segments, each rank reads a %d-sample halo before its segment, no collective"
                                % (wl["taps"] - 1 if wl["kind"] == "fir" else q - 1)) if time_sharded
                   else "independent channels per rank, no collective",
                   "l2": f"inputs {dx.numel() * dx.element_size() / 1e9:.2f} GB per GPU >> 126 MB L2, no flush needed",
                   "input": "integer LCG noise generated on the device (SURVEY.md 8d)",
                   **({"fir_algo": f"overlap-save, {fft_desc}" if fir_fft else "direct form"} if wl["kind"] == "fir" else {})},
        "roofline": {"kernel": kernel, "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic": {"bytes_per_output": bytes_per_out, "flop_per_output": flop_per_out,
                                     "outputs_per_launch": outs_per_step},
                     "fma_pipe": fma_pipe},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
    }
    if gather is not None:
        line["gather"] = gather
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c1(args):
    """BASELINE configs[0]: the example CLI's job (60 s mono 44.1 kHz s16 sweep -> 48 kHz) through the reference's
    own frame-by-frame entry points (llz_resample_filter_init / llz_resample) with host buffers.  It is 113 frames
    of 23,520 samples: a latency-bound drop-in path (one H2D, two launches and one D2H per frame), reported for
    completeness; the roofline discussion belongs to C2-C5."""
    import torch
    import llzlab_b200 as z
    import oracle
    torch.cuda.set_device(0)
    P, R = oracle.port(), oracle.ref()
    seconds, rate = 60, 44100
    t = np.arange(seconds * rate, dtype=np.float64) / rate
    pcm = np.round(0.5 * 32767 * np.sin(2 * np.pi * (20.0 * t + (20000.0 - 20.0) / (2 * seconds) * t * t))).astype(np.int16)
    r = z.Resampler(z.KIND_RESAMPLE, 160, 147, 1.0, z.BLACKMAN)
    num_in = r.bytes_in // 2
    frames = len(pcm) // num_in + 1                          # main.c:91-119
    x = np.zeros(frames * num_in, np.int16)
    x[:len(pcm)] = pcm
    times = []
    for it in range(args.warmup + args.steps):
        h = z.Resampler(z.KIND_RESAMPLE, 160, 147, 1.0, z.BLACKMAN)
        t0 = time.perf_counter()
        y = h.stream(x)
        dt = time.perf_counter() - t0
        h.close()
        if it >= args.warmup:
            times.append(dt)
    dt = statistics.median(times)
    t0 = time.perf_counter()
    ref_y = R.resample_stream(160, 147, 1.0, 1, x) if R is not None else P.resample_run(P.resample_plan(160, 147, 1), 1.0, x, len(y))
    ref_dt = time.perf_counter() - t0
    line = {"metric": METRIC, "value": len(y) / dt / 1e6, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None,
            "dtype": "s16 io / f64 acc", "data": "synthetic",
            "config": {"workload": "example/llz_resample job: 60 s mono 44.1 kHz s16 sine sweep -> 48 kHz (L=160, M=147, BLACKMAN, Q=45), "
                                   "113 frames through llz_resample with host buffers", "name": "c1", "timing": "host wall clock around the frame loop (synchronous API)"},
            "bit_identical_to_reference": bool(np.array_equal(y, ref_y)),
            "cpu_baseline": {"value": len(ref_y) / ref_dt / 1e6, "unit": UNIT, "cores": 1, "kind": "reference" if R is not None else "port",
                             "sample": "the whole job (2,892,800 outputs)"},
            "e2e": {"value": len(y) / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(y.nbytes)},
            "gpu_launches": 2 * frames * args.steps}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1"] + sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--algo", default="auto", choices=["auto", "direct", "fft"],
                    help="FIR kernel family (c2/c5): auto = overlap-save where it applies, else direct form")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gather", action="store_true", help="skip the NCCL all-gather of the outputs at N > 1")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    args = ap.parse_args()
    if args.workload == "c1":
        return run_c1(args) if int(os.environ.get("RANK", "0")) == 0 else 0
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
