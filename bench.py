#!/usr/bin/env python
"""bench.py -- throughput of the llzlab FIR / resampling hot path on B200 (libllzfilter_cuda).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--dtype f64|f32]
    python bench.py --impl reference ...        # the reference's own CPU code, all host threads

One "step" = one pass of the hot path over one batch of synthetic input.  The headline workload is
BASELINE.json configs[1] (C2): llz_fir 127-tap low-pass on 1024 independent channels x 10 s @ 48 kHz,
in the reference's sample type (double); the default run then measures every other config (C3, C4 at
its full hour, C5, and the f32 / fast variants) the same way and reports them under "workloads", each
with its own roofline, clocks and an oracle parity record.  With N > 1 (torchrun, one rank per GPU)
the BASELINE totals are SPLIT across the ranks through the library's multi-GPU job API
(llz_cuda_mgpu_*): C2 / C3 by channel, C4 / C5 into time segments with halo (strong scaling, no
data-path collective); the step is also timed with the result gathered on rank 0 (chunked NCCL
send/recv overlapped with compute, and kernels storing straight into rank 0's buffer over NVLink),
and every rank's shard is compared byte for byte with the one-GPU call.

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
    # C5 as doubles is 88.5 GB in + 88.5 GB out for the full hour: more than one 180 GB GPU holds beside the context, so the
    # f64 bench line filters 16 channels x 20 min (29.5 + 29.5 GB; the same total at every N, strong scaling); the f32
    # variant (--dtype f32) runs the same 20 min
    "c5": dict(kind="fir", desc="llz_fir 4095-tap lowpass (fc 0.11, KAISER), 16 channels x 230.4 M samples (20 min @ 192 kHz of the 1 h stream; the full hour as f64 is 177 GB)",
               channels=16, n=230_400_000, taps=4095, fc=0.11, win=2, seed=12345, shard="time"),
    "c3": dict(kind="resample", desc="llz_resample 48 kHz -> 16 kHz (L=1, M=3, BLACKMAN, Q=134), 64 channels x 28.8 M samples (10 min)",
               channels=64, n=28_800_000, L=1, M=3, k=0, win=1, seed=777, shard="channel"),
    # not a BASELINE config: the llz_interp sibling on the same kernel, for its throughput record
    "interp4": dict(kind="resample", interp=True,
                    desc="llz_interp x4 (12 kHz -> 48 kHz, L=4, BLACKMAN), 64 channels x 7,200,768 samples (10 min, 7032 frames of 1024)",
                    channels=64, n=1024 * 7032, L=4, M=1, k=0, win=1, seed=777, shard="channel"),
    "c4": dict(kind="resample", desc="llz_resample 44.1 kHz -> 96 kHz (L=320, M=147, BLACKMAN, 256-tap bank: k=128, Q=257), 8 channels x 158.76 M samples (the full 1 h stream, 3375 frames of 47040)",
               channels=8, n=47_040 * 3375, L=320, M=147, k=128, win=1, seed=777, shard="time"),
}


def umma_macs_per_output(L: int, M: int, Q: int, planes: int) -> float:
    """int8 multiply-accumulates the tcgen05 phase-bank kernel executes per output sample: the tile geometry of
    llzlab_b200/csrc/llz_umma_tables.h (replication, 64-phase tiles, K steps of 32 bytes, 2 x planes digit products)"""
    r = 1
    while (r * M) % 16:
        r *= 2
    while r * L < 64:
        r *= 2
    while (r * L) % 64 and r * L < 16384:
        r *= 2
    UL, UM = L * r, M * r
    ksteps = 0
    for p in range((UL + 63) // 64):
        l0 = 64 * p
        pbv = min(64, UL - l0)
        c_lo, c_hi = (l0 * UM) // UL, ((l0 + pbv - 1) * UM) // UL
        ksteps += (Q + (c_hi - c_lo) + (c_lo & 15) + 31) // 32
    return ksteps * 2 * planes * 64 * 32 / UL


def peaks_tensor():
    """dense bf16 TFLOP/s measured on this pool (MEASURED_PEAKS.json); kind::i8 runs at twice the bf16 MAC rate"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["bf16_tflops"]), "measured bf16 x 2 (MEASURED_PEAKS.json; tcgen05 kind::i8 = twice the bf16 rate)"
    return 1590.0, "fallback bf16 x 2 (B200_PROFILING.md)"


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
    wl_name = args.workload or "c2"
    wl = WORKLOADS[wl_name]
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
            "config": {"workload": wl["desc"], "name": wl_name, "timing": "host wall clock (CPU implementation, no device)"},
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
class Dist:
    """torch.distributed plumbing of one rank (one process per GPU under torchrun)"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))

    def init(self):
        torch, dist = self.torch, self.dist
        if self.world > 1:
            # NCCL prints its version banner on stdout when the first communicator comes up; keep stdout for the JSON line
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_obj(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def checksum64(t):
    """position-weighted 64-bit checksum of a tensor's bytes, computed on the device (wraps mod 2^64)"""
    import torch
    v = t.contiguous().view(torch.uint8).reshape(-1)
    pad = (-v.numel()) % 8
    if pad:
        v = torch.cat([v, v.new_zeros(pad)])
    w = v.view(torch.int64)
    idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64) * 2 + 1
    return int((w * idx).sum().item())


class Workload:
    """One BASELINE config on this rank: its share of the job (the whole job at N = 1; at N > 1 a channel shard or a
    time segment with halo through the library's multi-GPU job API, include/llz_cuda.h "Multi-GPU"), the timed step,
    the oracle window check and the byte-identity check against the one-GPU call."""

    def __init__(self, z, D: Dist, name: str, dtype: str, algo: str, mg):
        import torch
        self.z, self.D, self.torch, self.name, self.dtype = z, D, torch, name, dtype
        wl = self.wl = WORKLOADS[name]
        self.fir = wl["kind"] == "fir"
        world, rank = D.world, D.rank
        self.stream = torch.cuda.current_stream().cuda_stream
        C_, n = wl["channels"], wl["n"]
        self.C_total, self.n_total = C_, n
        f32 = dtype == "f32"
        self.mode = z.SHARD_TIME if wl["shard"] == "time" else z.SHARD_CHANNEL
        self.job = self.bank = None
        if self.fir:
            self.tdt, self.es, self.np_dt = (torch.float32, 4, np.float32) if f32 else (torch.float64, 8, np.float64)
            self.lcg_kind = 1 if f32 else 0
            bank_dtype = z.F32 if f32 else z.F64
            fir_algo = {"auto": z.FIR_AUTO, "direct": z.FIR_DIRECT, "fft": z.FIR_FFT}[algo]
            if world > 1:
                self.job = z.MgpuJob.fir(mg, C_, bank_dtype, self.mode, kind=z.LPF, flt_len=wl["taps"], fc1=wl["fc"], win=wl["win"])
                self.bank_handle = self.job.bank(0)
                if fir_algo != z.FIR_AUTO:
                    z._check(z.lib().llz_cuda_fir_bank_set_algo(self.bank_handle, fir_algo), "set_algo")
                self.fir_fft = z.lib().llz_cuda_fir_bank_get_algo(self.bank_handle) == z.FIR_FFT
            else:
                self.bank = z.FirBank(C_, bank_dtype, kind=z.LPF, flt_len=wl["taps"], fc1=wl["fc"], win=wl["win"], algo=fir_algo)
                self.fir_fft = self.bank.algo == z.FIR_FFT
            self.flop_per_out, self.bytes_per_out = 2.0 * wl["taps"], 2.0 * self.es
            self.dtype_name = "f32" if f32 else "f64"
            self.halo_full = wl["taps"] - 1
        else:
            self.tdt, self.es, self.np_dt, self.lcg_kind = torch.int16, 2, np.int16, 2
            acc = z.ACC_F32 if f32 else z.ACC_F64
            if wl.get("interp"):
                if world > 1:
                    raise RuntimeError("the interp workload is measured on one GPU only")
                self.bank = z.ResampleBank(z.KIND_INTERP, wl["L"], 1, C_, win=wl["win"], acc=acc)
                info = self.bank.info
            elif world > 1:
                self.job = z.MgpuJob.resample(mg, wl["L"], wl["M"], C_, self.mode, win=wl["win"], k_override=wl["k"], acc=acc)
                info = z.bank_info(self.job.bank(0))
            else:
                self.bank = z.ResampleBank(z.KIND_RESAMPLE, wl["L"], wl["M"], C_, win=wl["win"], k_override=wl["k"], acc=acc)
                info = self.bank.info
            self.q = info.taps_per_phase
            self.flop_per_out, self.bytes_per_out = 2.0 * self.q, 2.0 * (1.0 + wl["M"] / wl["L"])
            self.dtype_name = "s16 io / f32 acc" if f32 else "s16 io / f64 acc"
            self.fir_fft = False
            self.halo_full = self.q - 1
        # ---- this rank's shard ----
        if world > 1:
            sh = self.job.plan(n, rank)
            self.c0, self.cc = sh.first_channel, sh.n_channels
            self.in_start, self.in_count, self.halo = sh.seg.in_start, sh.seg.in_count, sh.seg.halo
            self.out_start, self.out_count = sh.seg.out_start, sh.seg.out_count
            self.total_out = self.job.out_len(n)
        else:
            self.c0, self.cc, self.in_start, self.in_count, self.halo, self.out_start = 0, C_, 0, n, 0, 0
            self.out_count = self.total_out = n if self.fir else self.bank.out_len(n)
        self.x_stride = self.halo + self.in_count
        self.dx_all = torch.empty(self.cc, self.x_stride, dtype=self.tdt, device="cuda")
        z.synth_lcg_at(self.dx_all, self.x_stride, self.cc, self.in_start - self.halo, self.x_stride, self.lcg_kind,
                       wl["seed"] + self.c0, self.stream)
        self.dy = torch.empty(self.cc, self.out_count, dtype=self.tdt, device="cuda")
        self.outs_rank = self.cc * self.out_count
        self.kernel, self.fft_instr_per_out, self.fft_desc = self._kernel_name()
        # launches per step: filter kernel(s) + history kernel (+ halo copy for a time segment)
        self.launches_per_step = (3 if self.fir_fft else 2) + (1 if self.halo else 0)

    def _kernel_name(self):
        wl, f32 = self.wl, self.dtype == "f32"
        if not self.fir:
            # large calls run the tcgen05 kernel (llz_cuda_polybank_umma.cu): 3 digit planes in the fast mode, 5 in the exact
            # mode; refresh_kernel() replaces this guess with what the library reports after the timed steps
            return ("poly_bank_umma_kernel<3>" if f32 else "poly_bank_umma_kernel<5>"), None, None
        t = "float" if f32 else "double"
        if self.fir_fft:
            # which overlap-save kernel: the library's work-item length says it (2 * (transform length - padded halo))
            handle = self.bank.handle if self.bank is not None else self.bank_handle
            blk = self.z._check(self.z.lib().llz_cuda_fir_bank_block_len(handle), "llz_cuda_fir_bank_block_len")
            halo16 = (wl["taps"] - 1 + 511) // 512 * 512
            halo8 = (wl["taps"] - 1 + 255) // 256 * 256
            if blk == 2 * (16384 - halo16):
                # 16384-point kernel (llz_cuda_fir_fft16k.cu): one CTA of 256 threads turns 2*B outputs out of 5512 (f32) /
                # 5736 (f64: outer twiddles of the last pass computed instead of read) FMA-pipe instructions per thread
                per_thread = 5512.0 if f32 else 5736.0
                return f"fir_fft16k_kernel<{t}>", per_thread * 256 / blk, "16384-point FFT per CTA in two rounds, half of the item in an L2-resident scratch"
            if blk == 2 * (8192 - halo8):
                # 8192-point overlap-save kernel (llz_cuda_fir_fft8k.cu): one CTA of 256 threads turns 2*B outputs out of
                # 2580 FMA-pipe instructions per thread
                return f"fir_fft8k_kernel<{t}>", 2580.0 * 256 / blk, "8192-point FFT per CTA"
        if self.fir_fft:
            # overlap-save kernel (llz_cuda_fir_fft.cu): one warp turns 2*B outputs out of 1928 FMA-pipe instructions per lane
            halo_pad = (wl["taps"] - 1 + 31) // 32 * 32
            return f"fir_fft_kernel<{t}>", 1928.0 * 32 / (2 * (1024 - halo_pad)), "1024-point FFT per warp"
        return f"fir_tile_kernel<{t}>", None, None

    # ---- the timed step -------------------------------------------------------------------------------
    def step(self, gather: int = 0, chunks: int = 0):
        z = self.z
        if chunks == 0:                                 # the copy-engine push exposes 1 / chunks of the transfer: use the most
            chunks = 8 if gather == z.GATHER_COPY else 4
        if self.job is not None:
            self.job.run(self.n_total, [self.dx_all], [self.x_stride], [self.dy], [self.out_count], self.total_out,
                         gather, chunks, [self.stream])
        elif self.fir:
            self.bank.reset()
            self.bank.run(self.dx_all, self.x_stride, self.dy, self.out_count, self.in_count, self.stream)
        else:
            self.bank.reset()
            self.bank.run(self.dx_all, self.x_stride, self.in_count, self.dy, self.out_count, self.stream)

    def refresh_kernel(self):
        """resampler workloads: the library reports the kernel that filtered the last call and its launches"""
        if self.fir:
            return
        import ctypes as C
        z = self.z
        handle = self.bank.handle if self.bank is not None else self.job.bank(0)
        n, buf = C.c_int(0), C.create_string_buffer(96)
        if z.lib().llz_cuda_resample_bank_last_run(handle, C.byref(n), buf, 96) == 0 and buf.value:
            self.kernel = buf.value.decode()
            self.launches_per_step = n.value + (1 if self.halo else 0)

    def time_steps(self, steps: int, warmup: int, gather: int = 0, sample_clocks: bool = True):
        """CUDA events on the launching stream around `steps` steps, barrier + synchronize on both sides, max over ranks"""
        torch, D = self.torch, self.D
        for _ in range(max(warmup, 3)):
            self.step(gather)
        sampler = ClockSampler(D.local) if sample_clocks else None   # NVML init happens here, outside the bracket
        D.barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step(gather)
        e1.record()
        if sampler:
            sampler.sample_now()                    # the launches above are still running: a sample under load
        D.barrier()
        clocks = sampler.result() if sampler else None
        self.refresh_kernel()
        ms_local = e0.elapsed_time(e1) / steps
        return D.max(ms_local), ms_local, clocks

    # ---- checks (never inside a timed region) -------------------------------------------------------------
    def parity(self):
        """windows of this rank's device-resident result against the oracle (reference restatement) on the same input:
        the first outputs of the shard (where the halo / zero history matters), the last ones, and two in between"""
        import oracle
        P = oracle.port()
        wl, torch = self.wl, self.torch
        W = 4096
        chans = sorted({0, self.cc // 2, self.cc - 1})
        rec = {"checked_outputs": 0, "windows": 0, "max_abs_diff": 0.0, "mismatches": 0}
        if self.fir:
            h = P.fir_design(0, wl["taps"], wl["fc"], 0.0, wl["win"])
            N = wl["taps"]
            scale = max(float(np.abs(h).sum()), 1.0)
            tol = (TOL_F32 if self.dtype == "f32" else TOL_F64) * scale
            rec.update(tolerance=tol, tolerance_rule="max |gpu - oracle| <= %g x max(1, sum|h|) of full scale (input in [-1, 1))"
                       % (TOL_F32 if self.dtype == "f32" else TOL_F64))
            starts = sorted({0, self.out_count // 3, 2 * self.out_count // 3 + 17, max(0, self.out_count - W)})
            num = den = 0.0
            for c in chans:
                for t0 in starts:
                    w = min(W, self.out_count - t0)
                    lo = self.halo + t0 - (N - 1)                      # index into dx_all of the first needed sample
                    xs = np.zeros(w + N - 1, np.float64)
                    a = max(lo, 0)
                    xs[a - lo:] = self.dx_all[c, a:self.halo + t0 + w].cpu().numpy().astype(np.float64)
                    want = P.fir_run(h, xs)[N - 1:]
                    got = self.dy[c, t0:t0 + w].cpu().numpy().astype(np.float64)
                    d = np.abs(got - want)
                    rec["max_abs_diff"] = max(rec["max_abs_diff"], float(d.max()))
                    num += float((want ** 2).sum())
                    den += float(((got - want) ** 2).sum())
                    rec["checked_outputs"] += w
                    rec["windows"] += 1
            rec["snr_db"] = float(10 * np.log10(num / den)) if den > 0 else float("inf")
            rec["ok"] = rec["max_abs_diff"] <= tol and (self.dtype != "f32" or rec["snr_db"] >= 120.0)
        elif wl.get("interp"):
            L_ = wl["L"]
            plan = P.interp_plan(L_, wl["win"])
            F = plan.num_in
            fast = self.dtype == "f32"
            rec.update(tolerance=1 if fast else 0, tolerance_rule="|diff| <= 1 LSB (fast mode)" if fast else "bit-exact int16")
            n_frames = self.in_count // F
            for c in chans:
                for f0 in sorted({0, n_frames // 3, 2 * n_frames // 3 + 1, n_frames - 2}):
                    xs = self.dx_all[c, f0 * F:(f0 + 2) * F].cpu().numpy()
                    want = P.interp_run(plan, 1.0, xs)
                    got = self.dy[c, f0 * F * L_:(f0 + 2) * F * L_].cpu().numpy()
                    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
                    rec["max_abs_diff"] = max(rec["max_abs_diff"], float(d.max()))
                    rec["mismatches"] += int((d != 0).sum())
                    rec["checked_outputs"] += len(want)
                    rec["windows"] += 1
            rec["ok"] = rec["max_abs_diff"] <= (1 if fast else 0)
        else:
            L_, M, Q = wl["L"], wl["M"], self.q
            plan = P.resample_plan(L_, M, wl["win"], wl["k"])
            pad = (Q - 1 + M - 1) // M * M
            fast = self.dtype == "f32"
            rec.update(tolerance=1 if fast else 0, tolerance_rule="|diff| <= 1 LSB (fast mode)" if fast else "bit-exact int16")
            cyc = self.out_count // L_
            for c in chans:
                for j0 in sorted({0, cyc // 3, 2 * cyc // 3 + 5, max(0, cyc - W // L_ - 1)}):
                    o0 = j0 * L_
                    w = min(W, self.out_count - o0)
                    p0 = self.halo + j0 * M                              # dx_all index of the sample the window's phase 0 reads
                    need = (w * M) // L_ + 2
                    xs = np.zeros(pad + need, np.int16)
                    a = max(p0 - pad, 0)
                    b = min(p0 + need, self.x_stride)
                    xs[a - (p0 - pad):b - (p0 - pad)] = self.dx_all[c, a:b].cpu().numpy()
                    want = P.resample_run(plan, 1.0, xs, w, m0=pad // M * L_)
                    got = self.dy[c, o0:o0 + w].cpu().numpy()
                    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
                    rec["max_abs_diff"] = max(rec["max_abs_diff"], float(d.max()))
                    rec["mismatches"] += int((d != 0).sum())
                    rec["checked_outputs"] += w
                    rec["windows"] += 1
            rec["ok"] = rec["max_abs_diff"] <= (1 if fast else 0)
        rec["channels_checked"] = [self.c0 + c for c in chans]
        rec["against"] = "oracle/llz_oracle.c (restatement pinned to the compiled reference, tests/test_oracle.py)"
        return rec

    def one_gpu_identity(self):
        """N > 1: this rank recomputes its channels from sample 0 to the end of its segment the way the one-GPU call does
        (zero history, no segment boundary) and compares its shard with that, byte for byte"""
        z, torch, wl = self.z, self.torch, self.wl
        if self.D.world == 1:
            return None
        end_in = self.in_start + self.in_count
        need = self.cc * end_in * self.es * (2 if self.fir else 1.0 + wl["L"] / wl["M"])
        free, _ = torch.cuda.mem_get_info()
        if need > 0.9 * free:
            return {"skipped": f"needs {need / 1e9:.1f} GB on this rank"}
        x = torch.empty(self.cc, end_in, dtype=self.tdt, device="cuda")
        z.synth_lcg_at(x, end_in, self.cc, 0, end_in, self.lcg_kind, wl["seed"] + self.c0, self.stream)
        if self.fir:
            one = z.FirBank(self.cc, z.F32 if self.dtype == "f32" else z.F64, kind=z.LPF, flt_len=wl["taps"], fc1=wl["fc"], win=wl["win"])
            z._check(z.lib().llz_cuda_fir_bank_set_algo(one.handle, z.lib().llz_cuda_fir_bank_get_algo(self.bank_handle)), "set_algo")
            y = torch.empty(self.cc, end_in, dtype=self.tdt, device="cuda")
            one.run(x, end_in, y, end_in, end_in, self.stream)
            n_out = end_in
        else:
            one = z.ResampleBank(z.KIND_RESAMPLE, wl["L"], wl["M"], self.cc, win=wl["win"], k_override=wl["k"],
                                 acc=z.ACC_F32 if self.dtype == "f32" else z.ACC_F64)
            n_out = one.out_len(end_in)
            y = torch.empty(self.cc, n_out, dtype=self.tdt, device="cuda")
            one.run(x, end_in, end_in, y, n_out, self.stream)
        torch.cuda.synchronize()
        same = bool(torch.equal(y[:, self.out_start:self.out_start + self.out_count], self.dy))
        one.close()
        del x, y
        return {"bit_identical": same, "compared_outputs": self.outs_rank,
                "against": "the one-GPU call over [0, end of this rank's segment) of the same channels"}

    def close(self):
        if self.job is not None:
            self.job.close()
        if self.bank is not None:
            self.bank.close()
        self.dx_all = self.dy = None
        self.torch.cuda.empty_cache()


TOL_F64, TOL_F32 = 1e-12, 1e-5


def measure(z, D: Dist, mg, name: str, dtype: str, algo: str, steps: int, warmup: int, args, headline: bool):
    """one workload -> the record that goes into the JSON line (headline) or under "workloads" """
    torch = D.torch
    W = Workload(z, D, name, dtype, algo, mg)
    wl = W.wl
    hbm_peak, peak_src = peaks()
    fma_dtype = z.F32 if dtype == "f32" else z.F64
    fma_peak_nominal = FP32_NOMINAL_TFLOPS if dtype == "f32" else FP64_NOMINAL_TFLOPS
    fma_peak_measured = z.probe_fma(fma_dtype)

    ms_step, ms_local, clocks = W.time_steps(steps, warmup, z.GATHER_NONE)
    outs_all = int(D.sum(float(W.outs_rank)))
    value = outs_all / (ms_step * 1e-3) / 1e6
    ach_gbs = W.outs_rank * W.bytes_per_out / (ms_local * 1e-3) / 1e9
    ach_tf = W.outs_rank * W.flop_per_out / (ms_local * 1e-3) / 1e12

    parity = W.parity()
    parity_all = D.gather_obj(parity)
    identity = D.gather_obj(W.one_gpu_identity()) if D.world > 1 else None

    # ---- N > 1: the same step with the result gathered on rank 0, two ways (library gather modes) ----
    gather = None
    if D.world > 1 and not args.no_gather:
        gather = {"note": "whole planar result on rank 0's device; compute_only is the timed step above (no collective: "
                          "shards carry their halo)", "compute_only_ms": ms_step,
                  "result_bytes": W.C_total * W.total_out * W.es, "root_ingress_bytes": (W.C_total * W.total_out - (W.outs_rank if D.rank == 0 else 0)) * W.es}
        try:
            mg.result_alloc(0, W.C_total * W.total_out * W.es)
            own = checksum64(W.dy)
            for mode, key in ((z.GATHER_NCCL, "nccl_chunked_send_recv"), (z.GATHER_PEER, "peer_store_fused"), (z.GATHER_COPY, "copy_engine_push")):
                try:
                    gms, _, _ = W.time_steps(max(3, min(steps, 10)), 2, mode, sample_clocks=False)
                    # every rank's region of the gathered result against that rank's own shard
                    sums = D.gather_obj((W.c0, W.cc, W.out_start, W.out_count, own))
                    ok = None
                    if D.rank == 0:
                        import ctypes
                        full = torch.empty(W.C_total, W.total_out, dtype=W.tdt, device="cuda")
                        rt = ctypes.CDLL("libcudart.so.12")
                        rt.cudaMemcpy(ctypes.c_void_p(full.data_ptr()), ctypes.c_void_p(mg.result_ptr(0)),
                                      ctypes.c_size_t(full.numel() * W.es), 3)
                        torch.cuda.synchronize()
                        ok = [checksum64(full[c0:c0 + cc, o0:o0 + oc]) == cs for c0, cc, o0, oc, cs in sums]
                        del full
                    ingress = gather["root_ingress_bytes"] if D.rank == 0 else 0
                    ingress = D.max(float(ingress))
                    floor_ms = max(ms_step, ingress / 900e9 * 1e3)
                    gather[key] = {"compute_plus_gather_ms": gms, "floor_ms": floor_ms, "over_floor": gms / floor_ms,
                                   "root_ingress_gbs": ingress / (gms * 1e-3) / 1e9,
                                   "regions_match_each_ranks_shard": ok}
                except Exception as ex:             # noqa: BLE001
                    gather[key] = {"error": repr(ex)}
                    D.barrier()
            gather["floor_rule"] = "max(compute-only step, bytes into rank 0 / 900 GB/s NVLink ingress)"
            D.barrier()
            mg.result_free()
        except Exception as ex:                     # noqa: BLE001
            gather["error"] = repr(ex)

    # ---- end to end through the C-ABI with host buffers (this rank's shard through *_run_host) ----
    e2e = None
    if headline or args.e2e_all:
        e2e = end_to_end(z, D, W, outs_all, min(steps, 5), args)

    rec_rank0 = None
    if D.rank == 0:
        if W.fir_fft:
            ginstr = W.outs_rank * W.fft_instr_per_out / (ms_local * 1e-3) / 1e12
            fma_pipe = {"achieved": ginstr, "unit": "T lane-instr/s", "instr_per_output": W.fft_instr_per_out,
                        "peak_measured": fma_peak_measured / 2 if fma_peak_measured else None,
                        "peak_nominal": fma_peak_nominal / 2,
                        "frac_of_measured": ginstr / (fma_peak_measured / 2) if fma_peak_measured else None,
                        "frac_of_nominal": ginstr / (fma_peak_nominal / 2),
                        "direct_form_equivalent_tflops": ach_tf,
                        "note": f"overlap-save ({W.fft_desc}, arithmetic in the bank's type): "
                                f"{W.fft_instr_per_out:.1f} FMA-pipe instructions per output instead of {wl['taps']} FMAs"}
        else:
            fma_pipe = {"achieved": ach_tf, "unit": "TFLOP/s", "peak_measured": fma_peak_measured,
                        "peak_nominal": fma_peak_nominal,
                        "frac_of_measured": ach_tf / fma_peak_measured if fma_peak_measured else None,
                        "frac_of_nominal": ach_tf / fma_peak_nominal,
                        "note": f"algorithmic flops ({W.flop_per_out:.0f} per output) against the FMA pipe of the accumulator type"}
            if W.kernel.startswith("poly_bank_umma_kernel"):
                fma_pipe["note"] += ("; the kernel evaluates the sums as exact integers on the tensor cores (tcgen05.mma.kind::i8, "
                                     "%d digit products per multiply-add), so the fraction is its speed relative to the FMA roof "
                                     "it no longer uses" % (6 if W.dtype == "f32" else 10))
        tensor = None
        if W.kernel.startswith("poly_bank_umma_kernel"):
            planes = 3 if W.dtype == "f32" else 5
            macs = umma_macs_per_output(wl["L"], wl["M"], W.q, planes)
            bf16, tsrc = peaks_tensor()
            tops = W.outs_rank * macs * 2.0 / (ms_local * 1e-3) / 1e12
            tensor = {"achieved": tops, "peak": 2.0 * bf16, "unit": "TOP/s", "frac": tops / (2.0 * bf16), "peak_source": tsrc,
                      "int8_macs_per_output": macs, "digit_products_per_mac": 2 * planes,
                      "note": "executed int8 operations (2 x MACs of the issued 128x64x32 MMAs); an N = 64 MMA is bound by its "
                              "shared-memory operand reads at 2/3 of the kind::i8 rate (tools/probe_umma_rate.cu: 48 cycles "
                              "against 32), so 0.67 is this tile shape's ceiling"}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(f"{name}_{dtype}")
        time_sharded = D.world > 1 and W.mode == z.SHARD_TIME
        rec_rank0 = {
            "value": value, "unit": UNIT, "ms_per_step": ms_step, "steps": steps, "dtype": W.dtype_name,
            "scaling": "strong" if D.world > 1 else "weak",
            "config": {"workload": wl["desc"], "name": name, "channels": W.C_total, "samples_per_channel": W.n_total,
                       "outputs_per_step": outs_all, "outputs_per_step_this_rank": W.outs_rank,
                       "sharding": ("whole job on one GPU" if D.world == 1 else
                                    ("%d time segments (llz_cuda_mgpu_*), each rank reads a %d-sample halo before its segment, no collective"
                                     % (D.world, W.halo_full)) if time_sharded else
                                    "%d channel shards of %d channels (llz_cuda_mgpu_*), no collective" % (D.world, W.cc)),
                       "l2": f"inputs {W.dx_all.numel() * W.es / 1e9:.2f} GB per GPU vs 126 MB L2"
                             + ("" if W.dx_all.numel() * W.es > 4 * 126e6 else "; outputs + inputs of consecutive steps still exceed L2" if (W.dx_all.numel() + W.dy.numel()) * W.es > 2 * 126e6 else " (fits: strong-scaled shard)"),
                       "input": "integer LCG noise generated on the device (SURVEY.md 8d)",
                       **({"fir_algo": f"overlap-save, {W.fft_desc}" if W.fir_fft else "direct form"} if W.fir else {})},
            "roofline": ({"kernel": W.kernel, "bound": "tensor", "achieved": tensor["achieved"], "peak": tensor["peak"],
                          "unit": "TOP/s", "frac": tensor["frac"], "traffic": traffic if D.world == 1 else None,
                          "peak_source": tensor["peak_source"], "tensor": tensor,
                          "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                  "peak_source": peak_src},
                          "algorithmic": {"bytes_per_output": W.bytes_per_out, "flop_per_output": W.flop_per_out,
                                          "outputs_per_launch": W.outs_rank},
                          "fma_pipe": fma_pipe} if tensor else
                         {"kernel": W.kernel, "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                          "frac": ach_gbs / hbm_peak, "traffic": traffic if D.world == 1 else None, "peak_source": peak_src,
                          "algorithmic": {"bytes_per_output": W.bytes_per_out, "flop_per_output": W.flop_per_out,
                                          "outputs_per_launch": W.outs_rank},
                          "fma_pipe": fma_pipe}),
            "parity": parity_all[0] if D.world == 1 else {"ok": all(p["ok"] for p in parity_all), "per_rank": parity_all},
            "gpu_launches": W.launches_per_step * steps, "clocks": clocks,
        }
        if identity is not None:
            rec_rank0["one_gpu_identity"] = {"per_rank": identity,
                                             "bit_identical": all(bool(i and i.get("bit_identical")) for i in identity)}
        if gather is not None:
            rec_rank0["gather"] = gather
        if e2e is not None:
            rec_rank0["e2e"] = e2e
    W.close()
    return rec_rank0


def end_to_end(z, D: Dist, W: Workload, outs_all: int, e_steps: int, args):
    """the same pass through llz_cuda_*_bank_run_host with page-locked HOST buffers: H2D, kernels and D2H of every chunk
    inside the timed region (host wall clock around synchronous calls), this rank's shard"""
    torch = D.torch
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        n, n_out, cc = W.in_count, W.out_count, W.cc
        in_bytes, out_bytes = cc * n * W.es, cc * n_out * W.es
        hx = z.host_alloc(in_bytes, W.np_dt).reshape(cc, n)
        hy = z.host_alloc(out_bytes, W.np_dt).reshape(cc, n_out)
        torch.from_numpy(hx).copy_(W.dx_all[:, W.halo:])       # this rank's own samples (the halo goes into the bank's history)
        torch.cuda.synchronize()
        if W.job is not None:
            handle = W.job.bank(0)
            bank = (z.FirBank if W.fir else z.ResampleBank).__new__(z.FirBank if W.fir else z.ResampleBank)
            bank.handle = handle
        else:
            bank = W.bank

        def e2e_step():
            if W.halo:
                bank.set_history(W.dx_all, W.x_stride, W.stream)
                torch.cuda.synchronize()
            else:
                bank.reset()
            if W.fir:
                bank.run_host(hx, n, hy, n_out, n)
            else:
                bank.run_host(hx, n, n, hy, n_out)
        e2e_step()                                  # warm-up: staging buffers, streams
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()                              # synchronous: result is in hy on return
        dt = D.max(time.perf_counter() - t0)
        # the result read back is the full output of the shard; compare a window with the device-resident result
        dev = W.dy[0, :4096].cpu().numpy()
        max_diff = float(np.abs(hy[0, :4096].astype(np.float64) - dev.astype(np.float64)).max())
        e2e = {"value": outs_all * e_steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(D.sum(float(in_bytes))), "d2h_bytes_per_step": int(D.sum(float(out_bytes))),
               "steps": e_steps, "matches_device_result": bool(np.array_equal(hy[0, :4096], dev)),
               "max_abs_diff_vs_device": max_diff,
               "gbs_per_direction_all_gpus": D.sum(float(in_bytes)) * e_steps / dt / 1e9,
               "api": "llz_cuda_fir_bank_run_host" if W.fir else "llz_cuda_resample_bank_run_host",
               "host_memory": "page-locked (llz_cuda_host_alloc)"}
        if W.job is not None:
            bank.handle = 0
        z.host_free(hx.reshape(-1))
        z.host_free(hy.reshape(-1))
        return e2e
    except Exception as ex:                         # noqa: BLE001
        return {"value": None, "unit": UNIT, "error": repr(ex)}


def run_cuda(args):
    import torch
    import llzlab_b200 as z

    D = Dist()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libllzfilter_cuda has no CPU path (use --impl reference for the CPU arm)")
    bind_to_gpu_numa_node(D.local)
    torch.cuda.set_device(D.local)
    D.init()
    z.lib()
    mg = None
    if D.world > 1:
        # the library's multi-GPU context, one rank per process: rank 0's NCCL id travels through torch.distributed
        ids = [z.mgpu_unique_id() if D.rank == 0 else None]
        D.dist.broadcast_object_list(ids, src=0)
        mg = z.Mgpu(unique_id=ids[0], world=D.world, rank=D.rank)

    head_name = args.workload or "c2"
    head = measure(z, D, mg, head_name, args.dtype, args.algo, args.steps, args.warmup, args, headline=True)

    # ---- every other BASELINE config, same run (device-timed, parity-checked; fewer steps) ----
    others = {}
    if args.workload is None and not args.headline_only:
        extra = (("interp4", "f64"), ("interp4", "f32")) if D.world == 1 else ()
        for name, dtype in (("c2", "f32"), ("c3", "f64"), ("c3", "f32"), ("c4", "f64"), ("c4", "f32"), ("c5", "f64"), ("c5", "f32")) + extra:
            key = f"{name}_{'f32' if dtype == 'f32' else ('f64' if WORKLOADS[name]['kind'] == 'fir' else 'exact')}"
            try:
                rec = measure(z, D, mg, name, dtype, "auto", max(3, min(args.steps, 5)), 3, args, headline=False)
            except Exception as ex:                 # noqa: BLE001
                rec = {"error": repr(ex)}
                D.barrier()
            if D.rank == 0:
                others[key] = rec

    if D.rank != 0:
        if mg is not None:
            mg.close()
        if D.world > 1:
            D.dist.destroy_process_group()
        return 0

    # ---- CPU baseline: the unmodified reference, one thread, bounded sample ----
    wl = WORKLOADS[head_name]
    cpu = None
    if D.world == 1 and not args.no_cpu:
        if wl["kind"] == "fir":
            n_cpu = 480_000 if wl["taps"] < 1000 else 40_960
            ch_cpu = max(1, int(1.0e8 * 127 / wl["taps"] / n_cpu))      # ~12 s at the probed 8.4 Msamples/s
        else:
            n_cpu, ch_cpu = reference_sample_shape(wl)[1] * 4, 6
        rate, kind, dt = cpu_reference_rate(wl, 1, ch_cpu, n_cpu)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
               "sample": f"{ch_cpu} channel(s) x {n_cpu} input samples of the workload, frame by frame through the reference API, 1 thread (the reference is single-threaded); host has {os.cpu_count()} logical cores"}

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None,
        "dtype": head["dtype"], "data": "synthetic", "config": head["config"], "roofline": head["roofline"],
        "cpu_baseline": cpu, "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
        "parity": head["parity"],
    }
    for k in ("one_gpu_identity", "gather"):
        if k in head:
            line[k] = head[k]
    if others:
        line["workloads"] = others
    print(json.dumps(line))
    if mg is not None:
        mg.close()
    if D.world > 1:
        D.dist.destroy_process_group()
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
    ap.add_argument("--workload", default=None, choices=["c1"] + sorted(WORKLOADS),
                    help="run ONE config as the headline; default: C2 as the headline and every other config under 'workloads'")
    ap.add_argument("--headline-only", action="store_true", help="default run without the other configs")
    ap.add_argument("--e2e-all", action="store_true", help="host-buffer end-to-end leg for the non-headline configs too")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--algo", default="auto", choices=["auto", "direct", "fft"],
                    help="FIR kernel family (c2/c5): auto = overlap-save where it applies, else direct form")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gather", action="store_true", help="skip the compute+gather measurements at N > 1")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    args = ap.parse_args()
    if args.workload == "c1":
        return run_c1(args) if int(os.environ.get("RANK", "0")) == 0 else 0
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
