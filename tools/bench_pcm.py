#!/usr/bin/env python
"""Micro-benchmark of the PCM layout kernels (HBM-bound): achieved GB/s against MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llzlab_b200 as z  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
for C, n_frames, fmt, ptype, name in [(2, 256 * 1024 * 1024, 0, 0, "stereo s16 -> planar s16"), (8, 64 * 1024 * 1024, 0, 2, "8ch s16 -> planar f64"),
                                     (64, 16 * 1024 * 1024, 0, 0, "64ch s16 -> planar s16"), (2, 128 * 1024 * 1024, 2, 1, "stereo f32 -> planar f32")]:
    bps, pes = (2, 3, 4)[fmt], (2, 4, 8)[ptype]
    frames = torch.empty(n_frames * C * bps, dtype=torch.uint8, device="cuda").random_(0, 255)
    planar = torch.empty(C * n_frames * pes, dtype=torch.uint8, device="cuda")
    nbytes = n_frames * C * (bps + pes)
    for direction in ("deinterleave", "interleave"):
        def run():
            if direction == "deinterleave":
                z.pcm_deinterleave(frames, fmt, C, n_frames, planar, ptype, n_frames, st)
            else:
                z.pcm_interleave(planar, ptype, n_frames, C, n_frames, frames, fmt, st)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = nbytes / ms / 1e6
        print(json.dumps({"kernel": "pcm_transpose_kernel", "case": name, "direction": direction, "ms": ms, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak}))
