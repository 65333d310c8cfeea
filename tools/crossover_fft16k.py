"""Where does the 16384-point kernel overtake the 8192-point one?  (tuning of kFirFft16kMinTapsAutoF64 / F32)
Run under gpurun: python tools/crossover_fft16k.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llzlab_b200 as z  # noqa: E402

C_, n = 16, 16_000_000
for dtype, tdt in ((z.F64, torch.float64), (z.F32, torch.float32)):
    dx = torch.randn(C_, n, dtype=tdt, device="cuda")
    dy = torch.empty_like(dx)
    for taps in (1025, 1537, 2049, 2561, 3073, 4095, 5121, 6145, 8191, 12289):
        row = [f"{'f64' if dtype == z.F64 else 'f32'} taps {taps:5d}:"]
        for size in ("8192", "16384"):
            if size == "8192" and taps > 6145:
                continue
            bank = z.FirBank(C_, dtype, kind=z.LPF, flt_len=taps, fc1=0.2, algo=z.FIR_FFT)
            bank.set_fft_size(int(size))
            for _ in range(3):
                bank.run(dx, n, dy, n, n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                bank.run(dx, n, dy, n, n)
            e1.record()
            torch.cuda.synchronize()
            row.append(f"{size} {C_ * n * 5 / e0.elapsed_time(e1) / 1e6:8.1f} Gs/s")
            bank.close()
        print("  ".join(row))
