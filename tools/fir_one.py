"""One overlap-save FIR bank call in a loop, for ncu and quick timing:
    python tools/fir_one.py <taps> <f64|f32> <fft_size> [channels] [samples] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llzlab_b200 as z  # noqa: E402

taps, dt, size = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
C_ = int(sys.argv[4]) if len(sys.argv) > 4 else 16
n = int(sys.argv[5]) if len(sys.argv) > 5 else 16_000_000
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dtype, tdt = (z.F64, torch.float64) if dt == "f64" else (z.F32, torch.float32)
dx = torch.randn(C_, n, dtype=tdt, device="cuda")
dy = torch.empty_like(dx)
bank = z.FirBank(C_, dtype, kind=z.LPF, flt_len=taps, fc1=0.2, algo=z.FIR_FFT)
bank.set_fft_size(size)
for k, v in (a.split("=") for a in os.environ.get("LLZ_TUNE", "").split(",") if a):
    z.tune(k, float(v))
for _ in range(2):
    bank.run(dx, n, dy, n, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    bank.run(dx, n, dy, n, n)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{dt} taps {taps} fft {size}: {ms:.3f} ms  {C_ * n / ms / 1e6:.1f} Gsamples/s")
bank.close()
