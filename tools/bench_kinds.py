import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llzlab_b200 as z
C_, n = 64, 4_000_000
dx = torch.randint(-20000, 20000, (C_, n), dtype=torch.int16, device="cuda")
for kind, L, M, name in ((z.KIND_INTERP, 3, 1, "interp x3"), (z.KIND_INTERP, 2, 1, "interp x2"), (z.KIND_DECIMATE, 1, 3, "decimate /3"), (z.KIND_RESAMPLE, 3, 1, "resample 3/1"), (z.KIND_RESAMPLE, 3, 2, "resample 3/2"), (z.KIND_RESAMPLE, 1, 3, "resample 1/3")):
    for acc in (z.ACC_F64, z.ACC_F32):
        bank = z.ResampleBank(kind, L, M, C_, acc=acc)
        info = bank.info
        nn = n // info.num_in * info.num_in
        n_out = bank.out_len(nn)
        dy = torch.empty(C_, n_out, dtype=torch.int16, device="cuda")
        for _ in range(2):
            bank.reset(); bank.run(dx, n, nn, dy, n_out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            bank.reset(); bank.run(dx, n, nn, dy, n_out)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:14s} acc={acc} Q={info.taps_per_phase:4d} n={info.n:5d}: {C_*n_out*3/e0.elapsed_time(e1)/1e6:8.1f} Gs/s out")
        bank.close()
