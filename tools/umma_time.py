"""times the tcgen05 phase-bank kernel on a slice of config C4 (8 channels), exact and fast mode, against the mma.sync tiles"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, llzlab_b200 as z
C_, frames = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 400
slabs = [float(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [96.0]
for acc, accname in ((z.ACC_F64, "exact"), (z.ACC_F32, "fast")):
    for tiles in (1, 4):
        bank = z.ResampleBank(z.KIND_RESAMPLE, 320, 147, C_, k_override=128, acc=acc)
        bank.set_tiles(tiles)
        n = bank.info.num_in * frames
        x = torch.empty(C_, n, dtype=torch.int16, device="cuda")
        z.synth_lcg(x, n, C_, n, 2, 777)
        n_out = bank.out_len(n)
        y = torch.empty(C_, n_out, dtype=torch.int16, device="cuda")
        for slab in ((96.0,) if tiles == 1 else slabs):
            z.tune("umma_slab_mib", slab)
            for _ in range(2):
                bank.reset(); bank.run(x, n, n, y, n_out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                bank.reset(); bank.run(x, n, n, y, n_out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{accname} tiles {tiles} slab {slab} MiB: {ms:.3f} ms  {C_ * n_out / ms / 1e6:.1f} Gsamples/s  checksum {int(y.view(torch.int16).to(torch.int64).sum())}", flush=True)
        bank.close()
