import sys, os
sys.path.insert(0, os.getcwd())
import torch, llzlab_b200 as z
C_, frames = 8, 400
acc = z.ACC_F64 if sys.argv[1] == "exact" else z.ACC_F32
bank = z.ResampleBank(z.KIND_RESAMPLE, 320, 147, C_, k_override=128, acc=acc)
bank.set_tiles(4)
n = bank.info.num_in * frames
x = torch.empty(C_, n, dtype=torch.int16, device="cuda")
z.synth_lcg(x, n, C_, n, 2, 777)
n_out = bank.out_len(n)
y = torch.empty(C_, n_out, dtype=torch.int16, device="cuda")
z.tune("umma_slab_mib", float(sys.argv[2]))
for _ in range(3):
    bank.reset(); bank.run(x, n, n, y, n_out)
torch.cuda.synchronize()
