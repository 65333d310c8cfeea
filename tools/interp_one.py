import sys, os
sys.path.insert(0, os.getcwd())
import torch, llzlab_b200 as z
C_, frames = 64, 2000
acc = z.ACC_F64 if sys.argv[1] == "exact" else z.ACC_F32
bank = z.ResampleBank(z.KIND_INTERP, 4, 1, C_, acc=acc)
n = 1024 * frames
x = torch.empty(C_, n, dtype=torch.int16, device="cuda")
z.synth_lcg(x, n, C_, n, 2, 777)
y = torch.empty(C_, n * 4, dtype=torch.int16, device="cuda")
for _ in range(3):
    bank.reset(); bank.run(x, n, n, y, n * 4)
torch.cuda.synchronize()
print(bank.last_run())
