"""One warm step + one profiled step of a bench.py workload (the same Workload object bench.py times), bracketed by
cudaProfilerStart / Stop so that ncu --profile-from-start off sees exactly one step:
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none --csv --log-file out.csv python tools/profile_step.py c4 f64"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import llzlab_b200 as z

name, dtype = sys.argv[1], sys.argv[2]
D = bench.Dist()
torch = D.torch
torch.cuda.set_device(0)
z.lib()
W = bench.Workload(z, D, name, dtype, "auto", None)
for _ in range(2):
    W.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
W.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
W.refresh_kernel()
print(name, dtype, W.kernel, W.launches_per_step, flush=True)
W.close()
