"""Host <-> device copy ceiling of one box with N GPUs busy at once (e2e at N = 8 scaled 0.18).
Run under torchrun with N ranks: every rank streams contiguous page-locked copies H2D and D2H simultaneously (two streams,
no kernels) for a few seconds after a barrier; rank 0 prints per-GPU and aggregate GB/s per direction, plus what the box
says about NUMA nodes and GPU affinity.  Variants: default pinned memory, portable + write-combined H2D source, and the
process bound to the GPU's NUMA node (when there is more than one).
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_multi.py [seconds]"""
import os, sys, time, subprocess
import torch
import torch.distributed as dist

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
CH = 64 << 20
rt = torch.cuda.cudart()


def probe(flags_in: int, label: str):
    import ctypes
    lib = ctypes.CDLL("libcudart.so.12")
    lib.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    hin, hout = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.cudaHostAlloc(ctypes.byref(hin), CH * 4, flags_in) == 0
    assert lib.cudaHostAlloc(ctypes.byref(hout), CH * 4, 0) == 0
    ctypes.memset(hin, 1, CH * 4); ctypes.memset(hout, 0, CH * 4)
    d_in = torch.empty(CH * 2, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(CH * 2, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    lib.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    out = {}
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n1 = n2 = 0
        while time.perf_counter() - t0 < secs:
            for k in range(4):
                if mode in ("h2d", "both"):
                    lib.cudaMemcpyAsync(ctypes.c_void_p(d_in.data_ptr() + (k % 2) * CH), ctypes.c_void_p(hin.value + k * CH), CH, 1, ctypes.c_void_p(s1.cuda_stream)); n1 += 1
                if mode in ("d2h", "both"):
                    lib.cudaMemcpyAsync(ctypes.c_void_p(hout.value + k * CH), ctypes.c_void_p(d_out.data_ptr() + (k % 2) * CH), CH, 2, ctypes.c_void_p(s2.cuda_stream)); n2 += 1
            s1.synchronize(); s2.synchronize()
        dt = time.perf_counter() - t0
        out[mode] = (n1 * CH / dt / 1e9, n2 * CH / dt / 1e9)
    lib.cudaFreeHost(hin); lib.cudaFreeHost(hout)
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, out)
    else:
        rows = [out]
    if rank == 0:
        for mode in ("h2d", "d2h", "both"):
            h = [r[mode][0] for r in rows]; d = [r[mode][1] for r in rows]
            print(f"N={world} {label:34s} {mode:4s}: H2D per GPU {min(h):5.1f}..{max(h):5.1f} GB/s, sum {sum(h):6.1f};  D2H per GPU {min(d):5.1f}..{max(d):5.1f}, sum {sum(d):6.1f}", flush=True)


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "ls -d /sys/devices/system/node/node* | wc -l; nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2"]):
        try:
            print(subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout[:2500], flush=True)
        except Exception as e:  # noqa: BLE001
            print("cannot run", cmd, e)
probe(0, "pinned (cudaHostAllocDefault)")
probe(1 | 4, "portable + write-combined source")
if world > 1:
    dist.destroy_process_group()
