import torch, time
n = 512*1024*1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device='cuda')
d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5, chunk=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d_in[o:o+chunk].copy_(h_in[o:o+chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_out[o:o+chunk].copy_(d_out[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
run(True, True, 1)
print("h2d only  GB/s", run(True, False))
print("d2h only  GB/s", run(False, True))
print("both (each dir) GB/s", run(True, True))
print("both, 64MiB chunks", run(True, True, chunk=64*1024*1024))
print("both, 8MiB chunks", run(True, True, chunk=8*1024*1024))
