"""Per-role clock stamps of the tcgen05 phase-bank kernel (CTA 0, tiles 1..6 of the last launch).  Needs the profiling build:
   make -C llzlab_b200/csrc EXTRA_NVFLAGS=-DLLZ_UMMA_TRACE OUT=$PWD/llzlab_b200/libllzfilter_cuda_trace.so OBJ=$PWD/llzlab_b200/csrc/_obj_trace
   LLZLAB_B200_LIB=$PWD/llzlab_b200/libllzfilter_cuda_trace.so python tools/umma_trace.py [slab MiB]"""
import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import numpy as np, torch, llzlab_b200 as z
slab = float(sys.argv[1]) if len(sys.argv) > 1 else 256.0
what = sys.argv[2] if len(sys.argv) > 2 else "c4"
if len(sys.argv) > 3:
    z.lib().llz_debug_umma_select(int(sys.argv[3]))          # which CTA stamps its roles (default 74)
C_, frames = (8, 352) if what == "c4" else (64, 500)
for acc, name in ((z.ACC_F64, "exact"), (z.ACC_F32, "fast")):
    if what == "c4":
        bank = z.ResampleBank(z.KIND_RESAMPLE, 320, 147, C_, k_override=128, acc=acc)
    elif what == "c3":
        bank = z.ResampleBank(z.KIND_RESAMPLE, 1, 3, C_, acc=acc)
    else:
        bank = z.ResampleBank(z.KIND_INTERP, 4, 1, C_, acc=acc)
    bank.set_tiles(4)
    n = bank.info.num_in * frames
    x = torch.empty(C_, n, dtype=torch.int16, device="cuda")
    z.synth_lcg(x, n, C_, n, 2, 777)
    n_out = bank.out_len(n)
    y = torch.empty(C_, n_out, dtype=torch.int16, device="cuda")
    z.tune("umma_slab_mib", slab)
    for _ in range(2):
        bank.reset(); bank.run(x, n, n, y, n_out)
    torch.cuda.synchronize()
    t = np.zeros(16 * 16 * 16, dtype=np.int64)
    assert z.lib().llz_debug_umma_trace(t.ctypes.data_as(ctypes.c_void_p)) == 0
    t = t.reshape(16, 16, 16)
    cta = np.zeros(2 * 148, dtype=np.int64)
    if hasattr(z.lib(), "llz_debug_umma_cta") and z.lib().llz_debug_umma_cta(cta.ctypes.data_as(ctypes.c_void_p)) == 0:
        dur = (cta[1::2] - cta[0::2]).astype(np.float64)
        dur = dur[dur > 0]
        order = np.argsort(-(cta[1::2] - cta[0::2]))
        print("  slowest CTAs (id: cycles):", [(int(i), int(cta[2 * i + 1] - cta[2 * i])) for i in order[:12]], " fastest:", [(int(i), int(cta[2 * i + 1] - cta[2 * i])) for i in order[-6:]])
        print(f"  CTA durations of the last launch (cycles, clocks are per SM): min {dur.min():.0f} mean {dur.mean():.0f} max {dur.max():.0f}  max/mean {dur.max() / dur.mean():.3f}")
    t0 = t[1, 0, 0]
    print(what, name, "slab", slab, "taps per phase", bank.info.taps_per_phase)
    for n_ in range(1, 7):
        print(f" tile {n_}: producer stage-free {[int(v - t0) for v in t[0, n_, :3]]}")
        print(f"         issuer t_empty-done {int(t[1, n_, 0] - t0)} s_full-done {[int(v - t0) for v in t[1, n_, 1:4]]} issued {int(t[1, n_, 8] - t0)}")
        for w in range(2, 10):
            print(f"         epi warp {w + 2}: wait-start {int(t[w, n_, 0] - t0)} t_full {int(t[w, n_, 1] - t0)} drained {int(t[w, n_, 2] - t0)} finished {int(t[w, n_, 3] - t0)}")
    bank.close()
