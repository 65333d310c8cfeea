"""llzlab_b200 -- ctypes view of ``libllzfilter_cuda.so`` (the product is the C library).

The reference (templeblock/llzlab) is C and so is this library's host side: tap design, polyphase
planning, handles and the C-ABI live in ``csrc/`` and ``include/``.  This module only binds the
exported C entry points so that the parity tests and ``bench.py`` can drive them from Python with
numpy / torch buffers.  Names follow the C API one to one (``llz_fir_filter``, ``llz_resample``,
``llz_cuda_fir_bank_run`` ...).

There is no fallback of any kind: if the shared library is missing or fails to load, importing
``llzlab_b200.lib()`` raises, and every data-path call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LLZLAB_B200_LIB: another build of the same library (profiling builds, tools/umma_trace.py)
LIB_PATH = os.environ.get("LLZLAB_B200_LIB") or os.path.join(HERE, "libllzfilter_cuda.so")
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

HAMMING, BLACKMAN, KAISER = 0, 1, 2
LPF, HPF, BPF, BSF = 0, 1, 2, 3
F64, F64_STRICT, F32 = 0, 1, 2
FIR_AUTO, FIR_DIRECT, FIR_FFT = 0, 1, 2
ACC_F64, ACC_F64_STRICT, ACC_F32 = 0, 1, 2
TILES_AUTO, TILES_INT8, TILES_FP64_TENSOR, TILES_CUDA_CORE, TILES_INT8_TCGEN05 = 0, 1, 2, 3, 4
SHARD_CHANNEL, SHARD_TIME = 0, 1
GATHER_NONE, GATHER_NCCL, GATHER_PEER, GATHER_COPY = 0, 1, 2, 3
KIND_DECIMATE, KIND_INTERP, KIND_RESAMPLE = 0, 1, 2
PCM_S16, PCM_S24, PCM_F32 = 0, 1, 2
PLANAR_S16, PLANAR_F32, PLANAR_F64 = 0, 1, 2
FAIL = C.c_ulong(-1).value

_dp = C.POINTER(C.c_double)
_ll = C.c_longlong
_vp = C.c_void_p
_ul = C.c_ulong


class LlzError(RuntimeError):
    pass


class ResampleInfo(C.Structure):
    _fields_ = [("kind", C.c_int), ("L", C.c_int), ("M", C.c_int), ("n", C.c_int),
                ("taps_per_phase", C.c_int), ("num_in", C.c_int), ("num_out", C.c_int),
                ("n_channels", C.c_int), ("acc", C.c_int)]


class Segment(C.Structure):
    _fields_ = [("in_start", _ll), ("in_count", _ll), ("halo", _ll), ("out_start", _ll),
                ("out_count", _ll)]


class Shard(C.Structure):
    _fields_ = [("first_channel", C.c_int), ("n_channels", C.c_int), ("seg", Segment)]


def build(verbose: bool = False) -> str:
    """Compile ``libllzfilter_cuda.so`` in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


# (name, restype, argtypes) for every exported function of include/*.h
_SIGNATURES = [
    # llz_fir.h
    ("llz_fir_filter_lpf_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_int]),
    ("llz_fir_filter_hpf_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_int]),
    ("llz_fir_filter_bandpass_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]),
    ("llz_fir_filter_bandstop_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]),
    ("llz_fir_filter_uninit", None, [_ul]),
    ("llz_fir_filter", C.c_int, [_ul, _vp, _vp, C.c_int]),
    ("llz_fir_filter_flush", C.c_int, [_ul, _vp]),
    ("llz_hamming", C.c_int, [_dp, C.c_int]),
    ("llz_blackman", C.c_int, [_dp, C.c_int]),
    ("llz_kaiser", C.c_int, [_dp, C.c_int]),
    ("llz_kaiser_beta", C.c_int, [_dp, C.c_int, C.c_double]),
    ("llz_kaiser_atten2beta", C.c_double, [C.c_double]),
    ("llz_hamming_cof_num", C.c_int, [C.c_double]),
    ("llz_blackman_cof_num", C.c_int, [C.c_double]),
    ("llz_kaiser_cof_num", C.c_int, [C.c_double, C.c_double]),
    ("llz_fir_lpf_cof", C.c_int, [C.POINTER(_dp), C.c_int, C.c_double, C.c_int]),
    ("llz_fir_hpf_cof", C.c_int, [C.POINTER(_dp), C.c_int, C.c_double, C.c_int]),
    ("llz_fir_bandpass_cof", C.c_int, [C.POINTER(_dp), C.c_int, C.c_double, C.c_double, C.c_int]),
    ("llz_fir_bandstop_cof", C.c_int, [C.POINTER(_dp), C.c_int, C.c_double, C.c_double, C.c_int]),
    ("llz_conv", C.c_double, [_dp, _dp, C.c_int]),
    # llz_resample.h
    ("llz_decimate_init", _ul, [C.c_int, C.c_double, C.c_int]),
    ("llz_decimate_uninit", None, [_ul]),
    ("llz_interp_init", _ul, [C.c_int, C.c_double, C.c_int]),
    ("llz_interp_uninit", None, [_ul]),
    ("llz_resample_filter_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_int]),
    ("llz_resample_filter_uninit", None, [_ul]),
    ("llz_get_resample_framelen_bytes", C.c_int, [_ul]),
    ("llz_decimate", C.c_int, [_ul, _vp, C.c_int, _vp, C.POINTER(C.c_int)]),
    ("llz_interp", C.c_int, [_ul, _vp, C.c_int, _vp, C.POINTER(C.c_int)]),
    ("llz_resample", C.c_int, [_ul, _vp, C.c_int, _vp, C.POINTER(C.c_int)]),
    # llz_cuda.h
    ("llz_cuda_last_error", C.c_char_p, []),
    ("llz_cuda_device_count", C.c_int, []),
    ("llz_cuda_build_info", C.c_char_p, []),
    ("llz_cuda_tune", C.c_int, [C.c_char_p, C.c_double]),
    ("llz_cuda_host_alloc", _vp, [C.c_size_t]),
    ("llz_cuda_host_free", None, [_vp]),
    ("llz_cuda_fir_bank_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_fir_bank_init_taps", _ul, [_dp, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_fir_bank_uninit", None, [_ul]),
    ("llz_cuda_fir_bank_flt_len", C.c_int, [_ul]),
    ("llz_cuda_fir_bank_set_algo", C.c_int, [_ul, C.c_int]),
    ("llz_cuda_fir_bank_get_algo", C.c_int, [_ul]),
    ("llz_cuda_fir_bank_set_fft_size", C.c_int, [_ul, C.c_int]),
    ("llz_cuda_fir_bank_block_len", _ll, [_ul]),
    ("llz_cuda_fir_bank_copy_taps", C.c_int, [_ul, _dp]),
    ("llz_cuda_fir_bank_reset", C.c_int, [_ul, _vp]),
    ("llz_cuda_fir_bank_set_history", C.c_int, [_ul, _vp, _ll, _vp]),
    ("llz_cuda_fir_bank_run", C.c_int, [_ul, _vp, _ll, _vp, _ll, _ll, _vp]),
    ("llz_cuda_fir_bank_flush", C.c_int, [_ul, _vp, _ll, _vp]),
    ("llz_cuda_fir_bank_run_host", C.c_int, [_ul, _vp, _ll, _vp, _ll, _ll]),
    ("llz_cuda_resample_bank_init", _ul, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_decimate_bank_init", _ul, [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_interp_bank_init", _ul, [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_resample_bank_uninit", None, [_ul]),
    ("llz_cuda_resample_bank_info", C.c_int, [_ul, C.POINTER(ResampleInfo)]),
    ("llz_cuda_resample_bank_copy_proto", C.c_int, [_ul, _dp]),
    ("llz_cuda_resample_bank_copy_bank", C.c_int, [_ul, _dp]),
    ("llz_cuda_resample_bank_out_len", _ll, [_ul, _ll]),
    ("llz_cuda_resample_bank_reset", C.c_int, [_ul, _vp]),
    ("llz_cuda_resample_bank_set_history", C.c_int, [_ul, _vp, _ll, _vp]),
    ("llz_cuda_resample_bank_run", C.c_int, [_ul, _vp, _ll, _ll, _vp, _ll, C.POINTER(_ll), _vp]),
    ("llz_cuda_resample_bank_run_host", C.c_int, [_ul, _vp, _ll, _ll, _vp, _ll, C.POINTER(_ll)]),
    ("llz_cuda_resample_bank_guard_count", _ll, [_ul]),
    ("llz_cuda_iir_bank_init", _ul, [C.c_int, _dp, C.c_int, _dp, C.c_int]),
    ("llz_cuda_iir_bank_uninit", None, [_ul]),
    ("llz_cuda_iir_bank_reset", C.c_int, [_ul, _vp]),
    ("llz_cuda_iir_bank_run", C.c_int, [_ul, _vp, _ll, _vp, _ll, _ll, _vp]),
    ("llz_iir_filter_init", _ul, [C.c_int, _dp, C.c_int, _dp]),
    ("llz_iir_filter_uninit", None, [_ul]),
    ("llz_iir_filter", C.c_int, [_ul, _dp, _dp, C.c_int]),
    ("llz_iir_filter_flush", C.c_int, [_ul, _dp]),
    ("llz_cuda_resample_bank_last_run", C.c_int, [_ul, C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    ("llz_cuda_resample_bank_run_pcm", C.c_int, [_ul, _vp, C.c_int, _ll, _vp, _ll, C.POINTER(_ll), _vp]),
    ("llz_cuda_resample_bank_run_pcm_host", C.c_int, [_ul, _vp, C.c_int, _ll, _vp, C.c_int, _ll, C.POINTER(_ll)]),
    ("llz_cuda_resample_bank_set_tiles", C.c_int, [_ul, C.c_int]),
    ("llz_cuda_resample_bank_set_guard_scale", C.c_int, [_ul, C.c_double]),
    ("llz_cuda_shard_channels", C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("llz_cuda_shard_fir_segments", C.c_int, [_ll, C.c_int, C.c_int, C.c_int, C.POINTER(Segment)]),
    ("llz_cuda_shard_fir_segments_aligned", C.c_int, [_ll, C.c_int, _ll, C.c_int, C.c_int, C.POINTER(Segment)]),
    ("llz_cuda_shard_resample_segments", C.c_int,
     [_ll, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Segment)]),
    ("llz_cuda_mgpu_unique_id", C.c_int, [C.c_char_p]),
    ("llz_cuda_mgpu_init_rank", _ul, [C.c_char_p, C.c_int, C.c_int]),
    ("llz_cuda_mgpu_init_all", _ul, [C.c_int, C.POINTER(C.c_int)]),
    ("llz_cuda_mgpu_uninit", None, [_ul]),
    ("llz_cuda_mgpu_world", C.c_int, [_ul]),
    ("llz_cuda_mgpu_local_count", C.c_int, [_ul]),
    ("llz_cuda_mgpu_local_rank", C.c_int, [_ul, C.c_int]),
    ("llz_cuda_mgpu_local_device", C.c_int, [_ul, C.c_int]),
    ("llz_cuda_mgpu_result_alloc", C.c_int, [_ul, C.c_int, C.c_size_t]),
    ("llz_cuda_mgpu_result_ptr", _vp, [_ul, C.c_int]),
    ("llz_cuda_mgpu_result_free", C.c_int, [_ul]),
    ("llz_cuda_mgpu_fir_init", _ul, [_ul, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_mgpu_resample_init", _ul, [_ul, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("llz_cuda_mgpu_job_uninit", None, [_ul]),
    ("llz_cuda_mgpu_job_bank", _ul, [_ul, C.c_int]),
    ("llz_cuda_mgpu_job_plan", C.c_int, [_ul, _ll, C.c_int, C.POINTER(Shard)]),
    ("llz_cuda_mgpu_job_out_len", _ll, [_ul, _ll]),
    ("llz_cuda_mgpu_job_run", C.c_int, [_ul, _ll, C.POINTER(_vp), C.POINTER(_ll), C.POINTER(_vp), C.POINTER(_ll), _ll,
                                        C.c_int, C.c_int, C.POINTER(_vp)]),
    ("llz_cuda_pcm_deinterleave", C.c_int, [_vp, C.c_int, C.c_int, _ll, _vp, C.c_int, _ll, _vp]),
    ("llz_cuda_pcm_interleave", C.c_int, [_vp, C.c_int, _ll, C.c_int, _ll, _vp, C.c_int, _vp]),
    ("llz_cuda_synth_lcg", C.c_int, [_vp, _ll, C.c_int, _ll, C.c_int, C.c_uint, _vp]),
    ("llz_cuda_synth_lcg_at", C.c_int, [_vp, _ll, C.c_int, _ll, _ll, C.c_int, C.c_uint, _vp]),
    ("llz_cuda_probe_fma", C.c_int, [C.c_int, _dp]),
]

EXPORTED = [s[0] for s in _SIGNATURES]

_lib = None


def lib() -> C.CDLL:
    """Load libllzfilter_cuda.so (raises if it is missing: there is no other implementation)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LlzError(f"{LIB_PATH} is not built; run `make -C {CSRC}` (or __graft_entry__.build()). "
                           "libllzfilter_cuda has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in _SIGNATURES:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        L._libc_free = C.CDLL(None).free
        L._libc_free.argtypes = [_vp]
        _lib = L
    return _lib


def last_error() -> str:
    return lib().llz_cuda_last_error().decode()


def tune(key: str, value: float):
    """llz_cuda_tune: process-wide measurement knob (see include/llz_cuda.h)"""
    _check(lib().llz_cuda_tune(key.encode(), float(value)), f"llz_cuda_tune({key})")


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise LlzError(f"{what} failed: {last_error()}")
    return rc


def _handle(h: int, what: str) -> int:
    if h == FAIL:
        raise LlzError(f"{what} failed: {last_error()}")
    return h


def _ptr(a) -> int | None:
    """Device / host address of a torch tensor, numpy array, int or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch.Tensor


# ---- host-side design (no GPU needed) ------------------------------------------------------------------
def window(N: int, win: int) -> np.ndarray:
    w = np.empty(N, dtype=np.float64)
    fn = (lib().llz_hamming, lib().llz_blackman, lib().llz_kaiser)[win]
    fn(w.ctypes.data_as(_dp), N)
    return w


def kaiser_beta(N: int, beta: float) -> np.ndarray:
    w = np.empty(N, dtype=np.float64)
    lib().llz_kaiser_beta(w.ctypes.data_as(_dp), N, beta)
    return w


def cof_num(win: int, ftrans: float, atten: float = 90.0) -> int:
    L = lib()
    if win == HAMMING:
        return L.llz_hamming_cof_num(ftrans)
    if win == BLACKMAN:
        return L.llz_blackman_cof_num(ftrans)
    return L.llz_kaiser_cof_num(ftrans, atten)


def fir_design(kind: int, N: int, fc1: float, fc2: float, win: int) -> np.ndarray:
    """llz_fir_{lpf,hpf,bandpass,bandstop}_cof: returns the malloc'd taps as an array (freed here)."""
    L = lib()
    hp = _dp()
    if kind == LPF:
        n = L.llz_fir_lpf_cof(C.byref(hp), N, fc1, win)
    elif kind == HPF:
        n = L.llz_fir_hpf_cof(C.byref(hp), N, fc1, win)
    elif kind == BPF:
        n = L.llz_fir_bandpass_cof(C.byref(hp), N, fc1, fc2, win)
    else:
        n = L.llz_fir_bandstop_cof(C.byref(hp), N, fc1, fc2, win)
    _check(n, "fir design")
    h = np.ctypeslib.as_array(hp, shape=(n,)).copy()
    L._libc_free(hp)
    return h


def conv(x: np.ndarray, newest: int, h: np.ndarray) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    px = C.cast(x.ctypes.data + 8 * newest, _dp)
    return lib().llz_conv(px, h.ctypes.data_as(_dp), len(h))


def shard_channels(n_channels: int, world: int, rank: int) -> tuple[int, int]:
    first, count = C.c_int(), C.c_int()
    _check(lib().llz_cuda_shard_channels(n_channels, world, rank, C.byref(first), C.byref(count)),
           "shard_channels")
    return first.value, count.value


def shard_fir_segments(n: int, flt_len: int, world: int, rank: int) -> Segment:
    seg = Segment()
    _check(lib().llz_cuda_shard_fir_segments(n, flt_len, world, rank, C.byref(seg)), "shard_fir_segments")
    return seg


def shard_fir_segments_aligned(n: int, flt_len: int, granule: int, world: int, rank: int) -> Segment:
    seg = Segment()
    _check(lib().llz_cuda_shard_fir_segments_aligned(n, flt_len, granule, world, rank, C.byref(seg)),
           "shard_fir_segments_aligned")
    return seg


def shard_resample_segments(n_in: int, L: int, M: int, taps_per_phase: int, frame_in: int, world: int,
                            rank: int) -> Segment:
    seg = Segment()
    _check(lib().llz_cuda_shard_resample_segments(n_in, L, M, taps_per_phase, frame_in, world, rank,
                                                  C.byref(seg)), "shard_resample_segments")
    return seg


# ---- drop-in mono handles (host buffers) ----------------------------------------------------------------
class FirFilter:
    """llz_fir_filter_*_init / llz_fir_filter / _flush / _uninit exactly as a C caller uses them."""

    def __init__(self, kind: int, frame_len: int, flt_len: int, fc1: float, fc2: float = 0.0,
                 win: int = HAMMING):
        L = lib()
        if kind == LPF:
            h = L.llz_fir_filter_lpf_init(frame_len, flt_len, fc1, win)
        elif kind == HPF:
            h = L.llz_fir_filter_hpf_init(frame_len, flt_len, fc1, win)
        elif kind == BPF:
            h = L.llz_fir_filter_bandpass_init(frame_len, flt_len, fc1, fc2, win)
        else:
            h = L.llz_fir_filter_bandstop_init(frame_len, flt_len, fc1, fc2, win)
        self.handle = _handle(h, "llz_fir_filter_*_init")
        self.frame_len = frame_len
        self.flt_len = L.llz_cuda_fir_bank_flt_len(self.handle)

    def filter(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        rc = lib().llz_fir_filter(self.handle, x.ctypes.data, y.ctypes.data, len(x))
        _check(rc, "llz_fir_filter")
        assert rc == len(x)
        return y

    def flush(self) -> np.ndarray:
        y = np.empty(self.flt_len - 1, dtype=np.float64)
        rc = _check(lib().llz_fir_filter_flush(self.handle, y.ctypes.data), "llz_fir_filter_flush")
        assert rc == self.flt_len - 1
        return y

    def close(self):
        if self.handle:
            lib().llz_fir_filter_uninit(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Resampler:
    """llz_{decimate,interp,resample_filter}_init + the matching frame call, as the CLI drives them."""

    def __init__(self, kind: int, L_: int, M: int, gain: float = 1.0, win: int = BLACKMAN):
        L = lib()
        self.kind = kind
        if kind == KIND_DECIMATE:
            h = L.llz_decimate_init(M, gain, win)
        elif kind == KIND_INTERP:
            h = L.llz_interp_init(L_, gain, win)
        else:
            h = L.llz_resample_filter_init(L_, M, gain, win)
        self.handle = _handle(h, "llz_*_init")
        self.bytes_in = L.llz_get_resample_framelen_bytes(self.handle)
        self.info = bank_info(self.handle)

    def frame(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.empty(self.info.num_out + 16, dtype=np.int16)
        osz = C.c_int(0)
        fn = (lib().llz_decimate, lib().llz_interp, lib().llz_resample)[self.kind]
        _check(fn(self.handle, x.ctypes.data, x.nbytes, out.ctypes.data, C.byref(osz)), "frame call")
        return out[:osz.value // 2].copy()

    def stream(self, x: np.ndarray) -> np.ndarray:
        n = self.bytes_in // 2
        assert len(x) % n == 0
        return np.concatenate([self.frame(x[i:i + n]) for i in range(0, len(x), n)])

    def close(self):
        if self.handle:
            lib().llz_resample_filter_uninit(self.handle)     # any uninit takes any kind (main.c:125)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bank_info(handle: int) -> ResampleInfo:
    info = ResampleInfo()
    _check(lib().llz_cuda_resample_bank_info(handle, C.byref(info)), "resample_bank_info")
    return info


# ---- banks (device pointers; torch tensors or raw addresses) ----------------------------------------------
class FirBank:
    def __init__(self, n_channels: int, dtype: int = F64, *, kind: int = LPF, flt_len: int = 0,
                 fc1: float = 0.0, fc2: float = 0.0, win: int = HAMMING, taps: np.ndarray | None = None,
                 algo: int = FIR_AUTO):
        L = lib()
        if taps is not None:
            taps = np.ascontiguousarray(taps, dtype=np.float64)
            h = L.llz_cuda_fir_bank_init_taps(taps.ctypes.data_as(_dp), len(taps), n_channels, dtype)
        else:
            h = L.llz_cuda_fir_bank_init(kind, flt_len, fc1, fc2, win, n_channels, dtype)
        self.handle = _handle(h, "llz_cuda_fir_bank_init")
        self.n_channels = n_channels
        self.dtype = dtype
        self.flt_len = L.llz_cuda_fir_bank_flt_len(self.handle)
        if algo != FIR_AUTO:
            self.set_algo(algo)

    def set_algo(self, algo: int):
        _check(lib().llz_cuda_fir_bank_set_algo(self.handle, algo), "llz_cuda_fir_bank_set_algo")

    def set_fft_size(self, size: int):
        _check(lib().llz_cuda_fir_bank_set_fft_size(self.handle, size), "llz_cuda_fir_bank_set_fft_size")

    @property
    def algo(self) -> int:
        """kernel family the next run uses: FIR_DIRECT or FIR_FFT"""
        return _check(lib().llz_cuda_fir_bank_get_algo(self.handle), "llz_cuda_fir_bank_get_algo")

    @property
    def block_len(self) -> int:
        """samples per work item of the kernel in use (1 for the direct form); see llz_cuda_fir_bank_block_len"""
        return _check(lib().llz_cuda_fir_bank_block_len(self.handle), "llz_cuda_fir_bank_block_len")

    def taps(self) -> np.ndarray:
        h = np.empty(self.flt_len, dtype=np.float64)
        _check(lib().llz_cuda_fir_bank_copy_taps(self.handle, h.ctypes.data_as(_dp)), "copy_taps")
        return h

    def reset(self, stream: int = 0):
        _check(lib().llz_cuda_fir_bank_reset(self.handle, stream), "fir_bank_reset")

    def set_history(self, d_hist, stride: int, stream: int = 0):
        _check(lib().llz_cuda_fir_bank_set_history(self.handle, _ptr(d_hist), stride, stream), "set_history")

    def run(self, d_in, in_stride: int, d_out, out_stride: int, n: int, stream: int = 0):
        _check(lib().llz_cuda_fir_bank_run(self.handle, _ptr(d_in), in_stride, _ptr(d_out), out_stride, n,
                                           stream), "llz_cuda_fir_bank_run")

    def flush(self, d_out, out_stride: int, stream: int = 0) -> int:
        return _check(lib().llz_cuda_fir_bank_flush(self.handle, _ptr(d_out), out_stride, stream),
                      "llz_cuda_fir_bank_flush")

    def run_host(self, h_in, in_stride: int, h_out, out_stride: int, n: int):
        _check(lib().llz_cuda_fir_bank_run_host(self.handle, _ptr(h_in), in_stride, _ptr(h_out), out_stride,
                                                n), "llz_cuda_fir_bank_run_host")

    def close(self):
        if self.handle:
            lib().llz_cuda_fir_bank_uninit(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IirBank:
    """n independent direct-form IIR filters (llz_cuda_iir_bank_*), planar doubles on the device"""

    def __init__(self, a, b, n_channels: int):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64) if b is not None else None
        h = lib().llz_cuda_iir_bank_init(len(a) - 1, a.ctypes.data_as(_dp), (len(b) - 1) if b is not None else 0,
                                         b.ctypes.data_as(_dp) if b is not None else None, n_channels)
        self.handle = _handle(h, "llz_cuda_iir_bank_init")
        self.n_channels = n_channels

    def run(self, d_x, x_stride: int, d_y, y_stride: int, n: int, stream: int = 0):
        _check(lib().llz_cuda_iir_bank_run(self.handle, _ptr(d_x), x_stride, _ptr(d_y), y_stride, n, stream), "llz_cuda_iir_bank_run")

    def reset(self, stream: int = 0):
        _check(lib().llz_cuda_iir_bank_reset(self.handle, stream), "llz_cuda_iir_bank_reset")

    def close(self):
        if self.handle:
            lib().llz_cuda_iir_bank_uninit(self.handle)
            self.handle = 0


class IirFilter:
    """the drop-in handle of llz_iir.h: host frames in, host frames out"""

    def __init__(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64) if b is not None else None
        self.N = (len(b) - 1) if b is not None else 0
        h = lib().llz_iir_filter_init(len(a) - 1, a.ctypes.data_as(_dp), self.N, b.ctypes.data_as(_dp) if b is not None else None)
        self.handle = _handle(h, "llz_iir_filter_init")

    def filter(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        if lib().llz_iir_filter(self.handle, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), len(x)) != len(x):
            raise LlzError("llz_iir_filter: " + last_error())
        return y

    def flush(self) -> np.ndarray:
        y = np.empty(max(self.N, 1), dtype=np.float64)
        n = lib().llz_iir_filter_flush(self.handle, y.ctypes.data_as(_dp))
        if n < 0:
            raise LlzError("llz_iir_filter_flush: " + last_error())
        return y[:n]

    def close(self):
        if self.handle:
            lib().llz_iir_filter_uninit(self.handle)
            self.handle = 0


class ResampleBank:
    def __init__(self, kind: int, L_: int, M: int, n_channels: int, *, gain: float = 1.0,
                 win: int = BLACKMAN, k_override: int = 0, acc: int = ACC_F64):
        L = lib()
        if kind == KIND_DECIMATE:
            h = L.llz_cuda_decimate_bank_init(M, gain, win, n_channels, acc)
        elif kind == KIND_INTERP:
            h = L.llz_cuda_interp_bank_init(L_, gain, win, n_channels, acc)
        else:
            h = L.llz_cuda_resample_bank_init(L_, M, gain, win, k_override, n_channels, acc)
        self.handle = _handle(h, "llz_cuda_*_bank_init")
        self.info = bank_info(self.handle)
        self.n_channels = n_channels

    def proto(self) -> np.ndarray:
        h = np.empty(self.info.n, dtype=np.float64)
        _check(lib().llz_cuda_resample_bank_copy_proto(self.handle, h.ctypes.data_as(_dp)), "copy_proto")
        return h

    def bank(self) -> np.ndarray:
        rows = self.info.M if self.info.kind == KIND_DECIMATE else self.info.L
        g = np.empty((rows, self.info.taps_per_phase), dtype=np.float64)
        _check(lib().llz_cuda_resample_bank_copy_bank(self.handle, g.ctypes.data_as(_dp)), "copy_bank")
        return g

    def out_len(self, n_in: int) -> int:
        return _check(lib().llz_cuda_resample_bank_out_len(self.handle, n_in), "out_len")

    def reset(self, stream: int = 0):
        _check(lib().llz_cuda_resample_bank_reset(self.handle, stream), "reset")

    def set_history(self, d_hist, stride: int, stream: int = 0):
        _check(lib().llz_cuda_resample_bank_set_history(self.handle, _ptr(d_hist), stride, stream),
               "set_history")

    def run(self, d_in, in_stride: int, n_in: int, d_out, out_stride: int, stream: int = 0) -> int:
        n_out = _ll(0)
        _check(lib().llz_cuda_resample_bank_run(self.handle, _ptr(d_in), in_stride, n_in, _ptr(d_out),
                                                out_stride, C.byref(n_out), stream),
               "llz_cuda_resample_bank_run")
        return n_out.value

    def run_host(self, h_in, in_stride: int, n_in: int, h_out, out_stride: int) -> int:
        n_out = _ll(0)
        _check(lib().llz_cuda_resample_bank_run_host(self.handle, _ptr(h_in), in_stride, n_in, _ptr(h_out),
                                                     out_stride, C.byref(n_out)),
               "llz_cuda_resample_bank_run_host")
        return n_out.value

    def run_pcm(self, d_frames, pcm_format: int, n_frames: int, d_out, out_stride: int, stream: int = 0) -> int:
        """interleaved PCM frames (device) in, planar int16 out; returns the outputs per channel"""
        n_out = _ll(0)
        _check(lib().llz_cuda_resample_bank_run_pcm(self.handle, _ptr(d_frames), pcm_format, n_frames, _ptr(d_out), out_stride,
                                                    C.byref(n_out), stream), "llz_cuda_resample_bank_run_pcm")
        return n_out.value

    def run_pcm_host(self, h_frames, in_format: int, n_frames: int, h_out, out_format: int, out_cap: int) -> int:
        """interleaved PCM frames (host) in and out; returns the output frames"""
        n_out = _ll(0)
        _check(lib().llz_cuda_resample_bank_run_pcm_host(self.handle, _ptr(h_frames), in_format, n_frames, _ptr(h_out), out_format,
                                                         out_cap, C.byref(n_out)), "llz_cuda_resample_bank_run_pcm_host")
        return n_out.value

    def last_run(self):
        """(kernel launches, name of the filtering kernel) of the last run call"""
        n, buf = C.c_int(0), C.create_string_buffer(96)
        _check(lib().llz_cuda_resample_bank_last_run(self.handle, C.byref(n), buf, 96), "last_run")
        return n.value, buf.value.decode()

    def guard_count(self) -> int:
        return _check(lib().llz_cuda_resample_bank_guard_count(self.handle), "guard_count")

    def set_tiles(self, tiles: int):
        _check(lib().llz_cuda_resample_bank_set_tiles(self.handle, tiles), "set_tiles")

    def set_guard_scale(self, scale: float):
        _check(lib().llz_cuda_resample_bank_set_guard_scale(self.handle, scale), "set_guard_scale")

    def close(self):
        if self.handle:
            lib().llz_cuda_resample_bank_uninit(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- multi-GPU contexts and sharded jobs ------------------------------------------------------------------
def mgpu_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().llz_cuda_mgpu_unique_id(buf), "llz_cuda_mgpu_unique_id")
    return buf.raw


class Mgpu:
    """llz_cuda_mgpu_init_all (one process, n GPUs) or llz_cuda_mgpu_init_rank (one process per GPU)."""

    def __init__(self, n_gpus: int = 0, devices=None, *, unique_id: bytes | None = None, world: int = 0, rank: int = 0):
        L = lib()
        if unique_id is not None:
            h = L.llz_cuda_mgpu_init_rank(unique_id, world, rank)
        else:
            arr = (C.c_int * n_gpus)(*devices) if devices is not None else None
            h = L.llz_cuda_mgpu_init_all(n_gpus, arr)
        self.handle = _handle(h, "llz_cuda_mgpu_init")
        self.world = L.llz_cuda_mgpu_world(self.handle)
        self.nlocal = L.llz_cuda_mgpu_local_count(self.handle)
        self.ranks = [L.llz_cuda_mgpu_local_rank(self.handle, i) for i in range(self.nlocal)]
        self.devices = [L.llz_cuda_mgpu_local_device(self.handle, i) for i in range(self.nlocal)]

    def result_alloc(self, root: int, nbytes: int):
        _check(lib().llz_cuda_mgpu_result_alloc(self.handle, root, nbytes), "llz_cuda_mgpu_result_alloc")

    def result_ptr(self, local_idx: int) -> int | None:
        return lib().llz_cuda_mgpu_result_ptr(self.handle, local_idx)

    def result_free(self):
        _check(lib().llz_cuda_mgpu_result_free(self.handle), "llz_cuda_mgpu_result_free")

    def close(self):
        if self.handle:
            lib().llz_cuda_mgpu_uninit(self.handle)
            self.handle = 0


class MgpuJob:
    """A sharded FIR or resampler job (llz_cuda_mgpu_fir_init / llz_cuda_mgpu_resample_init)."""

    def __init__(self, ctx: Mgpu, handle: int):
        self.ctx = ctx
        self.handle = _handle(handle, "llz_cuda_mgpu_*_init")

    @classmethod
    def fir(cls, ctx: Mgpu, n_channels: int, dtype: int, shard_mode: int, *, kind: int = LPF, flt_len: int,
            fc1: float, fc2: float = 0.0, win: int = HAMMING):
        return cls(ctx, lib().llz_cuda_mgpu_fir_init(ctx.handle, kind, flt_len, fc1, fc2, win, n_channels, dtype, shard_mode))

    @classmethod
    def resample(cls, ctx: Mgpu, L_: int, M: int, n_channels: int, shard_mode: int, *, gain: float = 1.0,
                 win: int = BLACKMAN, k_override: int = 0, acc: int = ACC_F64):
        return cls(ctx, lib().llz_cuda_mgpu_resample_init(ctx.handle, L_, M, gain, win, k_override, n_channels, acc, shard_mode))

    def bank(self, local_idx: int = 0) -> int:
        return _handle(lib().llz_cuda_mgpu_job_bank(self.handle, local_idx), "llz_cuda_mgpu_job_bank")

    def plan(self, n_total: int, rank: int) -> Shard:
        sh = Shard()
        _check(lib().llz_cuda_mgpu_job_plan(self.handle, n_total, rank, C.byref(sh)), "llz_cuda_mgpu_job_plan")
        return sh

    def out_len(self, n_total: int) -> int:
        return _check(lib().llz_cuda_mgpu_job_out_len(self.handle, n_total), "llz_cuda_mgpu_job_out_len")

    def run(self, n_total: int, d_in, in_stride, d_out, out_stride, result_stride: int = 0, gather: int = GATHER_NONE,
            chunks: int = 0, streams=None):
        n = self.ctx.nlocal
        vin = (_vp * n)(*[_ptr(a) for a in d_in])
        sin = (_ll * n)(*in_stride)
        vout = (_vp * n)(*[_ptr(a) for a in d_out]) if d_out is not None else None
        sout = (_ll * n)(*out_stride) if out_stride is not None else None
        vst = (_vp * n)(*streams) if streams is not None else None
        _check(lib().llz_cuda_mgpu_job_run(self.handle, n_total, vin, sin, vout, sout, result_stride, gather, chunks, vst),
               "llz_cuda_mgpu_job_run")

    def close(self):
        if self.handle:
            lib().llz_cuda_mgpu_job_uninit(self.handle)
            self.handle = 0


def pcm_deinterleave(d_frames, pcm_format: int, n_channels: int, n_frames: int, d_planar, planar_type: int,
                     planar_stride: int, stream: int = 0):
    _check(lib().llz_cuda_pcm_deinterleave(_ptr(d_frames), pcm_format, n_channels, n_frames, _ptr(d_planar),
                                           planar_type, planar_stride, stream), "llz_cuda_pcm_deinterleave")


def pcm_interleave(d_planar, planar_type: int, planar_stride: int, n_channels: int, n_frames: int, d_frames,
                   pcm_format: int, stream: int = 0):
    _check(lib().llz_cuda_pcm_interleave(_ptr(d_planar), planar_type, planar_stride, n_channels, n_frames,
                                         _ptr(d_frames), pcm_format, stream), "llz_cuda_pcm_interleave")


# ---- helpers for bench / tests -------------------------------------------------------------------------------
def synth_lcg(d_out, stride: int, n_channels: int, n: int, kind: int, seed0: int, stream: int = 0):
    _check(lib().llz_cuda_synth_lcg(_ptr(d_out), stride, n_channels, n, kind, seed0, stream), "synth_lcg")


def synth_lcg_at(d_out, stride: int, n_channels: int, first: int, n: int, kind: int, seed0: int, stream: int = 0):
    _check(lib().llz_cuda_synth_lcg_at(_ptr(d_out), stride, n_channels, first, n, kind, seed0, stream), "synth_lcg_at")


def probe_fma(dtype: int) -> float:
    v = C.c_double(0.0)
    _check(lib().llz_cuda_probe_fma(dtype, C.byref(v)), "probe_fma")
    return v.value


def host_alloc(nbytes: int, dtype) -> np.ndarray:
    """Page-locked numpy array (owned by the library allocator; free with host_free)."""
    p = lib().llz_cuda_host_alloc(nbytes)
    if not p:
        raise LlzError(f"llz_cuda_host_alloc failed: {last_error()}")
    buf = (C.c_char * nbytes).from_address(p)
    a = np.frombuffer(buf, dtype=dtype)
    return a


def host_free(a: np.ndarray):
    lib().llz_cuda_host_free(a.ctypes.data)
