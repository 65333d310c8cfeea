"""ctypes bindings for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Two checkers live here:

* ``port()``  -- ``oracle/_build/libllzoracle.so``: the C restatement in ``llz_oracle.c``.
* ``ref()``   -- ``oracle/_ref/libllzref.so``: the UNMODIFIED reference (``libllzfilter/llz_fir.c``
  + ``llz_resample.c``) compiled by ``oracle/Makefile`` with its symbols renamed ``ref_llz_*``.
  Present in the build container and shipped prebuilt to the GPU box; ``None`` if absent.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import
this package.  The product (``llzlab_b200`` / ``libllzfilter_cuda.so``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libllzoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libllzref.so")
REF_CLI = os.path.join(HERE, "_ref", "llz_resample_ref")

HAMMING, BLACKMAN, KAISER = 0, 1, 2
LPF, HPF, BPF, BSF = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_sp = C.POINTER(C.c_int16)


def build(ref: bool = True) -> None:
    """Compile the restatement (always) and, when /root/reference is mounted, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _Plan(C.Structure):
    _fields_ = [("L", C.c_int), ("M", C.c_int), ("n", C.c_int), ("rows", C.c_int),
                ("cols", C.c_int), ("num_in", C.c_int), ("num_out", C.c_int),
                ("h", _dp), ("bank", _dp)]


class Plan:
    """Python view of orc_plan_t (owns the C allocation)."""

    def __init__(self, raw: _Plan, lib):
        self._raw, self._lib = raw, lib
        self.L, self.M, self.n = raw.L, raw.M, raw.n
        self.rows, self.cols = raw.rows, raw.cols
        self.num_in, self.num_out = raw.num_in, raw.num_out
        self.h = np.ctypeslib.as_array(raw.h, shape=(raw.n,)).copy()
        self.bank = np.ctypeslib.as_array(raw.bank, shape=(raw.rows, raw.cols)).copy()

    def __del__(self):
        try:
            self._lib.orc_plan_free(C.byref(self._raw))
        except Exception:
            pass


class Port:
    """The C restatement (llz_oracle.c)."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.orc_window.argtypes = [_dp, C.c_int, C.c_int]
        L.orc_kaiser_beta.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_kaiser_atten2beta.argtypes = [C.c_double]
        L.orc_kaiser_atten2beta.restype = C.c_double
        L.orc_cof_num.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_fir_design.argtypes = [C.POINTER(_dp), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_fir_run.argtypes = [_dp, C.c_int, _dp, _dp, C.c_longlong, _dp, C.c_longlong]
        L.orc_fir_run.restype = None
        L.orc_iir_run.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, _dp]
        L.orc_iir_run.restype = None
        for f in (L.orc_resample_plan,):
            f.argtypes = [C.POINTER(_Plan), C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_decimate_plan.argtypes = [C.POINTER(_Plan), C.c_int, C.c_int]
        L.orc_interp_plan.argtypes = [C.POINTER(_Plan), C.c_int, C.c_int]
        L.orc_plan_free.argtypes = [C.POINTER(_Plan)]
        L.orc_plan_free.restype = None
        L.orc_resample_run.argtypes = [C.POINTER(_Plan), C.c_double, _sp, C.c_longlong, _sp,
                                       C.c_longlong, C.c_longlong]
        L.orc_resample_run.restype = None
        L.orc_decimate_run.argtypes = [C.POINTER(_Plan), C.c_double, _sp, C.c_longlong, _sp, C.c_longlong]
        L.orc_decimate_run.restype = None
        L.orc_interp_run.argtypes = [C.POINTER(_Plan), C.c_double, _sp, C.c_longlong, _sp]
        L.orc_interp_run.restype = None
        L.orc_fnv64.argtypes = [C.c_void_p, C.c_longlong]
        L.orc_fnv64.restype = C.c_uint64
        L.orc_lcg_s16.argtypes = [_sp, C.c_longlong, C.c_uint32]
        L.orc_lcg_s16.restype = None
        L.orc_lcg_f64.argtypes = [_dp, C.c_longlong, C.c_uint32]
        L.orc_lcg_f64.restype = None
        L.free = C.CDLL(None).free
        L.free.argtypes = [C.c_void_p]

    # -- design ------------------------------------------------------------------------------
    def window(self, N: int, win: int) -> np.ndarray:
        w = np.empty(N, dtype=np.float64)
        self.lib.orc_window(w.ctypes.data_as(_dp), N, win)
        return w

    def kaiser_beta(self, N: int, beta: float) -> np.ndarray:
        w = np.empty(N, dtype=np.float64)
        self.lib.orc_kaiser_beta(w.ctypes.data_as(_dp), N, beta)
        return w

    def atten2beta(self, atten: float) -> float:
        return self.lib.orc_kaiser_atten2beta(atten)

    def cof_num(self, win: int, ftrans: float, atten: float = 90.0) -> int:
        return self.lib.orc_cof_num(win, ftrans, atten)

    def fir_design(self, kind: int, N: int, fc1: float, fc2: float, win: int) -> np.ndarray:
        hp = _dp()
        n = self.lib.orc_fir_design(C.byref(hp), kind, N, fc1, fc2, win)
        h = np.ctypeslib.as_array(hp, shape=(n,)).copy()
        self.lib.free(hp)
        return h

    # -- data paths --------------------------------------------------------------------------
    def fir_run(self, h: np.ndarray, x: np.ndarray, n_out: int | None = None,
                hist: np.ndarray | None = None) -> np.ndarray:
        h = np.ascontiguousarray(h, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n_out = len(x) if n_out is None else n_out
        y = np.empty(n_out, dtype=np.float64)
        hp = None
        if hist is not None:
            hist = np.ascontiguousarray(hist, dtype=np.float64)
            assert len(hist) == len(h) - 1
            hp = hist.ctypes.data_as(_dp)
        self.lib.orc_fir_run(h.ctypes.data_as(_dp), len(h), hp, x.ctypes.data_as(_dp), len(x),
                             y.ctypes.data_as(_dp), n_out)
        return y

    def iir_run(self, a: np.ndarray, b: np.ndarray, x: np.ndarray | None, n: int | None = None, state=None):
        """llz_iir_filter over x (or n zeros when x is None: the flush); state = (xs, ys) arrays updated in place"""
        a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        M, N = len(a) - 1, len(b) - 1
        if state is None:
            state = (np.zeros(N + 1), np.zeros(M + 1))
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float64); n = len(x)
        y = np.empty(n, dtype=np.float64)
        self.lib.orc_iir_run(M, a.ctypes.data_as(_dp), N, b.ctypes.data_as(_dp), state[0].ctypes.data_as(_dp),
                             state[1].ctypes.data_as(_dp), x.ctypes.data_as(_dp) if x is not None else None, n, y.ctypes.data_as(_dp))
        return y

    def resample_plan(self, L: int, M: int, win: int, k_override: int = 0) -> Plan | None:
        raw = _Plan()
        if self.lib.orc_resample_plan(C.byref(raw), L, M, win, k_override) != 0:
            return None
        return Plan(raw, self.lib)

    def decimate_plan(self, M: int, win: int) -> Plan | None:
        raw = _Plan()
        if self.lib.orc_decimate_plan(C.byref(raw), M, win) != 0:
            return None
        return Plan(raw, self.lib)

    def interp_plan(self, L: int, win: int) -> Plan | None:
        raw = _Plan()
        if self.lib.orc_interp_plan(C.byref(raw), L, win) != 0:
            return None
        return Plan(raw, self.lib)

    def resample_run(self, plan: Plan, gain: float, x: np.ndarray, n_out: int, m0: int = 0) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.empty(n_out, dtype=np.int16)
        self.lib.orc_resample_run(C.byref(plan._raw), gain, x.ctypes.data_as(_sp), len(x),
                                  y.ctypes.data_as(_sp), m0, n_out)
        return y

    def decimate_run(self, plan: Plan, gain: float, x: np.ndarray, n_out: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.empty(n_out, dtype=np.int16)
        self.lib.orc_decimate_run(C.byref(plan._raw), gain, x.ctypes.data_as(_sp), len(x),
                                  y.ctypes.data_as(_sp), n_out)
        return y

    def interp_run(self, plan: Plan, gain: float, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.empty(len(x) * plan.L, dtype=np.int16)
        self.lib.orc_interp_run(C.byref(plan._raw), gain, x.ctypes.data_as(_sp), len(x),
                                y.ctypes.data_as(_sp))
        return y

    # -- helpers -----------------------------------------------------------------------------
    def fnv64(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        return int(self.lib.orc_fnv64(a.ctypes.data, a.nbytes))

    def lcg_s16(self, n: int, seed: int) -> np.ndarray:
        x = np.empty(n, dtype=np.int16)
        self.lib.orc_lcg_s16(x.ctypes.data_as(_sp), n, seed)
        return x

    def lcg_f64(self, n: int, seed: int) -> np.ndarray:
        x = np.empty(n, dtype=np.float64)
        self.lib.orc_lcg_f64(x.ctypes.data_as(_dp), n, seed)
        return x


class Ref:
    """The unmodified reference, frame-streaming exactly as example/llz_resample/main.c drives it."""

    def __init__(self, path: str = REF_SO):
        L = self.lib = C.CDLL(path)
        ul = C.c_ulong
        for name in ("lpf", "hpf"):
            f = getattr(L, f"ref_llz_fir_filter_{name}_init")
            f.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
            f.restype = ul
        for name in ("bandpass", "bandstop"):
            f = getattr(L, f"ref_llz_fir_filter_{name}_init")
            f.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
            f.restype = ul
        L.ref_llz_fir_filter_uninit.argtypes = [ul]
        L.ref_llz_fir_filter_uninit.restype = None
        L.ref_llz_fir_filter.argtypes = [ul, _dp, _dp, C.c_int]
        L.ref_llz_fir_filter_flush.argtypes = [ul, _dp]
        for name in ("hamming", "blackman", "kaiser"):
            getattr(L, f"ref_llz_{name}").argtypes = [_dp, C.c_int]
        L.ref_llz_kaiser_beta.argtypes = [_dp, C.c_int, C.c_double]
        L.ref_llz_kaiser_atten2beta.argtypes = [C.c_double]
        L.ref_llz_kaiser_atten2beta.restype = C.c_double
        L.ref_llz_hamming_cof_num.argtypes = [C.c_double]
        L.ref_llz_blackman_cof_num.argtypes = [C.c_double]
        L.ref_llz_kaiser_cof_num.argtypes = [C.c_double, C.c_double]
        L.ref_llz_fir_lpf_cof.argtypes = [C.POINTER(_dp), C.c_int, C.c_double, C.c_int]
        L.ref_llz_fir_hpf_cof.argtypes = [C.POINTER(_dp), C.c_int, C.c_double, C.c_int]
        L.ref_llz_fir_bandpass_cof.argtypes = [C.POINTER(_dp), C.c_int, C.c_double, C.c_double, C.c_int]
        L.ref_llz_fir_bandstop_cof.argtypes = [C.POINTER(_dp), C.c_int, C.c_double, C.c_double, C.c_int]
        L.ref_llz_conv.argtypes = [_dp, _dp, C.c_int]
        L.ref_llz_conv.restype = C.c_double
        L.ref_llz_decimate_init.argtypes = [C.c_int, C.c_double, C.c_int]
        L.ref_llz_decimate_init.restype = ul
        L.ref_llz_interp_init.argtypes = [C.c_int, C.c_double, C.c_int]
        L.ref_llz_interp_init.restype = ul
        L.ref_llz_resample_filter_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.ref_llz_resample_filter_init.restype = ul
        for name in ("decimate_uninit", "interp_uninit", "resample_filter_uninit"):
            f = getattr(L, f"ref_llz_{name}")
            f.argtypes = [ul]
            f.restype = None
        L.ref_llz_get_resample_framelen_bytes.argtypes = [ul]
        for name in ("decimate", "interp", "resample"):
            getattr(L, f"ref_llz_{name}").argtypes = [ul, C.c_void_p, C.c_int, C.c_void_p,
                                                      C.POINTER(C.c_int)]
        if hasattr(L, "ref_llz_iir_filter_init"):
            L.ref_llz_iir_filter_init.argtypes = [C.c_int, _dp, C.c_int, _dp]
            L.ref_llz_iir_filter_init.restype = ul
            L.ref_llz_iir_filter_uninit.argtypes = [ul]
            L.ref_llz_iir_filter_uninit.restype = None
            L.ref_llz_iir_filter.argtypes = [ul, _dp, _dp, C.c_int]
            L.ref_llz_iir_filter_flush.argtypes = [ul, _dp]
        self._free = C.CDLL(None).free
        self._free.argtypes = [C.c_void_p]

    def iir_stream(self, a: np.ndarray, b: np.ndarray, x: np.ndarray, frame: int = 1024, flush: bool = True) -> np.ndarray:
        """the reference's llz_iir_filter frame by frame (+ llz_iir_filter_flush)"""
        a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        M, N = len(a) - 1, len(b) - 1
        h = self.lib.ref_llz_iir_filter_init(M, a.ctypes.data_as(_dp), N, b.ctypes.data_as(_dp))
        out = []
        for t0 in range(0, len(x), frame):
            xi = np.ascontiguousarray(x[t0:t0 + frame]); yo = np.empty(len(xi))
            self.lib.ref_llz_iir_filter(h, xi.ctypes.data_as(_dp), yo.ctypes.data_as(_dp), len(xi))
            out.append(yo)
        if flush and N > 0:
            yo = np.empty(N)
            self.lib.ref_llz_iir_filter_flush(h, yo.ctypes.data_as(_dp))
            out.append(yo)
        self.lib.ref_llz_iir_filter_uninit(h)
        return np.concatenate(out) if out else np.empty(0)

    # -- design ------------------------------------------------------------------------------
    def window(self, N: int, win: int) -> np.ndarray:
        w = np.empty(N, dtype=np.float64)
        fn = (self.lib.ref_llz_hamming, self.lib.ref_llz_blackman, self.lib.ref_llz_kaiser)[win]
        fn(w.ctypes.data_as(_dp), N)
        return w

    def kaiser_beta(self, N: int, beta: float) -> np.ndarray:
        w = np.empty(N, dtype=np.float64)
        self.lib.ref_llz_kaiser_beta(w.ctypes.data_as(_dp), N, beta)
        return w

    def atten2beta(self, atten: float) -> float:
        return self.lib.ref_llz_kaiser_atten2beta(atten)

    def cof_num(self, win: int, ftrans: float, atten: float = 90.0) -> int:
        if win == HAMMING:
            return self.lib.ref_llz_hamming_cof_num(ftrans)
        if win == BLACKMAN:
            return self.lib.ref_llz_blackman_cof_num(ftrans)
        return self.lib.ref_llz_kaiser_cof_num(ftrans, atten)

    def fir_design(self, kind: int, N: int, fc1: float, fc2: float, win: int) -> np.ndarray:
        hp = _dp()
        if kind == LPF:
            n = self.lib.ref_llz_fir_lpf_cof(C.byref(hp), N, fc1, win)
        elif kind == HPF:
            n = self.lib.ref_llz_fir_hpf_cof(C.byref(hp), N, fc1, win)
        elif kind == BPF:
            n = self.lib.ref_llz_fir_bandpass_cof(C.byref(hp), N, fc1, fc2, win)
        else:
            n = self.lib.ref_llz_fir_bandstop_cof(C.byref(hp), N, fc1, fc2, win)
        h = np.ctypeslib.as_array(hp, shape=(n,)).copy()
        self._free(hp)
        return h

    # -- streaming data paths (full frames, as the reference requires) -------------------------
    def fir_stream(self, kind: int, N: int, fc1: float, fc2: float, win: int, x: np.ndarray,
                   frame_len: int, flush: bool = False) -> np.ndarray:
        """llz_fir_filter over whole frames of ``frame_len`` (+ optional flush tail)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert len(x) % frame_len == 0
        L = self.lib
        if kind == LPF:
            hd = L.ref_llz_fir_filter_lpf_init(frame_len, N, fc1, win)
        elif kind == HPF:
            hd = L.ref_llz_fir_filter_hpf_init(frame_len, N, fc1, win)
        elif kind == BPF:
            hd = L.ref_llz_fir_filter_bandpass_init(frame_len, N, fc1, fc2, win)
        else:
            hd = L.ref_llz_fir_filter_bandstop_init(frame_len, N, fc1, fc2, win)
        n_eff = N if (kind == LPF or N & 1) else N + 1
        y = np.empty(len(x) + (n_eff - 1 if flush else 0), dtype=np.float64)
        for f in range(len(x) // frame_len):
            xin = x[f * frame_len:(f + 1) * frame_len]
            yout = y[f * frame_len:(f + 1) * frame_len]
            L.ref_llz_fir_filter(hd, xin.ctypes.data_as(_dp), yout.ctypes.data_as(_dp), frame_len)
        if flush:
            assert frame_len >= n_eff - 1, "quirk F2: flush over-reads when N-1 > frame_len"
            tail = y[len(x):]
            L.ref_llz_fir_filter_flush(hd, tail.ctypes.data_as(_dp))
        L.ref_llz_fir_filter_uninit(hd)
        return y

    def _stream_s16(self, init, run, uninit, x: np.ndarray, pad_after: int = 0):
        x = np.ascontiguousarray(x, dtype=np.int16)
        L = self.lib
        hd = init()
        if hd == C.c_ulong(-1).value:
            return None
        bytes_in = L.ref_llz_get_resample_framelen_bytes(hd)
        num_in = bytes_in // 2
        assert len(x) % num_in == 0, (len(x), num_in)
        outs = []
        osz = C.c_int(0)
        # generous output buffer: ratio <= 16
        obuf = np.zeros(num_in * 16 + 16, dtype=np.int16)
        ibuf = np.zeros(num_in + pad_after, dtype=np.int16)
        for f in range(len(x) // num_in):
            ibuf[:num_in] = x[f * num_in:(f + 1) * num_in]
            run(hd, ibuf.ctypes.data, bytes_in, obuf.ctypes.data, C.byref(osz))
            outs.append(obuf[:osz.value // 2].copy())
        uninit(hd)
        return np.concatenate(outs) if outs else np.zeros(0, np.int16)

    def resample_stream(self, L_: int, M: int, gain: float, win: int, x: np.ndarray):
        L = self.lib
        return self._stream_s16(lambda: L.ref_llz_resample_filter_init(L_, M, gain, win),
                                L.ref_llz_resample, L.ref_llz_resample_filter_uninit, x)

    def decimate_stream(self, M: int, gain: float, win: int, x: np.ndarray):
        L = self.lib
        return self._stream_s16(lambda: L.ref_llz_decimate_init(M, gain, win),
                                L.ref_llz_decimate, L.ref_llz_decimate_uninit, x)

    def interp_stream(self, L_: int, gain: float, win: int, x: np.ndarray, pad: int = 4096):
        """llz_interp reads K-1 samples past each frame (quirk R4); the harness supplies zeros."""
        L = self.lib
        return self._stream_s16(lambda: L.ref_llz_interp_init(L_, gain, win),
                                L.ref_llz_interp, L.ref_llz_interp_uninit, x, pad_after=pad)

    def resample_framelen(self, L_: int, M: int, win: int) -> int:
        hd = self.lib.ref_llz_resample_filter_init(L_, M, 1.0, win)
        n = self.lib.ref_llz_get_resample_framelen_bytes(hd) // 2
        self.lib.ref_llz_resample_filter_uninit(hd)
        return n


_port = None
_ref = None


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref() -> Ref | None:
    """The compiled reference, or None when oracle/_ref was never built (and cannot be)."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = Ref()
    return _ref
