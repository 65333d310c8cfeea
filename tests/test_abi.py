"""The C-ABI boundary: every function declared in include/*.h is exported by libllzfilter_cuda.so,
the library has no dependency on the oracle, and it fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the 30 symbols of the reference's libllzfilter for this path
REFERENCE_SYMBOLS = """llz_hamming llz_blackman llz_kaiser llz_kaiser_beta llz_kaiser_atten2beta llz_hamming_cof_num
llz_blackman_cof_num llz_kaiser_cof_num llz_fir_lpf_cof llz_fir_hpf_cof llz_fir_bandpass_cof llz_fir_bandstop_cof
llz_conv llz_fir_filter_lpf_init llz_fir_filter_hpf_init llz_fir_filter_bandpass_init llz_fir_filter_bandstop_init
llz_fir_filter_uninit llz_fir_filter llz_fir_filter_flush llz_decimate_init llz_decimate_uninit llz_decimate
llz_interp_init llz_interp_uninit llz_interp llz_resample_filter_init llz_resample_filter_uninit llz_resample
llz_get_resample_framelen_bytes""".split()


def declared_functions():
    names = set()
    for hdr in ("llz_fir.h", "llz_resample.h", "llz_iir.h", "llz_cuda.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
        for m in re.finditer(r"\b(llz_\w+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_every_declared_symbol_is_exported(zlib):
    lib = zlib.lib()
    decl = declared_functions()
    assert len(decl) >= 60
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    for name in REFERENCE_SYMBOLS:
        assert name in decl and hasattr(lib, name)
    assert sorted(zlib.EXPORTED) == decl, "python binding and headers disagree"


def test_headers_compile_as_c_and_match_reference_names():
    src = '#include "llz_fir.h"\n#include "llz_resample.h"\n#include "llz_iir.h"\n#include "llz_cuda.h"\n' \
          "int main(void){win_t w=BLACKMAN; return (int)w + HAMMING + KAISER + LLZ_RATIO_MAX - LLZ_DEFAULT_FRAMELEN;}\n"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    "-x", "c", "-"], input=src.encode(), check=True)


def test_library_does_not_link_the_oracle(zlib):
    out = subprocess.run(["ldd", zlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "llzref" not in out
    syms = subprocess.run(["nm", "-D", zlib.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in syms and "ref_llz" not in syms


def test_no_cpu_fallback_without_a_device(zlib):
    lib = zlib.lib()
    if lib.llz_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert lib.llz_fir_filter_lpf_init(1024, 127, 0.23, 0) == zlib.FAIL
    assert "no usable CUDA device" in zlib.last_error()
    assert lib.llz_resample_filter_init(160, 147, 1.0, 1) == zlib.FAIL
    assert lib.llz_cuda_fir_bank_init(0, 127, 0.23, 0.0, 0, 4, 0) == zlib.FAIL
    # the range check precedes everything else, as in the reference (llz_resample.c:375-378)
    assert lib.llz_resample_filter_init(17, 1, 1.0, 1) == zlib.FAIL
    assert "ratio" in zlib.last_error()


def test_missing_library_is_a_loud_error(monkeypatch, zlib):
    monkeypatch.setattr(zlib, "_lib", None)
    monkeypatch.setattr(zlib, "LIB_PATH", "/nonexistent/libllzfilter_cuda.so")
    with pytest.raises(zlib.LlzError):
        zlib.lib()
