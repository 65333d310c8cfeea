"""Host emulation of the overlap-save FIR kernels (llz_cuda_fir_fft.cu, llz_cuda_fir_fft8k.cu).

llz_fft32.cuh -- the register-resident DFT-32 / DFT-8 networks with folded twiddles and the host-side table
builders -- compiles for the host; tests/cpu/*.cpp run the kernels' exact sequence of transforms, table lookups and
exchanges lane by lane and compare two blocks of outputs with the direct sum  y[t] = sum_i h[i] x[t-i]
(libllzfilter/llz_fir.c:411-426).  This pins the algorithm (index maps, table layouts, twiddle signs) without a GPU;
the device parity tests are in test_gpu_fir.py.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llzlab_b200", "csrc")


def build(tmp_path_factory, name):
    out = str(tmp_path_factory.mktemp("emu") / name)
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "cpu", name + ".cpp"), "-o", out],
                   check=True)
    return out


@pytest.fixture(scope="module")
def emu1k(tmp_path_factory):
    return build(tmp_path_factory, "fft_emulate")


@pytest.fixture(scope="module")
def emu8k(tmp_path_factory):
    return build(tmp_path_factory, "fft8k_emulate")


@pytest.mark.parametrize("ntaps", [1, 2, 48, 127, 513, 897])
@pytest.mark.parametrize("f32", [0, 1])
def test_warp_level_1024_point_overlap_save(emu1k, ntaps, f32):
    r = subprocess.run([emu1k, str(ntaps), str(f32)], capture_output=True, text=True)
    err = float(r.stdout.strip())
    assert r.returncode == 0 and err < (2e-6 if f32 else 1e-14), (ntaps, f32, err)


@pytest.mark.parametrize("ntaps", [898, 4095, 6145])
@pytest.mark.parametrize("f32", [0, 1])
def test_cta_level_8192_point_overlap_save(emu8k, ntaps, f32):
    r = subprocess.run([emu8k, str(ntaps), str(f32)], capture_output=True, text=True)
    err = float(r.stdout.strip())
    assert r.returncode == 0 and err < (2e-6 if f32 else 1e-14), (ntaps, f32, err)


@pytest.fixture(scope="module")
def emu16k(tmp_path_factory):
    return build(tmp_path_factory, "fft16k_emulate")


@pytest.mark.parametrize("ntaps", [2305, 3700, 4095, 12289])
@pytest.mark.parametrize("f32", [0, 1])
def test_16384_point_overlap_save_in_two_rounds(emu16k, ntaps, f32):
    r = subprocess.run([emu16k, str(ntaps), str(f32)], capture_output=True, text=True)
    err = float(r.stdout.strip())
    assert r.returncode == 0 and err < (2e-6 if f32 else 1e-14), (ntaps, f32, err)
