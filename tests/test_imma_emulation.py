"""Host emulation of the integer tensor-core phase-bank kernel (llz_cuda_polybank_imma.cu).

llz_imma_tables.h -- the tile geometry and the host-side builder of the int8 tap-digit tables -- compiles for the host;
tests/cpu/imma_emulate.cpp reads the tables back with the kernel's index algebra, accumulates the digit products per
weight class in integers as the IMMAs do, combines them as the epilogue does and compares with the long-double dot
product of llz_resample's inner loop (libllzfilter/llz_resample.c:586-592).  What it pins without a GPU: the table
layout, that no 32-bit accumulator can overflow, and that the error bound `eps` the near-integer guard band is built
from is a true bound (never exceeded) and a tight one (an adversarial input reaches > 90 % of it).
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llzlab_b200", "csrc")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "imma_emulate")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "cpu", "imma_emulate.cpp"), "-o", out],
                   check=True)
    return out


@pytest.mark.parametrize("L,M,Q,planes", [(320, 147, 257, 5), (160, 147, 45, 5), (147, 160, 47, 5), (64, 32, 45, 5),
                                          (64, 2048, 1220, 5), (441, 320, 45, 5), (320, 147, 257, 3)])
def test_digit_planes_reproduce_the_dot_product_within_the_guard_bound(emu, L, M, Q, planes):
    r = subprocess.run([emu, str(L), str(M), str(Q), str(planes)], capture_output=True, text=True)
    worst, adversarial, acc_max = r.stdout.split()
    assert r.returncode == 0, r.stdout
    assert float(worst) <= 1.0                      # the bound holds ...
    assert float(adversarial) >= 0.85               # ... and is not slack
    assert int(acc_max) < 2 ** 31 - 1               # exact s32 accumulation
