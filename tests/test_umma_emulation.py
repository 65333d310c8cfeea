"""Host emulation of the tcgen05 phase-bank kernel (llz_cuda_polybank_umma.cu).

llz_umma_tables.h -- replication, tile geometry, the K-major SWIZZLE_128B layout and the host-side builder of the int8
tap-digit planes -- compiles for the host; tests/cpu/umma_emulate.cpp reads the tables back with the kernel's index
algebra against a sample operand laid out as overlapping-row views of byte planes, accumulates the digit products per
weight class as the UTCIMMAs do, combines them, and compares with the long-double dot product of llz_resample's inner
loop (libllzfilter/llz_resample.c:586-592); then runs the kernel's integer finish step against the reference's
(llz_resample.c:594-601).  Pinned without a GPU: the replication is output-preserving, the table layout, that no 32-bit
accumulator can overflow, that `eps` is a true and tight bound, and that every output the integer finish gets wrong or
that lies inside the band is flagged for the guard.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llzlab_b200", "csrc")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "umma_emulate")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "cpu", "umma_emulate.cpp"), "-o", out],
                   check=True)
    return out


@pytest.mark.parametrize("L,M,Q,planes,gain", [(320, 147, 257, 5, 1.0), (160, 147, 45, 5, 1.0), (147, 160, 47, 5, 0.8),
                                               (1, 3, 134, 5, 1.0), (1, 4, 180, 5, -1.5), (4, 1, 45, 5, 1.0), (3, 2, 45, 5, 3.0),
                                               (513, 512, 40, 5, 1.0), (320, 147, 257, 3, 1.0), (1, 3, 134, 3, 1.0)])
def test_digit_planes_and_integer_finish(emu, L, M, Q, planes, gain):
    r = subprocess.run([emu, str(L), str(M), str(Q), str(planes), str(gain)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    worst, adversarial, acc_max, finish_bad, rep = r.stdout.split()
    assert float(worst) <= 1.0                      # the bound holds ...
    assert float(adversarial) >= 0.85               # ... and is not slack
    assert int(acc_max) < 2 ** 31 - 1               # exact s32 accumulation
    assert int(finish_bad) == 0                     # the integer finish is right wherever the guard does not look
    assert (int(rep) * M) % 16 == 0 and int(rep) * L >= 64
