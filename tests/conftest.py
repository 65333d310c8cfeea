"""Shared fixtures.  `-m "not gpu"` runs everywhere; `-m gpu` needs a B200 and calls the C-ABI."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
WIN = {"HAMMING": 0, "BLACKMAN": 1, "KAISER": 2}
KIND = {"LPF": 0, "HPF": 1, "BPF": 2, "BSF": 3}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def zlib():
    """The product library (built on demand; host-side functions work without a GPU)."""
    import llzlab_b200 as z
    if not os.path.exists(z.LIB_PATH):
        z.build()
    z.lib()
    return z


@pytest.fixture(scope="session")
def port():
    """CPU restatement of the reference algorithm (oracle/llz_oracle.c)."""
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (skips when it is not available)."""
    import oracle
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref/libllzref.so not built (reference tree not mounted)")
    return r


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def vectors():
    return np.load(os.path.join(GOLDEN, "vectors.npz"))


@pytest.fixture(scope="session")
def cuda(zlib):
    """torch + a CUDA device, for the gpu-marked tests."""
    import torch
    if not torch.cuda.is_available() or zlib.lib().llz_cuda_device_count() < 1:
        pytest.fail("gpu-marked test run without a usable CUDA device")
    torch.cuda.set_device(0)
    return torch


@pytest.fixture
def small_pipe_slots(zlib):
    """4 MiB staging slots for the *_run_host pipelines (llz_cuda_tune), restored afterwards"""
    zlib.tune("pipe_slot_mib", 4)
    yield
    zlib.tune("pipe_slot_mib", 64)
