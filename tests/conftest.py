import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import tetra_testlib as T
    return T.Oracle()


@pytest.fixture(scope="session")
def ref():
    import tetra_testlib as T
    T.ensure_oracle_built()
    if not T.have_ref():
        pytest.skip("oracle/_ref/libtetra_ref.so not built (needs /root/reference at build time)")
    return T.Ref()


@pytest.fixture(scope="session")
def emu():
    """the CUDA sources compiled for the CPU SIMT emulator (kernel + host logic checks)"""
    import tetra_testlib as T
    return T.B200(emulate=True)


@pytest.fixture(scope="session")
def gpu():
    """the real library on cuda:0"""
    import torch
    import tetra_testlib as T
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return T.B200(emulate=False)
