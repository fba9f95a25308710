import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


def pytest_sessionfinish(session, exitstatus):
    """GPU runs: persist the achieved-error table of every oracle comparison (also when a test failed)."""
    try:
        import torch
        import parity_common as pc
        if torch.cuda.is_available() and pc.PARITY_RECORD:
            pc.dump_record(os.path.join(ROOT, "gpurun_out", "parity_r2.json"))
    except Exception:      # noqa: BLE001 -- bookkeeping must never turn a green run red
        pass


@pytest.fixture(scope="session")
def emu_lib():
    """The kernel-emulation build of csrc (g++, fibers).  TEST-ONLY -- see tests/emu/cuda_emu.h."""
    import ctypes
    import __graft_entry__ as ge
    from ecog2txt_b200 import _lib
    path = ge.build_emu()
    return _lib.bind(ctypes.CDLL(path))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU; fails (does not skip) when it cannot be used."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import __graft_entry__ as ge
    from ecog2txt_b200 import _lib
    if not os.path.exists(ge.LIB):
        ge.build_cuda()
    return _lib.load()
