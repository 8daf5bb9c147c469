import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: tens of seconds of CPU oracle work (the 8K comparison against cv2)")
    config.addinivalue_line("markers", "multigpu: spawns torchrun over >= 2 GPUs (skipped on a 1-GPU box)")


@pytest.fixture(scope="session")
def emu_lib():
    """The g++ -DSCB_EMU build of the kernel + driver sources (CI checker; tests/emu/README)."""
    import __graft_entry__ as ge

    if os.environ.get("SCB_EMU_LIBRARY"):  # e.g. the AddressSanitizer build made by tools/asan_emu.sh
        return os.environ["SCB_EMU_LIBRARY"]
    return ge.build_emu()


@pytest.fixture(scope="session")
def cuda_lib():
    """libscb.so, guaranteed to be built from the sources in the tree: the library carries a hash of its sources
    (scb_source_hash) which is compared with the tree here -- on the GPU box too, where file times mean nothing after
    the snapshot copy -- and the library is rebuilt when they differ."""
    import __graft_entry__ as ge

    if os.environ.get("SCB_TEST_LIBRARY"):  # a compile-time variant made by tools/build_variant.sh (A/B runs only)
        return os.environ["SCB_TEST_LIBRARY"]
    return ge.build_cuda()
