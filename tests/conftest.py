import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def emu_lib():
    """The g++ -DSCB_EMU build of the kernel + driver sources (CI checker; tests/emu/README)."""
    import __graft_entry__ as ge

    if os.environ.get("SCB_EMU_LIBRARY"):  # e.g. the AddressSanitizer build made by tools/asan_emu.sh
        return os.environ["SCB_EMU_LIBRARY"]
    return ge.build_emu()


@pytest.fixture(scope="session")
def cuda_lib():
    import __graft_entry__ as ge

    if not os.path.exists(ge.LIB):
        ge.build_cuda()
    else:
        import torch

        if not torch.cuda.is_available():  # authoring container: keep the library in step with the sources (mtime check);
            ge.build_cuda()                # on the GPU box the shipped library is used as it is
    return ge.LIB
