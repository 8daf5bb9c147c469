"""CPU: the CUDA C-ABI library loads without a GPU, exports every symbol include/scb.h declares,
and fails loudly (no fallback) when asked to compute without a device."""
import os
import re

import pytest

from seamlesscloneoptimization_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "scb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b((?:scb|my_seamlessclone_api_imp)_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_cuda_library_exports_every_declared_symbol(cuda_lib):
    lib = capi.load(cuda_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_a_device(cuda_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import seamlesscloneoptimization_b200 as scb

    with pytest.raises(scb.ScbError) as e:
        scb.Context(0, lib_path=cuda_lib)
    assert e.value.code in (capi.SCB_ERR_NO_DEVICE, capi.SCB_ERR_CUDA)


def test_missing_library_is_an_import_error(tmp_path):
    with pytest.raises(ImportError):
        capi.load(str(tmp_path / "nope.so"))


def test_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "seamlesscloneoptimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "emu_cuda.h" not in text.replace('#include "emu_cuda.h"', "") or f == "scb_platform.h", f
