"""Kernel-against-kernel checks on the emulator build (g++ -DSCB_EMU): the re-scheduled kernels must reproduce, bit for bit, the kernels
the parity tests pin against the oracle.
  rhs_fold2_kernel       (packed 16-bit lanes)        == rhs_fold_kernel<2>   digit planes, byte for byte
  tri_solve_smem_kernel  (column tile in shared mem.) == tri_solve_kernel     Ct (float) and Y64 (float64), bit for bit
  tri_lowproj2_kernel    (no shared-memory atomics)   ~= tri_lowproj_kernel   W within 2e-9 of its largest entry (float64 sums reordered)
The same equalities are asserted end to end on the GPU by tools/ab_select.py (output bytes of the whole clone, per variant)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "seamlesscloneoptimization_b200", "csrc")


@pytest.mark.parametrize("name", ["test_rhs_fold2", "test_tri_smem", "test_lowproj2"])
def test_kernel_equals_its_reference_kernel(name):
    exe = os.path.join(EMU, "_build", name)
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-DSCB_EMU", "-x", "c++", "-I" + EMU, "-I" + CSRC, "-ffp-contract=off", "-Wno-unknown-pragmas",
                    os.path.join(EMU, name + ".cpp"), "-o", exe, "-lpthread"], check=True, cwd=ROOT)
    r = subprocess.run([exe], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 failed" in r.stdout, r.stdout[-500:]
