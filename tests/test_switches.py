"""The kernel switches: every generation of a hot kernel must give the bytes of the default one.
The switches are read once per process, so each setting runs the drop-in call in its own interpreter -- against the emulator build on
CPU, against libscb.so on a B200 (-m gpu).  This also keeps the kernels that are no longer the default (rhs_fold_kernel<2>,
tri_solve_kernel for whole solves) and the opt-in ones (tri_solve_smem2_kernel, i8_gemm_p2kernel, tri_lowproj2_kernel) under test on
the GPU.  tools/ab_select.py applies the same gate at the bench sizes before a default is changed."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, json, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import seamlesscloneoptimization_b200 as scb
from oracle import seamless_oracle as so
out = {}
with scb.Context(0) as ctx:
    out["variants"] = ctx.lib.scb_kernel_variants().decode()
    for name, seed in json.loads(sys.argv[2]):
        src, dst, mask, p = so.make_config(name, seed)
        blend = ctx.seamless_clone(src, dst, mask, p)
        np.save(sys.argv[3] + f"/{name}_{seed}.npy", blend)
        out[f"{name}_{seed}"] = hashlib.md5(np.ascontiguousarray(blend)).hexdigest()
print(json.dumps(out))
"""

# (environment, bit-exact against the defaults?)
SWITCHES = [
    ({"SCB_RHS_FOLD": "1"}, True),        # rhs_fold_kernel<2>: the previous stencil
    ({"SCB_TRI_SMEM": "0"}, True),        # tri_solve_kernel: the column solve on global memory
    ({"SCB_TRI_SMEM": "2"}, True),        # tri_solve_smem2_kernel
    ({"SCB_I8_PERSISTENT": "2"}, True),   # i8_gemm_p2kernel, both passes
    ({"SCB_I8_PERSISTENT": "3"}, True),   # i8_gemm_p2kernel, forward pass only
    ({"SCB_LOWPROJ": "2"}, False),        # tri_lowproj2_kernel: float64 sums in another order
]


def run(lib, env_extra, cases, outdir):
    env = dict(os.environ, SCB_LIBRARY=lib)
    for k in ("SCB_RHS_FOLD", "SCB_TRI_SMEM", "SCB_I8_PERSISTENT", "SCB_LOWPROJ", "SCB_I8_KB"):
        env.pop(k, None)
    env.update(env_extra)
    os.makedirs(outdir, exist_ok=True)
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, json.dumps(cases), str(outdir)], env=env, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("backend", ["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_every_kernel_generation_gives_the_default_bytes(tmp_path, request, backend):
    lib = request.getfixturevalue("emu_lib" if backend == "emu" else "cuda_lib")
    cases = [["small", 2]] if backend == "emu" else [["small", 2], ["cfg1", 0]]
    base = run(lib, {}, cases, tmp_path / "base")
    assert "rhs_fold=2" in base["variants"] and "tri_smem=1" in base["variants"], base["variants"]
    for env_extra, exact in SWITCHES:
        got = run(lib, env_extra, cases, tmp_path / "v")
        k, v = next(iter(env_extra.items()))
        if backend == "cuda" or not k.startswith("SCB_I8_"):  # (the emulator build has one stand-in for the tensor-core kernels)
            assert f"{k[4:].lower()}={v}" in got["variants"], (env_extra, got["variants"])
        for name, seed in cases:
            key = f"{name}_{seed}"
            if exact:
                assert got[key] == base[key], (env_extra, key)
            else:
                a, b = np.load(tmp_path / "base" / f"{key}.npy").astype(np.int16), np.load(tmp_path / "v" / f"{key}.npy").astype(np.int16)
                d = np.abs(a - b)
                assert d.max() <= 1 and (d != 0).mean() < 1e-4, (env_extra, key, int(d.max()), float((d != 0).mean()))
