"""The reference-shaped command line tools (tools/seamless_clone_cli.py = seamlessClone_main / seamlessClone_OpenCV,
tools/compare_vs.py = compare/vs.py), run against the emulator build on CPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("backend", ["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_cli_and_compare_report(tmp_path, request, backend):
    """Runs twice: against the emulator build on CPU and against libscb.so on a B200 (-m gpu)."""
    lib = request.getfixturevalue("emu_lib" if backend == "emu" else "cuda_lib")
    z = common.load_golden("small_ellipse")
    for k in ("src", "dst", "mask"):
        np.save(tmp_path / f"{k}.npy", z[k])
    px, py = (int(v) for v in z["p"])
    env = dict(os.environ, SCB_LIBRARY=lib)
    out = tmp_path / "blend.npy"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "seamless_clone_cli.py"), str(tmp_path / "src.npy"), str(tmp_path / "dst.npy"),
                        str(tmp_path / "mask.npy"), str(px), str(py), "--out", str(out)], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert "Compute stage performance time=" in r.stdout  # the reference's timing line (seamlessClone_imp.cu:343-346)
    blend = np.load(out)
    x, y, w, h, rx, ry = (int(v) for v in z["geom"])
    expect = z["dst"].copy()
    expect[ry : ry + h, rx : rx + w] = z["blend_roi"]
    np.save(tmp_path / "expect.npy", expect)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_vs.py"), str(out), str(tmp_path / "expect.npy")], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert "diff sum" in r.stdout
    d = np.abs(blend.astype(int) - expect)
    assert d.max() <= 1 and (d != 0).sum() <= common.allowed_mismatches(3 * (w - 2) * (h - 2))


def test_ab_selector_bookkeeping(tmp_path):
    """tools/ab_select.py with synthetic worker results (AB_FAKE=1): a variant whose bytes differ from the baseline's is never selected,
    one winner per switch group is combined, and the on-top round extends the selection."""
    import json

    env = dict(os.environ, AB_FAKE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_select.py"), "--out", str(tmp_path), "--workloads", "cfg2", "cfg1"], env=env, capture_output=True,
                       text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    sel = json.load(open(tmp_path / "selected.json"))
    res = json.load(open(tmp_path / "results.json"))
    assert res["i8_p2"]["_correct"] is False and "i8_p2" not in sel["winners"]
    assert sel["combined_ok"] and {"tri_smem2", "lowproj2"} <= set(sel["winners"])
    assert sel["env"]["SCB_TRI_SMEM"] == "2" and sel["env"].get("SCB_I8_PERSISTENT") != "2"
    assert "export SCB_TRI_SMEM=2" in open(tmp_path / "selected.env").read()
