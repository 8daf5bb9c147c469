"""Shared parity checks: the SAME assertions run against the CUDA library on a B200 (-m gpu) and
against the emulator build of the same sources on CPU (-m "not gpu")."""
from __future__ import annotations

import glob
import os

import numpy as np

import seamlesscloneoptimization_b200 as scb
from oracle import seamless_oracle as so
from seamlesscloneoptimization_b200 import _capi as capi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances of BASELINE.json's north_star
FLOAT_REL_TOL = 1e-4   # float intermediates: relative L-inf error
U8_MAX_ABS = 1         # final image: +-1 LSB ...
U8_MIN_EXACT = 99.9    # ... with at least 99.9 % of the solved bytes exact


def golden_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["flags"] = int(d["flags"]) if "flags" in d else so.NORMAL_CLONE  # fixtures older than the flag support are NORMAL_CLONE
    return d


def interior(img, geom):
    x, y, w, h, rx, ry = (int(v) for v in geom[:6])
    return img[ry + 1 : ry + h - 1, rx + 1 : rx + w - 1]


def allowed_mismatches(n_bytes: int) -> int:
    """>= 99.9 % exact, but never fewer than 2 bytes on tiny images (one truncation flip is one byte)."""
    return max(2, int(n_bytes * (100.0 - U8_MIN_EXACT) / 100.0))


def check_against_golden(ctx: scb.Context, name: str, check_float: bool = True):
    z = load_golden(name)
    src, dst, mask, p, geom = z["src"], z["dst"], z["mask"], tuple(int(v) for v in z["p"]), z["geom"]
    plan = ctx.plan(mask, src.shape[:2], dst.shape[:2], p, clone_flags=z["flags"])
    g = plan.geometry
    assert [g.x, g.y, g.w, g.h, g.rx, g.ry] == [int(v) for v in geom], "ROI geometry differs from OpenCV's"
    plan.set_debug(True)
    dst_before = dst.copy()
    mask_before = mask.copy()
    blend = plan.execute(src, dst)
    assert np.array_equal(dst, dst_before), "dst was modified"
    assert np.array_equal(mask, mask_before), "mask was modified (OpenCV does that; we must not)"
    # mask preparation and the integer stencil are bit-exact
    E = plan.intermediate(capi.INT_ERODED_MASK)[0].astype(np.uint8)
    assert np.array_equal(E, z["eroded"]), "eroded mask differs"
    rhs = plan.intermediate(capi.INT_RHS).transpose(1, 2, 0)
    assert np.array_equal(rhs, z["rhs"]), f"RHS not bit-exact, max diff {np.abs(rhs - z['rhs']).max()}"
    if check_float:
        u = plan.intermediate(capi.INT_SOLVED).transpose(1, 2, 0)
        if plan.engine not in (capi.ENGINE_TRI, capi.ENGINE_I8):  # the tridiagonal solve along y never forms the 2-D spectrum
            spec = plan.intermediate(capi.INT_SPECTRUM).transpose(2, 1, 0)
            assert so.rel_linf(spec, z["spectrum"]) < FLOAT_REL_TOL
        assert so.rel_linf(u, z["solved"]) < FLOAT_REL_TOL
    # final image
    x, y, w, h, rx, ry = (int(v) for v in geom)
    expect = dst.copy()
    expect[ry : ry + h, rx : rx + w] = z["blend_roi"]
    outside = np.ones(dst.shape[:2], bool)
    outside[ry + 1 : ry + h - 1, rx + 1 : rx + w - 1] = False
    assert np.array_equal(blend[outside], dst[outside]), "pixels outside the ROI interior must equal dst"
    cmp = so.compare_u8(interior(blend, geom), interior(expect, geom))
    assert cmp["max_abs"] <= U8_MAX_ABS, cmp
    assert cmp["n_diff"] <= allowed_mismatches(interior(blend, geom).size), cmp
    plan.close()
    return cmp
