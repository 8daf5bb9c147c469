#!/usr/bin/env python
"""A small, torch-free pass over every kernel of the default path for `compute-sanitizer --tool memcheck`:
     compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py
   (host-resident numpy images through the C ABI: all three clone modes, a banded transfer, a device plan reuse, a batch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import seamlesscloneoptimization_b200 as scb
from seamlesscloneoptimization_b200 import workloads

rng = np.random.default_rng(0)
with scb.Context(0) as ctx:
    for (hs, ws, H, W) in [(61, 83, 140, 170), (200, 333, 400, 600), (35, 700, 90, 800)]:
        src = workloads.smooth_rand(rng, hs, ws, 2.0)
        dst = workloads.smooth_rand(rng, H, W, 2.0)
        mask = workloads.ellipse_mask(hs, ws, ws / 2.0, hs / 2.0, ws * 0.45, hs * 0.45, 0.2)
        for flags in (1, 2, 3, 9):
            out = ctx.seamless_clone(src, dst, mask, (W // 2, H // 2), flags)
            assert out.shape == dst.shape
    os.environ["SCB_BANDS"] = "3"
    out = ctx.seamless_clone(src, dst, mask, (W // 2, H // 2), 1)
    plan = ctx.plan(mask, src.shape[:2], dst.shape[:2], (W // 2, H // 2))
    for _ in range(2):
        plan.execute(src, dst)
    plan.close()
print("sanitize_smoke ok")
