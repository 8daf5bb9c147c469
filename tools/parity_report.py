#!/usr/bin/env python
"""Byte parity of libscb.so against cv2.seamlessClone on the bench workloads: % of solved bytes exact, max |diff|.
  python tools/parity_report.py [cfg1 cfg2 cfg5 cfg4]        (SCB_LIBRARY selects a library variant, PARITY_SEEDS="0 1 2" the image seeds)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np

import seamlesscloneoptimization_b200 as scb
from seamlesscloneoptimization_b200 import workloads

with scb.Context(0) as ctx:
    for wl, seed in ((w, int(sd)) for w in (sys.argv[1:] or ["cfg1", "cfg2", "cfg5"]) for sd in os.environ.get("PARITY_SEEDS", "0").split()):
        src, dst, mask, p = workloads.make_config(wl, seed=seed)
        ref = cv2.seamlessClone(src, dst, mask.copy(), p, cv2.NORMAL_CLONE)
        plan = ctx.plan(mask, src.shape[:2], dst.shape[:2], p)
        got = plan.execute(src, dst)
        g = plan.geometry
        a = got[g.ry + 1 : g.ry + g.h - 1, g.rx + 1 : g.rx + g.w - 1].astype(np.int16)
        b = ref[g.ry + 1 : g.ry + g.h - 1, g.rx + 1 : g.rx + g.w - 1].astype(np.int16)
        d = np.abs(a - b)
        print(f"{wl} seed {seed}: ROI {g.w}x{g.h}  exact {100.0 * (d == 0).mean():.4f} %  max |diff| {int(d.max())}  differing bytes {int((d != 0).sum())}")
        plan.close()
