#!/usr/bin/env python
"""Step-by-step probe of the tensor-core pass on a B200 (prints after every case; run under `timeout`)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seamlesscloneoptimization_b200 as scb  # noqa: E402

cases = [(64, 128, False), (64, 128, True), (65, 100, False), (256, 128, False), (513, 200, False), (1000, 300, True), (1339, 384, False), (1808, 300, False), (4092, 128, False)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
with scb.Context(0) as c:
    for n, lines, tr in cases:
        t0 = time.time()
        print(f"n={n} lines={lines} transposed={tr} ...", end=" ", flush=True)
        err = c.tc_selftest(n, lines, tr)
        print(f"max rel err {err:.3e}  ({time.time() - t0:.2f}s)", flush=True)
