#!/usr/bin/env python
"""The reference's compare/vs.py (/root/reference/compare/vs.py:36-86) as a reusable report: absolute difference of two
result images -- here normally cv2.seamlessClone's against libscb.so's -- with the figures the reference's write-up
tabulates ("Diff sum", "Diff max", % of channels different; SeamlessClone Project Overview.pdf p.3, p.15).

  python tools/compare_vs.py A.png B.png [--diff-out diff.png]          two stored results
  python tools/compare_vs.py --clone SRC DST [MASK] CX CY [--flags F]   run both implementations on the same inputs
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def report(a: np.ndarray, b: np.ndarray, diff_out=None) -> dict:
    if a.shape != b.shape:
        raise SystemExit(f"shapes differ: {a.shape} vs {b.shape}")
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    n = int((d != 0).sum())
    out = {"diff_sum": int(d.sum()), "diff_count": n, "diff_max": int(d.max()), "diff_min_nonzero": int(d[d != 0].min()) if n else 0,
           "pct_channels_different": 100.0 * n / d.size}
    print("diff sum {diff_sum}, diff count {diff_count}, min/max {diff_min_nonzero}/{diff_max}, {pct_channels_different:.4f} % of channels different".format(**out))
    if diff_out:
        import cv2

        cv2.imwrite(diff_out, np.clip(d * 30, 0, 255).astype(np.uint8))  # x30 gain, like compare/vs.py
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("args", nargs="+")
    ap.add_argument("--clone", action="store_true")
    ap.add_argument("--flags", type=int, default=1)
    ap.add_argument("--diff-out")
    a = ap.parse_args()
    import cv2

    if not a.clone:
        if len(a.args) != 2:
            ap.error("expected two images")
        x, y = (np.load(f) if f.endswith(".npy") else cv2.imread(f) for f in a.args)
        report(x, y, a.diff_out)
        return
    from tools.seamless_clone_cli import load

    import seamlesscloneoptimization_b200 as scb

    src, dst = load(a.args[0]), load(a.args[1])
    rest = a.args[2:]
    mask = load(rest[0], gray=True) if len(rest) == 3 else np.full(src.shape[:2], 255, np.uint8)
    cx, cy = int(rest[-2]), int(rest[-1])
    ref = cv2.seamlessClone(src, dst, mask.copy(), (cx, cy), a.flags)  # cv2 overwrites its mask argument
    with scb.Context(0) as ctx:
        got = ctx.seamless_clone(src, dst, mask, (cx, cy), a.flags)
    report(ref, got, a.diff_out)


if __name__ == "__main__":
    main()
