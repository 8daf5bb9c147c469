#!/usr/bin/env python
"""Command-line driver in the shape of the reference's two CLIs
  /root/reference/seamlessClone-CUDA/seamlessClone_main.cu:69-94      seamlessClone_main src.yml dst.yml mask.yml cx cy gpu
  /root/reference/seamlessClone-OpenCV/seamlessClone_OpenCV.cpp:41-124  seamlessClone_OpenCV src.jpg dst.jpg cx cy
on top of libscb.so:

  python tools/seamless_clone_cli.py SRC DST [MASK] CX CY [--gpu N] [--flags 1|2|3|9|10|11] [--out blend.png] [--loops K]

SRC / DST / MASK are image files (anything cv2.imread reads), .npy arrays, or OpenCV FileStorage .yml/.yaml files holding one Mat
(the reference's interchange format, README.md:59).  Without MASK the whole patch is cloned (all-255 mask, as the reference's
drivers do).  Prints the reference's timing line ("Compute stage performance time= ... msec, patch size=WxH",
seamlessClone_imp.cu:343-346) for the median of K runs after one warm-up.  Image decoding uses cv2; the clone itself never does."""
import argparse
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def load(path, gray=False):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        a = np.load(path)
    elif ext in (".yml", ".yaml", ".xml"):
        import cv2

        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        root = fs.root()
        a = root.getNode(root.keys()[0]).mat()
        fs.release()
    else:
        import cv2

        a = cv2.imread(path, cv2.IMREAD_GRAYSCALE if gray else cv2.IMREAD_COLOR)
    if a is None:
        raise SystemExit(f"cannot read {path}")
    return np.ascontiguousarray(a.astype(np.uint8))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("rest", nargs="+", help="[MASK] CX CY")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--flags", type=int, default=1)
    ap.add_argument("--out", default="blend.png")
    ap.add_argument("--loops", type=int, default=1)
    args = ap.parse_args()
    if len(args.rest) == 3:
        mask_path, cx, cy = args.rest[0], int(args.rest[1]), int(args.rest[2])
    elif len(args.rest) == 2:
        mask_path, cx, cy = None, int(args.rest[0]), int(args.rest[1])
    else:
        ap.error("expected [MASK] CX CY")
    src, dst = load(args.src), load(args.dst)
    mask = load(mask_path, gray=True) if mask_path else np.full(src.shape[:2], 255, np.uint8)
    if mask.ndim == 3:
        mask = mask[:, :, 0]

    import seamlesscloneoptimization_b200 as scb

    with scb.Context(args.gpu) as ctx:
        blend = ctx.seamless_clone(src, dst, mask, (cx, cy), args.flags)  # warm-up, like the reference's run()
        ts = []
        for _ in range(max(1, args.loops)):
            t0 = time.perf_counter()
            blend = ctx.seamless_clone(src, dst, mask, (cx, cy), args.flags)
            ts.append((time.perf_counter() - t0) * 1e3)
    print(f"Compute stage performance time= {statistics.median(ts):.3f} msec, patch size={src.shape[1]}x{src.shape[0]}")
    if args.out.endswith(".npy"):
        np.save(args.out, blend)
    else:
        import cv2

        cv2.imwrite(args.out, blend)
    print(f"wrote {args.out}")


if __name__ == "__main__":
    main()
