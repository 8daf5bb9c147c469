"""Generate the golden vectors in tests/golden/*.npz by running the REAL reference arithmetic here:
`cv2.seamlessClone` (OpenCV 4.13.0 wheel) for the final image and the bit-exact restatement
(oracle.seamless_oracle.restate(transform="cv"), built from cv2.dft) for the float intermediates.

The reference repo holds no golden outputs (.MISSING_LARGE_BLOBS), so these fixtures are what pins
the oracle and the CUDA path when cv2 or /root/reference are not around (the GPU box has no
/root/reference).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import seamless_oracle as so  # noqa: E402


def save(name, src, dst, mask, p, flags=so.NORMAL_CLONE):
    import cv2

    ref = so.cv_reference(src, dst, mask, p, flags)
    tr = so.restate(src, dst, mask, p, flags=flags, transform="cv")
    assert np.array_equal(ref, tr.blend), name
    g = tr.geom
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        src=src, dst=dst, mask=mask, p=np.array(p, np.int32),
        geom=np.array([g.x, g.y, g.w, g.h, g.rx, g.ry], np.int32),
        blend_roi=ref[g.ry : g.ry + g.h, g.rx : g.rx + g.w],
        eroded=tr.eroded, rhs=tr.rhs, spectrum=tr.spectrum, solved=tr.solved,
        cv_version=np.array(cv2.__version__), flags=np.array(flags, np.int32),
    )
    print(name, g, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


def main():
    import cv2

    src, dst, mask, p = so.make_config("small", seed=0)
    save("small_ellipse", src, dst, mask, p)

    rng = np.random.default_rng(11)
    src = so.smooth_rand(rng, 37, 53, 3.0)
    dst = so.smooth_rand(rng, 90, 101, 3.0)
    save("odd_full", src, dst, np.full((37, 53), 255, np.uint8), (50, 45))

    src, dst, mask, p = so.make_config("small", seed=5)
    grey = (mask > 0) * rng.integers(1, 256, size=mask.shape).astype(np.uint8)
    save("grey_mask", src, dst, grey.astype(np.uint8), p)

    # U[0,255] noise: exercises the saturation branches of the compose step
    src = rng.integers(0, 256, size=(40, 64, 3), dtype=np.uint8)
    dst = rng.integers(0, 256, size=(80, 96, 3), dtype=np.uint8)
    save("noise_saturating", src, dst, so.ellipse_mask(40, 64, 32, 20, 28, 17, 0.0), (48, 40))

    # the other cv::seamlessClone flags (same solver, different gradient selection / ROI placement)
    rng = np.random.default_rng(23)
    src = so.smooth_rand(rng, 61, 83, 2.0)
    dst = so.smooth_rand(rng, 140, 170, 2.0)
    mask = so.ellipse_mask(61, 83, 36.0, 27.0, 30.0, 20.0, 0.3)
    save("flags_mixed_ellipse", src, dst, mask, (88, 71), so.MIXED_CLONE)
    grey = ((mask > 0) * rng.integers(1, 256, size=mask.shape)).astype(np.uint8)
    save("flags_monochrome_grey_mask", src, dst, grey, (88, 71), so.MONOCHROME_TRANSFER)
    save("flags_mixed_wide_ellipse", src, dst, mask, (88, 71), so.MIXED_CLONE_WIDE)

    ref_dir = "/root/reference/seamlessClone-OpenCV/images"
    if os.path.isdir(ref_dir):
        # the reference's own test pair (SeamlessClone_test.py:8-16, p=(800,150), full-255 mask);
        # dst is cropped around the ROI (651..949 x 54..246) so the fixture stays small: same ROI pixels, same result
        face = cv2.imread(os.path.join(ref_dir, "airplane.jpg"))
        body = cv2.imread(os.path.join(ref_dir, "sky.jpg"))
        full = so.cv_reference(face, body, np.full(face.shape[:2], 255, np.uint8), (800, 150))
        x0, y0 = 600, 20
        crop = np.ascontiguousarray(body[y0:280, x0:1000])
        mask = np.full(face.shape[:2], 255, np.uint8)
        save("airplane_sky_crop", face, crop, mask, (800 - x0, 150 - y0))
        z = np.load(os.path.join(HERE, "airplane_sky_crop.npz"))
        gx, gy, gw, gh, rx, ry = z["geom"]
        assert np.array_equal(z["blend_roi"], full[ry + y0 : ry + y0 + gh, rx + x0 : rx + x0 + gw])


if __name__ == "__main__":
    main()
