"""Synthetic workloads of BASELINE.json's configs: seeded, textured u8 images and masks.
Pure numpy; no GPU, no cv2.  Shared by tests, bench.py and the oracle so that all see identical inputs."""
from __future__ import annotations

import math

import numpy as np

def smooth_rand(rng: np.random.Generator, h: int, w: int, sigma: float = 8.0) -> np.ndarray:
    """Textured u8 image: Gaussian-blurred uniform noise rescaled to [28,228] + integer noise
    U[-6,6].  The blur is a separable FFT-free box approximation (3 box passes) so that no cv2 is
    needed on the generator path."""
    a = rng.random((h, w, 3), dtype=np.float32)
    r = max(1, int(round(sigma * 0.9)))
    for _ in range(3):
        for axis in (0, 1):
            c = np.cumsum(a, axis=axis, dtype=np.float64)
            n = a.shape[axis]
            idx_hi = np.minimum(np.arange(n) + r, n - 1)
            idx_lo = np.maximum(np.arange(n) - r - 1, -1)
            hi = np.take(c, idx_hi, axis=axis)
            lo = np.where(
                (idx_lo >= 0).reshape([-1 if i == axis else 1 for i in range(3)]),
                np.take(c, np.maximum(idx_lo, 0), axis=axis),
                0.0,
            )
            cnt = (idx_hi - idx_lo).reshape([-1 if i == axis else 1 for i in range(3)])
            a = ((hi - lo) / cnt).astype(np.float32)
    a -= a.min()
    a /= max(float(a.max()), 1e-9)
    img = 28.0 + 200.0 * a + rng.integers(-6, 7, size=a.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def ellipse_mask(h: int, w: int, cx: float, cy: float, ax: float, ay: float, deg: float) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    t = math.radians(deg)
    xr = (xx - cx) * math.cos(t) + (yy - cy) * math.sin(t)
    yr = -(xx - cx) * math.sin(t) + (yy - cy) * math.cos(t)
    return (((xr / ax) ** 2 + (yr / ay) ** 2) <= 1.0).astype(np.uint8) * 255


def make_config(name: str, seed: int = 0):
    """Return (src, dst, mask, p) for BASELINE.json's configs."""
    rng = np.random.default_rng(seed)
    if name == "cfg1":  # 512x384 full mask into 1080p, ROI origin (800,150)
        src = smooth_rand(rng, 384, 512)
        dst = smooth_rand(rng, 1080, 1920)
        mask = np.full((384, 512), 255, np.uint8)
        return src, dst, mask, (1055, 341)
    if name == "cfg2":  # 2048x1536 irregular mask into 4K
        src = smooth_rand(rng, 1536, 2048)
        dst = smooth_rand(rng, 2160, 3840)
        mask = ellipse_mask(1536, 2048, 1024, 768, 900, 650, 15.0)
        disc = ellipse_mask(1536, 2048, 300, 300, 200, 200, 0.0)
        return src, dst, np.maximum(mask, disc), (1920, 1080)
    if name == "cfg4":  # 4096^2 full mask into 8K
        src = smooth_rand(rng, 4096, 4096)
        dst = smooth_rand(rng, 4320, 7680)
        return src, dst, np.full((4096, 4096), 255, np.uint8), (3840, 2160)
    if name == "cfg5":  # 1080p stream, fixed elliptic mask
        src = smooth_rand(rng, 720, 1280)
        dst = smooth_rand(rng, 1080, 1920)
        return src, dst, ellipse_mask(720, 1280, 640, 360, 600, 330, 0.0), (960, 540)
    if name == "small":
        src = smooth_rand(rng, 61, 83, sigma=3.0)
        dst = smooth_rand(rng, 120, 160, sigma=3.0)
        return src, dst, ellipse_mask(61, 83, 41, 30, 30, 22, 20.0), (80, 60)
    raise ValueError(f"unknown config {name!r}")


def make_batch_jobs(n_jobs: int, seed: int = 0, dst_hw=(1080, 1920), w_range=(64, 1024), h_range=(64, 768)):
    """cfg3: independent clone jobs with varied patch sizes/offsets and full/elliptic masks.
    Returns a list of dicts(src_hw, mask_kind, p, seed); images are generated lazily by
    `materialise_job` so that 512 jobs need not live in memory at once."""
    rng = np.random.default_rng(seed)
    H, W = dst_hw
    jobs = []
    for j in range(n_jobs):
        ws = int(rng.integers(w_range[0], w_range[1] + 1))
        hs = int(rng.integers(h_range[0], h_range[1] + 1))
        kind = "full" if rng.random() < 0.5 else "ellipse"
        # bbox of a ring-zeroed full mask is (ws-2)x(hs-2), an ellipse's is smaller: this p is safe for both
        px = int(rng.integers(ws // 2 + 1, W - ws // 2 - 1))
        py = int(rng.integers(hs // 2 + 1, H - hs // 2 - 1))
        jobs.append(dict(src_hw=(hs, ws), mask_kind=kind, p=(px, py), seed=int(seed * 100003 + j)))
    return jobs


def materialise_job(job, dst_hw=(1080, 1920), sigma: float = 6.0):
    rng = np.random.default_rng(job["seed"])
    hs, ws = job["src_hw"]
    src = smooth_rand(rng, hs, ws, sigma)
    dst = smooth_rand(rng, dst_hw[0], dst_hw[1], sigma)
    if job["mask_kind"] == "full":
        mask = np.full((hs, ws), 255, np.uint8)
    else:
        mask = ellipse_mask(hs, ws, ws / 2.0, hs / 2.0, ws * 0.45, hs * 0.45, 0.0)
    return src, dst, mask, job["p"]


