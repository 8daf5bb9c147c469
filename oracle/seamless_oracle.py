"""CPU oracle for OpenCV's seamlessClone(NORMAL_CLONE) hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product path (seamlesscloneoptimization_b200/) never does; it fails loudly when
its CUDA library is missing.

What is restated
----------------
The arithmetic lives in an un-vendored third-party dependency of the reference: OpenCV's `photo`
module (modules/photo/src/seamless_cloning.cpp + seamless_cloning_impl.cpp), pinned by the
reference to OpenCV 3.4.5 built from source (/root/reference/README.md:41,72-76; call sites
/root/reference/seamlessClone-OpenCV/seamlessClone_OpenCV.cpp:104,110).  The reference's CUDA
kernels mirror it and keep chunks of it as comments; each function below cites those lines
(paths relative to /root/reference/seamlessClone-CUDA/).

Pinning
-------
The reference holds NO golden outputs (.MISSING_LARGE_BLOBS), so the oracle is
pinned against outputs of the reference arithmetic itself run here: `cv2.seamlessClone` from the
installed OpenCV 4.13.0 wheel.  `restate(..., transform="cv")` is required to be BIT-EXACT
(0 differing bytes) against `cv2.seamlessClone` -- tests/test_oracle.py checks that live whenever
cv2 is importable and against the committed fixtures in tests/golden/ (made by
tests/golden/make_golden.py) otherwise.

Two transform back ends:
  transform="cv"   odd-extension + cv2.dft exactly as OpenCV's Cloning::dst does  (bit-exact pin)
  transform="f64"  scipy DST-I in float64 with OpenCV's float32 eigen-denominator  (noise floor)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

NORMAL_CLONE = 1
MIXED_CLONE = 2
MONOCHROME_TRANSFER = 3


class OracleError(ValueError):
    pass


@dataclass
class Geometry:
    """ROI bookkeeping of cv::seamlessClone's prologue (reference initMask, imp.cpp:978-1071)."""

    x: int  # bbox of ring-zeroed mask inside src/mask coordinates
    y: int
    w: int
    h: int
    rx: int  # ROI origin inside dst
    ry: int
    empty: bool = False

    @property
    def nx(self) -> int:
        return self.w - 2

    @property
    def ny(self) -> int:
        return self.h - 2


@dataclass
class Trace:
    """Every intermediate the parity tests look at (float32 unless stated)."""

    geom: Geometry
    eroded: np.ndarray | None = None  # u8  h x w
    vx: np.ndarray | None = None  # h x w x 3  blended forward-difference gradient (x)
    vy: np.ndarray | None = None
    rhs: np.ndarray | None = None  # ny x nx x 3   OpenCV's mod_diff, reference's `g`
    spectrum: np.ndarray | None = None  # ny x nx x 3   forward 2-D DST (before the division)
    solved: np.ndarray | None = None  # ny x nx x 3   u (before clamp/truncate)
    blend: np.ndarray | None = None  # u8 H x W x 3
    den: np.ndarray | None = None  # ny x nx float32
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------
# mask preparation
# --------------------------------------------------------------------------------------
def normalise_mask(mask: np.ndarray | None, src_shape) -> np.ndarray:
    """cv::seamlessClone accepts an empty mask (= all 255) or a 1/3/4-channel one; multi-channel
    masks are converted to grey first.  Only 0 / non-0 matters for the bbox, the values matter for
    the erosion.  """
    if mask is None or mask.size == 0:
        return np.full(src_shape[:2], 255, np.uint8)
    m = np.asarray(mask)
    if m.ndim == 3 and m.shape[2] == 1:
        m = m[:, :, 0]
    elif m.ndim == 3:
        import cv2  # colour -> grey uses OpenCV's fixed-point weights; only the cv path needs it

        m = cv2.cvtColor(m, cv2.COLOR_BGR2GRAY if m.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)
    if m.dtype != np.uint8:
        raise OracleError("mask must be uint8")
    return np.ascontiguousarray(m)


def ring_zero(mask: np.ndarray) -> np.ndarray:
    """Outermost 1-px ring -> 0 (OpenCV copyMakeBorder(inner, 1,1,1,1, CONSTANT, 0); reference
    setMaskBoundaryToConstant, imp.cpp:967-976)."""
    out = mask.copy()
    out[0, :] = 0
    out[-1, :] = 0
    out[:, 0] = 0
    out[:, -1] = 0
    return out


def bounding_box(mask_rz: np.ndarray):
    """cv::boundingRect of the non-zero pixels (reference calBoundingBox, imp.cpp:927-963)."""
    ys, xs = np.nonzero(mask_rz)
    if ys.size == 0:
        return None
    x0, x1 = int(xs.min()), int(xs.max())
    y0, y1 = int(ys.min()), int(ys.max())
    return x0, y0, x1 - x0 + 1, y1 - y0 + 1


def erode3(mask_rz: np.ndarray) -> np.ndarray:
    """erode(3x3 ones, iterations=3) == 7x7 min filter, run on the FULL ring-zeroed mask (OpenCV
    erodes a ROI view with its parent as context, so the zeros just outside the ROI take part).
    Outside the image OpenCV's erosion border is +inf (does not lower the min).
    (reference myErode, imp.cpp:892-925 -- binarising variant; OpenCV's min filter is the target.)"""
    h, w = mask_rz.shape
    pad = np.full((h + 6, w + 6), 255, np.uint8)
    pad[3:-3, 3:-3] = mask_rz
    # separable 7-tap min
    tmp = pad[:, 0:w].copy()
    for k in range(1, 7):
        np.minimum(tmp, pad[:, k : k + w], out=tmp)
    out = tmp[0:h].copy()
    for k in range(1, 7):
        np.minimum(out, tmp[k : k + h], out=out)
    return out


def plan_geometry(mask_gray: np.ndarray, dst_shape, p) -> tuple[Geometry, np.ndarray]:
    """Prologue of cv::seamlessClone: ring-zero, bbox, ROI placement (truncating integer
    division on the BBOX size), bounds check.  Returns geometry + ring-zeroed mask."""
    mrz = ring_zero(mask_gray)
    bb = bounding_box(mrz)
    if bb is None:
        return Geometry(0, 0, 0, 0, 0, 0, empty=True), mrz
    x, y, w, h = bb
    px, py = int(p[0]), int(p[1])
    rx, ry = px - w // 2, py - h // 2
    H, W = dst_shape[:2]
    if not (0 <= rx and rx + w <= W and 0 <= ry and ry + h <= H):
        raise OracleError("ROI outside dst (OpenCV: -215 Assertion failed ... roi)")
    if w < 3 or h < 3:
        raise OracleError("mask bounding box must be at least 3x3")
    return Geometry(x, y, w, h, rx, ry), mrz


# --------------------------------------------------------------------------------------
# stencils (reference pre_process_kernel_gradient imp.cpp:1920-1964, _lapXY :1966-2018)
# --------------------------------------------------------------------------------------
def blended_gradients(D: np.ndarray, S: np.ndarray, E: np.ndarray):
    """computeGradientX/Y (kernel [0,-1,1], BORDER_REFLECT_101) of dst-ROI and src-ROI, then
    arrayProduct with (255-E)/255 and E/255, then add.  float32, two rounded multiplies + one
    rounded add (NOT an FMA)."""
    f = np.float32
    Df = D.astype(f)
    Sf = S.astype(f)

    def gx(a):
        g = np.empty_like(a)
        g[:, :-1] = a[:, 1:] - a[:, :-1]
        g[:, -1] = a[:, -2] - a[:, -1]  # reflect-101; never reaches the interior
        return g

    def gy(a):
        g = np.empty_like(a)
        g[:-1] = a[1:] - a[:-1]
        g[-1] = a[-2] - a[-1]
        return g

    inv255 = f(1.0 / 255.0)
    m = (E.astype(f) * inv255)[:, :, None]
    mi = ((255 - E).astype(f) * inv255)[:, :, None]
    vx = (gx(Df) * mi).astype(f) + (gx(Sf) * m).astype(f)
    vy = (gy(Df) * mi).astype(f) + (gy(Sf) * m).astype(f)
    return vx.astype(f), vy.astype(f)


def rhs_from_gradients(vx: np.ndarray, vy: np.ndarray, D: np.ndarray) -> np.ndarray:
    """computeLaplacianX/Y (kernel [-1,1,0]) + sum, then OpenCV poissonSolver's
    mod_diff = lap - Laplacian(bound) on the interior (reference: the x==1 / y==1 / x==w-2 /
    y==h-2 branches, imp.cpp:1992-2007)."""
    f = np.float32
    h, w = D.shape[:2]
    lapx = (vx[1:-1, 1:-1] - vx[1:-1, 0:-2]).astype(f)
    lapy = (vy[1:-1, 1:-1] - vy[0:-2, 1:-1]).astype(f)
    lap = (lapx + lapy).astype(f)
    Df = D.astype(f)
    B = np.zeros((h - 2, w - 2, D.shape[2]), f)
    B[:, 0] += Df[1:-1, 0]
    B[:, -1] += Df[1:-1, w - 1]
    B[0, :] += Df[0, 1:-1]
    B[-1, :] += Df[h - 1, 1:-1]
    return (lap - B).astype(f)


# --------------------------------------------------------------------------------------
# eigen-denominator (reference initDSTMatrix_kernel imp.cpp:569-603 with float PI -- a deviation;
# OpenCV's recipe is the target: updateUij_kernel_fft imp.cpp:1642-1669 shows the use)
# --------------------------------------------------------------------------------------
def filters(w: int, h: int):
    f = np.float32
    sx = math.pi / (w - 1)
    sy = math.pi / (h - 1)
    fx = np.array([f(2.0) * f(math.cos(sx * (i + 1))) for i in range(w - 2)], f)
    fy = np.array([f(2.0) * f(math.cos(sy * (j + 1))) for j in range(h - 2)], f)
    return fx, fy


def denominator(w: int, h: int) -> np.ndarray:
    fx, fy = filters(w, h)
    return ((fx[None, :] + fy[:, None]).astype(np.float32) - np.float32(4)).astype(np.float32)


# --------------------------------------------------------------------------------------
# transforms
# --------------------------------------------------------------------------------------
def _dst_cv(a: np.ndarray, invert: bool) -> np.ndarray:
    """Cloning::dst exactly as OpenCV runs it (reference keeps the code as comments at
    imp.cpp:1696-1801): odd extension of every row to length 2n+2, complex DFT_ROWS, keep Im of
    bins 1..n, transpose, repeat, transpose back."""
    import cv2

    f = np.float32
    flag = cv2.DFT_ROWS | (cv2.DFT_SCALE | cv2.DFT_INVERSE if invert else 0)
    rows, cols = a.shape
    temp = np.zeros((rows, 2 * cols + 2), f)
    temp[:, 1 : cols + 1] = a
    temp[:, cols + 2 :] = -a[:, ::-1]
    cplx = cv2.merge([temp, np.zeros_like(temp)])
    cplx = cv2.dft(cplx, flags=flag)
    im = cplx[:, :, 1]
    temp = np.zeros((cols, 2 * rows + 2), f)
    t = np.ascontiguousarray(im[:, 1 : cols + 1].T)
    temp[:, 1 : rows + 1] = t
    temp[:, rows + 2 :] = -t[:, ::-1]
    cplx = cv2.merge([temp, np.zeros_like(temp)])
    cplx = cv2.dft(cplx, flags=flag)
    im = cplx[:, :, 1]
    return np.ascontiguousarray(im[:, 1 : rows + 1].T).astype(f)


def _dst_f64(a: np.ndarray, invert: bool) -> np.ndarray:
    """Same transform in float64: forward = 4*sum sin*sin, inverse = sum sin*sin / ((nx+1)(ny+1))."""
    from scipy.fft import dstn

    ny, nx = a.shape
    s = dstn(a.astype(np.float64), type=1)  # = 4 * sum a sin sin
    if invert:
        s = s / (4.0 * (nx + 1) * (ny + 1))
    return s


def dst2d(a: np.ndarray, invert: bool, transform: str):
    if transform == "cv":
        return _dst_cv(np.ascontiguousarray(a, np.float32), invert)
    if transform == "f64":
        return _dst_f64(a, invert)
    raise OracleError(f"unknown transform {transform!r}")


def solve_channel(rhs: np.ndarray, den: np.ndarray, transform: str):
    """Cloning::solve (reference solve() imp.cpp:1814-1896): forward DST, divide by the float32
    eigen-denominator, inverse DST."""
    spec = dst2d(rhs, False, transform)
    if transform == "cv":
        q = (spec / den).astype(np.float32)
    else:
        q = spec / den.astype(np.float64)
    u = dst2d(q, True, transform)
    return spec, u


def compose_u8(u: np.ndarray) -> np.ndarray:
    """v<0 -> 0, v>255 -> 255, else truncate toward zero (reference post_processing
    imp.cpp:2078-2103; OpenCV static_cast<uchar>, quoted at imp.cpp:1866-1874)."""
    v = np.where(u < 0, 0, np.where(u > 255, 255, u))
    return np.trunc(v).astype(np.uint8)


# --------------------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------------------
def restate(src, dst, mask, p, flags: int = NORMAL_CLONE, transform: str = "cv") -> Trace:
    """Restatement of cv::seamlessClone(src, dst, mask, p, blend, NORMAL_CLONE) returning every
    intermediate.  Never mutates its inputs (the real function overwrites the caller's mask)."""
    if flags != NORMAL_CLONE:
        raise OracleError("only NORMAL_CLONE is in scope")
    src = np.asarray(src)
    dst = np.asarray(dst)
    if src.ndim == 2:
        src = np.repeat(src[:, :, None], 3, axis=2)
    if dst.ndim != 3 or dst.shape[2] != 3 or dst.dtype != np.uint8 or src.dtype != np.uint8:
        raise OracleError("src/dst must be 8UC3 (src may be 8UC1)")
    mg = normalise_mask(mask, src.shape)
    if mg.shape != src.shape[:2]:
        raise OracleError("mask and src sizes differ")
    geom, mrz = plan_geometry(mg, dst.shape, p)
    tr = Trace(geom)
    if geom.empty:
        tr.blend = dst.copy()
        return tr
    x, y, w, h, rx, ry = geom.x, geom.y, geom.w, geom.h, geom.rx, geom.ry
    D = dst[ry : ry + h, rx : rx + w]
    S = src[y : y + h, x : x + w]  # OpenCV zeroes S outside the mask; irrelevant to the result
    E = erode3(mrz)[y : y + h, x : x + w]
    tr.eroded = E
    tr.vx, tr.vy = blended_gradients(D, S, E)
    tr.rhs = rhs_from_gradients(tr.vx, tr.vy, D)
    tr.den = denominator(w, h)
    spec = np.empty(tr.rhs.shape, np.float64 if transform == "f64" else np.float32)
    sol = np.empty_like(spec)
    for c in range(3):
        spec[:, :, c], sol[:, :, c] = solve_channel(tr.rhs[:, :, c], tr.den, transform)
    tr.spectrum, tr.solved = spec, sol
    out = D.copy()
    out[1:-1, 1:-1] = compose_u8(sol)
    blend = dst.copy()
    blend[ry : ry + h, rx : rx + w] = out
    tr.blend = blend
    return tr


def cv_reference(src, dst, mask, p):
    """The real thing: cv2.seamlessClone with a COPY of the mask (it mutates its mask argument)."""
    import cv2

    m = normalise_mask(mask, np.asarray(src).shape).copy()
    return cv2.seamlessClone(np.ascontiguousarray(src), np.ascontiguousarray(dst), m, (int(p[0]), int(p[1])), cv2.NORMAL_CLONE)


# --------------------------------------------------------------------------------------
# synthetic workloads -- shared by tests and bench so both see identical inputs
# --------------------------------------------------------------------------------------
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
    raise OracleError(f"unknown config {name!r}")


def make_batch_jobs(n_jobs: int, seed: int = 0, dst_hw=(1080, 1920)):
    """cfg3: independent 1080p clone jobs with varied patch sizes/offsets and full/elliptic masks.
    Returns a list of dicts(src_hw, mask_kind, p, seed); images are generated lazily by
    `materialise_job` so that 512 jobs need not live in memory at once."""
    rng = np.random.default_rng(seed)
    H, W = dst_hw
    jobs = []
    for j in range(n_jobs):
        ws = int(rng.integers(64, 1025))
        hs = int(rng.integers(64, 769))
        kind = "full" if rng.random() < 0.5 else "ellipse"
        # bbox of ring-zeroed full mask is (ws-2)x(hs-2); ellipse bbox is smaller: keep p safe for both
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


def compare_u8(a: np.ndarray, b: np.ndarray) -> dict:
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return dict(max_abs=int(d.max()), n_diff=int(np.count_nonzero(d)), pct_exact=100.0 * float(np.mean(d == 0)), diff_sum=int(d.sum()))


def rel_linf(a: np.ndarray, ref: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(np.asarray(a, np.float64) - ref)) / max(np.max(np.abs(ref)), 1e-30))
