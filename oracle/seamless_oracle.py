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
NORMAL_CLONE_WIDE = 9  # OpenCV >= 4.11: the centre of src (not of the mask bounding box) goes to p; same arithmetic
MIXED_CLONE_WIDE = 10
MONOCHROME_TRANSFER_WIDE = 11


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


def plan_geometry(mask_gray: np.ndarray, dst_shape, p, wide: bool = False) -> tuple[Geometry, np.ndarray]:
    """Prologue of cv::seamlessClone: ring-zero, bbox, ROI placement (truncating integer
    division on the BBOX size), bounds check.  Returns geometry + ring-zeroed mask."""
    mrz = ring_zero(mask_gray)
    bb = bounding_box(mrz)
    if bb is None:
        return Geometry(0, 0, 0, 0, 0, 0, empty=True), mrz
    x, y, w, h = bb
    px, py = int(p[0]), int(p[1])
    rx, ry = px - w // 2, py - h // 2
    if wide:  # PROBE (cv2 4.13): flags 9/10/11 == flags 1/2/3 with the ROI at p - src_size/2 + bbox origin
        rx, ry = px - mask_gray.shape[1] // 2 + x, py - mask_gray.shape[0] // 2 + y
    H, W = dst_shape[:2]
    if not (0 <= rx and rx + w <= W and 0 <= ry and ry + h <= H):
        raise OracleError("ROI outside dst (OpenCV: -215 Assertion failed ... roi)")
    if w < 3 or h < 3:
        raise OracleError("mask bounding box must be at least 3x3")
    return Geometry(x, y, w, h, rx, ry), mrz


# --------------------------------------------------------------------------------------
# stencils (reference pre_process_kernel_gradient imp.cpp:1920-1964, _lapXY :1966-2018)
# --------------------------------------------------------------------------------------
def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY) for 8-bit images: 15-bit fixed point, (B*3735 + G*19235 + R*9798 + 2^14) >> 15
    (pinned bit-exact against cv2.cvtColor in tests/test_oracle.py)."""
    b, g, r = (img[:, :, k].astype(np.int64) for k in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def blended_gradients(D: np.ndarray, S: np.ndarray, E: np.ndarray, flags: int = NORMAL_CLONE):
    """computeGradientX/Y (kernel [0,-1,1], BORDER_REFLECT_101) of dst-ROI and src-ROI, then
    arrayProduct with (255-E)/255 and E/255, then add.  float32, two rounded multiplies + one
    rounded add (NOT an FMA).
      MIXED_CLONE          per pixel and channel the patch gradient pair is replaced by dst's unless
                           |gxS - gyS| > |gxD - gyD|   (Cloning::normalClone, OpenCV seamless_cloning_impl.cpp)
      MONOCHROME_TRANSFER  the patch gradients are those of cvtColor(patch, BGR2GRAY), replicated to 3 channels"""
    f = np.float32
    Df = D.astype(f)
    if flags == MONOCHROME_TRANSFER:
        S = np.repeat(bgr2gray_u8(S)[:, :, None], 3, axis=2)
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
    gxD, gyD, gxS, gyS = gx(Df), gy(Df), gx(Sf), gy(Sf)
    if flags == MIXED_CLONE:
        keep = np.abs(gxS - gyS) > np.abs(gxD - gyD)
        gxS = np.where(keep, gxS, gxD)
        gyS = np.where(keep, gyS, gyD)
    vx = (gxD * mi).astype(f) + (gxS * m).astype(f)
    vy = (gyD * mi).astype(f) + (gyS * m).astype(f)
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
    """Restatement of cv::seamlessClone(src, dst, mask, p, blend, flags) returning every intermediate;
    flags = NORMAL_CLONE (the hot path), MIXED_CLONE or MONOCHROME_TRANSFER (the same solver behind a different
    gradient selection).  Never mutates its inputs (the real function overwrites the caller's mask)."""
    wide = flags >= NORMAL_CLONE_WIDE
    if wide:
        flags -= NORMAL_CLONE_WIDE - NORMAL_CLONE
    if flags not in (NORMAL_CLONE, MIXED_CLONE, MONOCHROME_TRANSFER):
        raise OracleError("flags must be NORMAL_CLONE, MIXED_CLONE, MONOCHROME_TRANSFER or their _WIDE variants")
    src = np.asarray(src)
    dst = np.asarray(dst)
    if src.ndim == 2:
        src = np.repeat(src[:, :, None], 3, axis=2)
    if dst.ndim != 3 or dst.shape[2] != 3 or dst.dtype != np.uint8 or src.dtype != np.uint8:
        raise OracleError("src/dst must be 8UC3 (src may be 8UC1)")
    mg = normalise_mask(mask, src.shape)
    if mg.shape != src.shape[:2]:
        raise OracleError("mask and src sizes differ")
    geom, mrz = plan_geometry(mg, dst.shape, p, wide)
    tr = Trace(geom)
    if geom.empty:
        tr.blend = dst.copy()
        return tr
    x, y, w, h, rx, ry = geom.x, geom.y, geom.w, geom.h, geom.rx, geom.ry
    D = dst[ry : ry + h, rx : rx + w]
    S = src[y : y + h, x : x + w]  # OpenCV zeroes S outside the mask; irrelevant to the result
    E = erode3(mrz)[y : y + h, x : x + w]
    tr.eroded = E
    tr.vx, tr.vy = blended_gradients(D, S, E, flags)
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


def cv_reference(src, dst, mask, p, flags: int = NORMAL_CLONE):
    """The real thing: cv2.seamlessClone with a COPY of the mask (it mutates its mask argument)."""
    import cv2

    m = normalise_mask(mask, np.asarray(src).shape).copy()
    return cv2.seamlessClone(np.ascontiguousarray(src), np.ascontiguousarray(dst), m, (int(p[0]), int(p[1])), int(flags))


# --------------------------------------------------------------------------------------
# synthetic workloads live in the package so that tests, bench and oracle see the
# same inputs; re-exported here for convenience
# --------------------------------------------------------------------------------------
from seamlesscloneoptimization_b200.workloads import (ellipse_mask, make_batch_jobs, make_config, materialise_job, smooth_rand)  # noqa: E402,F401


def compare_u8(a: np.ndarray, b: np.ndarray) -> dict:
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return dict(max_abs=int(d.max()), n_diff=int(np.count_nonzero(d)), pct_exact=100.0 * float(np.mean(d == 0)), diff_sum=int(d.sum()))


def rel_linf(a: np.ndarray, ref: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(np.asarray(a, np.float64) - ref)) / max(np.max(np.abs(ref)), 1e-30))
