"""Host-side mirror of the reference's operator interface for the NORMAL_CLONE path.

  seamlessClone(src, dst, mask, p, flags)           cv2.seamlessClone-shaped free function
      (OpenCV C++: seamlessClone(src, dst, mask, Point p, Mat& blend, NORMAL_CLONE);
       reference call sites /root/reference/seamlessClone-OpenCV/seamlessClone_OpenCV.cpp:104,110)
  class SeamlessClone                               the reference's Boost.Python class, same method names
      (/root/reference/seamlessClone-CUDA/seamlessClone-python-binding/SeamlessClone.h:80-98,
       usage SeamlessClone_test.py:5-26)
  Context / Plan                                    thin objects over the C ABI (include/scb.h)

Everything here is plumbing over the CUDA library; there is no CPU implementation to fall back to.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import (EXEC_BLEND_PREFILLED, EXEC_DEFAULT, MEM_DEVICE, MEM_HOST, MIXED_CLONE, MONOCHROME_TRANSFER, NORMAL_CLONE, ScbError)


def _gray_mask(mask: np.ndarray | None, src_hw) -> np.ndarray:
    """cv::seamlessClone accepts an empty mask (= all 255) and 1/3/4-channel masks (colour masks go
    through cvtColor BGR2GRAY, 4.x fixed point: (3735 B + 19235 G + 9798 R + 16384) >> 15)."""
    if mask is None or getattr(mask, "size", 0) == 0:
        return np.full(tuple(src_hw), 255, np.uint8)
    m = np.asarray(mask)
    if m.dtype != np.uint8:
        raise ScbError(capi.SCB_ERR_INVALID_ARGUMENT, "mask must be uint8")
    if m.ndim == 3 and m.shape[2] == 1:
        m = m[:, :, 0]
    elif m.ndim == 3 and m.shape[2] in (3, 4):
        b, g, r = (m[:, :, k].astype(np.uint32) for k in range(3))
        m = ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)
    elif m.ndim != 2:
        raise ScbError(capi.SCB_ERR_INVALID_ARGUMENT, "mask must be HxW, HxWx1, HxWx3 or HxWx4")
    if m.strides[-1] != 1:
        m = np.ascontiguousarray(m)
    return m


def _bgr(img: np.ndarray, name: str) -> np.ndarray:
    a = np.asarray(img)
    if a.dtype != np.uint8:
        raise ScbError(capi.SCB_ERR_INVALID_ARGUMENT, f"{name} must be uint8")
    if a.ndim == 2 or (a.ndim == 3 and a.shape[2] == 1):  # OpenCV accepts a grey src: replicate
        a = np.repeat(a.reshape(a.shape[0], a.shape[1], 1), 3, axis=2)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ScbError(capi.SCB_ERR_INVALID_ARGUMENT, f"{name} must be HxWx3")
    if a.strides[2] != 1 or a.strides[1] != 3:
        a = np.ascontiguousarray(a)
    return a


class Context:
    """One device + one stream + grow-only workspace + table cache (scb_context)."""

    def __init__(self, device: int = 0, stream: int | None = None, lib_path: str | None = None):
        self.lib = capi.load(lib_path)
        h = C.c_void_p()
        rc = self.lib.scb_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != capi.SCB_OK:
            raise ScbError(rc, (self.lib.scb_last_error(None) or b"").decode())
        self.handle = h
        self.device = int(device)

    def _check(self, rc: int):
        if rc != capi.SCB_OK:
            raise ScbError(rc, (self.lib.scb_last_error(self.handle) or b"").decode())

    def sync(self):
        self._check(self.lib.scb_sync(self.handle))

    def set_engine(self, engine: int):
        """DST engine of plans created from now on: capi.ENGINE_AUTO / ENGINE_TRI / ENGINE_FFT / ENGINE_TC."""
        self._check(self.lib.scb_set_engine(self.handle, int(engine)))

    def set_orientation(self, orientation: int):
        """Tridiagonal engine: -1 = FFT axis chosen per plan by cost (default), 0 = FFT along x, 1 = FFT along y."""
        self._check(self.lib.scb_set_orientation(self.handle, int(orientation)))

    def tc_selftest(self, n: int, lines: int, transposed: bool = False) -> float:
        err = C.c_double()
        self._check(self.lib.scb_tc_selftest(self.handle, int(n), int(lines), int(bool(transposed)), C.byref(err)))
        return float(err.value)

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.scb_kernel_launches(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.scb_stream(self.handle) or 0)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.scb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- one-shot -----------------------------------------------------------------------------
    def seamless_clone(self, src, dst, mask, p, flags: int = NORMAL_CLONE) -> np.ndarray:
        """cv2.seamlessClone semantics on host arrays: returns a fresh blend, never touches dst or mask."""
        s, d = _bgr(src, "src"), _bgr(dst, "dst")
        m = _gray_mask(mask, s.shape[:2])
        blend = np.empty_like(d, order="C")
        vs, vd, vm, vb = capi.host_view(s), capi.host_view(d), capi.host_view(m), capi.host_view(blend)
        self._check(self.lib.scb_seamless_clone(self.handle, C.byref(vs), C.byref(vd), C.byref(vm), int(p[0]), int(p[1]), C.byref(vb), int(flags), MEM_HOST))
        return blend

    def plan(self, mask, src_hw, dst_hw, p, mem_kind: int = MEM_HOST, clone_flags: int = NORMAL_CLONE) -> "Plan":
        return Plan(self, mask, src_hw, dst_hw, p, mem_kind, clone_flags)


class Plan:
    """Everything that depends only on (mask, sizes, p): ROI, eroded mask, DST tables (scb_plan)."""

    def __init__(self, ctx: Context, mask, src_hw, dst_hw, p, mem_kind: int = MEM_HOST, clone_flags: int = NORMAL_CLONE):
        self.ctx = ctx
        self.lib = ctx.lib
        if mem_kind == MEM_HOST:
            m = _gray_mask(mask, src_hw)
            vm = capi.host_view(m)
        else:
            vm = mask if isinstance(mask, capi.ScbImage) else capi.tensor_view(mask)
        h = C.c_void_p()
        ctx._check(self.lib.scb_plan_create_ex(ctx.handle, C.byref(vm), mem_kind, int(src_hw[0]), int(src_hw[1]), int(dst_hw[0]), int(dst_hw[1]), int(p[0]), int(p[1]),
                                               int(clone_flags), C.byref(h)))
        self.handle = h
        g = capi.ScbGeometry()
        ctx._check(self.lib.scb_plan_geometry(self.handle, C.byref(g)))
        self.geometry = g
        self.engine = int(self.lib.scb_plan_engine(self.handle))

    def execute(self, src, dst, blend=None, mem_kind: int = MEM_HOST, flags: int = EXEC_DEFAULT):
        if mem_kind == MEM_HOST:
            s, d = _bgr(src, "src"), _bgr(dst, "dst")
            if blend is None:
                blend = np.empty_like(d, order="C")
            vs, vd, vb = capi.host_view(s), capi.host_view(d), capi.host_view(blend)
        else:
            if blend is None:
                raise ScbError(capi.SCB_ERR_INVALID_ARGUMENT, "device execution needs a blend tensor")
            vs, vd, vb = (x if isinstance(x, capi.ScbImage) else capi.tensor_view(x) for x in (src, dst, blend))
        self.ctx._check(self.lib.scb_plan_execute(self.handle, C.byref(vs), C.byref(vd), C.byref(vb), mem_kind, flags))
        return blend

    def execute_graph(self, src, dst, blend, flags: int = EXEC_DEFAULT):
        """Device-resident execute replayed as one CUDA graph launch (fixed-mask streams)."""
        vs, vd, vb = (x if isinstance(x, capi.ScbImage) else capi.tensor_view(x) for x in (src, dst, blend))
        self.ctx._check(self.lib.scb_plan_execute_graph(self.handle, C.byref(vs), C.byref(vd), C.byref(vb), flags))
        return blend

    STAGES = ("copy_in", "rhs", "lowfreq", "rows_fwd", "cols", "rows_inv", "copy_out")
    I8_KERNELS = ("i8_digitize_fwd", "i8_gemm_fwd", "i8_digitize_inv", "i8_gemm_inv", "i8_compose")

    def execute_timed(self, src, dst, blend, mem_kind: int = MEM_HOST, flags: int = EXEC_DEFAULT) -> dict:
        """execute() with CUDA events between the stages; returns {stage: ms} (syncs the stream)."""
        if mem_kind == MEM_HOST:
            vs, vd, vb = capi.host_view(_bgr(src, "src")), capi.host_view(_bgr(dst, "dst")), capi.host_view(blend)
        else:
            vs, vd, vb = (x if isinstance(x, capi.ScbImage) else capi.tensor_view(x) for x in (src, dst, blend))
        ms, i8 = (C.c_float * 7)(), (C.c_float * 5)()
        self.ctx._check(self.lib.scb_plan_execute_timed_i8(self.handle, C.byref(vs), C.byref(vd), C.byref(vb), mem_kind, flags, ms, i8))
        out = dict(zip(self.STAGES, (float(v) for v in ms)))
        if self.engine == capi.ENGINE_I8:  # the kernels of the INT8 passes on their own event pairs
            out.update(zip(self.I8_KERNELS, (float(v) for v in i8)))
        return out

    def set_debug(self, on: bool = True):
        self.ctx._check(self.lib.scb_plan_set_debug(self.handle, int(on)))

    def intermediate(self, which: int) -> np.ndarray:
        g = self.geometry
        shape = {
            capi.INT_GRADIENT_X: (3, g.h, g.w), capi.INT_GRADIENT_Y: (3, g.h, g.w), capi.INT_RHS: (3, g.ny, g.nx),
            capi.INT_SPECTRUM: (3, g.nx, g.ny), capi.INT_SOLVED: (3, g.ny, g.nx), capi.INT_ERODED_MASK: (1, g.h, g.w),
        }[which]
        out = np.empty(shape, np.float32)
        n = C.c_size_t()
        self.ctx._check(self.lib.scb_plan_get_intermediate(self.handle, which, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        assert n.value == out.size
        return out

    def close(self):
        if getattr(self, "handle", None) and getattr(self.ctx, "handle", None):
            self.lib.scb_plan_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def seamlessClone(src, dst, mask, p, flags: int = NORMAL_CLONE, device: int = 0) -> np.ndarray:
    """Drop-in for cv2.seamlessClone(src, dst, mask, p, flags) on numpy arrays; flags = NORMAL_CLONE (the hot path),
    MIXED_CLONE, MONOCHROME_TRANSFER or their _WIDE variants (9, 10, 11)."""
    return default_context(device).seamless_clone(src, dst, mask, p, flags)


class SeamlessClone:
    """The reference's Python class (SeamlessClone.h:80-98): load the three Mats, then run.

        sc = SeamlessClone()
        sc.loadMatsInSeamlessClone(face, body, mask, centerX, centerY, gpu_id)
        blended = sc.seamlessClone();  sc.sync();  sc.destroy()

    face = patch (src), body = destination; the instance is created lazily on the first
    seamlessClone() call like the reference does (SeamlessClone.cpp:110-113).  gpu_id is honoured
    (the reference never calls cudaSetDevice); an id past the visible devices wraps to device 0 so the
    reference's own test script (gpu_id=1) runs on a single-GPU box.
    """

    def __init__(self, lib_path: str | None = None):
        self._lib_path = lib_path
        self._ctx: Context | None = None
        self.face = self.body = self.mask = self.blendedMat = None
        self.centerX = self.centerY = 0
        self.gpu_id = 0

    def loadMatsInSeamlessClone(self, face, body, mask, centerX: int, centerY: int, gpu_id: int = 0):
        self.face, self.body, self.mask = face, body, mask
        self.centerX, self.centerY, self.gpu_id = int(centerX), int(centerY), int(gpu_id)

    def seamlessClone(self) -> np.ndarray:
        if self.face is None:
            raise ScbError(capi.SCB_ERR_INVALID_ARGUMENT, "call loadMatsInSeamlessClone first")
        if self._ctx is None:
            lib = capi.load(self._lib_path)
            ndev = max(1, int(lib.scb_device_count()))
            self._ctx = Context(self.gpu_id % ndev, lib_path=self._lib_path)
        self.blendedMat = self._ctx.seamless_clone(self.face, self.body, self.mask, (self.centerX, self.centerY))
        return self.blendedMat

    def sync(self):
        if self._ctx is not None:
            self._ctx.sync()

    def destroy(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    # template leftovers of the reference binding, kept as trivial shims
    def mat2py(self, mat):
        return np.asarray(mat)

    def py2mat(self, o, m=None):
        return np.asarray(o)

    def loadImageInCpp_Demo(self, imagePath: str):
        import cv2  # demo helper only

        return cv2.imread(imagePath)
