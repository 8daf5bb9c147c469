"""ctypes binding of include/scb.h.  Loads the CUDA library built in-tree by __graft_entry__.build()
(seamlesscloneoptimization_b200/lib/libscb.so) and nothing else: there is no CPU fallback -- a
missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "lib", "libscb.so")

SCB_OK = 0
SCB_ERR_INVALID_ARGUMENT = 1
SCB_ERR_ROI_OUT_OF_BOUNDS = 2
SCB_ERR_UNSUPPORTED = 3
SCB_ERR_CUDA = 4
SCB_ERR_NO_DEVICE = 5
SCB_ERR_OUT_OF_MEMORY = 6

NORMAL_CLONE = 1
MIXED_CLONE = 2
MONOCHROME_TRANSFER = 3
NORMAL_CLONE_WIDE, MIXED_CLONE_WIDE, MONOCHROME_TRANSFER_WIDE = 9, 10, 11

MEM_HOST = 0
MEM_DEVICE = 1

EXEC_DEFAULT = 0
ENGINE_AUTO, ENGINE_FFT, ENGINE_TC, ENGINE_TRI, ENGINE_I8 = 0, 1, 2, 3, 4
EXEC_BLEND_PREFILLED = 1

INT_GRADIENT_X, INT_GRADIENT_Y, INT_RHS, INT_SPECTRUM, INT_SOLVED, INT_ERODED_MASK = range(6)

# every symbol include/scb.h declares (tests check that the library exports all of them)
EXPORTS = [
    "scb_create", "scb_destroy", "scb_sync", "scb_stream", "scb_last_error", "scb_status_string",
    "scb_kernel_launches", "scb_device_count", "scb_source_hash", "scb_kernel_variants", "scb_set_engine", "scb_set_orientation", "scb_tc_selftest", "scb_host_alloc", "scb_host_free",
    "scb_plan_create", "scb_plan_create_ex", "scb_plan_destroy", "scb_plan_geometry", "scb_plan_engine", "scb_plan_execute", "scb_plan_execute_timed", "scb_plan_execute_timed_i8", "scb_plan_execute_graph", "scb_plan_set_debug",
    "scb_plan_get_intermediate", "scb_seamless_clone", "scb_plan_cache_stats", "scb_clone_batch",
    "scb_plan_rows_forward", "scb_plan_cols", "scb_plan_lowfreq_finish", "scb_plan_rows_inverse", "scb_plan_lowk",
    "scb_plan_tri_layout", "scb_plan_tri_forward", "scb_plan_tri_finish", "scb_plan_tri_finish_slots",
    "my_seamlessclone_api_imp_create_instance", "my_seamlessclone_api_imp_run",
    "my_seamlessclone_api_imp_destroy", "my_seamlessclone_api_imp_sync",
]


class ScbImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("channels", C.c_int32), ("stride", C.c_int64)]


class ScbGeometry(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("x", "y", "w", "h", "rx", "ry", "nx", "ny", "empty", "log2m_x", "log2m_y")]


class ScbJob(C.Structure):
    _fields_ = [("src", ScbImage), ("dst", ScbImage), ("mask", ScbImage), ("blend", ScbImage), ("px", C.c_int32), ("py", C.c_int32), ("status", C.c_int32), ("flags", C.c_int32)]


class ScbError(RuntimeError):
    """Mirrors cv2.error for the drop-in entry points: .code is the SCB_ERR_* status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"scb error {code}: {msg}")
        self.code = code
        self.msg = msg


_libs: dict[str, C.CDLL] = {}


def load(path: str | None = None) -> C.CDLL:
    path = os.path.abspath(path or os.environ.get("SCB_LIBRARY", DEFAULT_LIB))
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "seamlesscloneoptimization_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    P = C.POINTER
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    sig = {
        "scb_create": (i, [i, vp, P(vp)]),
        "scb_destroy": (i, [vp]),
        "scb_sync": (i, [vp]),
        "scb_stream": (vp, [vp]),
        "scb_last_error": (C.c_char_p, [vp]),
        "scb_status_string": (C.c_char_p, [i]),
        "scb_kernel_launches": (C.c_uint64, [vp]),
        "scb_device_count": (i, []),
        "scb_source_hash": (C.c_char_p, []),
        "scb_kernel_variants": (C.c_char_p, []),
        "scb_set_engine": (i, [vp, i]),
        "scb_set_orientation": (i, [vp, i]),
        "scb_tc_selftest": (i, [vp, i, i, i, P(C.c_double)]),
        "scb_plan_engine": (i, [vp]),
        "scb_host_alloc": (i, [P(vp), sz]),
        "scb_host_free": (i, [vp]),
        "scb_plan_create": (i, [vp, P(ScbImage), i, i, i, i, i, i, i, P(vp)]),
        "scb_plan_create_ex": (i, [vp, P(ScbImage), i, i, i, i, i, i, i, i, P(vp)]),
        "scb_plan_destroy": (i, [vp]),
        "scb_plan_geometry": (i, [vp, P(ScbGeometry)]),
        "scb_plan_execute": (i, [vp, P(ScbImage), P(ScbImage), P(ScbImage), i, i]),
        "scb_plan_execute_timed": (i, [vp, P(ScbImage), P(ScbImage), P(ScbImage), i, i, P(C.c_float)]),
        "scb_plan_execute_timed_i8": (i, [vp, P(ScbImage), P(ScbImage), P(ScbImage), i, i, P(C.c_float), P(C.c_float)]),
        "scb_plan_execute_graph": (i, [vp, P(ScbImage), P(ScbImage), P(ScbImage), i]),
        "scb_plan_set_debug": (i, [vp, i]),
        "scb_plan_get_intermediate": (i, [vp, i, vp, sz, P(sz)]),
        "scb_seamless_clone": (i, [vp, P(ScbImage), P(ScbImage), P(ScbImage), i, i, P(ScbImage), i, i]),
        "scb_plan_cache_stats": (i, [vp, P(C.c_uint64), P(C.c_uint64)]),
        "scb_clone_batch": (i, [vp, P(ScbJob), i, i]),
        "scb_plan_rows_forward": (i, [vp, P(ScbImage), P(ScbImage), i, i, i, vp, vp]),
        "scb_plan_cols": (i, [vp, i, i, vp, vp, vp]),
        "scb_plan_lowfreq_finish": (i, [vp, vp, vp]),
        "scb_plan_rows_inverse": (i, [vp, vp, P(ScbImage), i, i, i]),
        "scb_plan_lowk": (i, [vp, P(i), P(i)]),
        "scb_plan_tri_layout": (i, [vp, P(i), P(i), P(sz), P(sz), P(sz)]),
        "scb_plan_tri_forward": (i, [vp, P(ScbImage), P(ScbImage), i, i, i, vp, vp, vp]),
        "scb_plan_tri_finish": (i, [vp, P(ScbImage), i, i, i, vp, vp, vp]),
        "scb_plan_tri_finish_slots": (i, [vp, P(ScbImage), i, i, i, vp, vp, vp, i]),
        "my_seamlessclone_api_imp_create_instance": (vp, [i]),
        "my_seamlessclone_api_imp_run": (i, [vp, P(ScbImage), P(ScbImage), P(ScbImage), i, i, i, i, P(ScbImage)]),
        "my_seamlessclone_api_imp_destroy": (None, [vp]),
        "my_seamlessclone_api_imp_sync": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib


def host_view(a: np.ndarray) -> ScbImage:
    """POD view of a numpy uint8 image (H,W), (H,W,1) or (H,W,3); the last axis must be dense."""
    if a.dtype != np.uint8:
        raise TypeError("images must be uint8")
    if a.ndim == 2:
        ch = 1
    elif a.ndim == 3:
        ch = a.shape[2]
    else:
        raise TypeError("images must be HxW or HxWxC")
    if a.strides[1] != ch or (a.ndim == 3 and a.strides[2] != 1):
        raise TypeError("pixel rows must be dense (only the row stride may be padded)")
    return ScbImage(a.ctypes.data, a.shape[0], a.shape[1], ch, a.strides[0])


def device_view(ptr: int, rows: int, cols: int, channels: int, stride: int) -> ScbImage:
    return ScbImage(ptr, rows, cols, channels, stride)


def tensor_view(t) -> ScbImage:
    """POD view of a CUDA torch.uint8 tensor (H,W) or (H,W,C)."""
    if t.dim() == 2:
        ch = 1
    else:
        ch = t.shape[2]
        if t.stride(2) != 1:
            raise TypeError("channel axis must be dense")
    if t.stride(1) != ch:
        raise TypeError("pixel rows must be dense")
    return ScbImage(t.data_ptr(), t.shape[0], t.shape[1], ch, t.stride(0))
