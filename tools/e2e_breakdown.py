#!/usr/bin/env python
"""Where the end-to-end (host buffers) time of one 4K clone goes: plan creation, execute, raw PCIe copies."""
import ctypes as C
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import seamlesscloneoptimization_b200 as scb
from seamlesscloneoptimization_b200 import _capi as capi
from seamlesscloneoptimization_b200 import workloads

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
src, dst, mask, p = workloads.make_config(wl, seed=0)
pin = lambda a: torch.from_numpy(a.copy()).pin_memory()
h_src, h_dst, h_mask = pin(src), pin(dst), pin(mask)
h_blend = torch.empty(dst.shape, dtype=torch.uint8).pin_memory()
ctx = scb.Context(0)
vs, vd, vm, vb = (capi.host_view(t.numpy()) for t in (h_src, h_dst, h_mask, h_blend))


def timeit(f, n=30, warm=5):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def full():
    ctx._check(ctx.lib.scb_seamless_clone(ctx.handle, C.byref(vs), C.byref(vd), C.byref(vm), p[0], p[1], C.byref(vb), 1, 0))


plans = []


def plan_only():
    h = C.c_void_p()
    ctx._check(ctx.lib.scb_plan_create(ctx.handle, C.byref(vm), 0, src.shape[0], src.shape[1], dst.shape[0], dst.shape[1], p[0], p[1], C.byref(h)))
    ctx._check(ctx.lib.scb_plan_destroy(h))


plan = scb.Plan(ctx, mask, src.shape[:2], dst.shape[:2], p, scb.MEM_HOST)
g = plan.geometry


def exec_only():
    ctx._check(ctx.lib.scb_plan_execute(plan.handle, C.byref(vs), C.byref(vd), C.byref(vb), 0, 0))


def exec_prefilled():
    ctx._check(ctx.lib.scb_plan_execute(plan.handle, C.byref(vs), C.byref(vd), C.byref(vb), 0, 1))


d_in = torch.empty(2 * 3 * g.w * g.h + mask.size, dtype=torch.uint8, device="cuda")
h_in = torch.empty(2 * 3 * g.w * g.h + mask.size, dtype=torch.uint8).pin_memory()
d_out = torch.empty(3 * g.nx * g.ny, dtype=torch.uint8, device="cuda")
h_out = torch.empty(3 * g.nx * g.ny, dtype=torch.uint8).pin_memory()


def pcie():
    d_in.copy_(h_in, non_blocking=True)
    h_out.copy_(d_out, non_blocking=True)


def hostcopy():
    h_blend.numpy()[...] = h_dst.numpy()


print(f"{wl}: ROI {g.w}x{g.h}")
print(f"  scb_seamless_clone (HOST)            {timeit(full):.3f} ms")
print(f"  scb_plan_create + destroy            {timeit(plan_only):.3f} ms")
print(f"  scb_plan_execute (HOST)              {timeit(exec_only):.3f} ms")
print(f"  scb_plan_execute (HOST, prefilled)   {timeit(exec_prefilled):.3f} ms   (no host copy of dst -> blend)")
print(f"  raw H2D {d_in.numel()/1e6:.1f} MB + D2H {d_out.numel()/1e6:.1f} MB   {timeit(pcie):.3f} ms")
print(f"  numpy copy dst -> blend {dst.nbytes/1e6:.1f} MB        {timeit(hostcopy, 10, 2):.3f} ms (single thread)")
