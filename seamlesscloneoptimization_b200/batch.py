"""Batches of independent clone jobs (BASELINE cfg3) over the C ABI's scb_clone_batch.

One process per GPU; across GPUs the jobs are split by rank with no data-path collective
(embarrassingly parallel).  The reference has no batch interface: its driver runs one
clone per call (/root/reference/seamlessClone-CUDA/seamlessClone_imp.cu:265-352).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _capi as capi
from .api import Context, _bgr, _gray_mask


def shard_jobs(costs: Sequence[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of job indices to `world` ranks.
    `costs` is any per-job work estimate (solved pixels).  Deterministic, so every rank computes the
    same partition without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    for r in range(world):
        out[r].sort()
    return out


def clone_batch_host(ctx: Context, jobs: Sequence[tuple]) -> list[np.ndarray]:
    """jobs: (src, dst, mask, (px, py)) host arrays.  Returns one fresh blend per job (OpenCV semantics).
    Raises ScbError for the first failed job after the whole batch has run."""
    arr = (capi.ScbJob * len(jobs))()
    keep = []
    for k, (src, dst, mask, p) in enumerate(jobs):
        s, d = _bgr(src, "src"), _bgr(dst, "dst")
        m = _gray_mask(mask, s.shape[:2])
        b = np.empty_like(d, order="C")
        keep.append((s, d, m, b))
        arr[k].src, arr[k].dst, arr[k].mask, arr[k].blend = capi.host_view(s), capi.host_view(d), capi.host_view(m), capi.host_view(b)
        arr[k].px, arr[k].py = int(p[0]), int(p[1])
    ctx._check(ctx.lib.scb_clone_batch(ctx.handle, arr, len(jobs), capi.MEM_HOST))
    return [k[3] for k in keep]


def make_device_jobs(views: Sequence[tuple]):
    """views: (src_view, dst_view, mask_view, blend_view, (px, py)) of ScbImage device views.
    Returns the ctypes job array for clone_batch_device (build it once, replay it per step)."""
    arr = (capi.ScbJob * len(views))()
    for k, (vs, vd, vm, vb, p) in enumerate(views):
        arr[k].src, arr[k].dst, arr[k].mask, arr[k].blend = vs, vd, vm, vb
        arr[k].px, arr[k].py = int(p[0]), int(p[1])
    return arr


def clone_batch_device(ctx: Context, job_array) -> None:
    """Runs a device-resident batch; returns after the context's lanes have drained."""
    ctx._check(ctx.lib.scb_clone_batch(ctx.handle, job_array, len(job_array), capi.MEM_DEVICE))
    for j in job_array:
        if j.status != capi.SCB_OK:
            raise capi.ScbError(j.status, "a job of the batch failed")
