#!/usr/bin/env python
"""bench.py -- solved Mpix/s of seamlessClone(NORMAL_CLONE) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg5|cfg4] [--impl reference]

One "step" = one whole NORMAL_CLONE of the workload (default cfg2: 3840x2160 dst, 2048x1536 irregular
mask patch, ROI 1810x1341 -> 2.42 M solved RGB pixels).  With N > 1 (torchrun, one rank per GPU)
every rank clones its own independent job of that shape (jobs shard with no collective): weak scaling,
value = N * solved pixels / max-over-ranks time.

  value   device-resident: src/dst/mask/blend already in HBM, plan (mask prep + tables) made once,
          timed with CUDA events on the library's stream, L2 flushed between steps
  e2e     the drop-in call: scb_seamless_clone on pinned HOST buffers -- mask upload, bbox, erosion,
          ROI H2D, solve, ROI D2H, dst->blend host copy all inside the timed region (wall clock)
  roofline  the dominant kernel (columns pass), algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline  cv2.seamlessClone (OpenCV 4.13 wheel = the reference arithmetic) on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from seamlesscloneoptimization_b200 import workloads  # noqa: E402

METRIC = "solved_mpix_per_s"
UNIT = "Mpix/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (NVML, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def solved_pixels(g) -> int:
    return int(g.nx) * int(g.ny)


def cpu_reference_run(src, dst, mask, p, n_calls: int, threads: int | None):
    """cv2.seamlessClone = OpenCV's own CPU implementation of the path (the reference arithmetic)."""
    import cv2

    if threads:
        cv2.setNumThreads(threads)
    cv2.seamlessClone(src, dst, mask.copy(), p, cv2.NORMAL_CLONE)  # warm-up
    ts = []
    for _ in range(n_calls):
        m = mask.copy()
        t0 = time.perf_counter()
        cv2.seamlessClone(src, dst, m, p, cv2.NORMAL_CLONE)
        ts.append(time.perf_counter() - t0)
    return ts, cv2.getNumThreads(), cv2.__version__


def roi_of(mask):
    ys, xs = np.nonzero(mask[1:-1, 1:-1])
    w, h = int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)
    return w, h


def run_reference(args, rank, world):
    """--impl reference: OpenCV's CPU seamlessClone on the box's host cores, same config/metric/unit."""
    if rank != 0:
        return
    src, dst, mask, p = workloads.make_config(args.workload, seed=0)
    w, h = roi_of(mask)
    px = (w - 2) * (h - 2)
    ncores = os.cpu_count() or 1
    try:
        import cv2  # noqa: F401
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 not importable: {e}"}))
        return
    cpu_reference_run(src, dst, mask, p, max(0, args.warmup - 1), ncores)
    ts, nthreads, ver = cpu_reference_run(src, dst, mask, p, args.steps, ncores)
    total = sum(ts)
    value = px * len(ts) / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "solved_pixels_per_step": px, "roi": [w, h]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "reference",
                         "sample": f"cv2.seamlessClone (OpenCV {ver} wheel), {len(ts)} whole {args.workload} clones, {nthreads} threads of {ncores} host cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "p50_ms": 1e3 * statistics.median(ts),
    }
    print(json.dumps(line))


def workload_name(w):
    return {
        "cfg1": "cfg1: 512x384 full-mask patch into 1920x1080 at ROI origin (800,150)",
        "cfg2": "cfg2: 3840x2160 dst, 2048x1536 irregular-mask patch (ellipse+disc, ROI 1810x1341), p=(1920,1080)",
        "cfg4": "cfg4: 7680x4320 dst, 4096x4096 full-mask patch (ROI 4094x4094)",
        "cfg5": "cfg5: 1920x1080 dst, 1280x720 elliptic-mask patch (ROI 1201x661), fixed mask/offset",
    }[w]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg4", "cfg5"])
    ap.add_argument("--cpu-baseline-calls", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import seamlesscloneoptimization_b200 as scb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: seamlesscloneoptimization_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank clones its own independent job of the same shape (different seed per rank)
    src, dst, mask, p = workloads.make_config(args.workload, seed=rank)
    stream = torch.cuda.Stream(device=dev)
    ctx = scb.Context(local_rank, stream=stream.cuda_stream)
    d_src, d_dst, d_mask = (torch.from_numpy(a).to(dev) for a in (src, dst, mask))
    d_blend = torch.empty_like(d_dst)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    plan = scb.Plan(ctx, d_mask, src.shape[:2], dst.shape[:2], p, scb.MEM_DEVICE)
    g = plan.geometry
    px = solved_pixels(g)

    def device_step():
        plan.execute(d_src, d_dst, d_blend, scb.MEM_DEVICE)

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            device_step()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = ctx.kernel_launches
        evs = []
        for _ in range(args.steps):
            flush.fill_(1)  # evict L2 between timed steps (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            device_step()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        launches = ctx.kernel_launches - launches0
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(step_ms)

        # per-stage CUDA-event times of the same call (dominant kernel for the roofline)
        stage_acc = {}
        n_prof = min(10, max(3, args.steps))
        for _ in range(n_prof):
            flush.fill_(1)
            st = plan.execute_timed(d_src, d_dst, d_blend, scb.MEM_DEVICE)
            for k, v in st.items():
                stage_acc.setdefault(k, []).append(v)
        sampler.stop_flag = True
        sampler.join(timeout=2)
    stages = {k: statistics.mean(v) for k, v in stage_acc.items()}

    # ---- end to end through the drop-in call, host buffers ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = a
        return t

    h_src, h_dst, h_mask = pinned(src), pinned(dst), pinned(mask)
    h_blend = torch.empty(dst.shape, dtype=torch.uint8, pin_memory=True)
    from seamlesscloneoptimization_b200 import _capi as capi
    import ctypes as C

    vs, vd, vm, vb = (capi.host_view(t.numpy()) for t in (h_src, h_dst, h_mask, h_blend))

    def e2e_step():
        rc = ctx.lib.scb_seamless_clone(ctx.handle, C.byref(vs), C.byref(vd), C.byref(vm), p[0], p[1], C.byref(vb), scb.NORMAL_CLONE, scb.MEM_HOST)
        if rc:
            ctx._check(rc)

    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    e2e_ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        e2e_step()  # returns when blend is complete on the host
        e2e_ts.append(time.perf_counter() - t0)
    barrier()
    e2e_total = sum(e2e_ts)
    # sanity: host path and device path agree bit for bit
    assert np.array_equal(h_blend.numpy(), d_blend.cpu().numpy()), "host and device paths disagree"

    t_dev = torch.tensor([total_ms, e2e_total * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = (float(x) for x in t_dev.cpu())

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        value = world * px * args.steps / (total_ms_max * 1e-3) / 1e6
        e2e_value = world * px * args.steps / (e2e_ms_max * 1e-3) / 1e6
        dom = max(("rows_fwd", "cols", "rows_inv"), key=lambda k: stages.get(k, 0.0))
        alg_bytes = {"rows_fwd": 19, "cols": 24, "rows_inv": 15}[dom] * px
        achieved = alg_bytes / (stages[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "solved_pixels_per_step": px, "roi": [g.w, g.h], "fft_len": [1 << g.log2m_x, 1 << g.log2m_y],
                       "l2": "256 MiB flush write between timed steps", "jobs_per_step_per_gpu": 1},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(mask.size + 2 * 3 * g.w * g.h), "d2h_bytes_per_step": int(3 * g.nx * g.ny),
                    "ms_per_step": e2e_ms_max / args.steps, "p50_ms": 1e3 * statistics.median(e2e_ts), "call": "scb_seamless_clone(HOST pinned buffers): mask prep + ROI H2D + solve + ROI D2H + dst->blend host copy"},
            "gpu_launches": int(launches),
            "p50_ms_device": statistics.median(step_ms),
            "stages_ms": stages,
            "roofline": {"bound": "hbm", "kernel": {"rows_fwd": "rows_fwd_kernel", "cols": "cols_kernel", "rows_inv": "rows_inv_kernel"}[dom],
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_launch": alg_bytes, "traffic": None},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                ncores = os.cpu_count() or 1
                ts, nthreads, ver = cpu_reference_run(src, dst, mask, p, args.cpu_baseline_calls, ncores)
                line["cpu_baseline"] = {"value": px * len(ts) / sum(ts) / 1e6, "unit": UNIT, "cores": nthreads, "kind": "reference",
                                        "p50_ms": 1e3 * statistics.median(ts),
                                        "sample": f"cv2.seamlessClone (OpenCV {ver} wheel), {len(ts)} whole {args.workload} clones after 1 warm-up, {nthreads} threads of {ncores} host cores"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line))
    plan.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
