#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/sharded_check.py [--size S]

NCCL check of the row/column-sharded solve on N GPUs of one box: every rank runs ShardedSolve on a
full-mask S x S patch, gathers the row slabs and compares with the single-GPU solve of the same plan
(must be bit-identical: the passes are the same kernels on row/column ranges).  Prints per-rank times.
"""
import argparse
import contextlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seamlesscloneoptimization_b200 as scb  # noqa: E402
from seamlesscloneoptimization_b200 import _capi as capi, sharded, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--engine", default="tri", choices=["tri", "fft"], help="tri: segment scheme (2 all-reduces); fft: transpose scheme (2 all-to-alls)")
    ap.add_argument("--plain-context", action="store_true", help="Context on its own stream, caller on torch's default stream (the ordering ShardedSolve must provide itself)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S = args.size
    rng = np.random.default_rng(7)
    src = workloads.smooth_rand(rng, S, S)
    dst = workloads.smooth_rand(rng, S + 200, S + 300)
    mask = np.full((S, S), 255, np.uint8)
    p = ((S + 300) // 2, (S + 200) // 2)
    stream = torch.cuda.Stream(device=dev)
    ctx = scb.Context(local) if args.plain_context else scb.Context(local, stream=stream.cuda_stream)
    if args.plain_context:
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ctx.set_engine(capi.ENGINE_TRI if args.engine == "tri" else capi.ENGINE_FFT)  # single and sharded run the same engine: bit-identical
    d_src, d_dst, d_mask = (torch.from_numpy(a).to(dev) for a in (src, dst, mask))
    torch.cuda.synchronize()
    with (contextlib.nullcontext() if args.plain_context else torch.cuda.stream(stream)):
        plan = scb.Plan(ctx, d_mask, src.shape[:2], dst.shape[:2], p, scb.MEM_DEVICE)
        single = torch.empty_like(d_dst)
        plan.execute(d_src, d_dst, single, scb.MEM_DEVICE)
        blend = d_dst.clone()
        torch.cuda.synchronize()
        solve = sharded.ShardedSolve(ctx, plan, dev)
        vs, vd, vb = capi.tensor_view(d_src), capi.tensor_view(d_dst), capi.tensor_view(blend)
        solve.run(vs, vd, vb)
        solve.gather_rows(blend)
        torch.cuda.synchronize()
        same = bool(torch.equal(blend, single))
        if args.engine == "tri":  # the same solve replayed as ONE CUDA graph (passes + NCCL exchange captured together)
            blend2 = d_dst.clone()
            vb2 = capi.tensor_view(blend2)
            torch.cuda.synchronize()
            solve.capture(vs, vd, vb2)
            solve.run_graph()
            solve.run_graph()
            solve.gather_rows(blend2)
            torch.cuda.synchronize()
            same = same and bool(torch.equal(blend2, single))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record(stream)
        for _ in range(5):
            solve.run(vs, vd, vb)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
    print(f"rank {rank}/{world}: engine {args.engine}{' (plain context)' if args.plain_context else ''}: sharded == single: {same}; {ms:.3f} ms per sharded solve of {S}x{S}", flush=True)
    solve.graph = None  # the captured NCCL kernels go before the communicator does
    torch.cuda.synchronize()
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    plan.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
