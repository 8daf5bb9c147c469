#!/usr/bin/env python
"""Stage times of the FFT engine over a sweep of square ROIs (run with SCB_QUAD=0 / 1 to compare the modes)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seamlesscloneoptimization_b200 as scb  # noqa: E402
from seamlesscloneoptimization_b200 import workloads  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [170, 340, 680, 1360, 2700]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = scb.Context(0, stream=stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(0)
with torch.cuda.stream(stream):
    for n in sizes:
        S = n + 4
        src = workloads.smooth_rand(rng, S, S, 4.0)
        dst = workloads.smooth_rand(rng, S + 8, S + 8, 4.0)
        mask = np.full((S, S), 255, np.uint8)
        p = ((S + 8) // 2, (S + 8) // 2)
        d_src, d_dst, d_mask = (torch.from_numpy(a).to(dev) for a in (src, dst, mask))
        d_blend = torch.empty_like(d_dst)
        plan = scb.Plan(ctx, d_mask, src.shape[:2], dst.shape[:2], p, scb.MEM_DEVICE)
        g = plan.geometry
        acc = {}
        for it in range(8):
            flush.fill_(1)
            st = plan.execute_timed(d_src, d_dst, d_blend, scb.MEM_DEVICE)
            if it >= 3:
                for k, v in st.items():
                    acc.setdefault(k, []).append(v)
        m = {k: float(np.mean(v)) * 1e3 for k, v in acc.items()}
        print(f"quad={os.environ.get('SCB_QUAD', '1')} n={g.nx}x{g.ny} M={1 << g.log2m_x}: rows_fwd {m['rows_fwd']:.1f} us  cols {m['cols']:.1f} us  rows_inv {m['rows_inv']:.1f} us  rhs {m['rhs']:.1f} us", flush=True)
        plan.close()
ctx.close()
