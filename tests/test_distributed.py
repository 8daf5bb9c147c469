"""World-size-2 gloo test of the row/column-sharded solve (BASELINE cfg4 structure at toy size).

Each rank loads the emulator build of the CUDA sources (tests/emu: 'device' memory is host memory),
runs its passes through the C ABI and exchanges At / Ct with dist.all_to_all over gloo on 127.0.0.1.
The sharded result must equal the single-context result bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import seamlesscloneoptimization_b200 as scb
from oracle import seamless_oracle as so
from seamlesscloneoptimization_b200 import _capi as capi
from seamlesscloneoptimization_b200 import batch, sharded


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lib_path, out_dir, engine):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src, dst, mask, p = so.make_config("small", 12)
        ctx = scb.Context(0, lib_path=lib_path)
        ctx.set_engine(engine)
        hm, hs, hd = (np.ascontiguousarray(a) for a in (mask, src, dst))
        plan = scb.Plan(ctx, capi.host_view(hm), src.shape[:2], dst.shape[:2], p, scb.MEM_DEVICE)
        blend = torch.from_numpy(dst.copy())
        solve = sharded.ShardedSolve(ctx, plan, torch.device("cpu"))
        solve.run(capi.host_view(hs), capi.host_view(hd), capi.host_view(blend.numpy()))
        ctx.sync()
        solve.gather_rows(blend)
        np.save(os.path.join(out_dir, f"blend_{rank}.npy"), blend.numpy())
        plan.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("engine", [capi.ENGINE_I8, capi.ENGINE_TRI, capi.ENGINE_FFT], ids=["i8-segments", "tri-segments", "fft-transpose"])
def test_sharded_solve_over_gloo(tmp_path, emu_lib, world, engine):
    """tri: row shards = segment groups of the partitioned Thomas solve, two small all-reduces;  fft: two all-to-all transposes."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, emu_lib, str(tmp_path), engine), nprocs=world, join=True)
    src, dst, mask, p = so.make_config("small", 12)
    with scb.Context(0, lib_path=emu_lib) as ctx:
        ctx.set_engine(engine)
        single = ctx.seamless_clone(src, dst, mask, p)
    for r in range(world):
        got = np.load(tmp_path / f"blend_{r}.npy")
        assert np.array_equal(got, single), f"rank {r}: sharded result differs from the single-context result"


def test_split_and_lpt_partition():
    assert sharded.split(4092, 8) == [0, 512, 1024, 1536, 2048, 2559, 3070, 3581, 4092]
    assert sharded.split(5, 8)[-1] == 5
    costs = [9, 1, 8, 2, 7, 3, 6, 4]
    parts = batch.shard_jobs(costs, 2)
    assert sorted(parts[0] + parts[1]) == list(range(8))
    loads = [sum(costs[i] for i in part) for part in parts]
    assert abs(loads[0] - loads[1]) <= 1
    assert batch.shard_jobs(costs, 2) == parts  # deterministic: every rank derives the same partition
