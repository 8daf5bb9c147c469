"""One large solve sharded over the ranks of a torch.distributed group (BASELINE cfg4).

Two schemes, chosen by the plan's engine:

  tridiagonal engine (default) -- ShardedSolve runs scb_plan_tri_forward / scb_plan_tri_finish: rank r keeps its rows from the
  stencil to the composed bytes.  The column solve is a partitioned (SPIKE) Thomas solve whose segments follow the row shards,
  so the ranks only combine the segment-end values (3 x 16 x 2 x nx floats) and the 32 x 32 low-frequency projections with two
  small all-reduces.  No transpose, no all-to-all: 1.5 MB instead of 2 x 200 MB at 8K.

  FFT engine -- the transpose scheme below (two all-to-alls):

  rank r owns interior rows [ys[r], ys[r+1]) for the row passes and columns [xs[r], xs[r+1]) for the
  column pass; between the passes the row-transformed field is exchanged with an all-to-all:

    pass A   stencil + forward DST along x on the own rows          scb_plan_rows_forward
    A2A      At blocks [3][own cols][rows of rank s]  <-  rank s    grouped isend/irecv (NCCL all-to-all over NVLink; gloo in CI)
    pass B   forward DST along y, / eigenvalues, inverse DST        scb_plan_cols
    A2A      Ct blocks [3][own rows][cols of rank s]  <-  rank s
    pass C   inverse DST along x + compose on the own rows          scb_plan_rows_inverse

The stencil needs no halo exchange: every rank holds the (u8) inputs and reads one extra row either side.
The exact low-frequency refinement sums are completed with one small all-reduce (3 x 8 x ny doubles).
The reference has no multi-GPU path at all; its single-GPU pipeline is
/root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:2105-2135.

torch is used for buffers, pack/unpack copies and the collective: plumbing around the C ABI calls.

Stream contract: the C-ABI passes run on the Context's own stream (scb_stream), which is NOT torch's current stream unless
the caller made it so.  Every torch op and collective of this module is therefore issued under
`torch.cuda.stream(ExternalStream(ctx.stream))`, i.e. on the very stream the passes run on, so the passes, the pack/unpack
copies and the NCCL calls are ordered by the stream itself whatever stream the caller is on.  (On CPU tensors -- the gloo /
emulator tests -- everything is synchronous and the guard is a no-op.)  Between scb_plan_tri_forward and scb_plan_tri_finish
the field lives in the CONTEXT's workspace: run no other plan of the same context in between.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi as capi
from .api import Context, Plan


def split(n: int, parts: int) -> list[int]:
    """Boundaries of `parts` nearly equal contiguous ranges of [0, n): the first n % parts get one more."""
    q, r = divmod(n, parts)
    out = [0]
    for k in range(parts):
        out.append(out[-1] + q + (1 if k < r else 0))
    return out


class ShardedSolve:
    def __init__(self, ctx: Context, plan: Plan, device: torch.device, group=None):
        self.ctx, self.plan, self.device, self.group = ctx, plan, device, group
        # all torch work of this object is issued on the context's stream (see the module docstring)
        self._stream = torch.cuda.ExternalStream(ctx.stream, device=device) if device.type == "cuda" else None
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        g = plan.geometry
        if g.empty:
            raise capi.ScbError(capi.SCB_ERR_INVALID_ARGUMENT, "sharded solve: empty plan")
        self.nx, self.ny = int(g.nx), int(g.ny)
        self.ys, self.xs = split(self.ny, self.world), split(self.nx, self.world)
        self.tri = plan.engine in (capi.ENGINE_TRI, capi.ENGINE_I8)
        if self.tri:
            seg_len, n_segs = C.c_int(), C.c_int()
            n32, n64, nw = C.c_size_t(), C.c_size_t(), C.c_size_t()
            ctx._check(ctx.lib.scb_plan_tri_layout(plan.handle, C.byref(seg_len), C.byref(n_segs), C.byref(n32), C.byref(n64), C.byref(nw)))
            self.seg_len, self.n_segs = seg_len.value, n_segs.value
            self.segs = split(self.n_segs, self.world)  # rank r owns segments [segs[r], segs[r+1])
            self.ys = [min(self.ny, s * self.seg_len) for s in self.segs]
            # ONE exchange buffer: ends32 | ends64 | one W slot per rank.  Every rank fills only its own segments and its own W
            # slot, the rest is zero, so the supports are disjoint and a single integer all-reduce(SUM) combines them exactly
            # (x + 0 == x for any bit pattern; the W partials are summed by the kernel that consumes them).
            self.n32, self.n64, self.nw = n32.value, n64.value, nw.value
            n32p = (self.n32 + 1) // 2 * 2  # keep the float64 part 8-byte aligned
            self.pack = torch.zeros((4 * n32p + 8 * (self.n64 + self.nw * self.world)) // 8, dtype=torch.int64, device=device)
            self._e32 = self.pack.data_ptr()
            self._e64 = self._e32 + 4 * n32p
            self._w0 = self._e64 + 8 * self.n64
            self.exchange_bytes = self.pack.numel() * 8
            self.graph = None
            return
        lkx, lky = C.c_int(), C.c_int()
        ctx._check(ctx.lib.scb_plan_lowk(plan.handle, C.byref(lkx), C.byref(lky)))
        self.lowkx, self.lowky = lkx.value, lky.value
        n = 3 * self.nx * self.ny
        self.At = torch.zeros(n, dtype=torch.float32, device=device)
        self.Ct = torch.zeros(n, dtype=torch.float32, device=device)
        self.lowrows = torch.zeros(3 * self.lowkx * self.ny, dtype=torch.float64, device=device)
        self.lowspec = torch.zeros(3 * self.lowkx * self.lowky, dtype=torch.float32, device=device)

    # -- exchanges ---------------------------------------------------------------------------------
    def _all_to_all(self, send: list[torch.Tensor]) -> list[torch.Tensor]:
        recv_shapes = self._recv_shapes
        recv = [torch.empty(s, dtype=torch.float32, device=self.device) for s in recv_shapes]
        if self.world == 1:
            recv[0].copy_(send[0])
            return recv
        # grouped send/recv pairs = NCCL's all-to-all (ncclGroupStart .. ncclSend/ncclRecv .. ncclGroupEnd);
        # the same call works over gloo, which has no alltoall primitive
        recv[self.rank].copy_(send[self.rank])
        peer = (lambda s: dist.get_global_rank(self.group, s)) if self.group is not None else (lambda s: s)
        ops = []
        for s in range(self.world):
            if s == self.rank:
                continue
            ops.append(dist.P2POp(dist.isend, send[s], peer(s), group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv[s], peer(s), group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return recv

    def exchange_rows_to_cols(self):
        """At [3][nx][ny]: own rows (y range) of every column  ->  every row of the own columns."""
        r, xs, ys = self.rank, self.xs, self.ys
        At3 = self.At.view(3, self.nx, self.ny)
        send = [At3[:, xs[s] : xs[s + 1], ys[r] : ys[r + 1]].contiguous() for s in range(self.world)]
        self._recv_shapes = [(3, xs[r + 1] - xs[r], ys[s + 1] - ys[s]) for s in range(self.world)]
        recv = self._all_to_all(send)
        for s in range(self.world):
            At3[:, xs[r] : xs[r + 1], ys[s] : ys[s + 1]].copy_(recv[s])
        if self.world > 1:  # low-frequency row sums: every rank filled only its own y range
            dist.all_reduce(self.lowrows, group=self.group)

    def exchange_cols_to_rows(self):
        """Ct [3][ny][nx]: own columns of every row  ->  every column of the own rows."""
        r, xs, ys = self.rank, self.xs, self.ys
        Ct3 = self.Ct.view(3, self.ny, self.nx)
        send = [Ct3[:, ys[s] : ys[s + 1], xs[r] : xs[r + 1]].contiguous() for s in range(self.world)]
        self._recv_shapes = [(3, ys[r + 1] - ys[r], xs[s + 1] - xs[s]) for s in range(self.world)]
        recv = self._all_to_all(send)
        for s in range(self.world):
            Ct3[:, ys[r] : ys[r + 1], xs[s] : xs[s + 1]].copy_(recv[s])

    # -- the solve ---------------------------------------------------------------------------------
    # bench hook: with `time_collectives` set, CUDA events bracket the exchange of every run(); collective_ms() averages them
    time_collectives = False

    def _mark(self, start=None):
        if not self.time_collectives or self._stream is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(self._stream)
        if start is None:
            return e
        self._spans = getattr(self, "_spans", []) + [(start, e)]
        return None

    def collective_ms(self):
        spans = getattr(self, "_spans", [])
        self._spans = []
        return sum(a.elapsed_time(b) for a, b in spans) / len(spans) if spans else None

    def _on_ctx_stream(self):
        return torch.cuda.stream(self._stream) if self._stream is not None else contextlib.nullcontext()

    def run(self, src_view, dst_view, blend_view) -> None:
        """All images device resident (ScbImage views).  On return (order of the context's stream) the own interior rows
        of `blend` are solved; blend must already hold a copy of dst (or alias it).  Inputs written on another stream
        must be complete (or ordered by an event) before the call, as for any scb_plan_execute(DEVICE)."""
        with self._on_ctx_stream():
            self._run(src_view, dst_view, blend_view)

    def _run(self, src_view, dst_view, blend_view) -> None:
        lib, ph, r = self.ctx.lib, self.plan.handle, self.rank
        chk = self.ctx._check
        if self.tri:
            s0, s1 = self.segs[r], self.segs[r + 1]
            if self.world > 1:
                self.pack.zero_()  # the other ranks' parts of the previous solve
            chk(lib.scb_plan_tri_forward(ph, C.byref(src_view), C.byref(dst_view), capi.MEM_DEVICE, s0, s1, self._e32, self._e64, self._w0 + 8 * self.nw * r))
            if self.world > 1:
                ev = self._mark()
                dist.all_reduce(self.pack, group=self.group)
                self._mark(ev)
            chk(lib.scb_plan_tri_finish_slots(ph, C.byref(blend_view), capi.MEM_DEVICE, s0, s1, self._e32, self._e64, self._w0, self.world))
            return
        y0, y1, x0, x1 = self.ys[r], self.ys[r + 1], self.xs[r], self.xs[r + 1]
        self.lowrows.zero_()
        chk(lib.scb_plan_rows_forward(ph, C.byref(src_view), C.byref(dst_view), capi.MEM_DEVICE, y0, y1, self.At.data_ptr(), self.lowrows.data_ptr()))
        self.exchange_rows_to_cols()
        chk(lib.scb_plan_lowfreq_finish(ph, self.lowrows.data_ptr(), self.lowspec.data_ptr()))
        chk(lib.scb_plan_cols(ph, x0, x1, self.At.data_ptr(), self.Ct.data_ptr(), self.lowspec.data_ptr()))
        self.exchange_cols_to_rows()
        chk(lib.scb_plan_rows_inverse(ph, self.Ct.data_ptr(), C.byref(blend_view), capi.MEM_DEVICE, y0, y1))

    def capture(self, src_view, dst_view, blend_view) -> None:
        """Captures one whole sharded solve -- the C-ABI passes AND the NCCL exchange between them -- into a CUDA graph on the
        context's stream; run_graph() then replays it with a single launch and no host round trip per phase.  Call after at least
        one plain run() (workspace growth, NCCL connection set-up and table builds cannot be captured)."""
        if self._stream is None:
            raise capi.ScbError(capi.SCB_ERR_UNSUPPORTED, "CUDA graphs need a CUDA device")
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            self._run(src_view, dst_view, blend_view)
        self.graph = g

    def run_graph(self) -> None:
        self.graph.replay()

    def gather_rows(self, blend: torch.Tensor) -> None:
        """Make every rank's `blend` (H,W,3 u8 tensor) complete: all-gather the solved interior row slabs."""
        if self.world == 1:
            return
        with self._on_ctx_stream():
            self._gather_rows(blend)

    def _gather_rows(self, blend: torch.Tensor) -> None:
        g = self.plan.geometry
        x0, x1 = g.rx + 1, g.rx + g.w - 1
        slabs = [torch.empty((self.ys[s + 1] - self.ys[s], x1 - x0, 3), dtype=torch.uint8, device=self.device) for s in range(self.world)]
        mine = blend[g.ry + 1 + self.ys[self.rank] : g.ry + 1 + self.ys[self.rank + 1], x0:x1].contiguous()
        dist.all_gather(slabs, mine, group=self.group) if len({tuple(s.shape) for s in slabs}) == 1 else self._uneven_gather(slabs, mine)
        for s in range(self.world):
            blend[g.ry + 1 + self.ys[s] : g.ry + 1 + self.ys[s + 1], x0:x1].copy_(slabs[s])

    def _uneven_gather(self, slabs, mine):
        for s in range(self.world):
            if s == self.rank:
                slabs[s].copy_(mine)
            dist.broadcast(slabs[s], src=dist.get_global_rank(self.group, s) if self.group is not None else s, group=self.group)
