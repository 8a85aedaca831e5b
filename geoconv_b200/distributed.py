"""Multi-GPU use of the hot path: one process per GPU, `torch.distributed` (NCCL over
NVLink on the box, gloo in CPU tests) for the only two exchanges the path has.

1. Batches of meshes, data parallel: every rank runs whole meshes; weight gradients are summed with
   ONE flat all-reduce per step (`allreduce_gradients`).  The reference has no multi-GPU mode; the
   result equals the single-process sum of per-mesh gradients up to floating-point reassociation.

2. One very large mesh, partitioned by vertex rows (`VertexShardPlan` / `halo_exchange`): rank p owns
   signal / barycentric / output rows [lo_p, hi_p).  The rows its barycentric indices reference on
   other ranks (the halo) are fetched before the layer with one all-to-all of packed rows; indices
   are remapped (exactly, integers) to `local rows ++ halo rows`.  In backward the halo part of
   dSignal is sent back and added at the owner in fixed rank order (deterministic).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------
# data parallel over meshes
# ------------------------------------------------------------------------------------------------
def allreduce_gradients(params, group=None, average: bool = False) -> None:
    """Sum (or average) `.grad` of `params` over the process group: gradients above 1 MiB are reduced in place
    (no staging copy), the small ones travel together in one flat buffer; the collectives of one call are issued
    back to back and complete asynchronously on the collective stream (the compute stream waits, not the host)."""
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    big = [p for p in params if p.grad.numel() * p.grad.element_size() >= (1 << 20) and p.grad.is_contiguous()]
    small = [p for p in params if not any(p is q for q in big)]
    works = [dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True) for p in big]
    flat = None
    if small:
        flat = torch.cat([p.grad.reshape(-1) for p in small])
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    if flat is not None:
        off = 0
        for p in small:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    if average:
        for p in params:
            p.grad /= world


class PeerAllReduce:
    """In-place sum of a few gradient tensors over the ranks with libgeoconv_b200's own NVLink kernel
    (`gcb_peer_allreduce`: two-shot all-reduce on directly mapped peer buffers, include/geoconv_b200.h) instead of
    NCCL.  The peer-mapped staging buffers come from torch's symmetric memory; one buffer per total element count,
    allocated (collectively - every rank must reach the first call together) on first use."""

    def __init__(self, group=None, n_ctas: int = 12):
        self.group = group if group is not None else dist.group.WORLD
        self.n_ctas = n_ctas
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self._bufs: dict[int, tuple] = {}
        self.stream = None
        self.calls = 0

    def _buffer(self, n: int, device):
        hit = self._bufs.get(n)
        if hit is None:
            import torch.distributed._symmetric_memory as symm
            from . import _native
            lib = _native.load()
            nbytes = lib.gcb_peer_allreduce_bytes(n, self.world)
            buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
            buf.zero_()
            hdl = symm.rendezvous(buf, self.group)
            torch.cuda.synchronize(device)
            hdl.barrier()                       # every rank's flags are zero before anybody writes one
            torch.cuda.synchronize(device)
            hit = (buf, hdl, nbytes)
            self._bufs[n] = hit
        return hit

    def error_word(self, n: int) -> int:
        buf, _hdl, nbytes = self._bufs[n]
        n_pad = (n + 3) // 4 * 4
        off = (2 * n_pad * 4 + 255) // 256 * 256 + (2 * self.world + 4) * 4
        return int(buf[off:off + 4].view(torch.int32).item())

    def staging(self, sizes, device):
        """Views of this rank's peer-readable staging area, one per size: a producer that writes its gradients there
        (instead of into private tensors) saves the all-reduce kernel both of its local copies (`reduce_staged`)."""
        n = sum(sizes)
        buf, _hdl, _nbytes = self._buffer(n, device)
        flat = buf[:n * 4].view(torch.float32)
        out, off = [], 0
        for k in sizes:
            out.append(flat[off:off + k])
            off += k
        return out

    def reduce_staged(self, n: int, device, stream=None):
        """All-reduce what `staging` views hold; the sums are in `result(n)` once `stream` has passed the kernel."""
        from . import _native
        lib = _native.load()
        _buf, hdl, _nbytes = self._buffer(n, device)
        st = stream if stream is not None else torch.cuda.current_stream(device)
        with torch.cuda.device(device):
            _native.check(lib.gcb_peer_allreduce_staged(n, hdl.buffer_ptrs_dev, self.rank, self.world, self.n_ctas,
                                                        st.cuda_stream), "gcb_peer_allreduce_staged")
        self.calls += 1

    def result(self, n: int, device):
        buf, _hdl, _nbytes = self._buffer(n, device)
        return buf[n * 4:2 * n * 4].view(torch.float32)

    def __call__(self, tensors, stream=None):
        """Sum `tensors` (<= 3 contiguous fp32 CUDA tensors, each a multiple of 4 elements) over the ranks in place,
        on `stream` (default: the current stream)."""
        from . import _native
        lib = _native.load()
        ts = [t for t in tensors if t.numel()]
        if not 1 <= len(ts) <= 3:
            raise ValueError("PeerAllReduce takes one to three non-empty tensors")
        for t in ts:
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() % 4 == 0):
                raise ValueError("PeerAllReduce needs contiguous fp32 CUDA tensors whose sizes are multiples of 4")
        n = sum(t.numel() for t in ts)
        dev = ts[0].device
        _buf, hdl, _nbytes = self._buffer(n, dev)
        ts = ts + [None] * (3 - len(ts))
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            _native.check(lib.gcb_peer_allreduce(ts[0].data_ptr(), ts[0].numel(),
                                                 ts[1].data_ptr() if ts[1] is not None else None,
                                                 ts[1].numel() if ts[1] is not None else 0,
                                                 ts[2].data_ptr() if ts[2] is not None else None,
                                                 ts[2].numel() if ts[2] is not None else 0,
                                                 hdl.buffer_ptrs_dev, self.rank, self.world, self.n_ctas, st.cuda_stream),
                          "gcb_peer_allreduce")
        self.calls += 1


class overlapped_grad_sync:
    """Context manager: inside it, every geoconv_b200 convolution whose backward takes the AMP-sparse route
    all-reduces its weight gradients from INSIDE the backward, between the weight-gradient stage and the dSignal GEMM
    (ops.GradSync), so the collective runs beside that GEMM instead of after the step.  Valid when every rank runs
    ONE backward per step and layer (the sum of per-rank gradients is what is reduced); with several meshes per rank
    accumulate locally and call `allreduce_gradients` once instead.  `synced(params)` tells which parameters need no
    further reduction."""

    def __init__(self, group=None, max_ctas: int | None = 8, peer: bool = True):
        from . import ops
        self._ops = ops
        reducer = None
        if peer and dist.is_initialized() and dist.get_backend(group) == "nccl" and dist.get_world_size(group) > 1:
            # the library's own all-reduce kernel over NVLink peer memory; NCCL remains the fallback
            try:
                reducer = PeerAllReduce(group, n_ctas=12)
            except Exception:  # noqa: BLE001 - no symmetric memory in this build / on this machine
                reducer = None
        if group is None and max_ctas and dist.is_initialized() and dist.get_backend() == "nccl":
            group = _low_cta_group(max_ctas)
        self.sync = ops.GradSync(group, peer_reducer=reducer)
        self._prev = None

    def __enter__(self):
        self._prev = self._ops.GRAD_SYNC
        self._ops.GRAD_SYNC = self.sync
        self._yield(1)
        return self.sync

    def __exit__(self, *exc):
        self.sync.join()
        self._ops.GRAD_SYNC = self._prev
        self._yield(0)
        return False

    @staticmethod
    def _yield(on):
        try:
            from . import _native
            _native.load().gcb_yield_sms(on)     # the dSignal GEMM leaves SMs to the collective
        except Exception:  # noqa: BLE001 - CPU-only process (gloo tests): nothing to tell
            pass


_LOW_CTA_GROUPS: dict = {}


def _low_cta_group(max_ctas: int):
    """A process group over all ranks whose NCCL kernels use at most `max_ctas` CTAs: the collective that runs beside
    the dSignal GEMM has to fit on the SMs that GEMM leaves free (13 of 148 at the FAUST shape)."""
    if max_ctas not in _LOW_CTA_GROUPS:
        group = None
        try:
            # high-priority stream: the collective's few CTAs are placed as soon as an SM has room instead of
            # queueing behind the compute grid that was launched beside it
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            opts.config.max_ctas = int(max_ctas)
            opts.config.min_ctas = 1
            group = dist.new_group(backend="nccl", pg_options=opts)
        except Exception:  # noqa: BLE001 - option not available in this build: the default group
            group = None
        _LOW_CTA_GROUPS[max_ctas] = group
    return _LOW_CTA_GROUPS[max_ctas]


# ------------------------------------------------------------------------------------------------
# vertex-row sharding with halo exchange
# ------------------------------------------------------------------------------------------------
def row_ranges(n_rows: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [lo, hi) row ranges."""
    base, rem = divmod(n_rows, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


@dataclass
class VertexShardPlan:
    """What rank `rank` needs to run the layer on its rows of a row-partitioned mesh."""
    rank: int
    world: int
    ranges: list                      # [(lo, hi)] per rank
    bary_local: torch.Tensor          # (V_local, R, A, 3, 2): indices remapped to local ++ halo rows
    recv_rows: list                   # per peer: global row ids (sorted) this rank receives from it
    send_rows: list                   # per peer: LOCAL row ids this rank must send to it
    n_local: int
    n_halo: int


def build_shard_plans(bary: torch.Tensor, world: int) -> list[VertexShardPlan]:
    """Partition a whole-mesh barycentric tensor (V,R,A,3,2) by vertex rows (host-side, once per mesh).

    Index arithmetic is done on integers; the remapped indices are written back in the tensor's own
    floating dtype exactly as the reference stores them (exact below 2**24 rows in fp32)."""
    v = bary.shape[0]
    ranges = row_ranges(v, world)
    idx_all = bary[..., 0].to(torch.float32).long()
    # Zero-weight entries (the (0, 0) fallback of barycentric_coordinates.py:69 and anything else with weight 0)
    # contribute nothing to a finite signal: they are pointed at the shard's own first row instead of pulling global
    # row 0 into every rank's halo.  (With a non-finite signal 0 * inf differs from the unpartitioned result.)
    dead = bary[..., 1] == 0
    owner_bounds = torch.tensor([hi for _, hi in ranges])
    plans = []
    needs = []  # needs[p][q] = sorted global rows p needs from q
    for p, (lo, hi) in enumerate(ranges):
        idx_all[lo:hi][dead[lo:hi]] = lo
        idx = idx_all[lo:hi]
        uniq = torch.unique(idx)
        remote = uniq[(uniq < lo) | (uniq >= hi)]
        owner = torch.bucketize(remote, owner_bounds, right=True)
        needs.append([remote[owner == q] for q in range(world)])
    for p, (lo, hi) in enumerate(ranges):
        recv = needs[p]
        halo = torch.cat(recv) if world > 1 else torch.empty(0, dtype=torch.long)
        n_local = hi - lo
        # global row -> position in (local rows ++ halo rows ordered by peer, then by row id)
        lut = torch.full((v,), -1, dtype=torch.long)
        lut[lo:hi] = torch.arange(n_local)
        lut[halo] = n_local + torch.arange(halo.numel())
        local = bary[lo:hi].clone()
        local[..., 0] = lut[idx_all[lo:hi]].to(bary.dtype)
        send = [needs[q][p] - lo for q in range(world)]
        plans.append(VertexShardPlan(rank=p, world=world, ranges=ranges, bary_local=local, recv_rows=recv,
                                     send_rows=send, n_local=n_local, n_halo=int(halo.numel())))
    return plans


def build_local_shard_plan(bary_rows: torch.Tensor, v_total: int, group=None) -> VertexShardPlan:
    """The plan of THIS rank only, from this rank's own rows of the barycentric tensor (global indices): every rank
    finds the remote rows it references and the ranks exchange those lists (one all_gather of index lists), so no
    rank ever holds the whole mesh.  Same result as `build_shard_plans(whole, world)[rank]`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ranges = row_ranges(v_total, world)
    lo, hi = ranges[rank]
    if bary_rows.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns rows [{lo}, {hi}) but was given {bary_rows.shape[0]} barycentric rows")
    idx = bary_rows[..., 0].to(torch.float32).long()
    idx[bary_rows[..., 1] == 0] = lo                       # zero-weight entries stay local (see build_shard_plans)
    uniq = torch.unique(idx)
    remote = uniq[(uniq < lo) | (uniq >= hi)]
    owner = torch.bucketize(remote, torch.tensor([h_ for _, h_ in ranges]), right=True)
    recv = [remote[owner == q] for q in range(world)]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, [r_.tolist() for r_ in recv], group=group)
        send = [torch.tensor(gathered[q][rank], dtype=torch.long) - lo for q in range(world)]
    else:
        send = [torch.empty(0, dtype=torch.long)]
    halo = torch.cat(recv) if world > 1 else torch.empty(0, dtype=torch.long)
    n_local = hi - lo
    # global row -> position in (local rows ++ halo rows): local rows by offset, halo rows by search in the sorted list
    local = bary_rows.clone()
    is_local = (idx >= lo) & (idx < hi)
    pos = torch.where(is_local, idx - lo, torch.zeros_like(idx))
    if halo.numel():
        order = torch.argsort(halo)
        where = torch.searchsorted(halo[order], idx.clamp(min=0))
        where = where.clamp(max=halo.numel() - 1)
        pos = torch.where(is_local, pos, n_local + order[where])
    local[..., 0] = pos.to(bary_rows.dtype)
    return VertexShardPlan(rank=rank, world=world, ranges=ranges, bary_local=local, recv_rows=recv, send_rows=send,
                           n_local=n_local, n_halo=int(halo.numel()))


def _pack(src: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    if src.is_cuda:
        from . import ops
        return ops.pack_rows(src, rows)
    return src.index_select(0, rows.to(src.device))        # gloo / CPU tests of the exchange logic


class _HaloExchange(torch.autograd.Function):
    @staticmethod
    def forward(ctx, signal_local, plan: VertexShardPlan, group):
        world, f = plan.world, signal_local.shape[1]
        dev = signal_local.device
        send_counts = [int(r.numel()) for r in plan.send_rows]
        recv_counts = [int(r.numel()) for r in plan.recv_rows]
        send_idx = torch.cat(plan.send_rows).to(dev) if sum(send_counts) else torch.empty(0, dtype=torch.long, device=dev)
        send_buf = _pack(signal_local.contiguous(), send_idx) if send_idx.numel() else signal_local.new_empty((0, f))
        recv_buf = signal_local.new_empty((sum(recv_counts), f))
        if world > 1:
            dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=recv_counts,
                                   input_split_sizes=send_counts, group=group)
        ctx.plan, ctx.group = plan, group
        ctx.send_idx, ctx.send_counts, ctx.recv_counts = send_idx, send_counts, recv_counts
        return torch.cat([signal_local, recv_buf], dim=0)

    @staticmethod
    def backward(ctx, grad_full):
        plan = ctx.plan
        n_local = plan.n_local
        grad_local = grad_full[:n_local].clone()
        grad_halo = grad_full[n_local:].contiguous()
        back = grad_full.new_empty((sum(ctx.send_counts), grad_full.shape[1]))
        if plan.world > 1:
            dist.all_to_all_single(back, grad_halo, output_split_sizes=ctx.send_counts,
                                   input_split_sizes=ctx.recv_counts, group=ctx.group)
        # contributions arrive grouped by peer rank; add them in that fixed order
        off = 0
        for q, cnt in enumerate(ctx.send_counts):
            if cnt:
                grad_local.index_add_(0, ctx.send_idx[off:off + cnt], back[off:off + cnt])
                off += cnt
        return grad_local, None, None


def interior_rows(plan: VertexShardPlan) -> tuple[int, int]:
    """[lo, hi): the longest run of consecutive local rows whose template vertices only reference LOCAL signal rows.
    With a locality-preserving vertex order these are all rows but a thin band at either end of the shard; they can be
    convolved while the halo rows are still in flight."""
    idx = plan.bary_local[..., 0].reshape(plan.n_local, -1)
    local_only = (idx.max(dim=1).values < plan.n_local)
    if not bool(local_only.any()):
        return 0, 0
    # longest run of True
    z = torch.cat([torch.tensor([False]), local_only, torch.tensor([False])]).to(torch.int8)
    d = z[1:] - z[:-1]
    starts = (d == 1).nonzero().flatten()
    ends = (d == -1).nonzero().flatten()
    k = int(torch.argmax(ends - starts))
    return int(starts[k]), int(ends[k])


class ShardedForward:
    """Forward of one geoconv_b200 convolution on this rank's rows of a row-partitioned mesh with the halo exchange
    OVERLAPPED with the interior rows:

      1. the rows peers asked for are packed and the all-to-all is started asynchronously; it writes the halo rows
         straight behind the local rows of one preallocated signal buffer (no concatenation);
      2. meanwhile the layer runs on the interior rows - those whose template vertices reference local rows only -
         against the local part of the buffer;
      3. when the collective has finished the layer runs on the remaining boundary rows against the whole buffer.

    `__call__` returns the row blocks (Y_head, Y_interior, Y_tail) in local row order; `assemble` concatenates them
    (a copy; a consumer that accepts row blocks does not need it).  The layer is called with `row_offset`, so that
    a block's centre term reads its own vertices.  Forward only: training uses `sharded_layer_forward`."""

    def __init__(self, layer, plan: VertexShardPlan, n_features: int, device, group=None, max_ctas: int | None = 4):
        if group is None and max_ctas and plan.world > 1 and dist.is_initialized() and dist.get_backend() == "nccl":
            # the exchange runs BESIDE the interior rows' kernels: few CTAs on a high-priority stream, or its CTAs
            # queue behind the compute grid while their peers spin (N = 8: 11.1 ms overlapped against 7.3 ms serial
            # with the default communicator, profiles/r02_notes.md)
            group = _low_cta_group(max_ctas)
        self.layer, self.plan, self.group = layer, plan, group
        self.lo, self.hi = interior_rows(plan)
        bl = plan.bary_local
        self.bary_int = bl[self.lo:self.hi].contiguous().to(device)
        self.bary_head = bl[:self.lo].contiguous().to(device)       # boundary rows: they reference halo rows
        self.bary_tail = bl[self.hi:].contiguous().to(device)
        self.full = torch.empty((plan.n_local + plan.n_halo, n_features), dtype=torch.float32, device=device)
        self.send_counts = [int(r.numel()) for r in plan.send_rows]
        self.recv_counts = [int(r.numel()) for r in plan.recv_rows]
        self.send_idx = torch.cat(plan.send_rows).to(device) if sum(self.send_counts) else None
        self.halo_bytes = plan.n_halo * n_features * 4

    def load(self, signal_local: torch.Tensor) -> None:
        self.full[:self.plan.n_local].copy_(signal_local)

    def __call__(self, overlap: bool = True):
        p = self.plan
        local = self.full[:p.n_local]
        work = None
        if p.world > 1 and (p.n_halo or sum(self.send_counts)):
            send = _pack(local, self.send_idx) if self.send_idx is not None else local.new_empty((0, local.shape[1]))
            work = dist.all_to_all_single(self.full[p.n_local:], send, output_split_sizes=self.recv_counts,
                                          input_split_sizes=self.send_counts, group=self.group, async_op=True)
            if not overlap:
                work.wait()
                work = None
        # (row_offset: the block's rows are vertices lo .. hi of the buffer - its centre term reads those rows)
        y_int = self.layer([local, self.bary_int], row_offset=self.lo) if self.hi > self.lo else None
        if work is not None:
            work.wait()                      # the compute stream waits for the collective, the host does not
        y_head = self.layer([self.full, self.bary_head], row_offset=0) if self.lo else None
        y_tail = self.layer([self.full, self.bary_tail], row_offset=self.hi) if self.hi < p.n_local else None
        return y_head, y_int, y_tail

    def assemble(self, y_head, y_int, y_tail):
        return torch.cat([y_ for y_ in (y_head, y_int, y_tail) if y_ is not None])


def halo_exchange(signal_local: torch.Tensor, plan: VertexShardPlan, group=None) -> torch.Tensor:
    """(V_local, F) -> (V_local + n_halo, F): local rows followed by the halo rows this rank's
    barycentric indices reference.  Differentiable (halo gradients return to their owners)."""
    return _HaloExchange.apply(signal_local, plan, group)


def sharded_layer_forward(layer, signal_local: torch.Tensor, plan: VertexShardPlan, group=None,
                          bary_local_device: torch.Tensor | None = None):
    """Run a geoconv_b200 ConvIntrinsic on this rank's rows: halo exchange, then the layer with
    V_out = local rows and V_src = local + halo rows."""
    full = halo_exchange(signal_local, plan, group)
    bary = bary_local_device if bary_local_device is not None else plan.bary_local.to(signal_local.device)
    return layer([full, bary])
