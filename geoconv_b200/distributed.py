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
    """Sum (or average) `.grad` of `params` over the process group with a single flat all-reduce."""
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


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
    owner_bounds = torch.tensor([hi for _, hi in ranges])
    plans = []
    needs = []  # needs[p][q] = sorted global rows p needs from q
    for p, (lo, hi) in enumerate(ranges):
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
