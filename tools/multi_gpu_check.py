"""Real-NCCL check of the two multi-GPU modes, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--v-big 500000]

1. S3 - data parallel over meshes: 3 x (ConvGeodesic + AngularMaxPooling) + Linear, one flat all-reduce of the
   weight gradients; compared with the same batch run on one rank.
2. S4 - one large mesh partitioned by vertex rows: halo exchange of signal rows (all_to_all over NVLink), forward
   of ConvGeodesic 128 -> 128; compared with the unpartitioned forward (bit-exact: every output row sees the
   same arithmetic), plus a backward check at a smaller size.
Rank 0 prints one JSON line per check.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from geoconv_b200 import synthetic as S
from geoconv_b200.distributed import (ShardedForward, allreduce_gradients, build_shard_plans, overlapped_grad_sync,
                                      sharded_layer_forward)
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic


class Net(torch.nn.Module):
    def __init__(self, dims, r, a, n_classes):
        super().__init__()
        self.convs = torch.nn.ModuleList(
            ConvGeodesic(input_shape=[(None, fi), (None, r, a, 3, 2)], amt_templates=fo, template_radius=0.028)
            for fi, fo in zip(dims[:-1], dims[1:]))
        self.amp = AngularMaxPooling()
        self.head = torch.nn.Linear(dims[-1], n_classes)

    def forward(self, x, bary):
        for conv in self.convs:
            x = self.amp(conv([x, bary]))
        return self.head(x)


def nerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--v-big", type=int, default=500000)
    ap.add_argument("--meshes", type=int, default=8)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = []

    # ---------------- S3: data parallel over meshes ----------------
    v, r, a, n_cls = 6890, 5, 8, 32
    dims = [64, 96, 256, 384]
    torch.manual_seed(0)
    net = Net(dims, r, a, n_cls).to(dev)
    meshes = []
    for m in range(args.meshes):
        g = torch.Generator().manual_seed(1000 + m)
        meshes.append((S.make_signal(v, dims[0], seed=100 + m).to(dev), S.make_bary(v, r, a, seed=100 + m).to(dev),
                       torch.randint(0, n_cls, (v,), generator=g).to(dev)))
    params = [p for p in net.parameters()]
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")

    def run(ids):
        for p in params:
            p.grad = None
        for m in ids:
            x, bary, y = meshes[m]
            loss_fn(net(x, bary), y).backward()

    def dp_step():
        run(range(rank, args.meshes, world))
        allreduce_gradients(params)

    dp_step()
    dp = [p.grad.clone() for p in params]
    run(range(args.meshes))                      # the whole batch on this rank
    ref = [p.grad.clone() for p in params]
    err = max(nerr(x, y) for x, y in zip(dp, ref))
    ms = timed(dp_step)
    out.append({"check": "S3 data-parallel weight gradients vs one rank", "world": world, "meshes": args.meshes,
                "max_normalised_err": err, "ok": err < 1e-5, "ms_per_step": ms,
                "vertices_per_s": args.meshes * v / (ms * 1e-3)})

    # ---------------- S1: weight-gradient all-reduce started inside the backward ----------------
    # (the library's own NVLink peer-memory kernel when symmetric memory is available, else NCCL on a low-CTA group)
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling as AMP
    f1, t1 = 544, 96
    torch.manual_seed(3)
    layer1 = ConvGeodesic(input_shape=[(None, f1), (None, r, a, 3, 2)], amt_templates=t1, template_radius=0.028).to(dev)
    x1 = S.make_signal(v, f1, seed=200 + rank).to(dev)
    b1 = S.make_bary(v, r, a, seed=200 + rank).to(dev)
    gz1 = torch.randn((v, t1), generator=torch.Generator().manual_seed(300 + rank)).to(dev)

    def one_step():
        for p in layer1.parameters():
            p.grad = None
        xs = x1.clone().requires_grad_(True)
        AMP()(layer1([xs, b1])).backward(gz1)
        return xs.grad
    dx_plain = one_step()
    want = [p.grad.clone() for p in layer1.parameters()]
    for g_ in want:
        dist.all_reduce(g_)
    with overlapped_grad_sync() as sync:
        for _ in range(3):                     # repeated steps reuse the peer buffers (device-side epoch)
            dx_sync = one_step()
        torch.cuda.synchronize()
        got = [p.grad.clone() for p in layer1.parameters()]
        used_peer = sync.peer_reducer is not None and sync.peer_reducer.calls > 0
        peer_failed = sync.peer_failed
    err = max(nerr(x_, y_) for x_, y_ in zip(got, want))
    flags = torch.tensor([err, nerr(dx_sync, dx_plain)], device=dev, dtype=torch.float64)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    out.append({"check": "S1 weight gradients reduced inside the backward vs NCCL all-reduce of the local gradients",
                "world": world, "peer_memory_kernel": used_peer, "peer_fallback_reason": peer_failed,
                "max_normalised_err": flags[0].item(), "d_signal_unchanged_err": flags[1].item(),
                "ok": flags[0].item() < 1e-6 and flags[1].item() == 0.0})

    # ---------------- S4: vertex-row shards + halo exchange ----------------
    f = t = 128
    for vb, backward in ((20000, True), (args.v_big, False)):
        torch.manual_seed(1)
        # tanh for the gradient check: the shards tile the contraction differently from the whole mesh, the
        # forward then differs in the last bits and a ReLU derivative would flip for pre-activations at ~0
        layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028,
                             activation="tanh" if backward else "relu").to(dev)
        bary = S.make_bary_local(vb, r, a, window=2048, seed=5)
        signal = S.make_signal(vb, f, seed=5, kind="randn")
        plan = build_shard_plans(bary, world)[rank]
        lo, hi = plan.ranges[rank]
        x_loc = signal[lo:hi].to(dev).requires_grad_(backward)
        bary_loc = plan.bary_local.to(dev)
        x_all = signal.to(dev).requires_grad_(backward)
        bary_all = bary.to(dev)
        y_loc = sharded_layer_forward(layer, x_loc, plan, bary_local_device=bary_loc)
        y_all = layer([x_all, bary_all])
        exact = bool(torch.equal(y_loc, y_all[lo:hi]))
        rec = {"check": "S4 row-sharded forward vs unpartitioned", "world": world, "v": vb, "rows_local": hi - lo,
               "halo_rows": plan.n_halo, "bit_exact": exact, "max_normalised_err": nerr(y_loc, y_all[lo:hi])}
        if backward:
            gy = torch.randn(y_all.shape, generator=torch.Generator().manual_seed(9)).to(dev)
            for p in layer.parameters():
                p.grad = None
            y_loc.backward(gy[lo:hi])
            allreduce_gradients(list(layer.parameters()))
            g_sh = [p.grad.clone() for p in layer.parameters()]
            dx_sh = x_loc.grad.clone()
            for p in layer.parameters():
                p.grad = None
            y_all.backward(gy)
            rec["grad_err_weights"] = max(nerr(x, p.grad) for x, p in zip(g_sh, layer.parameters()))
            rec["grad_err_signal"] = nerr(dx_sh, x_all.grad[lo:hi])
            rec["ok"] = (rec["max_normalised_err"] < 1e-5 and rec["grad_err_weights"] < 1e-5 and
                         rec["grad_err_signal"] < 1e-5)
        else:
            with torch.no_grad():
                ms = timed(lambda: sharded_layer_forward(layer, x_loc, plan, bary_local_device=bary_loc))
            rec["ms_forward_sharded"] = ms
            rec["vertices_per_s"] = vb / (ms * 1e-3)
            rec["ok"] = exact
        flags = torch.tensor([1 if rec["ok"] else 0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        rec["ok_all_ranks"] = bool(flags.item())
        out.append(rec)
        del x_all, bary_all, y_all, y_loc
        torch.cuda.empty_cache()

    if rank == 0:
        for rec in out:
            print(json.dumps(rec), flush=True)
    ok = all(rec.get("ok_all_ranks", rec["ok"]) for rec in out)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
