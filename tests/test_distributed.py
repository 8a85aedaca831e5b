"""Multi-GPU host logic: vertex-row shard plans / halo exchange and the data-parallel gradient
all-reduce.  CPU tests run two ranks over gloo; the GPU test emulates the ranks one after the
other on one device (no concurrent spinning ranks on one GPU - B200_PROFILING.md)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import geoconv_oracle as O
from geoconv_b200 import distributed as D
from tests.helpers import TOL_FP32, nerr


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_row_ranges_cover():
    for n, w in ((10, 3), (2_000_000, 8), (5, 8)):
        r = D.row_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_oracle_equals_unsharded(world):
    """Index remapping is exact: running the CPU oracle shard by shard on (local ++ halo) rows
    reproduces the unsharded output rows bit for bit."""
    v, f, r, a, t = 203, 10, 3, 6, 5
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, kind="randn")
    bary = O.make_bary(v, r, a, seed=5, fallback_frac=0.05)
    y_ref = O.conv_forward(signal, bary, wn, ws, b, None, "relu", 1)
    plans = D.build_shard_plans(bary, world)
    assert sum(p.n_local for p in plans) == v
    for p in plans:
        lo, hi = p.ranges[p.rank]
        halo_rows = torch.cat(p.recv_rows) if world > 1 else torch.empty(0, dtype=torch.long)
        full = torch.cat([signal[lo:hi], signal[halo_rows]])
        idx_local = p.bary_local[..., 0].long()
        assert idx_local.min() >= 0 and idx_local.max() < full.shape[0]
        # every remapped index of a live entry points at the same signal row as before (zero-weight entries are
        # kept local: they contribute nothing and must not drag global row 0 into every halo)
        live = (bary[lo:hi, ..., 1] != 0).reshape(-1)
        assert torch.equal(full[idx_local.reshape(-1)][live], signal[bary[lo:hi, ..., 0].long().reshape(-1)][live])
        assert bool((idx_local.reshape(-1)[~live] == 0).all())
        # interpolation + contraction see identical operands => identical output rows
        interp = O.signal_retrieval(full, p.bary_local)
        assert torch.equal(interp, O.signal_retrieval(signal, bary)[lo:hi])
        # send lists mirror receive lists
        for q in range(world):
            assert torch.equal(plans[q].send_rows[p.rank] + p.ranges[q][0], p.recv_rows[q])
    assert y_ref.shape[0] == v


def test_interior_rows_of_a_local_mesh():
    from geoconv_b200 import synthetic as S
    bary = S.make_bary_local(1000, 2, 4, window=40, seed=1)
    for p_ in D.build_shard_plans(bary, 4):
        lo, hi = D.interior_rows(p_)
        idx = p_.bary_local[..., 0].reshape(p_.n_local, -1)
        assert hi - lo >= p_.n_local - 2 * 40 - 2
        assert int(idx[lo:hi].max()) < p_.n_local                         # interior rows reference local rows only
        if lo > 0:
            assert int(idx[lo - 1].max()) >= p_.n_local                    # the run cannot be extended
    one = D.build_shard_plans(bary, 1)[0]
    assert D.interior_rows(one) == (0, 1000)


def _worker(rank, world, port, v, f):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, a = 2, 4
        signal = O.make_signal(v, f, kind="randn")
        bary = O.make_bary(v, r, a, seed=9)
        plan = D.build_shard_plans(bary, world)[rank]
        lo, hi = plan.ranges[rank]
        local = signal[lo:hi].clone().requires_grad_(True)
        full = D.halo_exchange(local, plan)
        halo_rows = torch.cat(plan.recv_rows)
        assert torch.equal(full.detach(), torch.cat([signal[lo:hi], signal[halo_rows]]))
        # backward: d/dX of sum_i c_i * full_i where c depends on the GLOBAL row id
        glob_ids = torch.cat([torch.arange(lo, hi), halo_rows]).double()
        weight = (1.0 + glob_ids * (rank + 1)).unsqueeze(1).to(full.dtype)
        (full * weight).sum().backward()
        # expected: own rows get (1 + id*(rank+1)); plus for every peer q that holds row id as halo: (1 + id*(q+1))
        expect = (1.0 + torch.arange(lo, hi).double() * (rank + 1))
        all_plans = D.build_shard_plans(bary, world)
        for q in range(world):
            if q != rank:
                rows = all_plans[q].recv_rows[rank]
                expect[rows - lo] += 1.0 + rows.double() * (q + 1)
        assert torch.allclose(local.grad[:, 0].double(), expect, rtol=1e-6)
        # data-parallel gradient all-reduce
        p1 = torch.nn.Parameter(torch.zeros(3, 4))
        p2 = torch.nn.Parameter(torch.zeros(5))
        p1.grad = torch.full((3, 4), float(rank + 1))
        p2.grad = torch.arange(5.0) * (rank + 1)
        D.allreduce_gradients([p1, p2])
        tot = sum(range(1, world + 1))
        assert torch.equal(p1.grad, torch.full((3, 4), float(tot)))
        assert torch.equal(p2.grad, torch.arange(5.0) * tot)
        # a gradient above 1 MiB is reduced in place, beside the flat buffer of the small ones
        p3 = torch.nn.Parameter(torch.zeros(600, 512))
        p3.grad = torch.full((600, 512), float(rank + 1))
        p2.grad = torch.arange(5.0) * (rank + 1)
        D.allreduce_gradients([p3, p2], average=True)
        assert torch.equal(p3.grad, torch.full((600, 512), tot / world))
        assert torch.allclose(p2.grad, torch.arange(5.0) * tot / world)
        # the in-backward reduction object (ops.GradSync): asynchronous all-reduce, joined later
        from geoconv_b200 import ops
        with D.overlapped_grad_sync() as sync:
            assert ops.GRAD_SYNC is sync
            g1, g2 = torch.full((7,), float(rank + 1)), torch.full((2, 3), 2.0 * (rank + 1))
            sync.reduce_early([g1, g2, torch.empty(0)])
            sync.join()
            assert torch.equal(g1, torch.full((7,), float(tot))) and torch.equal(g2, torch.full((2, 3), 2.0 * tot))
        assert ops.GRAD_SYNC is None
        # forward of a row-sharded mesh with the halo exchange overlapped with the interior rows
        vb, fb, rb, ab, tb = 400, 6, 2, 4, 5
        from geoconv_b200 import synthetic as S
        sig_b = O.make_signal(vb, fb, kind="randn")
        bary_b = S.make_bary_local(vb, rb, ab, window=30, seed=3)
        wn, ws, b = O.make_weights(tb, rb, ab, fb, seed=1)
        y_ref = O.conv_forward(sig_b, bary_b, wn, ws, b, None, "relu", 1)
        plan_b = D.build_shard_plans(bary_b, world)[rank]
        lo_b, hi_b = plan_b.ranges[rank]
        def layer(inp, row_offset=0):      # the oracle on a row block: the block's own rows lead the centre term
            x_, b_ = inp
            return O.conv_forward(torch.cat([x_[row_offset:row_offset + b_.shape[0]], x_]),
                                  torch.stack([b_[..., 0] + b_.shape[0], b_[..., 1]], dim=-1), wn, ws, b, None, "relu", 1)
        # the plan built from this rank's rows alone (ranks exchange their needs) equals the whole-mesh plan
        mine = D.build_local_shard_plan(bary_b[lo_b:hi_b], vb)
        assert mine.n_halo == plan_b.n_halo and torch.equal(mine.bary_local, plan_b.bary_local)
        assert all(torch.equal(x_, y_) for x_, y_ in zip(mine.recv_rows, plan_b.recv_rows))
        assert all(torch.equal(x_, y_) for x_, y_ in zip(mine.send_rows, plan_b.send_rows))
        sf = D.ShardedForward(layer, plan_b, fb, torch.device("cpu"))
        assert 0 < sf.hi - sf.lo < plan_b.n_local
        assert sf.bary_int.shape[0] + sf.bary_head.shape[0] + sf.bary_tail.shape[0] == plan_b.n_local
        assert int(sf.bary_int[..., 0].max()) < plan_b.n_local
        sf.load(sig_b[lo_b:hi_b])
        for overlap in (True, False):
            assert torch.equal(sf.assemble(*sf(overlap=overlap)), y_ref[lo_b:hi_b])
    finally:
        dist.destroy_process_group()


def test_halo_exchange_and_allreduce_gloo_world2():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 57, 6), nprocs=2, join=True)


@pytest.mark.gpu
def test_sharded_layer_matches_unsharded_on_gpu():
    from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
    dev = "cuda:0"
    v, f, r, a, t = 1500, 64, 3, 8, 32
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, kind="randn")
    bary = O.make_bary(v, r, a, seed=5)
    layer = ConvDirac(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(dev)
    x_full = signal.to(dev).requires_grad_(True)
    y_full = layer([x_full, bary.to(dev)])
    g = torch.randn(y_full.shape, generator=torch.Generator().manual_seed(3)).to(dev)
    (y_full * g).sum().backward()
    ref_dx = x_full.grad.clone()
    ref_dw = layer._template_neighbor_weights.grad.clone()
    layer.zero_grad()
    world = 3
    plans = D.build_shard_plans(bary, world)
    dx = torch.zeros_like(ref_dx)
    for p in plans:                      # ranks emulated one after the other
        lo, hi = p.ranges[p.rank]
        halo_rows = torch.cat(p.recv_rows)
        full = torch.cat([signal[lo:hi], signal[halo_rows]]).to(dev).requires_grad_(True)
        y = layer([full, p.bary_local.to(dev)])
        assert nerr(y, y_full[lo:hi]) < 1e-6
        (y * g[lo:hi]).sum().backward()
        dx[lo:hi] += full.grad[: hi - lo]
        dx[halo_rows.to(dev)] += full.grad[hi - lo:]
    assert nerr(dx, ref_dx) < TOL_FP32
    assert nerr(layer._template_neighbor_weights.grad, ref_dw) < TOL_FP32     # accumulated over shards


@pytest.mark.gpu
def test_peer_allreduce_kernel_single_rank():
    """gcb_peer_allreduce / gcb_peer_allreduce_staged with world = 1 on an ordinary buffer: the kernel's phases, flags
    and device-side epoch run (three calls, then three CUDA-graph replays) and the sum over one rank is the identity.
    The multi-rank behaviour is checked over real NVLink by tools/peer_allreduce_check.py (profiles/)."""
    from geoconv_b200 import _native
    lib = _native.load()
    dev = torch.device("cuda:0")
    sizes = [4096 * 12, 512, 96]
    n = sum(sizes)
    buf = torch.zeros(lib.gcb_peer_allreduce_bytes(n, 1), dtype=torch.uint8, device=dev)
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        g = [torch.randn(k, device=dev) for k in sizes]
        want = [t.clone() for t in g]
        _native.check(lib.gcb_peer_allreduce(g[0].data_ptr(), sizes[0], g[1].data_ptr(), sizes[1], g[2].data_ptr(), sizes[2],
                                             ptrs.data_ptr(), 0, 1, 4, st), "gcb_peer_allreduce")
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(g, want))
    flat = buf[:n * 4].view(torch.float32)
    out = buf[n * 4:2 * n * 4].view(torch.float32)
    src = torch.randn(n, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        flat.copy_(src)
        _native.check(lib.gcb_peer_allreduce_staged(n, ptrs.data_ptr(), 0, 1, 4, side.cuda_stream), "staged")
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert torch.equal(out, src)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        flat.mul_(2.0)
        _native.check(lib.gcb_peer_allreduce_staged(n, ptrs.data_ptr(), 0, 1, 4, torch.cuda.current_stream().cuda_stream),
                      "staged (captured)")
    for rep in range(3):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, flat)
    n_pad = (n + 3) // 4 * 4
    off = (2 * n_pad * 4 + 255) // 256 * 256 + (2 * 1 + 4) * 4
    assert int(buf[off:off + 4].view(torch.int32).item()) == 0          # error word clear
    assert int(buf[off - 16:off - 12].view(torch.int32).item()) == 7      # epoch: 3 + 1 + 3 calls completed
