"""gcb_peer_allreduce (the library's own NVLink all-reduce kernel) against NCCL, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/peer_allreduce_check.py

Checks: the sums equal NCCL's to rounding, every rank holds the same bits, repeated calls and CUDA-graph replays keep
working (the epoch lives on the device), the error word stays clear; then times both on the FAUST gradient sizes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from geoconv_b200.distributed import PeerAllReduce


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sizes = [(96 * 40 * 544, 96 * 544, 96), (4096, 8, 4), (1000 * 4,)]
    par = PeerAllReduce(n_ctas=12)
    out = []
    ok = True
    for case, sz in enumerate(sizes):
        g = torch.Generator(device=dev).manual_seed(100 * case + rank)
        mine = [torch.randn(n, device=dev, generator=g) for n in sz]
        ref = [t.clone() for t in mine]
        for t in ref:
            dist.all_reduce(t)
        for rep in range(3):                       # repeated calls on the same buffers (epoch 1, 2, 3)
            got = [t.clone() for t in mine]
            par(got)
            torch.cuda.synchronize()
            err = max(((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
                      for a, b in zip(got, ref))
            # every rank must hold the same bits: compare with rank 0's copy
            same = True
            for t in got:
                r0 = t.clone()
                dist.broadcast(r0, src=0)
                same = same and bool(torch.equal(r0, t))
            ok = ok and err < 1e-6 and same
            out.append({"case": case, "rep": rep, "sizes": list(sz), "max_normalised_err_vs_nccl": err,
                        "bit_identical_across_ranks": same, "error_word": par.error_word(sum(sz))})
            ok = ok and out[-1]["error_word"] == 0
    # CUDA graph replay: the epoch is advanced on the device
    sz = sizes[0]
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    src = [torch.randn(n, device=dev, generator=g) for n in sz]
    work = [t.clone() for t in src]
    ref = [t.clone() for t in src]
    for t in ref:
        dist.all_reduce(t)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        par(work)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for w_, s_ in zip(work, src):
            w_.copy_(s_)
        par(work)
    for rep in range(3):
        graph.replay()
        torch.cuda.synchronize()
        err = max(((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
                  for a, b in zip(work, ref))
        ok = ok and err < 1e-6
        out.append({"graph_replay": rep, "max_normalised_err_vs_nccl": err, "error_word": par.error_word(sum(sz))})
    # timing
    def timed(fn, n=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    bufs = [t.clone() for t in src]
    flat = torch.cat([t.reshape(-1) for t in src])
    timing = {"world": world, "bytes": int(flat.numel() * 4)}
    for n_ctas in (4, 8, 12, 24, 48):
        par.n_ctas = n_ctas
        timing[f"peer_us_{n_ctas}_ctas"] = timed(lambda: par(bufs))
    timing["nccl_flat_us"] = timed(lambda: dist.all_reduce(flat))
    out.append(timing)
    if rank == 0:
        for rec in out:
            print(json.dumps(rec), flush=True)
        print("OK" if ok else "FAILED", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
