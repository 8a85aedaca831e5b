"""Times BatchNorm1d -> Dropout between two convolutions: geoconv_b200's fused chain against torch's two modules,
forward + backward, CUDA events, eager launches and CUDA-graph replays.   python tools/block_chain_bench.py"""
import json
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoconv_b200.pytorch.layers.batch_norm_dropout import BatchNormDropout  # noqa: E402


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def graphed(step):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    return g.replay


def main():
    for v, t in [(6890, 96), (6890, 384)]:
        z = torch.randn(v, t, device="cuda", requires_grad=True)
        g = torch.randn(v, t, device="cuda")
        ours = BatchNormDropout(t, p=0.2).cuda()
        ref = nn.Sequential(nn.BatchNorm1d(t), nn.Dropout(0.2)).cuda()

        def step(m):
            def f():
                z.grad = None
                m.zero_grad(set_to_none=True)
                m(z).backward(g)
            return f
        row = {"shape": [v, t], "ours_eager_us": round(timed(step(ours)), 1), "torch_eager_us": round(timed(step(ref)), 1)}
        row["ours_graph_us"] = round(timed(graphed(step(ours))), 1)
        row["torch_graph_us"] = round(timed(graphed(step(ref))), 1)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
