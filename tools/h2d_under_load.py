"""Pinned H2D bandwidth of one 24 MB copy: alone, and while the S1 step (CUDA graph) replays on another stream."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
dev = "cuda:0"
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = S.make_weights(t, r, a, f, seed=0)
x = S.make_signal(v, f, seed=1).to(dev).requires_grad_(True)
bary = S.make_bary(v, r, a, seed=1).to(dev)
gz = torch.randn((v, t)).to(dev)
layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028)
layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
layer = layer.to(dev); amp = AngularMaxPooling()
def step():
    x.grad = None
    for p in layer.parameters(): p.grad = None
    amp(layer([x, bary])).backward(gz)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
n = int(24.25 * (1 << 20))
src = torch.empty(n, dtype=torch.uint8).pin_memory(); dst = torch.empty(n, dtype=torch.uint8, device=dev)
copy = torch.cuda.Stream()
def copies(k, load):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if load:
        for _ in range(k + 4): g.replay()
    with torch.cuda.stream(copy):
        e0.record(copy)
        for _ in range(k): dst.copy_(src, non_blocking=True)
        e1.record(copy)
    torch.cuda.synchronize()
    return round(n * k / e0.elapsed_time(e1) / 1e6, 1)
for load in (False, True, False, True):
    print(json.dumps({"copy": "24.25 MB pinned H2D x 10", "S1 step replaying beside it": load, "GB_per_s": copies(10, load)}), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for load in (False, True):
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): g.replay()
    if load:
        with torch.cuda.stream(copy):
            for _ in range(20): dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"S1 step replay x 20, ms per step": round(e0.elapsed_time(e1) / 20, 4), "copies beside it": load}), flush=True)
