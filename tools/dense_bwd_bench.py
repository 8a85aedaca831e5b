"""Times the backward of a loss on Y (every rotation carries a gradient) at the FAUST shape: one pass through G
(gcb_conv_bwd_dense) against one sparse backward per rotation.   python tools/dense_bwd_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic

v, f, r, a, t = 6890, 544, 5, 8, 96
dev = "cuda:0"
wn, ws, b = S.make_weights(t, r, a, f, seed=0)
x = S.make_signal(v, f, seed=1).to(dev).requires_grad_(True)
bary = S.make_bary(v, r, a, seed=1).to(dev)
gy = torch.randn((v, a, t), generator=torch.Generator().manual_seed(2)).to(dev)
for route in ("1", "0"):
    os.environ["GCB_DENSE_BWD"] = route
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(dev)
    y = layer([x, bary])

    def bwd():
        x.grad = None
        for p in layer.parameters():
            p.grad = None
        y.backward(gy, retain_graph=True)
    for _ in range(3):
        bwd()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        bwd()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"route": "one pass through G" if route == "1" else "one sparse backward per rotation",
                      "shape": "S1: V=6890 F=544 T=96 R5xA8, gradient on all 8 rotations",
                      "backward_ms": round(e0.elapsed_time(e1) / 10, 3)}), flush=True)
