"""Run the S1 forward a few times (for ncu captures).  usage: run_fwd_once.py [precision] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = O.make_weights(t, r, a, f, seed=0)
layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision=prec)
layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
layer = layer.to("cuda:0")
x = O.make_signal(v, f, seed=1).to("cuda:0").requires_grad_(bwd)
bary = O.make_bary(v, r, a, seed=1).to("cuda:0")
gz = torch.randn(v, t, device="cuda:0")
for _ in range(iters):
    z = AngularMaxPooling()(layer([x, bary]))
    if bwd:
        z.backward(gz)
torch.cuda.synchronize()
print("ok", float(z.sum()))
