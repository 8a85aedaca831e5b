"""Three S1 steps (forward + AngularMaxPooling + backward), eager launches, for ncu captures of the third one:
    python tools/step_once.py [precision] [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
v, f, r, a, t = 6890, 544, 5, 8, 96
dev = "cuda:0"
wn, ws, b = S.make_weights(t, r, a, f, seed=0)
layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision=prec)
layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
layer = layer.to(dev)
x = S.make_signal(v, f, seed=1).to(dev).requires_grad_(True)
bary = S.make_bary(v, r, a, seed=1).to(dev)
gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2)).to(dev)
amp = AngularMaxPooling()
for _ in range(n_steps):
    x.grad = None
    for p in layer.parameters():
        p.grad = None
    z = amp(layer([x, bary]))
    z.backward(gz)
torch.cuda.synchronize()
print("ok", float(z.abs().max()), float(x.grad.abs().max()))
