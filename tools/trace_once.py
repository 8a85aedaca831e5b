import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
prec = sys.argv[1]
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = O.make_weights(t, r, a, f, seed=0)
layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision=prec)
layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
layer = layer.to("cuda:0")
x = O.make_signal(v, f, seed=1).to("cuda:0"); bary = O.make_bary(v, r, a, seed=1).to("cuda:0")
with torch.no_grad():
    os.environ.pop("GCB_TC_TRACE", None)
    for _ in range(2): layer([x, bary])
    torch.cuda.synchronize()
    os.environ["GCB_TC_TRACE"] = os.environ.get("GCB_TRACE_MODE", "1")
    layer([x, bary]); torch.cuda.synchronize()
