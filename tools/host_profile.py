"""cProfile of the host side of eager S1 steps (forward + pooling + backward): where the ~0.5 ms per step of host time go.
usage: host_profile.py [steps]"""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic

v, f, r, a, t = 6890, 544, 5, 8, 96
dev = "cuda:0"
wn, ws, b = S.make_weights(t, r, a, f, seed=0)
x = S.make_signal(v, f, seed=1).to(dev).requires_grad_(True)
bary = S.make_bary(v, r, a, seed=1).to(dev)
gz = torch.randn((v, t)).to(dev)
layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028)
layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
layer = layer.to(dev)
amp = AngularMaxPooling()


def step():
    x.grad = None
    for p in layer.parameters():
        p.grad = None
    amp(layer([x, bary])).backward(gz)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
    torch.cuda.synchronize()
pr.disable()
out = io.StringIO()
st = pstats.Stats(pr, stream=out)
st.sort_stats("tottime").print_stats(28)
print(out.getvalue().replace(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "."))
