"""fp32-mode (3xTF32) forward error at the FAUST shape against a float64 evaluation, for a signal kind."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
def nerr(x, y): return ((x.detach().cpu().double() - y.double()).abs().max() / y.double().abs().max()).item()
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = O.make_weights(t, r, a, f, seed=0)
prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.028))).float()
torch.set_num_threads(os.cpu_count())
for kind in sys.argv[1:] or ["shot", "randn"]:
    signal = O.make_signal(v, f, seed=1, kind=kind); bary = O.make_bary(v, r, a, seed=1)
    ref64 = O.conv_forward(signal.double(), bary, wn.double(), ws.double(), b.double(), prior.double(), "relu", 1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision="fp32")
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b}); layer = layer.cuda()
    with torch.no_grad():
        y = layer([signal.cuda(), bary.cuda()])
    print(f"ksplit={os.environ.get('GCB_TC_KSPLIT', 'auto')} {kind}: forward normalised error vs float64 {nerr(y, ref64):.3e}", flush=True)
