"""Backward (tensor-core GEMMs, GCB_BWD_PREC = fp32 | tf32x3 | f16x3) error at the FAUST shape behind an exact forward (CUDA-core fp32), so that
both sides select the same rotations and ReLU masks; compared with autograd over the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
def nerr(x, y): return ((x.detach().cpu().double() - y.double()).abs().max() / y.double().abs().max()).item()
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = O.make_weights(t, r, a, f, seed=0)
prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.028))).float()
torch.set_num_threads(os.cpu_count())
for kind in sys.argv[1:] or ["shot", "randn"]:
    signal = O.make_signal(v, f, seed=1, kind=kind); bary = O.make_bary(v, r, a, seed=1)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2))
    ref = O.conv_amp_forward_backward(signal.double(), bary, wn.double(), ws.double(), b.double(), prior.double(), gz.double(), "relu", 1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision="ffma")
    layer.backward_precision = os.environ.get("GCB_BWD_PREC", "fp32")
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b}); layer = layer.cuda()
    x = signal.cuda().requires_grad_(True)
    z = AngularMaxPooling()(layer([x, bary.cuda()])); z.backward(gz.cuda())
    print(f"{kind}: z {nerr(z, ref['z']):.2e} d_signal {nerr(x.grad, ref['d_signal']):.2e} d_w {nerr(layer._template_neighbor_weights.grad, ref['d_w_neighbor']):.2e} "
          f"d_w_self {nerr(layer._template_self_weights.grad, ref['d_w_self']):.2e} d_bias {nerr(layer._bias.grad, ref['d_bias']):.2e}", flush=True)
