"""One small forward + backward through every tensor-core kernel, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py

Shapes are chosen so that the forward runs as 2-CTA clusters (two rotation groups: k_tc_conv with CL = 2, DSMEM
bulk copies, multicast commits) and the backward takes the route through G (k_bw_build_g, k_tc_gemm_tn, k_tc_gemm)
in the fp16-split and the tf32-split modes; a second shape runs the unpaired forward (CL = 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic


def run(v, f, r, a, t, delta, precision):
    dev = "cuda:0"
    wn, ws, b = S.make_weights(t, r, a, f, seed=0)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                         rotation_delta=delta, precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(dev)
    xs = S.make_signal(v, f, seed=1, kind="randn").to(dev).requires_grad_(True)
    bary = S.make_bary(v, r, a, seed=1, fallback_frac=0.05).to(dev)
    gz = torch.randn(v, t, device=dev)
    z = AngularMaxPooling()(layer([xs, bary]))
    z.backward(gz)
    torch.cuda.synchronize()
    print(f"ok V={v} F={f} R{r}xA{a} T={t} delta={delta} {precision}: |z|={float(z.abs().max()):.4f} "
          f"|dx|={float(xs.grad.abs().max()):.4f}", flush=True)


if __name__ == "__main__":
    for prec in ("fp32", "tf32x3", "tf32"):
        run(300, 64, 2, 8, 96, 1, prec)      # 8 rotations x 96 templates: two rotation groups -> cluster pairs
    run(200, 32, 2, 6, 64, 2, "fp32")        # 3 rotations: one group, no cluster
