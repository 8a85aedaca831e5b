"""Repeat the tensor-core forward many times on the same inputs and count runs that are not bit-identical."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
dev = "cuda:0"
shapes = [(ConvDirac, 2000, 96, 5, 8, 96, 1), (ConvGeodesic, 6890, 544, 5, 8, 96, 1), (ConvGeodesic, 1500, 64, 3, 6, 32, 1),
          (ConvDirac, 900, 128, 5, 8, 128, 2), (ConvGeodesic, 3000, 160, 8, 12, 64, 1)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for cls, v, f, r, a, t, delta in shapes:
    for prec in ("tf32", "tf32x3"):
        torch.manual_seed(0)
        layer = cls(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03, rotation_delta=delta, precision=prec).to(dev)
        x = S.make_signal(v, f, seed=2).to(dev)
        bary = S.make_bary(v, r, a, seed=2).to(dev)
        with torch.no_grad():
            ref = layer([x, bary]).clone()
            bad = 0
            worst = 0.0
            for _ in range(reps):
                y = layer([x, bary])
                if not torch.equal(y, ref):
                    bad += 1
                    worst = max(worst, float((y - ref).abs().max() / ref.abs().max()))
        print(f"{cls.__name__} V{v} F{f} R{r} A{a} T{t} d{delta} {prec}: {bad}/{reps} runs differ, worst normalised diff {worst:.2e}", flush=True)
