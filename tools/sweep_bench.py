"""Throughput of the hot path over BASELINE config 3 (kernel-variant sweep on the 6890-vertex shape) and the layer
shapes of config 4 (3-layer net): fwd+bwd of layer + AngularMaxPooling, device-resident, CUDA events, L2 flushed
between iterations.  Prints one JSON line per case and a markdown table at the end."""
import itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_zero import ConvZero

dev = "cuda:0"
KINDS = {"geodesic": ConvGeodesic, "dirac": ConvDirac, "zero": ConvZero}
precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
amp = AngularMaxPooling()
rows = []


def run(tag, kind, v, f, t, r, a, delta, iters=5):
    torch.manual_seed(0)
    layer = KINDS[kind](input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028,
                        rotation_delta=delta, precision=precision).to(dev)
    x = S.make_signal(v, f, seed=1).to(dev).requires_grad_(True)
    bary = S.make_bary(v, r, a, seed=1).to(dev)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2)).to(dev)

    def step():
        x.grad = None
        for p_ in layer.parameters():
            p_.grad = None
        amp(layer([x, bary])).backward(gz)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    n_rot = len(range(0, a, delta))
    k = r * a * f
    flops = (2 * v * n_rot * k * t + 2 * v * f * t + 4 * v * k * t + 4 * v * f * t) if kind != "zero" else 6 * v * f * t
    rec = {"case": tag, "kind": kind, "V": v, "F": f, "T": t, "R": r, "A": a, "delta": delta, "n_rot": n_rot,
           "precision": precision, "ms_fwd_bwd": round(ms, 4), "vertices_per_s": round(v / ms * 1e3),
           "algorithmic_tflops": round(flops / ms / 1e9, 1)}
    rows.append(rec)
    print(json.dumps(rec), flush=True)


only_a = int(os.environ.get("GCB_SWEEP_ONLY_A", "0"))   # restrict the sweep to one angular size (A/B runs)
for kind, r, a, delta in itertools.product(("geodesic", "dirac", "zero"), (3, 5, 8), (6, 8, 12), (1, 2)):
    if only_a and (a != only_a or kind == "zero"):
        continue
    run("S2 sweep", kind, 6890, 544, 96, r, a, delta)
if only_a:
    sys.exit(0)
for fi, fo in ((64, 96), (96, 256), (256, 384)):
    run("S3 layer", "geodesic", 6890, fi, fo, 5, 8, 1)
run("S4 layer", "geodesic", 250000, 128, 128, 5, 8, 1, iters=3)
print("\n| case | kind | F->T | R x A | delta | ms fwd+bwd | M vertices/s | algorithmic TFLOP/s |\n|---|---|---|---|---|---|---|---|")
for r_ in rows:
    print(f"| {r_['case']} | {r_['kind']} | {r_['F']}->{r_['T']} | {r_['R']}x{r_['A']} | {r_['delta']} | {r_['ms_fwd_bwd']} | "
          f"{r_['vertices_per_s'] / 1e6:.2f} | {r_['algorithmic_tflops']} |")
