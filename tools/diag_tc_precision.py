"""GPU diagnostic: error of the tcgen05 forward (tf32 / tf32x3) vs an fp64 evaluation, as a
function of the contraction length, plus first kernel timings.  Not part of the product."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac

DEV = "cuda:0"

def ref64(signal, bary, wn, ws, b, offs):
    idx, w = O.split_bary(bary)
    idx, w = idx.to(DEV), w.to(DEV).double()
    x = signal.to(DEV).double()
    v, r, a, _ = idx.shape
    interp = (x[idx.reshape(-1)].reshape(v, r, a, 3, -1) * w.unsqueeze(-1)).sum(-2)
    wn = wn.to(DEV).double(); ws = ws.to(DEV).double(); b = b.to(DEV).double()
    centre = x @ ws[:, 0, :].T
    outs = [torch.einsum("txyf,kxyf->kt", torch.roll(wn, -o, 2), interp) for o in offs]
    return torch.stack(outs, 1) + centre[:, None, :] + b.reshape(1, 1, -1)

def run(v, f, r, a, t, sig_kind):
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind=sig_kind)
    bary = O.make_bary(v, r, a, seed=1)
    pre = ref64(signal, bary, wn, ws, b, list(range(a)))
    ref = torch.clamp_min(pre, 0)
    res = {}
    for prec in ("ffma", "tf32x3", "tf32"):
        layer = ConvDirac(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03, precision=prec)
        layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
        layer = layer.to(DEV)
        xs, bs = signal.to(DEV), bary.to(DEV)
        with torch.no_grad():
            y = layer([xs, bs])
            torch.cuda.synchronize()
            n_it = 3 if prec == "ffma" else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_it):
                y = layer([xs, bs])
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_it
        d = (y.double() - ref)
        mask = pre > 0
        nerr = d.abs().max().item() / ref.abs().max().item()
        bias = (d[mask] / ref[mask].clamp_min(1e-30)).mean().item()
        rms = (d[mask] / ref[mask].clamp_min(1e-30)).pow(2).mean().sqrt().item()
        res[prec] = (nerr, bias, rms, ms)
    K = r * a * f
    print(f"V={v} F={f} R={r} A={a} T={t} K={K} sig={sig_kind}: " +
          " | ".join(f"{p}: nerr {e:.2e} relbias {bb:+.2e} relrms {rr:.2e} {ms:.3f} ms" for p, (e, bb, rr, ms) in res.items()), flush=True)

if __name__ == "__main__":
    for r in (1, 2, 5):
        run(6890, 544, r, 8, 96, "shot")
    run(6890, 544, 5, 8, 96, "randn")
    run(6890, 64, 5, 8, 96, "shot")
