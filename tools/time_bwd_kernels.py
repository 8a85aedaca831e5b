"""Per-kernel CUDA-event times of one fwd+bwd step at the S1 shape (library profiling slots)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200 import _native
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
lib = _native.load()
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = O.make_weights(t, r, a, f, seed=0)
x = O.make_signal(v, f, seed=1).to("cuda:0").requires_grad_(True)
bary = O.make_bary(v, r, a, seed=1).to("cuda:0")
gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2)).to("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
amp = AngularMaxPooling()
names = ["fwd_conv", "bwd_dw", "bwd_dx_gemm", "bwd_build_g"]
for prec in sys.argv[1:] or ["fp32", "tf32"]:
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision=prec)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to("cuda:0")
    def step():
        x.grad = None
        for p_ in layer.parameters():
            p_.grad = None
        amp(layer([x, bary])).backward(gz)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.gcb_profile_enable(1)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    out = []
    for i, n in enumerate(names):
        ms = ctypes.c_double(0.0); cnt = ctypes.c_int64(0)
        lib.gcb_profile_read(i, ctypes.byref(ms), ctypes.byref(cnt))
        out.append(f"{n} {1e3 * ms.value / max(cnt.value, 1):.1f}us")
    lib.gcb_profile_enable(0)
    print(f"{prec}: step(eager, with event overhead) {tot / 10:.3f} ms | " + " | ".join(out), flush=True)
