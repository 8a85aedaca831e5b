"""Warm per-kernel durations of one S1 step (fwd + AMP + bwd) from torch.profiler (CUPTI activity records),
L2 flushed between steps.  usage: [GCB_KT_CLASS=geodesic|dirac|zero] [GCB_KT_V= GCB_KT_F= GCB_KT_T= GCB_KT_R= GCB_KT_A=] kernel_times.py [precision ...]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from geoconv_b200 import synthetic as S
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_zero import ConvZero
CLS = {"geodesic": ConvGeodesic, "dirac": ConvDirac, "zero": ConvZero}[os.environ.get("GCB_KT_CLASS", "geodesic")]

v, f, r, a, t = int(os.environ.get("GCB_KT_V", 6890)), int(os.environ.get("GCB_KT_F", 544)), int(os.environ.get("GCB_KT_R", 5)), int(os.environ.get("GCB_KT_A", 8)), int(os.environ.get("GCB_KT_T", 96))
dev = "cuda:0"
wn, ws, b = S.make_weights(t, r, a, f, seed=0)
x = S.make_signal(v, f, seed=1).to(dev).requires_grad_(True)
bary = S.make_bary(v, r, a, seed=1).to(dev)
gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
amp = AngularMaxPooling()
STEPS = 10
for prec in sys.argv[1:] or ["fp32"]:
    layer = CLS(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision=prec)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(dev)

    def step():
        x.grad = None
        for p in layer.parameters():
            p.grad = None
        amp(layer([x, bary])).backward(gz)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(STEPS):
            flush.zero_()
            step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name != "CUDA" or "FillFunctor" in ev.name and ev.device_time > 50:
            continue
        agg.setdefault(ev.name[:72], []).append(ev.device_time)
    total = 0.0
    print(f"== {CLS.__name__} {prec}: per-kernel device time per step (us), {STEPS} steps")
    for name, ts in agg.items():
        per_step = sum(ts) / STEPS
        total += per_step
        print(f"{per_step:9.1f}  x{len(ts) / STEPS:4.1f}  {name}")
    print(f"{total:9.1f}  sum of kernels per step")
    # the same steps without the profiler, CUDA events around each (L2 flush outside): kernels + launch gaps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        step()
        e1.record()
    torch.cuda.synchronize()
    print(f"{sum(a.elapsed_time(b) for a, b in ev) / STEPS * 1e3:9.1f}  step by CUDA events (eager launches)")
    # ... and with a host sync after every step (a loop that reads the loss each step): the GPU starts every step idle
    import time
    tot, host = 0.0, 0.0
    for _ in range(STEPS):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        step()
        e1.record()
        host += time.perf_counter() - t0
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f"{tot / STEPS * 1e3:9.1f}  step by CUDA events, host sync after every step (host time to enqueue a step: {host / STEPS * 1e6:.0f} us)")
