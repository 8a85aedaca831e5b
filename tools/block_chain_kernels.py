"""Per-kernel device times (CUPTI) of one forward + backward of BatchNormDropout and of torch's BatchNorm1d + Dropout."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from torch.profiler import ProfilerActivity, profile
from geoconv_b200.pytorch.layers.batch_norm_dropout import BatchNormDropout
v, t = 6890, int(sys.argv[1]) if len(sys.argv) > 1 else 96
z = torch.randn(v, t, device="cuda", requires_grad=True)
g = torch.randn(v, t, device="cuda")
for name, m in (("ours", BatchNormDropout(t, p=0.2).cuda()), ("torch", nn.Sequential(nn.BatchNorm1d(t), nn.Dropout(0.2)).cuda())):
    def step():
        z.grad = None
        m.zero_grad(set_to_none=True)
        m(z).backward(g)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name == "CUDA":
            agg.setdefault(ev.name[:90], []).append(ev.device_time)
    print(f"== {name} (V={v}, T={t})")
    tot = 0
    for k, ts in agg.items():
        print(f"{sum(ts)/10:8.1f} us x{len(ts)/10:3.1f}  {k}"); tot += sum(ts)/10
    print(f"{tot:8.1f} us total")
