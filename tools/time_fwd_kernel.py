"""Time only the forward contraction kernel (events inside the library) for the S1 shape."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoconv_oracle as O
from geoconv_b200 import _native
if os.environ.get('GCB_LIB'):
    _native.LIB_PATH = os.environ['GCB_LIB']
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
lib = _native.load()
v, f, r, a, t = 6890, 544, 5, 8, 96
wn, ws, b = O.make_weights(t, r, a, f, seed=0)
x = O.make_signal(v, f, seed=1).to("cuda:0")
bary = O.make_bary(v, r, a, seed=1).to("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
for prec in sys.argv[1:] or ["fp32", "tf32"]:
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028, precision=prec)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to("cuda:0")
    with torch.no_grad():
        for _ in range(3):
            layer([x, bary])
        torch.cuda.synchronize()
        lib.gcb_profile_enable(1)
        for _ in range(10):
            flush.zero_()
            layer([x, bary])
        torch.cuda.synchronize()
    ms = ctypes.c_double(0.0); cnt = ctypes.c_int64(0)
    lib.gcb_profile_read(0, ctypes.byref(ms), ctypes.byref(cnt))
    lib.gcb_profile_enable(0)
    print(f"{prec} debug={os.environ.get('GCB_TC_DEBUG','0')} cluster={os.environ.get('GCB_TC_CLUSTER','0')}: k_tc_conv {ms.value / max(cnt.value,1):.4f} ms over {cnt.value} launches", flush=True)
