"""torch.compile support (the reference demo compiles its model, training_demo.py:126-127).

Runs on CPU: the model lives on a *fake* CUDA device (FakeTensorMode), Dynamo traces it with ``fullgraph=True`` -
every device-side step must therefore be a ``geoconv_b200::*`` torch.library operator with a fake kernel - and the
operators' autograd formulas are exercised on fake tensors.  The numeric check of the compiled model is the GPU
test ``test_gpu_parity.py::test_survives_torch_compile``."""
import types

import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from geoconv_b200 import _native, ops
from geoconv_b200.pytorch.layers.angular_avg_pooling import AngularAvgPooling
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_zero import ConvZero

V, F, R, A, T = 300, 32, 3, 6, 32


class Net(torch.nn.Module):
    def __init__(self, conv_cls, pool_cls):
        super().__init__()
        self.down = torch.nn.Linear(F, F)
        self.conv = conv_cls(input_shape=[(None, F), (None, R, A, 3, 2)], amt_templates=T, template_radius=0.03)
        self.pool = pool_cls()
        self.head = torch.nn.Linear(T, 5)

    def forward(self, x, bary):
        return self.head(self.pool(self.conv([self.down(x), bary])))


def trace(conv_cls, pool_cls):
    net = Net(conv_cls, pool_cls)
    mode = FakeTensorMode(allow_non_fake_inputs=True)
    graphs = []

    def backend(gm, _inputs):
        graphs.append(gm)
        return gm.forward

    with mode:
        net = net.to_empty(device="cuda")
        x = torch.empty(V, F, device="cuda")
        bary = torch.empty(V, R, A, 3, 2, device="cuda")
        torch._dynamo.reset()
        out = torch.compile(net, backend=backend, fullgraph=True)(x, bary)
    assert len(graphs) == 1, "one graph, no graph break"
    targets = [str(n.target) for n in graphs[0].graph.nodes if n.op == "call_function"]
    return out, [t for t in targets if t.startswith("geoconv_b200")]


def test_fullgraph_trace_keeps_the_fused_pooling():
    out, ops_seen = trace(ConvGeodesic, AngularMaxPooling)
    assert tuple(out.shape) == (V, 5) and out.device.type == "cuda"
    assert ops_seen == ["geoconv_b200.prep_bary.default", "geoconv_b200.fold_prior.default",
                        "geoconv_b200.conv_fwd.default"], ops_seen   # AMP rides in conv_fwd: no pooling operator


def test_fullgraph_trace_other_layers():
    _, ops_seen = trace(ConvDirac, AngularAvgPooling)
    assert ops_seen == ["geoconv_b200.prep_bary.default", "geoconv_b200.conv_fwd.default",
                        "geoconv_b200.apool_fwd.default"], ops_seen
    _, ops_seen = trace(ConvZero, AngularMaxPooling)
    assert ops_seen == ["geoconv_b200.prep_bary.default", "geoconv_b200.conv_fwd.default"], ops_seen


def test_autograd_formulas_on_fake_tensors():
    """The backward formulas registered with torch.library run on fake tensors (what AOT autograd traces)."""
    n_rot = A
    with FakeTensorMode():
        dev = "cuda"
        x = torch.empty(V, F, device=dev)
        w_eff = torch.empty(T, R, A, F, device=dev)
        w_self = torch.empty(T, 1, F, device=dev)
        bias = torch.empty(1, T, device=dev)
        idx = torch.empty(V, R, A, 3, dtype=torch.int32, device=dev)
        w = torch.empty(V, R, A, 3, device=dev)
        rot = list(range(n_rot))
        y, z, arg, stats = ops.conv_fwd_op(x, w_eff, w_self, bias, idx, w, None, None, rot, 0, _native.PREC_F16X3,
                                           _native.PREC_F16X3, 0)
        assert tuple(y.shape) == (V, n_rot, T) and tuple(z.shape) == (V, T) and arg.dtype == torch.int32
        ctx = types.SimpleNamespace(needs_input_grad=[True] * 13, set_materialize_grads=lambda _f: None, saved=None)
        ctx.save_for_backward = lambda *ts: setattr(ctx, "saved_tensors", ts)
        ops._conv_setup(ctx, (x, w_eff, w_self, bias, idx, w, None, None, rot, 0, 3, 3, 0), (y, z, arg, stats))
        grads = ops._conv_backward(ctx, None, torch.empty_like(z), None, None)
        assert len(grads) == 13
        assert tuple(grads[0].shape) == (V, F) and tuple(grads[1].shape) == (T, R, A, F)
        assert tuple(grads[2].shape) == (T, 1, F) and tuple(grads[3].shape) == (1, T)
        assert all(g is None for g in grads[4:])
        # dense gradient through Y, no signal gradient, all-zero prior (anchor receives zeros)
        ctx.needs_input_grad = [False] + [True] * 12
        ops._conv_setup(ctx, (x, None, w_self, bias, idx, w, w_eff, None, rot, 0, 3, 3, 0), (y, z, arg, stats))
        grads = ops._conv_backward(ctx, torch.empty_like(y), None, None, None)
        assert grads[0] is None and grads[1] is None and tuple(grads[6].shape) == (T, R, A, F)
        # pooling and prior folding
        zz, aa = ops.apool_fwd_op(y, ops.POOL_MIN)
        pctx = types.SimpleNamespace()
        pctx.save_for_backward = lambda *ts: setattr(pctx, "saved_tensors", ts)
        ops._apool_setup(pctx, (y, ops.POOL_MIN), (zz, aa))
        gy, _none = ops._apool_backward(pctx, torch.empty_like(zz), None)
        assert tuple(gy.shape) == tuple(y.shape)
        prior = torch.empty(R, A, R, A, device=dev)
        fctx = types.SimpleNamespace()
        fctx.save_for_backward = lambda *ts: setattr(fctx, "saved_tensors", ts)
        ops._fold_setup(fctx, (w_eff, prior, 0), None)
        g = ops._fold_backward(fctx, torch.empty_like(w_eff))
        assert tuple(g[0].shape) == tuple(w_eff.shape) and g[1] is None and g[2] is None
