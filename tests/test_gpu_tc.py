"""tcgen05 forward kernel (TF32, 3xTF32 and the fp16-split 3xFP16) against the CPU oracle over the tilings it selects:
swizzle width KC in {32,16,8}, one or several rotation groups, template chunks, ragged vertex
counts, rotation_delta > 1."""
import pytest
import torch

from oracle import geoconv_oracle as O
from geoconv_b200 import _native, ops
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from tests.helpers import TOL_FP32, TOL_TF32, check_gradients, nerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    # kind, V, F, R, A, T, delta, act
    ("geodesic", 300, 64, 3, 8, 32, 1, "relu"),
    ("geodesic", 1000, 96, 5, 8, 96, 1, "relu"),
    ("dirac", 517, 40, 2, 6, 48, 2, "elu"),
    ("geodesic", 200, 48, 3, 12, 16, 1, "tanh"),
    ("dirac", 260, 32, 2, 8, 384, 4, "relu"),
    ("geodesic", 150, 128, 3, 8, 256, 1, "leaky_relu"),
    ("dirac", 129, 64, 1, 3, 64, 1, "sigmoid"),
]


@pytest.mark.parametrize("precision", ["tf32", "tf32x3", "f16x3"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "-".join(map(str, s)))
def test_tc_forward_matches_oracle(shape, precision):
    kind, v, f, r, a, t, delta, act = shape
    n_rot = len(range(0, a, delta))
    assert _native.load().gcb_tc_supported(r, a, f, t, n_rot), "shape should be served by tcgen05"
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="randn")
    bary = O.make_bary(v, r, a, seed=1, fallback_frac=0.05)
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.03))).float() if kind == "geodesic" else None
    y_ref = O.conv_forward(signal.double(), bary, wn.double(), ws.double(), b.double(),
                           None if prior is None else prior.double(), act, delta)
    cls = ConvGeodesic if kind == "geodesic" else ConvDirac
    layer = cls(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                activation=act, rotation_delta=delta, precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    y = layer([signal.to(DEV), bary.to(DEV)])
    tol = TOL_TF32 if precision == "tf32" else TOL_FP32
    err = nerr(y, y_ref)
    assert err < tol, f"normalised error {err:.3e} >= {tol}"
    # fused pooling outputs are consistent with the Y this kernel produced
    z, arg, _ = y._gcb_pooled
    z_ref, arg_ref = O.amp_forward(y.detach().cpu())
    assert torch.equal(arg.cpu(), arg_ref)
    assert torch.equal(z.cpu(), z_ref)


def test_tc_matches_ffma_on_device_and_is_deterministic():
    v, f, r, a, t = 2000, 96, 5, 8, 96
    wn, ws, b = O.make_weights(t, r, a, f, seed=3)
    signal = O.make_signal(v, f, seed=2, kind="shot").to(DEV)
    bary = O.make_bary(v, r, a, seed=2).to(DEV)
    outs = {}
    for precision in ("ffma", "tf32x3", "f16x3", "tf32"):
        layer = ConvDirac(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                          precision=precision)
        layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
        layer = layer.to(DEV)
        outs[precision] = layer([signal, bary]).detach()
        again = layer([signal, bary]).detach()
        assert torch.equal(outs[precision], again)
    assert nerr(outs["tf32x3"], outs["ffma"]) < TOL_FP32
    assert nerr(outs["f16x3"], outs["ffma"]) < TOL_FP32
    assert nerr(outs["tf32"], outs["ffma"]) < TOL_TF32


BWD_SHAPES = [
    ("geodesic", 300, 64, 3, 8, 32, 1, "relu"),
    ("geodesic", 1000, 96, 5, 8, 96, 1, "relu"),
    ("dirac", 260, 32, 2, 8, 384, 4, "elu"),
    ("geodesic", 150, 128, 3, 8, 256, 1, "tanh"),
    ("dirac", 777, 160, 2, 6, 64, 1, "relu"),
]


@pytest.mark.parametrize("precision", ["tf32", "tf32x3", "f16x3"])
@pytest.mark.parametrize("shape", BWD_SHAPES, ids=lambda s: "-".join(map(str, s)))
def test_tc_backward_matches_oracle(shape, precision):
    """AMP-sparse backward through the tcgen05 dW / dI kernels + CSR gather-reduce vs autograd over
    the CPU oracle.  The forward here runs the exact CUDA-core path so that both sides select the
    same rotations; only the backward arithmetic is under test."""
    kind, v, f, r, a, t, delta, act = shape
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="randn")
    bary = O.make_bary(v, r, a, seed=1, fallback_frac=0.05)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2))
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.03))).float() if kind == "geodesic" else None
    ref = O.conv_amp_forward_backward(signal, bary, wn, ws, b, prior, gz, act, delta)
    cls = ConvGeodesic if kind == "geodesic" else ConvDirac
    layer = cls(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                activation=act, rotation_delta=delta, precision="ffma")
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    layer.backward_precision = precision
    xs = signal.to(DEV).requires_grad_(True)
    y = layer([xs, bary.to(DEV)])
    z = AngularMaxPooling()(y)
    z.backward(gz.to(DEV))
    tol = TOL_TF32 if precision == "tf32" else TOL_FP32
    check_gradients(layer, xs.grad, z, y._gcb_pooled[1], ref, signal, bary, wn, ws, prior, gz, act,
                    range(0, a, delta), tol)


RANGE_CASES = [
    # signal scale, weight scale, upstream-gradient scale, outlier factor (one signal element multiplied by it)
    (1.0, 1.0, 1.0, 1.0),
    # (signal and weight scales cancel, so the pre-activations keep their magnitude and the activation does
    # not amplify or hide the contraction error)
    (3.7e6, 2.7e-7, 1.3e-7, 1.0),      # every operand far outside the fp16 range before scaling
    (4.1e-9, 2.4e8, 8.8e5, 1.0),
    (1.0, 1.0, 1.0, 1.0e3),            # one element a thousand times the rest
]


@pytest.mark.parametrize("case", RANGE_CASES, ids=lambda c: "x%g-w%g-g%g-o%g" % c)
def test_f16x3_operand_scaling(case):
    """The fp16 hi/lo split (GCB_PREC_F16X3, what precision='fp32' resolves to) rescales every operand by a
    power of two taken from device-side maxima: forward and gradients stay fp32-faithful for data far outside
    the fp16 range and for a wide dynamic range inside one tensor."""
    sx, sw, sg, outlier = case
    v, f, r, a, t = 700, 96, 3, 8, 96
    assert _native.load().gcb_tc_half_supported(r, a, f, t, a)
    wn, ws, b = O.make_weights(t, r, a, f, seed=5)
    wn, ws, b = wn * sw, ws * sw, b * (sx * sw)
    signal = O.make_signal(v, f, seed=6, kind="randn") * sx
    signal[17, 5] *= outlier
    bary = O.make_bary(v, r, a, seed=6, fallback_frac=0.03)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(7)) * sg
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.03))).float()
    ref = O.conv_amp_forward_backward(signal.double(), bary, wn.double(), ws.double(), b.double(), prior.double(),
                                      gz.double(), "tanh" if outlier == 1.0 else "elu", 1)
    act = "tanh" if outlier == 1.0 else "elu"
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                         activation=act, precision="f16x3")
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    assert ops.resolve_precision("f16x3", r, a, f, t, a) == _native.PREC_F16X3
    xs = signal.to(DEV).requires_grad_(True)
    y = layer([xs, bary.to(DEV)])
    z = AngularMaxPooling()(y)
    assert nerr(y, ref["y"]) < TOL_FP32
    z.backward(gz.to(DEV))
    check_gradients(layer, xs.grad, z, y._gcb_pooled[1], ref, signal, bary, wn, ws, prior, gz, act, range(a), TOL_FP32)


def test_f16x3_zero_and_constant_inputs():
    """All-zero signal (the scale kernel sees a zero bound) and all-zero weights give exact results."""
    v, f, r, a, t = 300, 64, 3, 8, 32
    wn, ws, b = O.make_weights(t, r, a, f, seed=8)
    bary = O.make_bary(v, r, a, seed=8).to(DEV)
    layer = ConvDirac(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                      activation="relu", precision="f16x3")
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    y0 = layer([torch.zeros(v, f, device=DEV), bary])
    expect = torch.relu(b.to(DEV)).reshape(1, 1, t).expand(v, a, t)
    assert torch.equal(y0, expect)
    layer.load_state_dict({"_template_neighbor_weights": wn * 0, "_template_self_weights": ws * 0, "_bias": b})
    y1 = layer([O.make_signal(v, f, seed=9, kind="randn").to(DEV), bary])
    assert torch.equal(y1, expect)


@pytest.mark.parametrize("shape", [(1000, 96, 5, 8, 96, 1), (700, 64, 3, 8, 32, 1), (517, 128, 2, 6, 256, 2), (300, 32, 2, 8, 384, 4)],
                         ids=lambda s: "-".join(map(str, s)))
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_fused_finish_equals_separate_finish(shape, precision, monkeypatch):
    """GCB_TC_FUSE=1: the split reduction, bias, activation and pooling run inside k_tc_conv (last-arriving split of
    a unit).  Same additions in the same order as k_tc_finish: Y, Z and argmax are bit-identical.  (Not the default:
    slower at the FAUST shape, profiles/r02_notes.md.)"""
    v, f, r, a, t, delta = shape
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="randn").to(DEV)
    bary = O.make_bary(v, r, a, seed=1, fallback_frac=0.05).to(DEV)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                         rotation_delta=delta, activation="elu", precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    outs = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("GCB_TC_FUSE", mode)
        with torch.no_grad():
            y = layer([signal, bary])
        outs[mode] = (y.clone(), y._gcb_pooled[0].clone(), y._gcb_pooled[1].clone())
    monkeypatch.delenv("GCB_TC_FUSE")
    for mode in ("1", "2"):
        for got, want in zip(outs[mode], outs["0"]):
            assert torch.equal(got, want), f"GCB_TC_FUSE={mode}"


DENSE_SHAPES = [
    # kind, V, F, R, A, T, delta, act
    ("geodesic", 700, 64, 5, 8, 96, 1, "relu"),
    ("geodesic", 450, 96, 3, 6, 64, 2, "elu"),
    ("dirac", 333, 64, 2, 8, 32, 1, "tanh"),
]


@pytest.mark.parametrize("precision", ["tf32x3", "f16x3"])
@pytest.mark.parametrize("shape", DENSE_SHAPES, ids=lambda s: "-".join(map(str, s)))
def test_dense_backward_one_pass(shape, precision, monkeypatch):
    """A loss on Y itself (the layer used without angular max-pooling): the gradient for every rotation goes through ONE
    operand G and one pair of GEMMs (gcb_conv_bwd_dense) - against float64 autograd over the oracle, and against the
    per-rotation route it replaces."""
    kind, v, f, r, a, t, delta, act = shape
    assert _native.load().gcb_conv_bwd_dense_supported(r, a, f, t, _native.PREC_TF32X3)
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="randn")
    bary = O.make_bary(v, r, a, seed=1, fallback_frac=0.05)
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.03))).float() if kind == "geodesic" else None
    n_rot = len(range(0, a, delta))
    gy = torch.randn((v, n_rot, t), generator=torch.Generator().manual_seed(5))
    xs = signal.double().requires_grad_(True)
    ps = [p.double().requires_grad_(True) for p in (wn, ws, b)]
    yo = O.conv_forward(xs, bary, ps[0], ps[1], ps[2], None if prior is None else prior.double(), act, delta)
    ref = torch.autograd.grad((yo * gy.double()).sum(), [xs] + ps)
    cls = ConvGeodesic if kind == "geodesic" else ConvDirac
    got = {}
    for route in ("dense", "loop"):
        monkeypatch.setenv("GCB_DENSE_BWD", "1" if route == "dense" else "0")
        layer = cls(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                    activation=act, rotation_delta=delta, precision=precision)
        layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
        layer = layer.to(DEV)
        x = signal.to(DEV).requires_grad_(True)
        launches = _native.load().gcb_launch_count()
        y = layer([x, bary.to(DEV)])
        (y * gy.to(DEV)).sum().backward()
        got[route] = (x.grad, layer._template_neighbor_weights.grad, layer._template_self_weights.grad, layer._bias.grad,
                      _native.load().gcb_launch_count() - launches)
    # activations with a kink flip a derivative bit where the pre-activation is within rounding of zero
    tol = TOL_FP32 if act != "relu" else 3 * TOL_FP32
    for i, name in enumerate(("d_signal", "d_w_neighbor", "d_w_self", "d_bias")):
        assert nerr(got["dense"][i], ref[i]) < tol, (name, nerr(got["dense"][i], ref[i]))
        assert nerr(got["dense"][i], got["loop"][i]) < TOL_FP32, name
    assert got["dense"][4] < got["loop"][4]          # fewer launches than n_rot sparse backward calls


def test_dense_backward_under_min_and_avg_pooling():
    from geoconv_b200.pytorch.layers.angular_avg_pooling import AngularAvgPooling
    from geoconv_b200.pytorch.layers.angular_min_pooling import AngularMinPooling
    v, f, r, a, t = 500, 64, 5, 8, 64
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="randn")
    bary = O.make_bary(v, r, a, seed=1)
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.03))).float()
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(7))
    for pool, ref_pool in ((AngularAvgPooling(), O.aavg_forward), (AngularMinPooling(), lambda y: O.amin_forward(y)[0])):
        xs = signal.double().requires_grad_(True)
        ps = [p.double().requires_grad_(True) for p in (wn, ws, b)]
        yo = O.conv_forward(xs, bary, ps[0], ps[1], ps[2], prior.double(), "tanh", 1)
        ref = torch.autograd.grad((ref_pool(yo) * gz.double()).sum(), [xs] + ps)
        layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                             activation="tanh", precision="fp32")
        layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
        layer = layer.to(DEV)
        x = signal.to(DEV).requires_grad_(True)
        pool(layer([x, bary.to(DEV)])).backward(gz.to(DEV))
        assert nerr(x.grad, ref[0]) < TOL_FP32
        assert nerr(layer._template_neighbor_weights.grad, ref[1]) < TOL_FP32
        assert nerr(layer._bias.grad, ref[3]) < TOL_FP32
