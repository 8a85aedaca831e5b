"""BASELINE.json configs 1, 4 and 5 on the CUDA path against the CPU oracle (config 2 = S1 is
test_gpu_parity.py::test_faust_shape_against_oracle, config 3 = test_gpu_sweep.py).

* config 1 (S0): icosphere-4 sized mesh, V = 2562, 3-D coordinate signal, ConvGeodesic 3 -> 32, R5 x A8, delta 1;
* config 4 layer shapes: the three layers of the segmentation net, 64 -> 96, 96 -> 256, 256 -> 384 on 6890 vertices;
* config 5 layer shape: 128 -> 128 (the row-sharded 2 M-vertex mesh runs this layer; here 6890 rows and a
  250 k-row property check against the same kernels on a row subset).
Forward and gradients, every precision mode, no skips: where a near-tie pooling decision differs the gradients
are checked in closed form at the kernel's own decisions (tests/helpers.py::check_gradients)."""
import os

import pytest
import torch

from oracle import geoconv_oracle as O
from geoconv_b200 import _native, ops
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from tests.helpers import TOL_FP32, TOL_TF32, argmax_admissible, check_gradients, nerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_case(v, f, t, r, a, precision, signal_kind, radius=0.028, act="relu"):
    torch.set_num_threads(os.cpu_count() or 1)
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind=signal_kind)
    bary = O.make_bary(v, r, a, seed=1)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2))
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, radius))).float()
    ref = O.conv_amp_forward_backward(signal, bary, wn, ws, b, prior, gz, act, 1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=radius,
                         activation=act, precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    code = ops.resolve_precision(precision, r, a, f, t, a)
    tol = TOL_TF32 if code == _native.PREC_TF32 else TOL_FP32
    xs = signal.to(DEV).requires_grad_(True)
    y = layer([xs, bary.to(DEV)])
    z = AngularMaxPooling()(y)
    assert nerr(y, ref["y"]) < tol
    arg = y._gcb_pooled[1]
    assert argmax_admissible(arg, ref["argmax"], ref["y"], 2e-6 if tol == TOL_FP32 else 4 * tol)
    if tol == TOL_TF32:
        # a TF32 forward flips ReLU bits near zero; check the TF32 backward arithmetic behind an exact forward
        layer.precision, layer.backward_precision = "ffma", precision
        xs = signal.to(DEV).requires_grad_(True)
        y = layer([xs, bary.to(DEV)])
        z = AngularMaxPooling()(y)
        arg = y._gcb_pooled[1]
    z.backward(gz.to(DEV))
    return check_gradients(layer, xs.grad, z, arg, ref, signal, bary, wn, ws, prior, gz, act, range(a), tol), code


@pytest.mark.parametrize("precision", ["fp32", "tf32", "ffma"])
def test_config1_s0_icosphere_sized(precision):
    """V = 2562, F = 3 (xyz-like randn), T = 32: no tensor-core tiling exists for F = 3, every mode resolves to the
    exact CUDA-core kernels (k_interp + k_sgemm) - the shape the reference's own CPU example runs."""
    _how, code = run_case(2562, 3, 32, 5, 8, precision, "randn")
    assert code == _native.PREC_FP32


@pytest.mark.parametrize("precision", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("dims", [(64, 96), (96, 256), (256, 384), (128, 128)], ids=lambda d: f"{d[0]}to{d[1]}")
def test_config45_layer_shapes(dims, precision):
    f, t = dims
    _how, code = run_case(6890, f, t, 5, 8, precision, "randn")
    assert code != _native.PREC_FP32, "these shapes run on tensor cores"


def test_config5_rows_of_a_large_mesh():
    """250 k rows of a locality-preserving mesh (the per-rank share of config 5), 128 -> 128: the first 384 rows
    against the float64 oracle, and the whole result against the same kernels run on two halves of the rows
    (row independence at a size where the forward runs many waves)."""
    from geoconv_b200 import synthetic as S
    v, f, t, r, a = 250_000, 128, 128, 5, 8
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=3, kind="randn")
    bary = S.make_bary_local(v, r, a, seed=3)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    xd, bd = signal.to(DEV), bary.to(DEV)
    with torch.no_grad():
        y = layer([xd, bd])
        half = v // 2
        y_lo = layer([xd, bd[:half].contiguous()])
        assert nerr(y_lo, y[:half]) < 4e-6      # (the K-split count may depend on the row count)
        del y_lo
        sub = 384
        prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.028)))
        torch.set_num_threads(os.cpu_count() or 1)
        y_ref = O.conv_forward(signal.double(), bary[:sub], wn.double(), ws.double(), b.double(), prior, "relu", 1)
        assert nerr(y[:sub], y_ref) < TOL_FP32
