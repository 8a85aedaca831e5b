"""BASELINE config 3: kernel-variant sweep ConvGeodesic / ConvDirac / ConvZero at n_radial in {3,5,8},
n_angular in {6,8,12}, rotation_delta in {1,2} on the 6890-vertex shape (F=544, T=96) - all 54 cases, every one
compared NUMERICALLY with the CPU oracle:

* forward: the full-size GPU result (default fp32-faithful precision: the tensor-core kernels) on the first
  SUB rows against a float64 evaluation of the oracle on those rows (`bary[:SUB]` against the full signal keeps
  the index range; rows of Y are independent of each other);
* backward: the same layer run on the SUB-row barycentric tensor against the full signal (V_out = SUB,
  V_src = 6890: same kernels, same index range) against autograd over the float64 oracle - or, where a near-tie
  pooling decision went the other way, against the closed form at the kernel's own decisions (no skips);
* size-independent properties of the full-size result: AMP output rows are rows of Y, argmax is the first
  maximum of the norms, ConvZero is rotation-invariant with argmax 0 and an all-zero neighbour gradient;
* the full-size gradients of the tensor-core path against the exact CUDA-core path where the latter is cheap."""
import os

import pytest
import torch

from oracle import geoconv_oracle as O
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_zero import ConvZero
from tests.helpers import TOL_FP32, argmax_admissible, check_gradients, nerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
V, F, T = 6890, 544, 96
SUB = 512
KINDS = {"geodesic": ConvGeodesic, "dirac": ConvDirac, "zero": ConvZero}


def build(kind, r, a, delta, precision, weights):
    wn, ws, b = weights
    layer = KINDS[kind](input_shape=[(None, F), (None, r, a, 3, 2)], amt_templates=T, template_radius=0.028,
                        rotation_delta=delta, precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    return layer.to(DEV)


def oracle_subset(kind, signal, bary_sub, weights, prior, gz_sub, delta):
    """float64 oracle forward + autograd backward on a row subset (all host cores)."""
    torch.set_num_threads(os.cpu_count() or 1)
    wn, ws, b = (p_.double() for p_ in weights)
    pr = None if prior is None else prior.double()
    if kind == "zero":
        # an all-zero prior removes the neighbour term exactly (conv_zero.py:12-15): evaluate the remaining
        # centre term directly instead of contracting 52 k zeros per row
        xs = signal.double().clone().requires_grad_(True)
        ps = [p_.clone().requires_grad_(True) for p_ in (ws, b)]
        n_rot = len(range(0, bary_sub.shape[2], delta))
        pre = xs[: bary_sub.shape[0]] @ ps[0][:, 0, :].T + ps[1]
        y = O.activation("relu", pre)[:, None, :].expand(-1, n_rot, -1)
        z, arg = O.amp_forward(y)
        gx, gws, gb = torch.autograd.grad(z, [xs] + ps, grad_outputs=gz_sub.double())
        return {"y": y.detach(), "z": z.detach(), "argmax": arg, "d_signal": gx, "d_w_neighbor": torch.zeros_like(wn),
                "d_w_self": gws, "d_bias": gb}
    return O.conv_amp_forward_backward(signal.double(), bary_sub, wn, ws, b, pr, gz_sub.double(), "relu", delta)


@pytest.mark.parametrize("delta", [1, 2])
@pytest.mark.parametrize("a", [6, 8, 12])
@pytest.mark.parametrize("r", [3, 5, 8])
@pytest.mark.parametrize("kind", ["geodesic", "dirac", "zero"])
def test_sweep(kind, r, a, delta):
    weights = O.make_weights(T, r, a, F, seed=0)
    signal_h = O.make_signal(V, F, seed=1, kind="shot")
    bary_h = O.make_bary(V, r, a, seed=1)
    gz_h = torch.randn((V, T), generator=torch.Generator().manual_seed(2))
    signal, bary, gz = signal_h.to(DEV), bary_h.to(DEV), gz_h.to(DEV)
    prior = {"geodesic": torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.028))).float(),
             "dirac": None, "zero": torch.zeros(r, a, r, a)}[kind]
    offsets = range(0, a, delta)
    n_rot = len(offsets)

    # ---- full size, default precision ----
    layer = build(kind, r, a, delta, "fp32", weights)
    xs = signal.clone().requires_grad_(True)
    y = layer([xs, bary])
    z = AngularMaxPooling()(y)
    z.backward(gz)
    y, z, arg = y.detach(), z.detach(), y._gcb_pooled[1]
    dx, dw = xs.grad, layer._template_neighbor_weights.grad
    assert y.shape == (V, n_rot, T) and z.shape == (V, T)
    norms = torch.linalg.vector_norm(y, dim=-1)
    # first maximum of the norms; rotations whose norms agree to ~1 ulp may be ordered differently by
    # a different summation order (torch.linalg.vector_norm vs the kernel's warp reduction)
    picked = norms[torch.arange(V, device=DEV), arg.long()]
    assert (picked >= norms.max(dim=1).values * (1 - 2e-6)).all()
    assert (arg.long() != torch.argmax(norms, dim=1)).sum() <= 3
    assert torch.equal(z, y[torch.arange(V, device=DEV), arg.long()])
    assert torch.isfinite(dx).all() and torch.isfinite(dw).all()
    if kind == "zero":
        assert torch.equal(arg, torch.zeros_like(arg))
        assert (y - y[:, :1, :]).abs().max() == 0
        assert dw.abs().max() == 0

    # ---- the oracle on a row subset: forward of the full-size run, backward of a subset run ----
    ref = oracle_subset(kind, signal_h, bary_h[:SUB], weights, prior, gz_h[:SUB], delta)
    err_y = nerr(y[:SUB], ref["y"])
    assert err_y < TOL_FP32, f"forward vs float64 oracle: {err_y:.3e}"
    assert argmax_admissible(arg[:SUB], ref["argmax"], ref["y"], 2e-6)
    layer_s = build(kind, r, a, delta, "fp32", weights)
    xs_s = signal.clone().requires_grad_(True)
    y_s = layer_s([xs_s, bary[:SUB].contiguous()])
    z_s = AngularMaxPooling()(y_s)
    # rows of Y do not depend on how many rows are computed (the K-split count may: compare to rounding)
    assert nerr(y_s, y[:SUB]) < 4e-6
    z_s.backward(gz[:SUB].contiguous())
    check_gradients(layer_s, xs_s.grad, z_s, y_s._gcb_pooled[1], ref, signal_h, bary_h[:SUB], weights[0], weights[1],
                    prior, gz_h[:SUB], "relu", offsets, TOL_FP32)

    # ---- full-size gradients: tensor-core backward vs the exact CUDA-core backward behind the same forward ----
    if kind == "zero" or r * a <= 40:
        layer_x = build(kind, r, a, delta, "ffma", weights)
        outs = {}
        for bwd in ("ffma", "fp32"):
            layer_x.backward_precision = bwd
            for p_ in layer_x.parameters():
                p_.grad = None
            xs_x = signal.clone().requires_grad_(True)
            y_x = layer_x([xs_x, bary])
            AngularMaxPooling()(y_x).backward(gz)
            outs[bwd] = (y_x.detach(), y_x._gcb_pooled[1], xs_x.grad, layer_x._template_neighbor_weights.grad.clone())
        assert nerr(y, outs["ffma"][0]) < TOL_FP32
        assert (arg != outs["ffma"][1]).sum() <= 4        # near-tie flips only
        assert nerr(outs["fp32"][2], outs["ffma"][2]) < TOL_FP32
        assert nerr(outs["fp32"][3], outs["ffma"][3]) < TOL_FP32 or kind == "zero"
