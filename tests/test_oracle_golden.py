"""Pin the CPU oracle against vectors produced by the executed reference (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import geoconv_oracle as O
from tests.helpers import TOL_FP32, golden_cases, load_golden, nerr

CASES = golden_cases("all")


def _prior(g):
    return g["prior"] if g["include_prior"] else None


def test_fixtures_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("name", CASES)
def test_template_and_prior(name):
    g = load_golden(name)
    r, a = g["bary"].shape[1:3]
    # template radius is not stored; recover it from the outermost ring
    radius = g["template"][-1, 0, 0].item()
    tm = O.template_matrix(r, a, radius)
    np.testing.assert_allclose(tm, g["template"].numpy(), rtol=0, atol=1e-15)
    args = (g["prior_param"],) if "prior_param" in g else ()   # exp_lambda / dof of the parametrised priors
    prior = O.PRIORS[g["kind"]](tm, *args).astype(np.float32)
    np.testing.assert_allclose(prior, g["prior"].numpy(), rtol=0, atol=1e-7)
    if g["kind"] != "dirac":
        assert O.prior_is_rotation_symmetric(O.PRIORS[g["kind"]](tm, *args), g["delta"])


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    g = load_golden(name)
    y, z, arg = O.conv_amp_forward(g["signal"], g["bary"], g["w_neighbor"], g["w_self"], g["bias"],
                                   _prior(g), g["act"], g["delta"])
    assert y.shape == g["y"].shape and z.shape == g["z"].shape
    assert nerr(y, g["y"]) < 2e-6
    assert torch.equal(arg, g["argmax"])          # integer output: bit exact
    assert nerr(z, g["z"]) < 2e-6
    yo = O.conv_forward(g["signal"], g["bary"], g["w_neighbor"], g["w_self"], g["bias"], _prior(g),
                        g["act"], g["delta"], orientations=g["orientations"])
    assert nerr(yo, g["y_orient"]) < 2e-6


@pytest.mark.parametrize("name", CASES)
def test_indices_bit_exact(name):
    g = load_golden(name)
    idx, w = O.split_bary(g["bary"])
    ref_idx = g["bary"].to(torch.float32)[..., 0].long()
    assert torch.equal(idx, ref_idx)
    assert idx.min() >= 0 and idx.max() < g["signal"].shape[0]


@pytest.mark.parametrize("name", CASES)
def test_autograd_backward_matches_reference(name):
    g = load_golden(name)
    out = O.conv_amp_forward_backward(g["signal"], g["bary"], g["w_neighbor"], g["w_self"], g["bias"],
                                      _prior(g), g["grad_z"], g["act"], g["delta"])
    for k in ("d_signal", "d_w_neighbor", "d_w_self", "d_bias"):
        assert nerr(out[k], g[k]) < TOL_FP32, k


@pytest.mark.parametrize("name", CASES)
def test_closed_form_matches_reference(name):
    """The folded-prior forward and the closed-form AMP-sparse backward (what the CUDA kernels
    implement) agree with the reference in float64."""
    g = load_golden(name)
    dd = torch.float64
    prior = _prior(g)
    a = g["bary"].shape[2]
    offs = O.rotation_offsets(a, g["delta"])
    w_eff = O.fold_prior(g["w_neighbor"].to(dd), None if prior is None else prior.to(dd))
    y = O.conv_forward_folded(g["signal"].to(dd), g["bary"], w_eff, g["w_self"].to(dd), g["bias"].to(dd),
                              g["act"], offs)
    assert nerr(y, g["y"]) < 2e-6
    z, arg = O.amp_forward(y)
    assert torch.equal(arg, g["argmax"])
    rot = torch.tensor(offs)[arg.long()]
    dx, dweff, dws, db = O.conv_amp_backward_closed_form(
        g["signal"].to(dd), g["bary"], w_eff, g["w_self"].to(dd), z, rot, g["grad_z"].to(dd), g["act"])
    dwn = O.unfold_prior_grad(dweff, None if prior is None else prior.to(dd))
    assert nerr(dx, g["d_signal"]) < TOL_FP32
    assert nerr(dwn, g["d_w_neighbor"]) < TOL_FP32
    assert nerr(dws, g["d_w_self"]) < TOL_FP32
    assert nerr(db, g["d_bias"]) < TOL_FP32


@pytest.mark.parametrize("act", sorted(O.ACT_IDS))
def test_activation_grad_from_output(act):
    x = torch.linspace(-3, 3, 101, dtype=torch.float64, requires_grad=True)
    y = O.activation(act, x)
    (gx,) = torch.autograd.grad(y.sum(), x)
    ref = getattr(torch.nn.functional, act)(x.detach())
    assert torch.allclose(y, ref, atol=1e-12)
    mask = x.detach().abs() > 1e-9   # relu/leaky kink
    assert torch.allclose(O.activation_grad_from_output(act, y.detach())[mask], gx[mask], atol=1e-12)
