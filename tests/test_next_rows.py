"""the other priors (ConvExp / ConvChiSquared / ConvStudentT) and poolings
(AngularMinPooling / AngularAvgPooling), pinned by fixtures from the executed reference."""
import numpy as np
import pytest
import torch

from oracle import geoconv_oracle as O
from tests.helpers import TOL_FP32, golden_cases, load_golden, nerr

NEXT = golden_cases("next")
LAYER_OF = {"exp": ("conv_exp", "ConvExp", "exp_lambda"), "chi_squared": ("conv_chi_squared", "ConvChiSquared", "dof"),
            "student_t": ("conv_student_t", "ConvStudentT", "dof")}


def _layer(g, **kw):
    import importlib
    mod, cls, key = LAYER_OF[g["kind"]]
    klass = getattr(importlib.import_module("geoconv_b200.pytorch.layers." + mod), cls)
    v, r, a = g["bary"].shape[:3]
    f, t = g["signal"].shape[1], g["bias"].shape[1]
    layer = klass(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.04,
                  activation=g["act"], rotation_delta=g["delta"], **{key: g["prior_param"]}, **kw)
    layer.load_state_dict({"_template_neighbor_weights": g["w_neighbor"], "_template_self_weights": g["w_self"],
                           "_bias": g["bias"]})
    return layer


def test_fixtures_present():
    assert len(NEXT) >= 7


@pytest.mark.parametrize("name", NEXT)
def test_layer_prior_matches_reference(name):
    """define_kernel_values of the drop-in classes (numpy, no GPU) against the reference's prior, including the
    reference's in-place normalisation of the template radii."""
    g = load_golden(name)
    layer = _layer(g)
    np.testing.assert_allclose(layer._kernel.numpy(), g["prior"].numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(layer._template_vertices.numpy(), g["template"].numpy(), rtol=0, atol=1e-15)
    assert layer._prior_is_symmetric


@pytest.mark.parametrize("name", NEXT)
def test_oracle_min_avg_pooling(name):
    g = load_golden(name)
    params = [g["signal"].clone().requires_grad_(True)] + [g[k].clone().requires_grad_(True)
                                                           for k in ("w_neighbor", "w_self", "bias")]
    y = O.conv_forward(params[0], g["bary"], params[1], params[2], params[3], g["prior"], g["act"], g["delta"])
    z_min, arg = O.amin_forward(y)
    assert torch.equal(arg, g["argmin"])
    assert nerr(z_min, g["z_min"]) < 2e-6
    z_avg = O.aavg_forward(y)
    assert nerr(z_avg, g["z_avg"]) < 2e-6
    for tag, z in (("min", z_min), ("avg", z_avg)):
        grads = torch.autograd.grad(z, params, grad_outputs=g["grad_z"], retain_graph=True)
        for key, gr in zip(("d_signal", "d_w_neighbor", "d_w_self", "d_bias"), grads):
            assert nerr(gr, g[f"{key}_{tag}"]) < TOL_FP32, (tag, key)


def test_min_pooling_module_keeps_the_reference_name():
    from geoconv_b200.pytorch.layers import angular_min_pooling as m
    assert m.AngularMaxPooling is m.AngularMinPooling   # sic: angular_min_pooling.py:6 of the reference


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "ffma"])
@pytest.mark.parametrize("name", NEXT)
def test_gpu_layers_and_poolings_match_reference(name, precision):
    from geoconv_b200.pytorch.layers.angular_avg_pooling import AngularAvgPooling
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.angular_min_pooling import AngularMinPooling
    g = load_golden(name)
    dev = "cuda:0"
    layer = _layer(g, precision=precision).to(dev)
    bary = g["bary"].to(dev)
    gz = g["grad_z"].to(dev)
    pools = (("", AngularMaxPooling(), "argmax"), ("_min", AngularMinPooling(), "argmin"), ("_avg", AngularAvgPooling(), None))
    for tag, pool, arg_key in pools:
        layer.zero_grad()
        x = g["signal"].to(dev).requires_grad_(True)
        y = layer([x, bary])
        assert nerr(y, g["y"]) < TOL_FP32
        z = pool(y)
        assert nerr(z, g["z" + tag]) < TOL_FP32, tag
        if arg_key:
            norms = torch.linalg.vector_norm(y, ord=2, dim=-1)
            arg = (torch.argmin if tag == "_min" else torch.argmax)(norms, dim=1).int().cpu()
            assert torch.equal(arg, g[arg_key]), tag
        z.backward(gz)
        assert nerr(x.grad, g["d_signal" + tag]) < TOL_FP32, tag
        assert nerr(layer._template_neighbor_weights.grad, g["d_w_neighbor" + tag]) < TOL_FP32, tag
        assert nerr(layer._template_self_weights.grad, g["d_w_self" + tag]) < TOL_FP32, tag
        assert nerr(layer._bias.grad, g["d_bias" + tag]) < TOL_FP32, tag
