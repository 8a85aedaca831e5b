"""CPU-side checks: the C-ABI library loads and exports what include/geoconv_b200.h declares, the
drop-in layers mirror the reference's construction (seed-for-seed), and there is no CPU path."""
import os
import re

import numpy as np
import pytest
import torch

from geoconv_b200 import _native
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic
from geoconv_b200.pytorch.layers.conv_zero import ConvZero
from geoconv_b200.preprocessing.barycentric_coordinates import create_template_matrix
from tests.helpers import golden_cases, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = {"geodesic": ConvGeodesic, "dirac": ConvDirac, "zero": ConvZero}


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "geoconv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    declared = _header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/geoconv_b200.h but not exported"
    assert sorted(_native.SIGNATURES) == declared
    assert lib.gcb_abi_version() == _native.ABI_VERSION


def test_error_reporting_without_gpu():
    lib = _native.load()
    # argument validation happens before any CUDA call
    rc = lib.gcb_prep_bary(None, 0, 10, 5, None, None, None, None)
    assert rc == 1
    assert b"null pointer" in lib.gcb_last_error()
    with pytest.raises(_native.GeoconvNativeError):
        _native.check(rc, "gcb_prep_bary")


def test_precision_resolution():
    """'fp32' resolves to the fastest fp32-faithful arithmetic the shape allows: fp16 hi/lo splits (F % 16 == 0),
    else tf32 splits, else CUDA cores; explicit modes degrade the same way (host-side tiling queries, no GPU)."""
    from geoconv_b200 import ops
    lib = _native.load()
    assert lib.gcb_tc_supported(5, 8, 544, 96, 8) == 1 and lib.gcb_tc_half_supported(5, 8, 544, 96, 8) == 1
    assert ops.resolve_precision("fp32", 5, 8, 544, 96, 8) == _native.PREC_F16X3
    assert ops.resolve_precision("f16x3", 5, 8, 544, 96, 8) == _native.PREC_F16X3
    assert ops.resolve_precision("tf32x3", 5, 8, 544, 96, 8) == _native.PREC_TF32X3
    assert ops.resolve_precision("tf32", 5, 8, 544, 96, 8) == _native.PREC_TF32
    assert ops.resolve_precision("ffma", 5, 8, 544, 96, 8) == _native.PREC_FP32
    # F = 40: tensor tiling exists (F % 8 == 0) but no fp16 row of 16 features
    assert lib.gcb_tc_supported(2, 6, 40, 48, 3) == 1 and lib.gcb_tc_half_supported(2, 6, 40, 48, 3) == 0
    assert ops.resolve_precision("fp32", 2, 6, 40, 48, 3) == _native.PREC_TF32X3
    assert ops.resolve_precision("f16x3", 2, 6, 40, 48, 3) == _native.PREC_TF32X3
    # F = 3 (xyz signal): no tensor tiling at all
    assert lib.gcb_tc_supported(5, 8, 3, 32, 8) == 0
    assert ops.resolve_precision("fp32", 5, 8, 3, 32, 8) == _native.PREC_FP32
    assert ops.resolve_precision("tf32", 5, 8, 3, 32, 8) == _native.PREC_FP32
    with pytest.raises(ValueError):
        ops.resolve_precision("bf16", 5, 8, 544, 96, 8)


def test_fp16_split_shape_check_in_the_c_abi():
    """gcb_conv_fwd refuses GCB_PREC_F16X3 for a shape without an fp16 tiling before any CUDA call, and the
    workspace query covers the fp16 mode's extra buffers (scales, ring schedules, split table)."""
    import ctypes as C
    lib = _native.load()
    v, r, a, f, t, n_rot = 256, 2, 6, 40, 48, 3
    need = lib.gcb_conv_fwd_workspace_bytes(v, v, r, a, f, t, n_rot, _native.PREC_F16X3)
    rot = (C.c_int32 * n_rot)(0, 2, 4)
    fake = C.c_void_p(256)   # never dereferenced: the shape check comes first
    rc = lib.gcb_conv_fwd(fake, fake, fake, fake, fake, fake, rot, n_rot, v, v, r, a, f, t, 0, _native.PREC_F16X3,
                          fake, None, None, None, None, None, None, 0, fake, C.c_size_t(max(need, 1) + (1 << 30)), None)
    assert rc == 3, lib.gcb_last_error()
    assert b"fp16 split" in lib.gcb_last_error()
    # FAUST shape: the fp16 mode needs less workspace than 3xTF32 with its 12 K splits, more than nothing
    w16 = lib.gcb_conv_fwd_workspace_bytes(6890, 6890, 5, 8, 544, 96, 8, _native.PREC_F16X3)
    w32 = lib.gcb_conv_fwd_workspace_bytes(6890, 6890, 5, 8, 544, 96, 8, _native.PREC_TF32X3)
    assert 0 < w16 < w32
    b16 = lib.gcb_conv_bwd_workspace_bytes(6890, 6890, 5, 8, 544, 96, _native.PREC_F16X3)
    b32 = lib.gcb_conv_bwd_workspace_bytes(6890, 6890, 5, 8, 544, 96, _native.PREC_TF32X3)
    assert b16 == b32   # the fp16 buffers live inside the 3xTF32 layout


def _make(g, **kw):
    v, r, a = g["bary"].shape[:3]
    f = g["signal"].shape[1]
    t = g["w_self"].shape[0]
    radius = g["template"][-1, 0, 0].item()
    torch.manual_seed(0)
    return LAYERS[g["kind"]](input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t,
                             template_radius=radius, activation=g["act"], rotation_delta=g["delta"], **kw)


@pytest.mark.parametrize("name", golden_cases())
def test_seeded_construction_matches_reference(name):
    g = load_golden(name)
    layer = _make(g)
    sd = layer.state_dict()
    assert list(sd.keys()) == ["_template_neighbor_weights", "_template_self_weights", "_bias"]
    assert torch.equal(sd["_template_neighbor_weights"], g["w_neighbor"])
    assert torch.equal(sd["_template_self_weights"], g["w_self"])
    assert torch.equal(sd["_bias"], g["bias"])
    np.testing.assert_allclose(layer._kernel.numpy(), g["prior"].numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(layer._template_vertices.numpy(), g["template"].numpy(), rtol=0, atol=1e-15)
    assert layer._template_size == tuple(g["bary"].shape[1:3])
    assert layer.include_prior == g["include_prior"]


def test_public_attributes_and_hook():
    layer = ConvGeodesic(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=3, template_radius=0.1,
                         activation="elu", rotation_delta=2, initializer="xavier_normal")
    assert (layer.activation_fn, layer.rotation_delta, layer.amt_templates) == ("elu", 2, 3)
    assert (layer.template_radius, layer.initializer, layer.include_prior) == (0.1, "xavier_normal", True)
    assert layer._rotation_offsets(None) == [0, 2]
    assert layer._rotation_offsets(torch.tensor([3, 1])) == [3, 1]
    assert ConvDirac(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=3, template_radius=0.1).include_prior is False
    assert ConvZero(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=3, template_radius=0.1)._prior_is_zero

    class Lopsided(ConvIntrinsic):
        def define_kernel_values(self, template_matrix):
            k = np.zeros(template_matrix.shape[:-1] + template_matrix.shape[:-1])
            k[0, 0, 0, 1] = 1.0
            return k

    assert Lopsided(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=3, template_radius=0.1)._prior_is_symmetric is False
    with pytest.raises(TypeError):
        ConvIntrinsic(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=3, template_radius=0.1)


def test_template_matrix_cartesian():
    pol = create_template_matrix(3, 4, 0.3)
    car = create_template_matrix(3, 4, 0.3, in_cart=True)
    np.testing.assert_allclose(car[..., 0], pol[..., 0] * np.cos(pol[..., 1]))
    np.testing.assert_allclose(np.hypot(car[..., 0], car[..., 1]), pol[..., 0])


def test_no_cpu_path():
    layer = ConvDirac(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=3, template_radius=0.1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer([torch.zeros(5, 6), torch.zeros(5, 2, 4, 3, 2)])
    with pytest.raises(RuntimeError, match="no CPU path"):
        AngularMaxPooling()(torch.zeros(5, 4, 3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "geoconv_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, fn)).read(), fn


def test_backward_workspace_covers_every_internal_layout():
    """the fp16 route of the backward picks its own vertex-split count for the G^T X partial sums; the
    workspace query must cover the tf32, fp16 and centre-only layouts for every shape (host-only self-check)."""
    import itertools
    from geoconv_b200 import _native
    lib = _native.load()
    for v, r, a, f, t in itertools.product([100, 6890, 14000, 20000, 50000, 250000], [1, 3, 4, 5, 8], [6, 8, 12, 16],
                                           [32, 64, 128, 384, 544], [32, 96, 128, 384]):
        for prec in (_native.PREC_TF32, _native.PREC_TF32X3, _native.PREC_F16X3):
            assert lib.gcb_conv_bwd_workspace_selfcheck(v, v, r, a, f, t, prec) == 1, (v, r, a, f, t, prec)


def test_check_gradients_helper_both_branches():
    """tests/helpers.py::check_gradients on CPU data: the autograd branch when the decisions agree, the closed
    form (row subset against the full signal included) when a pooling decision differs."""
    import types
    import torch
    from oracle import geoconv_oracle as O
    from tests.helpers import check_gradients
    v_src, v, f, r, a, t = 90, 40, 8, 2, 4, 8
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v_src, f, seed=1, kind="randn")
    bary = O.make_bary(v, r, a, seed=1, v_src=v_src)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(2))
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.1))).float()
    ref = O.conv_amp_forward_backward(signal, bary, wn, ws, b, prior, gz, "relu", 1)

    def fake_layer(gr):
        ns = types.SimpleNamespace
        return ns(_template_neighbor_weights=ns(grad=gr["d_w_neighbor"]), _template_self_weights=ns(grad=gr["d_w_self"]),
                  _bias=ns(grad=gr["d_bias"]))
    how = check_gradients(fake_layer(ref), ref["d_signal"], ref["z"], ref["argmax"], ref, signal, bary, wn, ws, prior, gz,
                          "relu", range(a), 1e-5)
    assert how == "autograd"
    # force another rotation for one vertex: the "kernel" result is then the closed form at that decision
    arg2 = ref["argmax"].clone()
    arg2[3] = (arg2[3] + 1) % a
    z2 = ref["y"][torch.arange(v), arg2.long()]
    xs = signal.clone().requires_grad_(True)
    ps = [p_.clone().requires_grad_(True) for p_ in (wn, ws, b)]
    y = O.conv_forward(xs, bary, ps[0], ps[1], ps[2], prior, "relu", 1)
    g2 = torch.autograd.grad(y[torch.arange(v), arg2.long()], [xs] + ps, grad_outputs=gz)
    got = {"d_signal": g2[0], "d_w_neighbor": g2[1], "d_w_self": g2[2], "d_bias": g2[3]}
    how = check_gradients(fake_layer(got), got["d_signal"], z2, arg2, ref, signal, bary, wn, ws, prior, gz, "relu",
                          range(a), 1e-5)
    assert how == "closed-form"
