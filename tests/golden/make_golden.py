"""Generate golden input/output vectors by EXECUTING the unmodified reference layers.

Run in the build container only (the reference is mounted read-only at /root/reference and does
not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/<case>.npz.  Every fixture stores the inputs (signal, bary, the three
parameters, the prior the reference built, upstream gradient) and the reference's outputs
(Y, Z, AMP argmax, and the four gradients from autograd).  The reference modules used are
  /root/reference/src/geoconv/pytorch/layers/conv_geodesic.py, conv_dirac.py, conv_zero.py
  /root/reference/src/geoconv/pytorch/layers/angular_max_pooling.py
and, for the "next" rows of SURVEY.md §8(f-4):
  conv_exp.py, conv_chi_squared.py, conv_student_t.py (its `np.math` does not exist in NumPy 2: the
  generator sets `np.math = math`, which is what older NumPy exposed), angular_min_pooling.py,
  angular_avg_pooling.py.
"""
import os
import sys

import math

import numpy as np
import torch

if not hasattr(np, "math"):
    np.math = math   # conv_student_t.py:25-27 was written against NumPy 1.x

sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from geoconv.pytorch.layers.angular_max_pooling import AngularMaxPooling  # noqa: E402
from geoconv.pytorch.layers.conv_dirac import ConvDirac  # noqa: E402
from geoconv.pytorch.layers.conv_geodesic import ConvGeodesic  # noqa: E402
from geoconv.pytorch.layers.conv_zero import ConvZero  # noqa: E402
from geoconv.pytorch.layers.conv_exp import ConvExp  # noqa: E402
from geoconv.pytorch.layers.conv_chi_squared import ConvChiSquared  # noqa: E402
from geoconv.pytorch.layers.conv_student_t import ConvStudentT  # noqa: E402
from geoconv.pytorch.layers.angular_min_pooling import AngularMaxPooling as AngularMinPooling  # noqa: E402  (sic: the reference names the min-pooling class AngularMaxPooling)
from geoconv.pytorch.layers.angular_avg_pooling import AngularAvgPooling  # noqa: E402

from oracle.geoconv_oracle import make_bary, make_signal  # noqa: E402

LAYERS = {"geodesic": ConvGeodesic, "dirac": ConvDirac, "zero": ConvZero, "exp": ConvExp,
          "chi_squared": ConvChiSquared, "student_t": ConvStudentT}
PRIOR_KWARG = {"exp": "exp_lambda", "chi_squared": "dof", "student_t": "dof"}

# name, kind, V, F, R, A, T, delta, activation, bary dtype, signal kind, pool, template_radius
CASES = [
    ("geodesic_small",      "geodesic", 61, 12, 3, 4, 8,  1, "relu",       torch.float32, "randn", None, 0.03),
    ("geodesic_r5a8",       "geodesic", 97, 20, 5, 8, 16, 1, "relu",       torch.float32, "shot",  None, 0.028),
    ("geodesic_delta2",     "geodesic", 83, 16, 5, 8, 12, 2, "relu",       torch.float32, "randn", None, 0.028),
    ("geodesic_a6_delta4",  "geodesic", 40, 9,  3, 6, 5,  4, "elu",        torch.float32, "randn", None, 0.05),
    ("geodesic_f64bary",    "geodesic", 50, 8,  2, 4, 6,  1, "tanh",       torch.float64, "randn", None, 0.05),
    ("geodesic_pool",       "geodesic", 70, 24, 4, 12, 10, 3, "leaky_relu", torch.float32, "randn", 7,   0.04),
    ("dirac_r5a8",          "dirac",    90, 20, 5, 8, 16, 1, "relu",       torch.float32, "shot",  None, 0.028),
    ("dirac_selu",          "dirac",    45, 7,  3, 6, 4,  1, "selu",       torch.float32, "randn", None, 0.028),
    ("dirac_sigmoid",       "dirac",    33, 5,  2, 5, 3,  2, "sigmoid",    torch.float32, "randn", None, 0.028),
    ("zero_r5a8",           "zero",     64, 20, 5, 8, 16, 1, "relu",       torch.float32, "shot",  None, 0.028),
    ("geodesic_single_vertex", "geodesic", 1, 4, 2, 3, 2, 1, "relu",       torch.float32, "randn", None, 0.028),
    # other priors (prior parameter appended) - these fixtures also hold min / avg pooling outputs and gradients
    ("exp_r5a8",            "exp",         72, 16, 5, 8, 12, 1, "relu",    torch.float32, "shot",  None, 0.028, 1),
    ("exp_lambda3",         "exp",         41, 10, 3, 6, 6,  2, "tanh",    torch.float32, "randn", None, 0.05, 3),
    ("chi_dof2",            "chi_squared", 66, 12, 4, 8, 8,  1, "relu",    torch.float32, "randn", None, 0.028, 2),
    ("chi_dof1",            "chi_squared", 38, 8,  3, 5, 5,  1, "elu",     torch.float32, "randn", None, 0.04, 1),
    ("chi_dof3",            "chi_squared", 30, 6,  2, 4, 4,  2, "sigmoid", torch.float32, "randn", None, 0.04, 3),
    ("student_dof2",        "student_t",   58, 12, 5, 8, 8,  1, "relu",    torch.float32, "shot",  None, 0.028, 2),
    ("student_dof3",        "student_t",   35, 9,  3, 6, 6,  3, "selu",    torch.float32, "randn", None, 0.05, 3),
]


def run_case(name, kind, v, f, r, a, t, delta, act, bdtype, skind, pool, radius, prior_param=None):
    torch.manual_seed(0)
    extra = {PRIOR_KWARG[kind]: prior_param} if prior_param is not None else {}
    layer = LAYERS[kind](input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t,
                         template_radius=radius, activation=act, rotation_delta=delta, **extra)
    amp = AngularMaxPooling()
    signal = make_signal(v, f, seed=1, kind=skind)
    bary = make_bary(v, r, a, seed=1, fallback_frac=0.05, pool=pool, dtype=bdtype)
    signal.requires_grad_(True)
    y = layer([signal, bary])
    z = amp(y)
    g = torch.Generator().manual_seed(2)
    grad_z = torch.randn(z.shape, generator=g)
    params = [layer._template_neighbor_weights, layer._template_self_weights, layer._bias]
    grads = torch.autograd.grad(z, [signal] + params, grad_outputs=grad_z, allow_unused=True, retain_graph=True)
    norms = torch.linalg.vector_norm(y, ord=2, dim=-1)
    arg = torch.argmax(norms, dim=1).int()
    out = {
        "kind": np.array(kind), "act": np.array(act), "delta": np.array(delta),
        "include_prior": np.array(bool(layer.include_prior)),
        "signal": signal.detach().numpy(), "bary": bary.numpy(),
        "w_neighbor": params[0].detach().numpy(), "w_self": params[1].detach().numpy(),
        "bias": params[2].detach().numpy(), "prior": layer._kernel.numpy(),
        "template": layer._template_vertices.numpy(),
        "grad_z": grad_z.numpy(),
        "y": y.detach().numpy(), "z": z.detach().numpy(), "argmax": arg.numpy(),
        "d_signal": grads[0].numpy(),
        "d_w_neighbor": (grads[1] if grads[1] is not None else torch.zeros_like(params[0])).numpy(),
        "d_w_self": grads[2].numpy(), "d_bias": grads[3].numpy(),
    }
    if prior_param is not None:
        out["prior_param"] = np.array(prior_param)
        for tag, pool_mod in (("min", AngularMinPooling()), ("avg", AngularAvgPooling())):
            zp = pool_mod(y)
            gp = torch.autograd.grad(zp, [signal] + params, grad_outputs=grad_z, allow_unused=True, retain_graph=True)
            out["z_" + tag] = zp.detach().numpy()
            out["d_signal_" + tag] = gp[0].numpy()
            out["d_w_neighbor_" + tag] = gp[1].numpy()
            out["d_w_self_" + tag] = gp[2].numpy()
            out["d_bias_" + tag] = gp[3].numpy()
        out["argmin"] = torch.argmin(norms, dim=1).int().numpy()
    # explicit orientations argument (forward only): reversed rotation list
    orient = torch.arange(0, a, delta).flip(0)
    out["orientations"] = orient.numpy()
    out["y_orient"] = layer([signal.detach(), bary], orientations=orient).detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: y{tuple(y.shape)} argmax hist {np.bincount(arg.numpy(), minlength=y.shape[1]).tolist()}")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
