"""CPU oracle for the ConvIntrinsic + AngularMaxPooling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``geoconv_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker or as the timed CPU baseline.

It restates, with stock ``torch`` CPU ops and in the reference's own evaluation order
(materialised gather -> weighted sum -> prior einsum -> one roll + einsum per rotation ->
stack/permute -> activation; norm -> argmax -> row select), what these reference functions compute:

* ``create_template_matrix``              src/geoconv/preprocessing/barycentric_coordinates.py:91-124
* ``ConvGeodesic.define_kernel_values``   src/geoconv/pytorch/layers/conv_geodesic.py:52-68 (+ :7-41)
* ``ConvDirac`` / ``ConvZero`` priors     src/geoconv/pytorch/layers/conv_dirac.py:13-26, conv_zero.py:12-15
* ``ConvIntrinsic._signal_retrieval``     src/geoconv/pytorch/layers/conv_intrinsic.py:198-240
* ``ConvIntrinsic._patch_operator``       src/geoconv/pytorch/layers/conv_intrinsic.py:173-196
* ``ConvIntrinsic.forward``               src/geoconv/pytorch/layers/conv_intrinsic.py:118-171
* ``AngularMaxPooling.forward``           src/geoconv/pytorch/layers/angular_max_pooling.py:9-29

Parity pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the *executed* reference, generated in the build container
by ``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every fixture.  Gradients come from autograd over this
restatement (exactly how the reference obtains them) and, independently, from the closed-form
expressions in :func:`conv_amp_backward_closed_form`.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# Activations: the six entries of the reference's ACTIVATIONS table (conv_intrinsic.py:10-17),
# all with torch's default hyper-parameters.  The integer ids are the C-ABI's ``act_id``.
# --------------------------------------------------------------------------------------------
ACT_IDS = {"relu": 0, "elu": 1, "leaky_relu": 2, "selu": 3, "sigmoid": 4, "tanh": 5}

_SELU_ALPHA = 1.6732632423543772848170429916717
_SELU_SCALE = 1.0507009873554804934193349852946


def activation(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "relu":
        return torch.clamp_min(x, 0.0)
    if name == "elu":
        return torch.where(x > 0, x, torch.expm1(x))
    if name == "leaky_relu":
        return torch.where(x > 0, x, 0.01 * x)
    if name == "selu":
        return _SELU_SCALE * torch.where(x > 0, x, _SELU_ALPHA * torch.expm1(x))
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    raise KeyError(name)


def activation_grad_from_output(name: str, y: torch.Tensor) -> torch.Tensor:
    """d act / d pre-activation written in terms of the activation's *output* y."""
    one = torch.ones_like(y)
    if name == "relu":
        return (y > 0).to(y.dtype)
    if name == "elu":
        return torch.where(y > 0, one, y + 1.0)
    if name == "leaky_relu":
        return torch.where(y > 0, one, 0.01 * one)
    if name == "selu":
        return torch.where(y > 0, _SELU_SCALE * one, y + _SELU_SCALE * _SELU_ALPHA)
    if name == "sigmoid":
        return y * (1.0 - y)
    if name == "tanh":
        return 1.0 - y * y
    raise KeyError(name)


# --------------------------------------------------------------------------------------------
# Build-time constants
# --------------------------------------------------------------------------------------------
def template_matrix(n_radial: int, n_angular: int, radius: float) -> np.ndarray:
    """Polar template grid (R, A, 2): rho_j = j*radius/R (j=1..R), theta_k = 2*pi*k/A (k=1..A).

    Follows barycentric_coordinates.py:91-124 (polar branch).
    """
    rho = (np.arange(1, n_radial + 1, dtype=np.float64) * radius) / n_radial
    theta = (2 * np.arange(1, n_angular + 1, dtype=np.float64) * np.pi) / n_angular
    out = np.zeros((n_radial, n_angular, 2), dtype=np.float64)
    out[:, :, 0] = rho[:, None]
    out[:, :, 1] = theta[None, :]
    return out


def geodesic_prior(template: np.ndarray) -> np.ndarray:
    """Gaussian-in-(rho, theta) interpolation coefficients, soft-maxed per (r, a).

    Follows conv_geodesic.py:52-68 with normal_pdf (:11-41) and angle_distance (:7-8):
    K[r,a,x,y] = softmax_{x,y}( c * exp(-0.5*((rho_x-rho_r)^2/var_rho + dtheta^2/var_theta)) ),
    dtheta = min(|t1-t2|, 2*pi-|t1-t2|), variances = population variance over the template.
    """
    rho = template[:, :, 0]
    theta = template[:, :, 1]
    var_rho = rho.var()
    var_theta = theta.var()
    coeff = 1.0 / np.sqrt((2 * np.pi) ** 2 * var_rho * var_theta)
    d_rho = rho[None, None, :, :] - rho[:, :, None, None]
    hi = np.maximum(theta[:, :, None, None], theta[None, None, :, :])
    lo = np.minimum(theta[:, :, None, None], theta[None, None, :, :])
    d_theta = np.minimum(hi - lo, lo + 2.0 * np.pi - hi)
    pdf = coeff * np.exp(-0.5 * (d_rho * d_rho / var_rho + d_theta * d_theta / var_theta))
    flat = pdf.reshape(pdf.shape[0], pdf.shape[1], -1)
    flat = flat - flat.max(axis=-1, keepdims=True)
    e = np.exp(flat)
    sm = e / e.sum(axis=-1, keepdims=True)
    return sm.reshape(pdf.shape)


def dirac_prior(template: np.ndarray) -> np.ndarray:
    """Identity on the (r, a) axis (conv_dirac.py:13-26; visualisation only - Dirac skips the prior)."""
    r, a = template.shape[:2]
    return np.eye(r * a, dtype=np.float64).reshape(r, a, r, a)


def zero_prior(template: np.ndarray) -> np.ndarray:
    """All-zero coefficients (conv_zero.py:12-15)."""
    r, a = template.shape[:2]
    return np.zeros((r, a, r, a), dtype=np.float64)


def _softmax_over_template(pdf: np.ndarray) -> np.ndarray:
    flat = pdf.reshape(pdf.shape[0], pdf.shape[1], -1)
    flat = flat - flat.max(axis=-1, keepdims=True)
    e = np.exp(flat)
    return (e / e.sum(axis=-1, keepdims=True)).reshape(pdf.shape)


def _polar_distances(template: np.ndarray):
    """|rho - mean_rho| and the wrap-around angular distance for every (mean, point) pair, after the
    in-place radial normalisation the three priors below apply (conv_exp.py:47, conv_chi_squared.py:79,
    conv_student_t.py:75: template_matrix[:, :, 0] /= max)."""
    rho = template[:, :, 0] / template[:, :, 0].max()
    theta = template[:, :, 1]
    d_rho = np.abs(rho[None, None, :, :] - rho[:, :, None, None])
    hi = np.maximum(theta[:, :, None, None], theta[None, None, :, :])
    lo = np.minimum(theta[:, :, None, None], theta[None, None, :, :])
    return d_rho, np.minimum(hi - lo, lo + 2.0 * np.pi - hi)


def exp_prior(template: np.ndarray, exp_lambda=1) -> np.ndarray:
    """conv_exp.py:8-38, 41-62: lambda^2 * exp(-lambda * (d_rho + d_theta)), soft-maxed per (r, a)."""
    d_rho, d_theta = _polar_distances(template)
    return _softmax_over_template(exp_lambda ** 2 * np.exp(-exp_lambda * (d_rho + d_theta)))


def _chi_gamma(dof):
    """The reference's own recursion (conv_chi_squared.py:8-28) - not the Gamma function, kept as is."""
    r = dof / 2
    if dof == 1 / 2:
        return np.sqrt(np.pi)
    if dof == 1:
        return 1.0
    return r * _chi_gamma(dof - 1)


def chi_squared_prior(template: np.ndarray, dof=2) -> np.ndarray:
    """conv_chi_squared.py:31-70, 73-94, including its dof == 1 special cases (weight 1 at zero distance)."""
    d_rho, d_theta = _polar_distances(template)
    gamma = (1 / (2 ** (dof / 2) * _chi_gamma(dof))) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        pdf = gamma * d_rho ** (dof / 2 - 1) * d_theta ** (dof / 2 - 1) * np.exp(-(d_rho + d_theta) / 2)
    if dof == 1:
        pdf = np.where((d_theta == 0) | (d_rho == 0), 1.0, pdf)
    return _softmax_over_template(pdf)


def _student_gamma(x):
    """conv_student_t.py:9-27 (factorial based; np.math there is math)."""
    n = math.floor(x)
    if x - n == 1 / 2:
        return (math.factorial(2 * n) / (math.factorial(n) * 4 ** n)) * np.sqrt(np.pi)
    return math.factorial(n)


def student_t_prior(template: np.ndarray, dof=2) -> np.ndarray:
    """conv_student_t.py:30-65, 68-89."""
    d_rho, d_theta = _polar_distances(template)
    t_theta = (1 + d_theta ** 2 / dof) ** (-(dof + 1) / 2)
    t_rho = (1 + d_rho ** 2 / dof) ** (-(dof + 1) / 2)
    quotient = (_student_gamma((dof + 1) / 2) / (np.sqrt(dof * np.pi) * _student_gamma(dof / 2))) ** 2
    return _softmax_over_template(quotient * t_rho * t_theta)


PRIORS = {"geodesic": geodesic_prior, "dirac": dirac_prior, "zero": zero_prior, "exp": exp_prior,
          "chi_squared": chi_squared_prior, "student_t": student_t_prior}


def rotation_offsets(n_angular: int, rotation_delta: int) -> list[int]:
    """arange(0, A, delta) - conv_intrinsic.py:152-154."""
    return list(range(0, n_angular, rotation_delta))


# --------------------------------------------------------------------------------------------
# Forward, stage by stage
# --------------------------------------------------------------------------------------------
def split_bary(bary: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """bary (V,R,A,3,2) -> (idx int64 (V,R,A,3), w float32 (V,R,A,3)).

    conv_intrinsic.py:213-221: cast to float32 first, then ``.long()`` (truncation toward zero).
    """
    b = bary.to(torch.float32)
    return b[..., 0].long(), b[..., 1]


def signal_retrieval(signal: torch.Tensor, bary: torch.Tensor) -> torch.Tensor:
    """I[k,r,a,:] = sum_j w[k,r,a,j] * X[idx[k,r,a,j], :]   (conv_intrinsic.py:198-240)."""
    x = signal.to(torch.float32) if signal.dtype != torch.float64 else signal
    idx, w = split_bary(bary)
    w = w.to(x.dtype)
    v, r, a, _ = idx.shape
    rows = x.index_select(0, idx.reshape(-1)).reshape(v, r, a, 3, x.shape[1])
    return (rows * w.unsqueeze(-1)).sum(dim=-2)


def patch_operator(interp: torch.Tensor, prior: torch.Tensor | None) -> torch.Tensor:
    """P[k,r,a,f] = sum_{x,y} K[r,a,x,y] * I[k,x,y,f]; identity when prior is None (:173-196)."""
    if prior is None:
        return interp
    return torch.einsum("raxy,kxyf->kraf", prior.to(interp.dtype), interp)


def conv_forward(signal, bary, w_neighbor, w_self, bias, prior, act="relu", rotation_delta=1,
                 orientations=None):
    """ConvIntrinsic.forward (:118-171).  Returns Y (V, n_rot, T).

    ``prior`` is the (R,A,R,A) coefficient tensor or ``None`` for include_prior=False.
    """
    a = bary.shape[2]
    # (the reference has as many barycentric rows as signal rows; a row SUBSET of the barycentric tensor against
    # the full signal - a vertex-row shard, or a test that evaluates a few rows of a large mesh - produces those
    # rows of the full result, the centre term reading the subset's own signal rows)
    centre = torch.einsum("tef,kf->ket", w_self, signal[: bary.shape[0]])
    patches = patch_operator(signal_retrieval(signal, bary), prior)
    shifts = rotation_offsets(a, rotation_delta) if orientations is None else [int(o) for o in orientations]
    per_rot = [
        torch.einsum("traf,kraf->kt", w_neighbor, torch.roll(patches, shifts=o, dims=2))
        for o in shifts
    ]
    neigh = torch.stack(per_rot).permute(1, 0, 2)
    return activation(act, centre + neigh + bias)


def amp_forward(y: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """AngularMaxPooling.forward (:27-29).  Returns (Z (V,T), argmax int32 (V,)); first max wins."""
    norms = torch.linalg.vector_norm(y, ord=2, dim=-1)
    arg = torch.argmax(norms, dim=1)
    return y[torch.arange(y.shape[0]), arg], arg.to(torch.int32)


def amin_forward(y: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """AngularMinPooling.forward (angular_min_pooling.py:27-29): smallest norm, first minimum wins."""
    norms = torch.linalg.vector_norm(y, ord=2, dim=-1)
    arg = torch.argmin(norms, dim=1)
    return y[torch.arange(y.shape[0]), arg], arg.to(torch.int32)


def aavg_forward(y: torch.Tensor) -> torch.Tensor:
    """AngularAvgPooling.forward (angular_avg_pooling.py:27)."""
    return torch.mean(y, dim=1)


def conv_amp_forward(signal, bary, w_neighbor, w_self, bias, prior, act="relu", rotation_delta=1):
    y = conv_forward(signal, bary, w_neighbor, w_self, bias, prior, act, rotation_delta)
    z, arg = amp_forward(y)
    return y, z, arg


def conv_amp_forward_backward(signal, bary, w_neighbor, w_self, bias, prior, grad_z,
                              act="relu", rotation_delta=1):
    """Forward + autograd backward, the way the reference obtains gradients.

    Returns dict with y, z, argmax, d_signal, d_w_neighbor, d_w_self, d_bias.
    """
    xs = signal.detach().clone().requires_grad_(True)
    wn = w_neighbor.detach().clone().requires_grad_(True)
    ws = w_self.detach().clone().requires_grad_(True)
    bs = bias.detach().clone().requires_grad_(True)
    y, z, arg = conv_amp_forward(xs, bary, wn, ws, bs, prior, act, rotation_delta)
    gx, gwn, gws, gb = torch.autograd.grad(z, [xs, wn, ws, bs], grad_outputs=grad_z, allow_unused=True)
    if gwn is None:
        gwn = torch.zeros_like(wn)
    return {"y": y.detach(), "z": z.detach(), "argmax": arg, "d_signal": gx,
            "d_w_neighbor": gwn, "d_w_self": gws, "d_bias": gb}


# --------------------------------------------------------------------------------------------
# Closed forms the CUDA kernels implement (SURVEY.md §8a "algebraic restatement"), in any dtype
# --------------------------------------------------------------------------------------------
def fold_prior(w_neighbor: torch.Tensor, prior: torch.Tensor | None) -> torch.Tensor:
    """Weff[t,x,y,f] = sum_{r,a} W[t,r,a,f] * K[r,a,x,y]  (identity when prior is None)."""
    if prior is None:
        return w_neighbor
    return torch.einsum("traf,raxy->txyf", w_neighbor, prior.to(w_neighbor.dtype))


def prior_is_rotation_symmetric(prior: np.ndarray, rotation_delta: int = 1, tol: float = 1e-12) -> bool:
    """K[r,a,x,y] == K[r,(a+s)%A,x,(y+s)%A] for every rotation step s - what makes folding legal."""
    a = prior.shape[1]
    for s in range(0, a, rotation_delta):
        if not np.allclose(prior, np.roll(np.roll(prior, s, axis=1), s, axis=3), atol=tol, rtol=0):
            return False
    return True


def conv_forward_folded(signal, bary, w_eff, w_self, bias, act="relu", offsets=(0,)):
    """Y[k,i,t] = act(X[k]·Wself[t] + sum_{x,y,f} Weff[t,x,(y+o_i)%A,f] * I[k,x,y,f] + b[t])."""
    interp = signal_retrieval(signal, bary).to(w_eff.dtype)
    centre = signal[: bary.shape[0]].to(w_eff.dtype) @ w_self[:, 0, :].T
    outs = []
    for o in offsets:
        outs.append(torch.einsum("txyf,kxyf->kt", torch.roll(w_eff, shifts=-int(o), dims=2), interp))
    pre = torch.stack(outs, dim=1) + centre[:, None, :] + bias.reshape(1, 1, -1)
    return activation(act, pre)


def conv_amp_backward_closed_form(signal, bary, w_eff, w_self, z, rot_offset, grad_z, act="relu"):
    """AMP-sparse backward in closed form.

    ``rot_offset[k]`` is the angular shift o of the rotation AMP selected for vertex k.
    Returns (d_signal, d_w_eff, d_w_self, d_bias).
    """
    dt = w_eff.dtype
    idx, w = split_bary(bary)
    w = w.to(dt)
    v, r, a, _ = idx.shape
    f = signal.shape[1]
    h = grad_z.to(dt) * activation_grad_from_output(act, z.to(dt))
    interp = signal_retrieval(signal.to(dt), bary)
    ar = torch.arange(a)
    # rotated interpolation seen by the weights: Irot[k,x,y',f] = I[k,x,(y'-o_k)%A,f]
    src = (ar[None, :] - rot_offset[:, None].long()) % a
    irot = torch.gather(interp, 2, src[:, None, :, None].expand(v, r, a, f))
    d_w_eff = torch.einsum("kt,kxyf->txyf", h, irot)
    # dI[k,x,y,f] = sum_t h[k,t] * Weff[t,x,(y+o_k)%A,f]
    d_full = torch.einsum("kt,txyf->kxyf", h, w_eff)
    dst = (ar[None, :] + rot_offset[:, None].long()) % a
    d_interp = torch.gather(d_full, 2, dst[:, None, :, None].expand(v, r, a, f))
    d_signal = torch.zeros((signal.shape[0], f), dtype=dt)
    d_signal[:v] = h @ w_self[:, 0, :]          # (v < signal rows: a row subset against the full signal)
    contrib = (d_interp.unsqueeze(3) * w.unsqueeze(-1)).reshape(-1, f)
    d_signal = d_signal.index_add(0, idx.reshape(-1), contrib)
    d_w_self = (h.T @ signal[:v].to(dt)).unsqueeze(1)
    d_bias = h.sum(dim=0, keepdim=True)
    return d_signal, d_w_eff, d_w_self, d_bias


def unfold_prior_grad(d_w_eff: torch.Tensor, prior: torch.Tensor | None) -> torch.Tensor:
    """dW[t,r,a,f] = sum_{x,y} dWeff[t,x,y,f] * K[r,a,x,y]."""
    if prior is None:
        return d_w_eff
    return torch.einsum("txyf,raxy->traf", d_w_eff, prior.to(d_w_eff.dtype))


# --------------------------------------------------------------------------------------------
# Seeded input generators: defined in geoconv_b200/synthetic.py so that bench.py's
# product arm never imports this module; re-exported here for the tests.
# --------------------------------------------------------------------------------------------
from geoconv_b200.synthetic import make_bary, make_signal, make_weights, xavier_uniform  # noqa: E402,F401
