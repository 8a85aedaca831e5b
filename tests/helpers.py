"""Shared helpers for the test-suite: golden fixture loading and error metrics."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


NEXT_PREFIXES = ("exp", "chi", "student")   # fixtures of the SURVEY.md §8(f-4) rows (other priors / poolings)


def golden_cases(which="path"):
    """Fixture names: "path" = the hot path's own layers (geodesic / dirac / zero), "next" = the §8(f-4)
    rows, "all" = both."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if which == "gpc":     # §8(f-1) fixtures: GPC systems -> barycentric coordinates
        return [n for n in names if n.startswith("gpc_")]
    if which == "wave":    # §8(f-1) fixtures: mesh -> GPC systems (the wavefront)
        return [n for n in names if n.startswith("wave_")]
    names = [n for n in names if not n.startswith(("gpc_", "wave_", "faust_reader_"))]   # (§8 f-3 reader fixtures)
    is_next = lambda n: n.split("_")[0] in NEXT_PREFIXES  # noqa: E731
    if which == "all":
        return names
    return [n for n in names if is_next(n) == (which == "next")]


def load_golden(name):
    raw = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in raw.files:
        a = raw[k]
        if a.dtype.kind in "US":
            out[k] = str(a)
        elif a.ndim == 0:
            out[k] = a.item()
        else:
            out[k] = torch.from_numpy(a)
    return out


def nerr(a, b):
    """Normalised max error  max|a-b| / max|b|  (elementwise relative error is
    meaningless next to ReLU zeros)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = torch.linalg.vector_norm(b).item()
    if denom == 0.0:
        return torch.linalg.vector_norm(a - b).item()
    return torch.linalg.vector_norm(a - b).item() / denom


def check_gradients(layer, signal_grad, z, arg, ref, signal, bary, w_neighbor, w_self, prior, grad_z, act, offsets,
                    tol, y_ref_max=None):
    """Gradient parity without escape hatches.

    `ref` holds the oracle's autograd result for the same inputs.  If this forward selected the same rotation
    for every vertex and no activation-derivative bit differs, the gradients are compared with that autograd
    result directly.  Otherwise (a near-tie in the angular max-pooling went the other way, or a ReLU
    pre-activation within rounding of zero changed sign - both admissible) the gradients
    legitimately belong to a different piecewise-linear branch than the oracle's: they are then compared with
    the closed-form backward (float64) evaluated at THIS forward's own outputs and selected rotations.
    Returns "autograd" or "closed-form" (which comparison ran)."""
    from oracle import geoconv_oracle as O
    zc = z.detach().cpu()
    argc = arg.cpu().long()
    same_rot = torch.equal(argc, ref["argmax"].long())
    dz_own = O.activation_grad_from_output(act, zc.double())
    dz_ref = O.activation_grad_from_output(act, ref["z"].double())
    same_mask = same_rot and bool(((dz_own - dz_ref).abs() <= 10 * tol).all())
    got = {"d_signal": signal_grad, "d_w_neighbor": layer._template_neighbor_weights.grad,
           "d_w_self": layer._template_self_weights.grad, "d_bias": layer._bias.grad}
    if same_rot and same_mask:
        want, how = {k: ref[k] for k in got}, "autograd"
    else:
        rot = torch.tensor([int(o) for o in offsets])[argc]
        w_eff = O.fold_prior(w_neighbor.double(), None if prior is None else prior.double())
        if prior is not None and not bool(prior.any()):
            w_eff = torch.zeros_like(w_eff)
        dx, dweff, dws, db = O.conv_amp_backward_closed_form(signal.double(), bary, w_eff, w_self.double(), zc.double(),
                                                             rot, grad_z.double(), act)
        want = {"d_signal": dx, "d_w_neighbor": O.unfold_prior_grad(dweff, None if prior is None else prior.double()),
                "d_w_self": dws, "d_bias": db}
        how = "closed-form"
    errs = {k: nerr(got[k], want[k]) for k in got if got[k] is not None}
    assert all(e < tol for e in errs.values()), (how, errs)
    for k in got:
        assert got[k] is None or tuple(got[k].shape) == tuple(want[k].shape), k
    return how


def argmax_admissible(arg, ref_arg, y_ref, tol):
    """Exact, except where the reference's best and runner-up norms are closer than `tol` (relative): there a
    different-but-equally-valid rotation may win."""
    arg = arg.cpu().long()
    ref_arg = ref_arg.long()
    bad = (arg != ref_arg).nonzero().flatten()
    if bad.numel() == 0:
        return True
    norms = torch.linalg.vector_norm(y_ref.double(), dim=-1)
    n_ref = norms[bad, ref_arg[bad]]
    n_got = norms[bad, arg[bad]]
    return bool(((n_ref - n_got).abs() <= tol * n_ref.abs().clamp_min(1e-30)).all())


# Tolerances stated by BASELINE.json's north_star (normalised metrics)
TOL_FP32 = 1e-5
TOL_TF32 = 2e-3


# --------------------------------------------------------------------------------------------
# Synthetic GPC systems for the f-1 row (GPC systems -> barycentric coordinates)
# --------------------------------------------------------------------------------------------
class _FakeGpcSystem:
    """The part of geoconv.preprocessing.gpc_system.GPCSystem that compute_barycentric_coordinates touches."""

    def __init__(self, tri_xy, faces):
        self._tri_xy = tri_xy
        self.faces = {(-1, -1): [list(map(int, f)) for f in faces]}

    def get_gpc_triangles(self, in_cart=False):
        assert in_cart
        return self._tri_xy


class _FakeGpcGroup:
    def __init__(self, systems):
        self.object_mesh_gpc_systems = np.array(systems, dtype=object)


def synthetic_gpc_systems(v, radius, seed, rings=3, spokes=9):
    """Seeded stand-in for a GPCSystemGroup: around every vertex a triangulated disc (centre + `rings` jittered
    rings of `spokes` points) reaching ~1.2 * radius, so some outer template vertices fall outside every
    triangle (the reference's zero fallback), with random mesh vertex ids on the corners and the faces in a
    shuffled order (the first containing triangle wins)."""
    rng = np.random.default_rng(seed)
    systems = []
    for _ in range(v):
        pts = [np.zeros(2)]
        for k in range(1, rings + 1):
            ang = np.sort(rng.uniform(0, 2 * np.pi, spokes))
            rad = radius * 1.2 * k / rings * rng.uniform(0.85, 1.0, spokes)
            pts.extend(np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=-1))
        pts = np.array(pts)
        ids = rng.integers(0, v, len(pts))
        faces = []
        ring = lambda k, i: 1 + (k - 1) * spokes + (i % spokes)  # noqa: E731
        for i in range(spokes):
            faces.append((0, ring(1, i), ring(1, i + 1)))
            for k in range(1, rings):
                faces.append((ring(k, i), ring(k + 1, i), ring(k + 1, i + 1)))
                faces.append((ring(k, i), ring(k + 1, i + 1), ring(k, i + 1)))
        faces = np.array(faces)[rng.permutation(len(faces))]
        systems.append(_FakeGpcSystem(pts[faces], ids[faces]))
    return _FakeGpcGroup(systems)
