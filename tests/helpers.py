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
    names = [n for n in names if not n.startswith("gpc_")]
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
