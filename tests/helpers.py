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
