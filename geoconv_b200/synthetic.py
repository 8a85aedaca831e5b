"""Seeded synthetic workloads of SURVEY.md §8d (S1-S4): signals, barycentric tensors, weights.

Plain torch CPU generators shared by bench.py and the tests; no compute path here.
"""
from __future__ import annotations

import math

import torch


def make_bary(v, r, a, seed=1, fallback_frac=0.02, pool=None, dtype=torch.float32, v_src=None):
    """Synthetic barycentric tensor (V,R,A,3,2): uniform-random indices stored as floats,
    weights normalised to sum 1, a fraction of (k,r,a) entries set to the (0,0) fallback
    (barycentric_coordinates.py:69).  ``pool``: each vertex draws from a private pool of that
    many random vertices (patch-local reuse)."""
    g = torch.Generator().manual_seed(seed)
    v_src = v if v_src is None else v_src
    if pool is None:
        idx = torch.randint(0, v_src, (v, r, a, 3), generator=g)
    else:
        pools = torch.randint(0, v_src, (v, pool), generator=g)
        pick = torch.randint(0, pool, (v, r * a * 3), generator=g)
        idx = torch.gather(pools, 1, pick).reshape(v, r, a, 3)
    w = torch.rand((v, r, a, 3), generator=g)
    w = w / w.sum(dim=-1, keepdim=True)
    if fallback_frac > 0:
        dead = torch.rand((v, r, a), generator=g) < fallback_frac
        idx[dead] = 0
        w[dead] = 0.0
    return torch.stack([idx.to(dtype), w.to(dtype)], dim=-1)


def make_signal(v, f, seed=1, kind="shot"):
    g = torch.Generator().manual_seed(seed + 7919)
    if kind == "shot":
        x = torch.rand((v, f), generator=g)
        return x / torch.linalg.vector_norm(x, dim=1, keepdim=True)
    return torch.randn((v, f), generator=g)


def xavier_uniform(shape, gen):
    """nn.init.xavier_uniform_ fan computation for the reference's parameter shapes."""
    recept = 1
    for s in shape[2:]:
        recept *= s
    fan_in, fan_out = shape[1] * recept, shape[0] * recept
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def make_weights(t, r, a, f, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (xavier_uniform((t, r, a, f), g), xavier_uniform((t, 1, f), g), xavier_uniform((1, t), g))


def make_bary_local(v, r, a, window=512, seed=1, fallback_frac=0.02, dtype=torch.float32):
    """Barycentric tensor of a mesh whose vertex order preserves locality: the three
    indices of every template vertex lie within `window` rows of the vertex itself (wrapping around),
    so a row-partitioned mesh only needs a thin halo."""
    g = torch.Generator().manual_seed(seed)
    own = torch.arange(v).view(v, 1, 1, 1)
    idx = (own + torch.randint(-window, window + 1, (v, r, a, 3), generator=g)) % v
    w = torch.rand((v, r, a, 3), generator=g)
    w = w / w.sum(dim=-1, keepdim=True)
    if fallback_frac > 0:
        dead = torch.rand((v, r, a), generator=g) < fallback_frac
        idx[dead] = 0
        w[dead] = 0.0
    return torch.stack([idx.to(dtype), w.to(dtype)], dim=-1)


def make_bary_local_rows(v_total, lo, hi, r, a, window=512, seed=1, fallback_frac=0.02, dtype=torch.float32):
    """Rows [lo, hi) of a locality-preserving barycentric tensor of a `v_total`-vertex mesh (global indices within
    `window` rows of the vertex, wrapping around), generated without the other rows: what one rank of a
    row-partitioned mesh holds.  Seeded per row block, so different partitions give different (equally valid) meshes."""
    g = torch.Generator().manual_seed(seed * 1_000_003 + lo)
    n = hi - lo
    own = torch.arange(lo, hi).view(n, 1, 1, 1)
    idx = (own + torch.randint(-window, window + 1, (n, r, a, 3), generator=g)) % v_total
    w = torch.rand((n, r, a, 3), generator=g)
    w = w / w.sum(dim=-1, keepdim=True)
    if fallback_frac > 0:
        dead = torch.rand((n, r, a), generator=g) < fallback_frac
        idx[dead] = 0
        w[dead] = 0.0
    return torch.stack([idx.to(dtype), w.to(dtype)], dim=-1)
