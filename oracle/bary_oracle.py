"""CPU oracle for SURVEY.md §8 f-1: GPC systems -> barycentric coordinates.

TEST INFRASTRUCTURE ONLY (same rule as geoconv_oracle.py).  Restates, loop for loop,
/root/reference/src/geoconv/preprocessing/barycentric_coordinates.py:7-43 (compute_barycentric),
:46-69 (interpolation: first containing triangle wins) and :127-175 (compute_barycentric_coordinates) on the
packed triangle arrays of geoconv_b200.preprocessing.barycentric_coordinates.pack_gpc_triangles.
Pinned by tests/golden/gpc_*.npz, which the executed reference function produced
(tests/golden/make_golden_gpc.py).
"""
import sys

import numpy as np


def template_cart(n_radial, n_angular, radius):
    """create_template_matrix(in_cart=True) (:111-124): polar grid, then (rho cos, rho sin)."""
    out = np.zeros((n_radial, n_angular, 2))
    for j in range(1, n_radial + 1):
        rho = (j * radius) / n_radial
        for k in range(1, n_angular + 1):
            theta = (2 * k * np.pi) / n_angular
            out[j - 1, k - 1] = [rho * np.cos(theta), rho * np.sin(theta)]
    return out


def barycentric(q, tri):
    v0, v1, v2 = tri[2] - tri[0], tri[1] - tri[0], q - tri[0]
    dot00, dot01, dot02 = v0.dot(v0), v0.dot(v1), v0.dot(v2)
    dot11, dot12 = v1.dot(v1), v1.dot(v2)
    den = dot00 * dot11 - dot01 * dot01
    if den == 0:
        den += sys.float_info.min
    w2 = (dot11 * dot02 - dot01 * dot12) / den
    w1 = (dot00 * dot12 - dot01 * dot02) / den
    w0 = 1 - w2 - w1
    return (w0, w1, w2), (w2 > 0 and w1 > 0 and w2 + w1 <= 1)


def barycentric_coordinates(tri_xy, tri_idx, tri_ptr, n_radial, n_angular, radius):
    tmpl = template_cart(n_radial, n_angular, radius)
    v = len(tri_ptr) - 1
    out = np.zeros((v, n_radial, n_angular, 3, 2))
    for s in range(v):
        for r in range(n_radial):
            for a in range(n_angular):
                for t in range(tri_ptr[s], tri_ptr[s + 1]):
                    w, inside = barycentric(tmpl[r, a], tri_xy[t])
                    if inside:
                        out[s, r, a, :, 0] = tri_idx[t]
                        out[s, r, a, :, 1] = w
                        break
    return out
