"""Golden vectors for SURVEY.md §8 f-1 from the EXECUTED reference function
/root/reference/src/geoconv/preprocessing/barycentric_coordinates.py::compute_barycentric_coordinates.

    python tests/golden/make_golden_gpc.py

The reference's GPCSystemGroup cannot be imported here (matplotlib / compiled extension), and the function only
touches `.object_mesh_gpc_systems[i].get_gpc_triangles(in_cart=True)` and `.faces[(-1, -1)]`
(barycentric_coordinates.py:160-168), so the fixtures feed it stand-in objects with exactly that surface,
filled with seeded synthetic GPC systems (tests/helpers.py::synthetic_gpc_systems).
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from geoconv.preprocessing.barycentric_coordinates import compute_barycentric_coordinates  # noqa: E402

from tests.helpers import synthetic_gpc_systems  # noqa: E402

CASES = [("gpc_small", 24, 2, 4, 0.05, 11), ("gpc_r5a8", 40, 5, 8, 0.028, 12), ("gpc_r3a6_sparse", 30, 3, 6, 0.08, 13)]

if __name__ == "__main__":
    for name, v, r, a, radius, seed in CASES:
        group = synthetic_gpc_systems(v, radius, seed)
        bary = compute_barycentric_coordinates(group, n_radial=r, n_angular=a, radius=radius)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), bary=bary, v=np.array(v), r=np.array(r), a=np.array(a),
                            radius=np.array(radius), seed=np.array(seed))
        hit = (bary[..., 1].sum(-1) > 0).mean()
        print(f"{name}: {bary.shape}, template vertices inside a triangle: {hit:.2f}")
