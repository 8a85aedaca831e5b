"""Times mesh -> GPC systems -> barycentric coordinates on the GPU on jittered icospheres.

    python tools/gpc_bench.py [subdivisions ...]     (default: 4 5 = 2 562 and 10 242 vertices)

Prints one JSON line per mesh: systems/s of gcb_gpc_wavefront (CUDA events, warm), patch statistics, the share of
systems flagged fragile, the time of the barycentric kernel behind it.  The reference needs ~0.09 s per system of
this patch size on one host core (tests/golden/make_golden_wavefront.py: 642 systems in 57 s)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoconv_b200.preprocessing.barycentric_coordinates import barycentric_from_packed  # noqa: E402
from geoconv_b200.preprocessing.gpc_system_group import GPCSystemGroup  # noqa: E402


def icosphere(subdivisions, jitter=0.05, seed=1):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(x, np.float64) / np.linalg.norm(x) for x in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    v, f = np.asarray(v), np.asarray(f, np.int64)
    edge = np.linalg.norm(v[f[:, 0]] - v[f[:, 1]], axis=1).mean()
    v = v + np.random.default_rng(seed).uniform(-1.0, 1.0, v.shape) * jitter * edge
    return v, f, edge


def main():
    subs = [int(a) for a in sys.argv[1:]] or [4, 5]
    for sub in subs:
        class Mesh:
            pass
        Mesh.vertices, Mesh.faces, edge = icosphere(sub)
        radius = 2.2 * edge
        group = GPCSystemGroup(Mesh())
        group.compute(u_max=radius / 0.75)   # warm-up; the reference's ratio kernel radius = 0.75 x GPC radius (mesh tables, allocations)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        group.compute(u_max=radius / 0.75)
        ev[1].record()
        tri_xy, tri_idx, tri_ptr = group.packed
        barycentric_from_packed(tri_xy, tri_idx, tri_ptr, 5, 8, radius)
        ev[2].record()
        bary = barycentric_from_packed(tri_xy, tri_idx, tri_ptr, 5, 8, radius)
        ev[3].record()
        torch.cuda.synchronize()
        n = Mesh.vertices.shape[0]
        nv = torch.diff(group.vertex_tables[3]).float()
        nf = torch.diff(tri_ptr).float()
        ms = ev[0].elapsed_time(ev[1])
        print(json.dumps({"mesh": f"icosphere({sub}) jittered", "vertices": n, "u_max_in_edges": round(2.2 / 0.75, 2),
                          "gpc_ms_all_systems (tables + wavefront + compaction)": round(ms, 3),
                          "systems_per_s": round(n / ms * 1e3), "patch_vertices_mean_max": [round(nv.mean().item(), 1), int(nv.max())],
                          "triangles_mean_max": [round(nf.mean().item(), 1), int(nf.max())],
                          "fragile_share": round((group.flags != 0).float().mean().item(), 3),
                          "bary_ms": round(ev[2].elapsed_time(ev[3]), 3),
                          "template_vertices_inside": round((bary[..., 1].sum(-1) > 0).float().mean().item(), 4)}), flush=True)


if __name__ == "__main__":
    main()
