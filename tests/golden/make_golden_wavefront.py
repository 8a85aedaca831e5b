"""Golden vectors for the GPC wavefront (second half) from the EXECUTED reference:
/root/reference/src/geoconv/preprocessing/gpc_system_group.py::GPCSystemGroup.compute_gpc_system (:42-122),
gpc_system.py::GPCSystem (:12-360), gpc_system_utils.py (:9-212) and the reference's own compiled
c_extension.c (:46-145), followed by barycentric_coordinates.py::compute_barycentric_coordinates (:127-175).

    oracle/build_ref.sh && python tests/golden/make_golden_wavefront.py

What is real and what is stubbed:
  * the reference's Python and C sources run unmodified, from where they lie; the C extension is compiled by
    oracle/build_ref.sh into oracle/_ref/ against the OpenBLAS that scipy bundles;
  * matplotlib (plotting only), pygeodesic and trimesh (imported at module level by utils/misc.py, never called
    on this path) are absent from the image and replaced by empty modules;
  * the mesh object is oracle/mesh_stub.py::StubMesh, which supplies the six attributes the GPC code reads, with
    trimesh's documented definitions (neighbours in ascending vertex order).
Stored per case: the mesh, u_max, eps, the oracle's per-system ``fragile`` flag, and for every source vertex the radial / angular coordinates of all mesh
vertices (inf / -1 = not reached) and the GPC system's triangle list in insertion order (the order decides which
triangle "contains" a template vertex first); plus the barycentric coordinates the reference derives from them.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, "/root/reference/src")


def install_stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")
    mpl.patches = mod("matplotlib.patches", Polygon=object)
    pg = mod("pygeodesic")
    pg.geodesic = mod("pygeodesic.geodesic")
    mod("trimesh")


def reference_gpc_systems(mesh, u_max, eps=1e-6):
    """Runs the reference wavefront for every source vertex of ``mesh``; returns (rho (V,V), theta (V,V),
    tri_idx (n,3) int64, tri_ptr (V+1,) int64, systems)."""
    from geoconv.preprocessing.gpc_system_group import GPCSystemGroup
    group = GPCSystemGroup(mesh, eps=eps, use_c=True, processes=1)
    n = mesh.vertices.shape[0]
    rho, theta, tri, ptr, systems = np.zeros((n, n)), np.zeros((n, n)), [], [0], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for v in range(n):
            s = group.compute_gpc_system(v, u_max)
            systems.append(s)
            rho[v], theta[v] = s.radial_coordinates, s.angular_coordinates
            faces = np.asarray(s.faces[(-1, -1)], dtype=np.int64).reshape(-1, 3)
            tri.append(faces)
            ptr.append(ptr[-1] + faces.shape[0])
    group.object_mesh_gpc_systems = np.array(systems)
    return rho, theta, np.concatenate(tri), np.asarray(ptr, np.int64), group


def cases():
    from oracle.mesh_stub import blob, bumpy_grid, icosphere, torus
    v, f = icosphere(2, jitter=0.05, seed=3)
    yield "wave_ico2_jitter", v, f, 0.75, 3, 6, 0.5
    v, f = bumpy_grid(11, 9, seed=5)
    yield "wave_grid_boundary", v, f, 0.26, 2, 5, 0.17
    v, f = icosphere(1, jitter=0.08, seed=7)
    yield "wave_ico1_wide", v, f, 1.6, 2, 4, 1.0
    v, f = icosphere(3, jitter=0.1, seed=11)
    yield "wave_ico3_patch60", v, f, 0.62, 5, 8, 0.45
    v, f = torus(16, 9, seed=2)
    yield "wave_torus", v, f, 0.75, 3, 8, 0.5
    v, f = blob(2, seed=4)
    yield "wave_blob", v, f, 0.8, 4, 6, 0.55


if __name__ == "__main__":
    only = set(sys.argv[1:])      # optional: regenerate the named cases only
    install_stub_modules()
    from geoconv.preprocessing.barycentric_coordinates import compute_barycentric_coordinates
    from oracle.mesh_stub import StubMesh
    for name, vertices, faces, u_max, n_radial, n_angular, radius in cases():
        if only and name not in only:
            continue
        mesh = StubMesh(vertices, faces)
        rho, theta, tri_idx, tri_ptr, group = reference_gpc_systems(mesh, u_max)
        bary = compute_barycentric_coordinates(group, n_radial=n_radial, n_angular=n_angular, radius=radius)
        # the oracle's "a guard verdict hung on rounding noise" flag per system (oracle/gpc_oracle.py::intersects),
        # stored so that GPU tests know which systems may legitimately differ without re-running the slow oracle
        from oracle import gpc_oracle
        o_rho, o_theta, o_tri, o_ptr, fragile = gpc_oracle.compute_all(mesh.vertices, mesh.faces, mesh.vertex_normals, u_max)
        assert np.array_equal(o_tri, tri_idx) and np.array_equal(o_ptr, tri_ptr), "oracle and reference disagree"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), vertices=mesh.vertices, faces=mesh.faces,
                            normals=mesh.vertex_normals, u_max=np.array(u_max), eps=np.array(1e-6), rho=rho, theta=theta,
                            tri_idx=tri_idx, tri_ptr=tri_ptr, bary=bary, fragile=fragile, n_radial=np.array(n_radial),
                            n_angular=np.array(n_angular), radius=np.array(radius))
        reached = np.isfinite(rho).sum(1)
        hit = (bary[..., 1].sum(-1) > 0).mean()
        print(f"{name}: V={mesh.vertices.shape[0]} F={mesh.faces.shape[0]} patch vertices {reached.min()}-{reached.max()} "
              f"(mean {reached.mean():.1f}), triangles per system {np.diff(tri_ptr).min()}-{np.diff(tri_ptr).max()}, "
              f"template vertices inside a triangle: {hit:.2f}")
