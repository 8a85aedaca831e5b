"""Drop-in for ``geoconv.preprocessing.barycentric_coordinates``
(/root/reference/src/geoconv/preprocessing/barycentric_coordinates.py).

* ``create_template_matrix`` (:91-124) - build-time helper of the hot path, numpy.
* ``compute_barycentric_coordinates`` (:127-175) - SURVEY.md §8 f-1, the step right before the hot path:
  the reference's triple Python loop (GPC system x radial x angular, each scanning the system's
  triangles) runs as one CUDA kernel (``gcb_bary_from_gpc``, one warp per template vertex).  CUDA only.
"""
import numpy as np


def create_template_matrix(n_radial, n_angular, radius, in_cart=False):
    """Template vertices K[j, k] = (rho, theta) with rho = (j+1)*radius/n_radial and
    theta = 2*pi*(k+1)/n_angular (so the last angular coordinate is 2*pi == 0).

    Returns a float64 array (n_radial, n_angular, 2); with ``in_cart`` the entries are
    (x, y) = (rho*cos(theta), rho*sin(theta)) instead.
    """
    rho = np.arange(1, n_radial + 1, dtype=np.float64) * radius / n_radial
    theta = 2.0 * np.arange(1, n_angular + 1, dtype=np.float64) * np.pi / n_angular
    grid = np.empty((n_radial, n_angular, 2), dtype=np.float64)
    grid[..., 0] = rho[:, None]
    grid[..., 1] = theta[None, :]
    if in_cart:
        grid = np.stack([grid[..., 0] * np.cos(grid[..., 1]), grid[..., 0] * np.sin(grid[..., 1])], axis=-1)
    return grid


def pack_gpc_triangles(gpc_systems):
    """Flatten a GPC-system group into the ragged arrays the kernel reads.

    ``gpc_systems`` is what the reference passes around (``GPCSystemGroup``): ``.object_mesh_gpc_systems`` is a
    sequence of systems, each with ``get_gpc_triangles(in_cart=True)`` -> (n_tri, 3, 2) float64 and
    ``faces[(-1, -1)]`` -> n_tri index triples (barycentric_coordinates.py:160-168).
    Returns (tri_xy float64 [n, 3, 2], tri_idx int32 [n, 3], tri_ptr int64 [V + 1])."""
    xy, idx, ptr = [], [], [0]
    for system in gpc_systems.object_mesh_gpc_systems:
        tri = np.asarray(system.get_gpc_triangles(in_cart=True), dtype=np.float64).reshape(-1, 3, 2)
        nodes = np.asarray(system.faces[(-1, -1)], dtype=np.int64).reshape(-1, 3)
        if tri.shape[0] != nodes.shape[0]:
            raise ValueError("a GPC system's triangle coordinates and face indices differ in length")
        xy.append(tri)
        idx.append(nodes)
        ptr.append(ptr[-1] + tri.shape[0])
    tri_xy = np.concatenate(xy) if xy else np.zeros((0, 3, 2))
    tri_idx = np.concatenate(idx) if idx else np.zeros((0, 3), dtype=np.int64)
    if tri_idx.size and (tri_idx.min() < 0 or tri_idx.max() >= 2 ** 31):
        raise ValueError("vertex indices must fit int32")
    return (np.ascontiguousarray(tri_xy, dtype=np.float64), np.ascontiguousarray(tri_idx, dtype=np.int32),
            np.asarray(ptr, dtype=np.int64))


def barycentric_from_packed(tri_xy, tri_idx, tri_ptr, n_radial, n_angular, radius, device="cuda"):
    """The kernel call on already packed triangles (torch tensors or numpy arrays); returns a float64 CUDA
    tensor (V, n_radial, n_angular, 3, 2) laid out like the reference's array."""
    import torch
    from geoconv_b200 import _native as N
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("geoconv_b200: barycentric coordinates are computed on a CUDA device (no CPU path)")
    as_t = lambda a, dt: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(dev, dt).contiguous()  # noqa: E731
    tri_xy, tri_idx, tri_ptr = as_t(tri_xy, torch.float64), as_t(tri_idx, torch.int32), as_t(tri_ptr, torch.int64)
    v = tri_ptr.numel() - 1
    tmpl = torch.from_numpy(create_template_matrix(n_radial, n_angular, radius, in_cart=True).reshape(-1, 2)).to(dev)
    out = torch.empty((v, n_radial, n_angular, 3, 2), dtype=torch.float64, device=dev)
    lib = N.load()
    with torch.cuda.device(dev):
        N.check(lib.gcb_bary_from_gpc(tri_xy.data_ptr(), tri_idx.data_ptr(), tri_ptr.data_ptr(), v, tmpl.data_ptr(),
                                      n_radial * n_angular, out.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream), "gcb_bary_from_gpc")
    return out


def compute_barycentric_coordinates(gpc_systems, n_radial=2, n_angular=4, radius=0.05, device="cuda"):
    """Same call and same result layout as the reference (barycentric_coordinates.py:127-175): a float64 numpy
    array B[v, r, a, :, 0] = vertex indices, B[v, r, a, :, 1] = barycentric weights of the first triangle of
    GPC system v that contains template vertex (r, a); zeros if none does."""
    packed = getattr(gpc_systems, "packed", None)   # geoconv_b200's GPCSystemGroup: already packed, on the device
    if packed is None:
        packed = pack_gpc_triangles(gpc_systems)
    return barycentric_from_packed(*packed, n_radial, n_angular, radius, device=device).cpu().numpy()
