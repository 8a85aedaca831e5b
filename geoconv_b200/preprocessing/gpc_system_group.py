"""Drop-in for ``geoconv.preprocessing.gpc_system_group`` / ``gpc_system``
(/root/reference/src/geoconv/preprocessing/gpc_system_group.py:13-122, gpc_system.py:12-360): geodesic polar
coordinate systems around every vertex of a triangle mesh - SURVEY.md §8 f-1, the wavefront half.

The reference grows one system per Python process call (a min-heap wavefront with a segment-crossing guard,
~0.1 s per vertex); here ALL systems are grown by one CUDA launch, one warp per source vertex
(``gcb_gpc_wavefront``, csrc/gcb_gpc.cu), and stay on the device in exactly the packed form
``gcb_bary_from_gpc`` reads, so ``compute_barycentric_coordinates(group, ...)`` never leaves the GPU.
CUDA only: there is no CPU path (the mesh tables are plain tensor code and run anywhere, for the tests).

Contract differences, all deliberate:
  * the mesh may be any object with ``vertices`` (V,3) and ``faces`` (F,3) (a ``trimesh.Trimesh`` works; trimesh is
    not required); ``vertex_normals`` is used when present, else corner-angle-weighted face normals are computed;
  * a vertex's neighbours are taken in ascending vertex id (what trimesh's adjacency graph yields for its
    lexicographically sorted unique edges); the first neighbour fixes the system's reference direction;
  * ``use_c`` / ``processes`` are accepted and ignored.
"""
import numpy as np
import torch

GPC_OVERFLOW, GPC_NONMANIFOLD, GPC_FRAGILE = 1, 2, 4


def vertex_normals(vertices, faces):
    """Unit vertex normals: adjacent face normals weighted by the face's corner angle at the vertex."""
    v, f = vertices, faces.long()
    p = v[f]
    fn = torch.linalg.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    fn = fn / fn.norm(dim=1, keepdim=True).clamp_min(1e-300)
    out = torch.zeros_like(v)
    for c in range(3):
        a, b = p[:, (c + 1) % 3] - p[:, c], p[:, (c + 2) % 3] - p[:, c]
        cosang = (a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1)).clamp_min(1e-300)
        out.index_add_(0, f[:, c], fn * torch.acos(cosang.clamp(-1.0, 1.0))[:, None])
    return out / out.norm(dim=1, keepdim=True).clamp_min(1e-300)


def mesh_tables(vertices, faces):
    """Adjacency in the layout ``gcb_gpc_wavefront`` reads (include/geoconv_b200.h): ``nbr_ptr`` (V+1,) int32,
    ``nbr`` (nnz,) int32 ascending per vertex, ``opposite`` (nnz, 2) int32 = third vertices of the edge's first /
    second face in ascending face index (-1 = none).  Replaces what the reference asks trimesh for
    (utils/misc.py:46-83: ``edges_sorted``, ``edges_face``, ``vertex_adjacency_graph``).  Any device."""
    f = faces.long()
    n_v, n_f = vertices.shape[0], f.shape[0]
    if f.numel() and (int(f.min()) < 0 or int(f.max()) >= n_v):
        raise ValueError("face indices outside [0, n_vertices)")
    if n_v >= 2 ** 31 - 1:
        raise ValueError("vertex ids must fit int32")
    # the six directed (vertex, neighbour, opposite) triples of a face, face by face
    perm = torch.tensor([[0, 1, 2], [1, 0, 2], [1, 2, 0], [2, 1, 0], [2, 0, 1], [0, 2, 1]], device=f.device)
    trip = f[:, perm].reshape(-1, 3)
    key = trip[:, 0] * n_v + trip[:, 1]
    order = torch.sort(key, stable=True).indices          # stable: equal keys stay in ascending face index
    key, opp = key[order], trip[order, 2]
    uniq, counts = torch.unique_consecutive(key, return_counts=True)
    if counts.numel() and int(counts.max()) > 2:
        raise NotImplementedError("non-manifold mesh: an edge with more than two faces")
    first = torch.cumsum(counts, 0) - counts
    opposite = torch.full((uniq.numel(), 2), -1, dtype=torch.int32, device=f.device)
    opposite[:, 0] = opp[first].int()
    two = counts == 2
    opposite[two, 1] = opp[first[two] + 1].int()
    owner = torch.div(uniq, n_v, rounding_mode="floor")
    nbr = (uniq - owner * n_v).int()
    nbr_ptr = torch.zeros(n_v + 1, dtype=torch.int64, device=f.device)
    nbr_ptr[1:] = torch.cumsum(torch.bincount(owner, minlength=n_v), 0)
    return nbr_ptr.int().contiguous(), nbr.contiguous(), opposite.contiguous()


class GPCSystem:
    """Host-side view of one finished system with the attributes the reference's consumers read
    (gpc_system.py:339-360, barycentric_coordinates.py:160-168)."""

    def __init__(self, source_point, n_vertices, vid, rho, theta, tri_idx):
        self.source_point = int(source_point)
        self.radial_coordinates = np.full((n_vertices,), np.inf)
        self.angular_coordinates = np.full((n_vertices,), -1.0)
        self.radial_coordinates[vid] = rho
        self.angular_coordinates[vid] = theta
        with np.errstate(invalid="ignore"):
            self.x_coordinates = self.radial_coordinates * np.cos(self.angular_coordinates)
            self.y_coordinates = self.radial_coordinates * np.sin(self.angular_coordinates)
        self.x_coordinates[np.isinf(self.radial_coordinates)] = np.inf
        self.y_coordinates[np.isinf(self.radial_coordinates)] = np.inf
        self.faces = {(-1, -1): [list(map(int, t)) for t in tri_idx]}

    def get_gpc_system(self):
        return np.stack([self.radial_coordinates, self.angular_coordinates], axis=1)

    def get_gpc_triangles(self, in_cart=False):
        tri = self.get_gpc_system()[np.array(self.faces[(-1, -1)], dtype=np.int64).reshape(-1, 3)]
        if in_cart:
            return np.stack([tri[..., 0] * np.cos(tri[..., 1]), tri[..., 0] * np.sin(tri[..., 1])], axis=-1)
        return tri


class GPCSystemGroup:
    def __init__(self, object_mesh, eps=0.000001, use_c=True, processes=1, max_patch=128, device="cuda",
                 batch=32768):
        self.object_mesh = object_mesh
        self.eps = float(eps)
        self.use_c = use_c
        self.processes = processes
        self.max_patch = int(max_patch)
        self.device = torch.device(device)
        self.batch = int(batch)
        self._systems = None
        self.packed = None       # (tri_xy f64 [n,3,2], tri_idx i32 [n,3], tri_ptr i64 [V+1]) on the device
        self.vertex_tables = None  # (vid i32 [m], rho f64 [m], theta f64 [m], ptr i64 [V+1]) on the device
        self.flags = None        # int32 [V]: GPC_FRAGILE bit per system

    # -- the launch -------------------------------------------------------------------------------------------
    def _run(self, tables, sources, u_max, max_patch):
        from geoconv_b200 import _native as N
        pos, nrm, nbr_ptr, nbr, opp = tables
        dev, n = pos.device, sources.numel()
        i32 = dict(dtype=torch.int32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        fc = 3 * max_patch
        nv, nf, flags = torch.empty(n, **i32), torch.empty(n, **i32), torch.empty(n, **i32)
        vid, rho, theta = torch.empty(n, max_patch, **i32), torch.empty(n, max_patch, **f64), torch.empty(n, max_patch, **f64)
        tri, xy = torch.empty(n, fc, 3, **i32), torch.empty(n, fc, 3, 2, **f64)
        lib = N.load()
        with torch.cuda.device(dev):
            N.check(lib.gcb_gpc_wavefront(pos.data_ptr(), nrm.data_ptr(), nbr_ptr.data_ptr(), nbr.data_ptr(),
                                          opp.data_ptr(), pos.shape[0], sources.data_ptr(), n, float(u_max), self.eps,
                                          max_patch, nv.data_ptr(), vid.data_ptr(), rho.data_ptr(), theta.data_ptr(),
                                          nf.data_ptr(), tri.data_ptr(), xy.data_ptr(), flags.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "gcb_gpc_wavefront")
        return nv, vid, rho, theta, nf, tri, xy, flags

    def compute(self, u_max=.04):
        """All GPC systems of the mesh with maximal radius ``u_max`` (gpc_system_group.py:21-40)."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("geoconv_b200: GPC systems are computed on a CUDA device (no CPU path)")
        as_t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a) if not torch.is_tensor(a) else a).to(dev, dt).contiguous()  # noqa: E731
        pos = as_t(self.object_mesh.vertices, torch.float64)
        faces = as_t(self.object_mesh.faces, torch.int64)
        given = getattr(self.object_mesh, "vertex_normals", None)
        nrm = as_t(given, torch.float64) if given is not None else vertex_normals(pos, faces)
        tables = (pos, nrm) + mesh_tables(pos, faces)
        n_v = pos.shape[0]
        parts = []
        for start in range(0, n_v, self.batch):
            sources = torch.arange(start, min(n_v, start + self.batch), dtype=torch.int32, device=dev)
            parts.append(self._compact(tables, sources, u_max))
        cat = lambda k: torch.cat([p[k] for p in parts]) if parts else None  # noqa: E731
        counts_f, counts_v = cat("nf").long(), cat("nv").long()
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        self.packed = (cat("xy"), cat("tri"), torch.cat([zero, torch.cumsum(counts_f, 0)]))
        self.vertex_tables = (cat("vid"), cat("rho"), cat("theta"), torch.cat([zero, torch.cumsum(counts_v, 0)]))
        self.flags = cat("flags")
        self._systems = None
        return self

    def _compact(self, tables, sources, u_max):
        """One batch of sources, compacted to ragged arrays; a batch in which a system outgrew its tables is run
        again with twice the capacity (rare: max_patch is a few times the usual patch size)."""
        patch = self.max_patch
        while True:
            nv, vid, rho, theta, nf, tri, xy, flags = self._run(tables, sources, u_max, patch)
            if bool((flags & GPC_NONMANIFOLD).any()):
                raise NotImplementedError("non-manifold mesh: an edge with more than two faces")
            if not bool((flags & GPC_OVERFLOW).any()):
                break
            if patch >= 1023:
                raise RuntimeError("geoconv_b200: a GPC system exceeds 1023 vertices; lower u_max")
            patch = min(1023, 2 * patch)
        keep_v = torch.arange(patch, device=nv.device)[None, :] < nv[:, None]
        keep_f = torch.arange(3 * patch, device=nv.device)[None, :] < nf[:, None]
        return {"nv": nv, "nf": nf, "flags": flags & GPC_FRAGILE, "vid": vid[keep_v], "rho": rho[keep_v],
                "theta": theta[keep_v], "tri": tri[keep_f], "xy": xy[keep_f]}

    # -- the reference's view of the result ---------------------------------------------------------------------
    @property
    def object_mesh_gpc_systems(self):
        if self._systems is None and self.packed is not None:
            vid, rho, theta, vptr = (t.cpu().numpy() for t in self.vertex_tables)
            tri, tptr = self.packed[1].cpu().numpy(), self.packed[2].cpu().numpy()
            n_v = len(vptr) - 1
            self._systems = np.array([
                GPCSystem(s, n_v, vid[vptr[s]:vptr[s + 1]], rho[vptr[s]:vptr[s + 1]], theta[vptr[s]:vptr[s + 1]],
                          tri[tptr[s]:tptr[s + 1]]) for s in range(n_v)], dtype=object)
        return self._systems

    @object_mesh_gpc_systems.setter
    def object_mesh_gpc_systems(self, value):
        self._systems = value

    def compute_gpc_system(self, source_point, u_max, gpc_system=None, plot_path=""):
        """One system (gpc_system_group.py:42-122); the group call amortises the launch over all vertices."""
        if plot_path:
            raise NotImplementedError("plotting the update steps is not part of the GPU path")
        one = GPCSystemGroup(self.object_mesh, self.eps, max_patch=self.max_patch, device=self.device)
        dev = self.device
        as_t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a) if not torch.is_tensor(a) else a).to(dev, dt).contiguous()  # noqa: E731
        pos, faces = as_t(self.object_mesh.vertices, torch.float64), as_t(self.object_mesh.faces, torch.int64)
        given = getattr(self.object_mesh, "vertex_normals", None)
        nrm = as_t(given, torch.float64) if given is not None else vertex_normals(pos, faces)
        res = one._compact((pos, nrm) + mesh_tables(pos, faces),
                           torch.tensor([int(source_point)], dtype=torch.int32, device=dev), u_max)
        return GPCSystem(source_point, pos.shape[0], res["vid"].cpu().numpy(), res["rho"].cpu().numpy(),
                         res["theta"].cpu().numpy(), res["tri"].cpu().numpy())
