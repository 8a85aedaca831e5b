"""CPU oracle for the GPC wavefront (second half): mesh -> geodesic polar coordinate systems.

TEST INFRASTRUCTURE ONLY (same rule as geoconv_oracle.py: imported by tests/, smoke() and bench.py's CPU leg,
never by the product).  Restates the reference's algorithm (Melvaer & Reimers, "Geodesic polar coordinates on
polygonal meshes", with the reference's intersection guard) on plain arrays / sets:

  * wavefront loop ................ /root/reference/src/geoconv/preprocessing/gpc_system_group.py:67-122
  * system initialisation ......... gpc_system.py:36-97
  * edge / face bookkeeping ....... gpc_system.py:110-170 (add_edge, add_face incl. its recursive closing of faces)
  * guarded update ................ gpc_system.py:172-254, segment intersection :304-337
  * candidate coordinates ......... gpc_system_utils.py:150-212 (both faces of an edge, cached faces first)
  * Euclidean update + angle ...... preprocessing/c_extension/c_extension.c:11-145
  * neighbours / faces of an edge . utils/misc.py:46-83

Pinned by tests/golden/wave_*.npz, produced by the EXECUTED reference (its Python sources and its compiled C
extension; tests/golden/make_golden_wavefront.py).  Level-1 BLAS calls of the C code are written out in plain
double arithmetic (OpenBLAS' dnrm2 accumulates in 80-bit registers), so radial coordinates agree with the fixtures
to 1e-12 and angles to 1e-7 (the reference direction's own angle is acos(1 - ulp) = 0 or 2.1e-8), decisions exactly
- except in systems the oracle flags ``fragile`` (a guard verdict on collinear segments = rounding noise).
"""
import heapq
import math
import sys
from fractions import Fraction

import numpy as np

_TWO_PI = 2.0 * math.pi


class MeshTables:
    """What the wavefront reads from a mesh: neighbours (ascending) and, per sorted edge, the third vertex of
    each incident face in ascending face index (misc.py:57-61 returns the faces in that order)."""

    def __init__(self, vertices, faces, normals):
        self.vertices = np.asarray(vertices, np.float64)
        self.faces = np.asarray(faces, np.int64)
        self.normals = np.asarray(normals, np.float64)
        n = self.vertices.shape[0]
        nb = [set() for _ in range(n)]
        self.edge_opposite = {}
        for a, b, c in self.faces.tolist():
            for p, q, r in ((a, b, c), (b, c, a), (c, a, b)):
                nb[p].add(q)
                nb[q].add(p)
                self.edge_opposite.setdefault((min(p, q), max(p, q)), []).append(r)
        self.neighbours = [sorted(s) for s in nb]

    def faces_of_edge(self, a, b):
        key = (min(a, b), max(a, b))
        return [tuple(sorted((key[0], key[1], k))) for k in self.edge_opposite.get(key, [])]


def _norm3(v):
    """cblas_dnrm2 as the OpenBLAS x86-64 kernel evaluates it: squares, sum and square root in 80-bit registers,
    one final rounding to double (measured: identical to scipy.linalg.blas.dnrm2 on 20 000 random vectors)."""
    ld = np.longdouble
    return float(np.sqrt(ld(v[0]) * ld(v[0]) + ld(v[1]) * ld(v[1]) + ld(v[2]) * ld(v[2])))


def _fma(a, b, c):
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))  # exact, then one rounding


def _dot3(a, b):
    """cblas_ddot for n = 3 as OpenBLAS evaluates it: fma(a2, b2, fma(a1, b1, a0 * b0)) (measured: identical to
    scipy.linalg.blas.ddot on 5 000 random vectors; the plain sum of rounded products matches only 66 %)."""
    return _fma(a[2], b[2], _fma(a[1], b[1], float(a[0]) * float(b[0])))


def _angle(a, b):  # c_extension.c:11-25
    x = _dot3(a, b) / (_norm3(a) * _norm3(b))
    return math.acos(min(1.0, max(-1.0, x)))


def angle_360(a, b, axis):  # c_extension.c:28-44
    ang = _angle(a, b)
    cross = (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
    return _TWO_PI - ang if _dot3(axis, cross) < 0.0 else ang


def dist_and_dir(p_i, p_j, p_k, u_j, u_k, th_j, th_k):  # c_extension.c:46-145
    e_j = [p_j[c] - p_i[c] for c in range(3)]
    e_k = [p_k[c] - p_i[c] for c in range(3)]
    e_kj = [p_k[c] - p_j[c] for c in range(3)]
    nj, nk, kj2 = _norm3(e_j), _norm3(e_k), _dot3(e_kj, e_kj)
    area = nj * nk * math.sin(_angle(e_j, e_k))
    radicand = (kj2 - (u_j - u_k) ** 2) * ((u_j + u_k) ** 2 - kj2)

    def along_edge():
        dj, dk = u_j + nj, u_k + nk
        return (dj, th_j) if dj <= dk else (dk, th_k)
    if radicand <= 0:
        return along_edge()
    h = math.sqrt(radicand)
    x_j = area * (kj2 + u_k * u_k - u_j * u_j) + _dot3(e_k, e_kj) * h
    x_k = area * (kj2 + u_j * u_j - u_k * u_k) - _dot3(e_j, e_kj) * h
    if x_j < 0 or x_k < 0:
        return along_edge()
    den = 2 * area * kj2
    x_j, x_k = x_j / den, x_k / den
    res = [x_k * e_k[c] + x_j * e_j[c] for c in range(3)]
    u = _norm3(res)
    s = [res[c] + p_i[c] for c in range(3)]
    q_k, q_j, q_i = ([p[c] - s[c] for c in range(3)] for p in (p_k, p_j, p_i))
    alpha = _angle(q_i, q_j) / _angle(q_k, q_j)
    if th_k <= th_j:
        if th_j - th_k >= math.pi:
            th_k += _TWO_PI
    elif th_k - th_j >= math.pi:
        th_j += _TWO_PI
    return u, math.fmod((1 - alpha) * th_j + alpha * th_k, _TWO_PI)


class GPCState:
    """One GPC system under construction (gpc_system.py:12-97)."""

    def __init__(self, mesh, source):
        self.mesh, self.source = mesh, source
        n = mesh.vertices.shape[0]
        self.rho = np.full(n, np.inf)
        self.theta = np.full(n, -1.0)
        self.x = np.full(n, np.inf)
        self.y = np.full(n, np.inf)
        self.edges, self.edge_set = [], set()        # edges[-1] of the reference, insertion order
        self.faces, self.face_set = [], set()        # faces[(-1, -1)], insertion order
        self.edge_faces = {}                         # faces[(a, b)], insertion order
        self.fragile = False                         # a guard decision hung on rounding noise (see intersects)
        nbs = mesh.neighbours[source]
        p0 = mesh.vertices[source]
        for nb in nbs:
            d = mesh.vertices[nb] - p0
            self.rho[nb] = math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        self.rho[source] = 0.0
        ref = mesh.vertices[nbs[0]] - p0
        for nb in nbs:
            self.theta[nb] = angle_360(ref, mesh.vertices[nb] - p0, mesh.normals[source])
        self.theta[source] = 0.0
        self.x[source] = self.y[source] = 0.0
        for nb in nbs:
            self.x[nb] = self.rho[nb] * np.cos(self.theta[nb])  # numpy's cos / sin like misc.py:142-157: the guard below
            self.y[nb] = self.rho[nb] * np.sin(self.theta[nb])  # can hinge on the last bit (collinear segments)
        for nb in nbs:
            self.add_edge((min(source, nb), max(source, nb)))
            for face in mesh.faces_of_edge(source, nb):
                self.add_face(face)

    def add_edge(self, edge):  # gpc_system.py:110-132
        if edge not in self.edge_set:
            self.edge_set.add(edge)
            self.edges.append(edge)

    def add_face(self, face):  # gpc_system.py:134-170
        if face not in self.face_set:
            self.face_set.add(face)
            self.faces.append(face)
        for edge in ((face[0], face[1]), (face[1], face[2]), (face[0], face[2])):
            known = self.edge_faces.setdefault(edge, [])
            if face not in known:
                known.append(face)
            # the edge's other mesh face: closed as soon as all three of its vertices carry coordinates
            for other in self.mesh.faces_of_edge(*edge):
                if other != face and other not in self.face_set and not any(math.isinf(self.rho[v]) for v in other):
                    self.update(other[0], self.rho[other[0]], self.theta[other[0]], other[1], [other[2]], store=False)

    def intersects(self, p, q):  # gpc_system.py:304-337
        e = np.asarray(self.edges)
        xs3, ys3, xs4, ys4 = self.x[e[:, 0]], self.y[e[:, 0]], self.x[e[:, 1]], self.y[e[:, 1]]
        (x1, y1), (x2, y2) = p, q
        den = (x1 - x2) * (ys3 - ys4) - (y1 - y2) * (xs3 - xs4)
        # Collinear segments (the along-edge fallback hands a vertex its neighbour's angle, so chains of vertices on
        # one ray through the origin are common): den and a numerator are both rounding noise, and the reference's
        # verdict hinges on the last bit of cos / sin / dnrm2.  Such systems are flagged ``fragile``, not compared.
        # Not noise: segments that share a stored end point (p1 = p3, p1 = p4, p2 = p3) - there t or u is exactly
        # 0 or 1 in any implementation that evaluates these expressions in this order without contraction.
        tiny = 1e-9
        c_den = np.abs(den) <= tiny * (abs(x1 - x2) * np.abs(ys3 - ys4) + abs(y1 - y2) * np.abs(xs3 - xs4))
        n1 = (x1 - xs3) * (ys3 - ys4) - (y1 - ys3) * (xs3 - xs4)
        n2 = (x1 - xs3) * (y1 - y2) - (y1 - ys3) * (x1 - x2)
        c_n1 = np.abs(n1) <= tiny * (np.abs((x1 - xs3) * (ys3 - ys4)) + np.abs((y1 - ys3) * (xs3 - xs4)))
        c_n2 = np.abs(n2) <= tiny * (np.abs((x1 - xs3) * (y1 - y2)) + np.abs((y1 - ys3) * (x1 - x2)))
        shared = ((x1 == xs3) & (y1 == ys3)) | ((x1 == xs4) & (y1 == ys4)) | ((x2 == xs3) & (y2 == ys3))
        if np.any(c_den & (c_n1 | c_n2) & ~shared):
            self.fragile = True
        if 0.0 in den:
            den = den + sys.float_info.min
        with np.errstate(all="ignore"):
            t = n1 / den
            u = n2 / den
        eps = 1e-5
        return bool(np.any((eps < t) & (t < 1.0 - eps) & (eps < u) & (u < 1.0 - eps)))

    def update(self, i, rho_i, theta_i, j, k_vertices, store=True):  # gpc_system.py:172-254
        for k in k_vertices:
            if math.isinf(self.rho[k]):
                continue
            face = tuple(sorted((i, j, k)))
            of_interest = [e for e in self.edges if i in e]
            for e in ((face[0], face[1]), (face[1], face[2]), (face[0], face[2])):
                if e not in of_interest:
                    of_interest.append(e)
            x, y = rho_i * np.cos(theta_i), rho_i * np.sin(theta_i)
            for a, b in of_interest:
                pa = (x, y) if a == i else (self.x[a], self.y[a])
                pb = (x, y) if b == i else (self.x[b], self.y[b])
                if self.intersects(pa, pb):
                    return False
            if store:
                self.rho[i], self.theta[i], self.x[i], self.y[i] = rho_i, theta_i, x, y
            for e in of_interest:
                self.add_edge(e)
            self.add_face(face)
        return True

    def candidate(self, i, j):  # gpc_system_utils.py:150-212
        key = (min(i, j), max(i, j))
        faces = self.edge_faces[key] if key in self.edge_faces else self.mesh.faces_of_edge(i, j)
        best, ks = None, []
        for face in faces:
            k = [v for v in face if v != i and v != j][0]
            ks.append(k)
            if self.rho[k] < np.inf and self.theta[k] >= 0.0:
                pv = self.mesh.vertices
                u, th = dist_and_dir(pv[i], pv[j], pv[k], self.rho[j], self.rho[k], self.theta[j], self.theta[k])
                if best is None or (u, th, k) < best:
                    best = (u, th, k)
        if best is None:
            return math.inf, -1.0, None
        return best[0], best[1], ks


def compute_gpc_system(mesh, source, u_max, eps=1e-6):
    """gpc_system_group.py:67-122 for one source vertex; returns the finished GPCState."""
    state = GPCState(mesh, source)
    heap = [(state.rho[nb], nb) for nb in mesh.neighbours[source]]
    heapq.heapify(heap)
    while heap:
        _, j = heapq.heappop(heap)
        for i in mesh.neighbours[j]:
            if i == source:
                continue
            u, th, ks = state.candidate(i, j)
            if u < u_max and state.rho[i] / u > 1 + eps:
                if state.update(i, u, th, j, ks):
                    heapq.heappush(heap, (u, i))
    return state


def compute_all(vertices, faces, normals, u_max, eps=1e-6):
    """Every source vertex; returns (rho (V,V), theta (V,V), tri_idx (n,3), tri_ptr (V+1,), fragile (V,) bool)."""
    mesh = MeshTables(vertices, faces, normals)
    n = mesh.vertices.shape[0]
    rho, theta, tri, ptr, fragile = np.zeros((n, n)), np.zeros((n, n)), [], [0], np.zeros(n, bool)
    for v in range(n):
        s = compute_gpc_system(mesh, v, u_max, eps)
        rho[v], theta[v], fragile[v] = s.rho, s.theta, s.fragile
        tri.extend(s.faces)
        ptr.append(len(tri))
    return rho, theta, np.asarray(tri, np.int64).reshape(-1, 3), np.asarray(ptr, np.int64), fragile


def triangles_cart(rho_row, theta_row, tri_idx):
    """get_gpc_triangles(in_cart=True) (gpc_system.py:347-360, misc.py:142-157): (n, 3, 2) x / y."""
    r, t = rho_row[tri_idx], theta_row[tri_idx]
    return np.stack([r * np.cos(t), r * np.sin(t)], axis=-1)
