// Mesh -> geodesic polar coordinate (GPC) systems on the GPU (the wavefront half): the step
// before gcb_bary_from_gpc.  Replaces, for ALL source vertices at once,
//   GPCSystemGroup.compute_gpc_system ... /root/reference/src/geoconv/preprocessing/gpc_system_group.py:42-122
//   GPCSystem (init, add_edge, add_face, update, line_segment_intersection) ... gpc_system.py:12-337
//   compute_distance_and_angle ......... gpc_system_utils.py:150-212
//   compute_dist_and_dir / compute_angle(_360) ... preprocessing/c_extension/c_extension.c:11-145
// (Melvaer & Reimers, "Geodesic polar coordinates on polygonal meshes", plus the reference's guard that rejects an
// update whose edges would cross edges already in the system).
//
// One WARP per source vertex; the reference runs one Python process per vertex (~0.1 s each).  The algorithm is
// a sequential min-heap wavefront, so the scalar control flow runs uniformly in all 32 lanes (same registers,
// broadcast shared-memory reads; lane 0 writes), and the lanes split what is wide:
//   * the guard: every candidate edge against every edge of the system (float64 segment test, 32 edges per step),
//   * look-ups in the per-system tables (vertex / edge / face by key: compare + ballot),
//   * the candidate list (arg-min by (distance, vertex id) = the pop order of Python's heapq on those tuples).
// All per-system state lives in shared memory: vertex table (id, rho, theta, x, y), edge table (end points, "in the
// edge list" flag, the <= 2 faces cached under the edge IN INSERTION ORDER - the reference prefers cached faces
// over the mesh's, gpc_system_utils.py:181-186), face list in insertion order (it decides which triangle
// "contains" a template vertex first), candidate list, and an explicit stack for add_face's recursion
// (gpc_system.py:155-170: a face whose three vertices all carry coordinates is closed immediately, depth-first).
// Arithmetic follows the reference operation by operation in float64 without contraction (__dmul_rn / __dadd_rn),
// including what its BLAS does: ddot(n=3) = fma(a2,b2, fma(a1,b1, a0*b0)) and dnrm2 = square root of the sum of
// squares accumulated in extended precision (here: double-double, correctly rounded).  What cannot be matched
// bit for bit is libm / numpy cos, sin, acos; the guard's verdict on COLLINEAR segments is rounding noise in the
// reference too, so systems where such a verdict was taken are flagged (GCB_GPC_FRAGILE) - DESIGN.md, f-1.
#include <float.h>
#include <math.h>

#include "gcb_common.cuh"

#ifndef GCB_WARP
#define GCB_WARP 32   // lanes that share one system (the CPU test suite compiles the text below with 1, tests/emul/)
#endif

namespace gcb {

#define FULL 0xffffffffu
constexpr double kTwoPi = 6.283185307179586476925286766559;   // 2 * M_PI
constexpr double kPi = 3.14159265358979323846;

#ifndef GCB_GPC_HEAP_MULT
#define GCB_GPC_HEAP_MULT 2   // candidate-list capacity in units of the vertex capacity (measured peak: below 1)
#endif

__host__ __device__ inline int64_t gpc_smem_per_system(int P) {
  const int64_t EC = 4 * (int64_t)P, FC = 3 * (int64_t)P, HC = GCB_GPC_HEAP_MULT * (int64_t)P, SC = 3 * (int64_t)P;
  const int64_t b = 8 * (4 * P + HC) + 4 * (P + HC) + 4 * (EC + FC + SC) + 2 * 2 * EC + EC + EC;
  return (b + 15) / 16 * 16;
}

struct Wave {
  int P, EC, FC, HC, SC, lane;
  // mesh tables
  const double* pos; const double* nrm; const int32_t* nptr; const int32_t* nbr; const int32_t* opp;
  // vertex table
  int32_t* vid; double* rho; double* th; double* x; double* y; int nv;
  // edge table: key = slot_a << 16 | slot_b with vid[slot_a] < vid[slot_b]
  uint32_t* ekey; uint8_t* e_in; uint8_t* e_nf; uint16_t* e_f; int ne;
  // faces: key = slot0 << 20 | slot1 << 10 | slot2, slots ordered by vertex id
  uint32_t* fkey; int nf;
  // candidates
  double* hd; int32_t* hv; int nh;
  // add_face stack: face << 4 | edge << 2 | phase
  uint32_t* stk; int ns;
  int flags;
};

__device__ __forceinline__ void wsync() { __syncwarp(); }

// ---------------------------------------------------------------------------------------------- arithmetic
__device__ __forceinline__ double dot3(const double* a, const double* b) {   // OpenBLAS ddot, n = 3
  return __fma_rn(a[2], b[2], __fma_rn(a[1], b[1], __dmul_rn(a[0], b[0])));
}
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  s = __dadd_rn(a, b);
  const double bb = __dsub_rn(s, a);
  e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
__device__ __forceinline__ double norm3(const double* v) {   // OpenBLAS dnrm2 (x87 accumulation), n = 3
  double p0 = __dmul_rn(v[0], v[0]), p1 = __dmul_rn(v[1], v[1]), p2 = __dmul_rn(v[2], v[2]);
  const double e0 = __fma_rn(v[0], v[0], -p0), e1 = __fma_rn(v[1], v[1], -p1), e2 = __fma_rn(v[2], v[2], -p2);
  double s, t, s2, t2;
  two_sum(p0, p1, s, t);
  two_sum(s, p2, s2, t2);
  const double lo = __dadd_rn(__dadd_rn(t, t2), __dadd_rn(__dadd_rn(e0, e1), e2));
  double hi, l2;
  two_sum(s2, lo, hi, l2);
  if (!(hi > 0.0)) return 0.0;
  const double r = sqrt(hi);
  const double res = __dadd_rn(__fma_rn(-r, r, hi), l2);   // hi + l2 - r*r
  return __dadd_rn(r, __ddiv_rn(res, __dmul_rn(2.0, r)));
}
__device__ __forceinline__ double angle3(const double* a, const double* b) {   // c_extension.c:11-25
  double c = __ddiv_rn(dot3(a, b), __dmul_rn(norm3(a), norm3(b)));
  if (c > 1.0) c = 1.0;
  if (c < -1.0) c = -1.0;
  return acos(c);
}
__device__ double angle360(const double* a, const double* b, const double* axis) {   // c_extension.c:28-44
  const double ang = angle3(a, b);
  double cr[3];
  cr[0] = __dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1]));
  cr[1] = __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2]));
  cr[2] = __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]));
  return dot3(axis, cr) < 0.0 ? __dsub_rn(kTwoPi, ang) : ang;
}

// c_extension.c:46-145
__device__ void dist_and_dir(const double* pi, const double* pj, const double* pk, double u_j, double u_k,
                             double th_j, double th_k, double& u_out, double& th_out) {
  double ej[3], ek[3], ekj[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ej[c] = __dsub_rn(pj[c], pi[c]);
    ek[c] = __dsub_rn(pk[c], pi[c]);
    ekj[c] = __dsub_rn(pk[c], pj[c]);
  }
  const double nj = norm3(ej), nk = norm3(ek), kj2 = dot3(ekj, ekj);
  const double area = __dmul_rn(__dmul_rn(nj, nk), sin(angle3(ej, ek)));
  const double dm = __dsub_rn(u_j, u_k), dp = __dadd_rn(u_j, u_k);
  const double radicand = __dmul_rn(__dsub_rn(kj2, __dmul_rn(dm, dm)), __dsub_rn(__dmul_rn(dp, dp), kj2));
  bool along = radicand <= 0.0;
  double xj = 0.0, xk = 0.0;
  if (!along) {
    const double h = sqrt(radicand);
    const double uj2 = __dmul_rn(u_j, u_j), uk2 = __dmul_rn(u_k, u_k);
    xj = __dadd_rn(__dmul_rn(area, __dsub_rn(__dadd_rn(kj2, uk2), uj2)), __dmul_rn(dot3(ek, ekj), h));
    xk = __dsub_rn(__dmul_rn(area, __dsub_rn(__dadd_rn(kj2, uj2), uk2)), __dmul_rn(dot3(ej, ekj), h));
    along = xj < 0.0 || xk < 0.0;
  }
  if (along) {
    const double dj = __dadd_rn(u_j, nj), dk = __dadd_rn(u_k, nk);
    if (dj <= dk) { u_out = dj; th_out = th_j; } else { u_out = dk; th_out = th_k; }
    return;
  }
  const double den = __dmul_rn(__dmul_rn(2.0, area), kj2);
  xj = __ddiv_rn(xj, den);
  xk = __ddiv_rn(xk, den);
  double res[3], s[3], qk[3], qj[3], qi[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    res[c] = __dadd_rn(__dmul_rn(ek[c], xk), __dmul_rn(ej[c], xj));
    s[c] = __dadd_rn(res[c], pi[c]);
    qk[c] = __dsub_rn(pk[c], s[c]);
    qj[c] = __dsub_rn(pj[c], s[c]);
    qi[c] = __dsub_rn(pi[c], s[c]);
  }
  u_out = norm3(res);
  const double alpha = __ddiv_rn(angle3(qi, qj), angle3(qk, qj));
  if (th_k <= th_j) {
    if (__dsub_rn(th_j, th_k) >= kPi) th_k = __dadd_rn(th_k, kTwoPi);
  } else {
    if (__dsub_rn(th_k, th_j) >= kPi) th_j = __dadd_rn(th_j, kTwoPi);
  }
  th_out = fmod(__dadd_rn(__dmul_rn(__dsub_rn(1.0, alpha), th_j), __dmul_rn(alpha, th_k)), kTwoPi);
}

// ---------------------------------------------------------------------------------------------- tables
__device__ int find_vertex(const Wave& w, int gid) {
  for (int base = 0; base < w.nv; base += GCB_WARP) {
    const int s = base + w.lane;
    const unsigned b = __ballot_sync(FULL, s < w.nv && w.vid[s] == gid);
    if (b) return base + __ffs((int)b) - 1;
  }
  return -1;
}
__device__ int get_vertex(Wave& w, int gid) {   // the reference's arrays hold inf / -1 for every untouched vertex
  int s = find_vertex(w, gid);
  if (s >= 0) return s;
  if (w.nv >= w.P) { w.flags |= GCB_GPC_OVERFLOW; return -1; }
  s = w.nv++;
  if (w.lane == 0) {
    w.vid[s] = gid;
    w.rho[s] = INFINITY; w.th[s] = -1.0; w.x[s] = INFINITY; w.y[s] = INFINITY;
  }
  wsync();
  return s;
}
__device__ __forceinline__ uint32_t edge_key(const Wave& w, int sa, int sb) {
  return w.vid[sa] < w.vid[sb] ? ((uint32_t)sa << 16 | (uint32_t)sb) : ((uint32_t)sb << 16 | (uint32_t)sa);
}
__device__ int find_edge(const Wave& w, uint32_t key) {
  for (int base = 0; base < w.ne; base += GCB_WARP) {
    const int e = base + w.lane;
    const unsigned b = __ballot_sync(FULL, e < w.ne && w.ekey[e] == key);
    if (b) return base + __ffs((int)b) - 1;
  }
  return -1;
}
__device__ int get_edge(Wave& w, uint32_t key) {
  int e = find_edge(w, key);
  if (e >= 0) return e;
  if (w.ne >= w.EC) { w.flags |= GCB_GPC_OVERFLOW; return -1; }
  e = w.ne++;
  if (w.lane == 0) { w.ekey[e] = key; w.e_in[e] = 0; w.e_nf[e] = 0; }
  wsync();
  return e;
}
__device__ void add_edge(Wave& w, uint32_t key) {   // gpc_system.py:110-132
  const int e = get_edge(w, key);
  if (e < 0) return;
  if (w.lane == 0) w.e_in[e] = 1;
  wsync();
}
__device__ void sort3(const Wave& w, int& a, int& b, int& c) {   // slots by vertex id
  int t;
  if (w.vid[a] > w.vid[b]) { t = a; a = b; b = t; }
  if (w.vid[b] > w.vid[c]) { t = b; b = c; c = t; }
  if (w.vid[a] > w.vid[b]) { t = a; a = b; b = t; }
}
__device__ __forceinline__ uint32_t face_key(int a, int b, int c) { return (uint32_t)a << 20 | (uint32_t)b << 10 | (uint32_t)c; }
__device__ int find_face(const Wave& w, uint32_t key) {
  for (int base = 0; base < w.nf; base += GCB_WARP) {
    const int f = base + w.lane;
    const unsigned b = __ballot_sync(FULL, f < w.nf && w.fkey[f] == key);
    if (b) return base + __ffs((int)b) - 1;
  }
  return -1;
}
// the <= 2 mesh faces of edge {ga, gb} as their third vertices, in ascending face index (misc.py:46-64)
__device__ void mesh_faces_of_edge(const Wave& w, int ga, int gb, int& k0, int& k1) {
  k0 = k1 = -1;
  const int beg = w.nptr[ga], end = w.nptr[ga + 1];
  for (int base = beg; base < end; base += GCB_WARP) {
    const int p = base + w.lane;
    const unsigned b = __ballot_sync(FULL, p < end && w.nbr[p] == gb);
    if (b) {
      const int q = base + __ffs((int)b) - 1;
      k0 = w.opp[2 * q];
      k1 = w.opp[2 * q + 1];
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------- the guard
// gpc_system.py:304-337: does segment p1-p2 properly cross any edge of the system?  32 edges per step.
__device__ bool intersects(Wave& w, double x1, double y1, double x2, double y2) {
  const double dx12 = __dsub_rn(x1, x2), dy12 = __dsub_rn(y1, y2);
  const double lo = 1e-5, hi = __dsub_rn(1.0, 1e-5);
  bool any_zero = false, hit_plain = false, hit_shift = false, noise = false;
  for (int base = 0; base < w.ne; base += GCB_WARP) {
    const int e = base + w.lane;
    bool zero = false, hp = false, hs = false, nz = false;
    if (e < w.ne && w.e_in[e]) {
      const uint32_t key = w.ekey[e];
      const int sa = key >> 16, sb = key & 0xffffu;
      const double x3 = w.x[sa], y3 = w.y[sa], x4 = w.x[sb], y4 = w.y[sb];
      const double dx34 = __dsub_rn(x3, x4), dy34 = __dsub_rn(y3, y4);
      const double dx13 = __dsub_rn(x1, x3), dy13 = __dsub_rn(y1, y3);
      const double a0 = __dmul_rn(dx12, dy34), a1 = __dmul_rn(dy12, dx34);
      const double b0 = __dmul_rn(dx13, dy34), b1 = __dmul_rn(dy13, dx34);
      const double c0 = __dmul_rn(dx13, dy12), c1 = __dmul_rn(dy13, dx12);
      const double den = __dsub_rn(a0, a1), n1 = __dsub_rn(b0, b1), n2 = __dsub_rn(c0, c1);
      zero = den == 0.0;
      const double t = __ddiv_rn(n1, den), u = __ddiv_rn(n2, den);
      hp = lo < t && t < hi && lo < u && u < hi;
      const double den2 = __dadd_rn(den, DBL_MIN);   // "if 0. in denominators: denominators += float_info.min"
      const double t2 = __ddiv_rn(n1, den2), u2 = __ddiv_rn(n2, den2);
      hs = lo < t2 && t2 < hi && lo < u2 && u2 < hi;
      // collinear segments: den and a numerator are rounding noise
      const double tiny = 1e-9;
      const bool c_den = fabs(den) <= __dmul_rn(tiny, __dadd_rn(fabs(a0), fabs(a1)));
      const bool c_n1 = fabs(n1) <= __dmul_rn(tiny, __dadd_rn(fabs(b0), fabs(b1)));
      const bool c_n2 = fabs(n2) <= __dmul_rn(tiny, __dadd_rn(fabs(c0), fabs(c1)));
      const bool shared = (x1 == x3 && y1 == y3) || (x1 == x4 && y1 == y4) || (x2 == x3 && y2 == y3);
      nz = c_den && (c_n1 || c_n2) && !shared;
    }
    any_zero |= __any_sync(FULL, zero);
    hit_plain |= __any_sync(FULL, hp);
    hit_shift |= __any_sync(FULL, hs);
    noise |= __any_sync(FULL, nz);
  }
  if (noise) w.flags |= GCB_GPC_FRAGILE;
  return any_zero ? hit_shift : hit_plain;
}

// every edge of the system that contains slot si, then the three edges of face (f0, f1, f2); (x, y) stands in
// for si's coordinates (gpc_system.py:199-235)
__device__ bool any_crossing(Wave& w, int si, double x, double y, int f0, int f1, int f2) {
  for (int base = 0; base < w.ne; base += GCB_WARP) {
    const int e = base + w.lane;
    bool m = false;
    if (e < w.ne && w.e_in[e]) {
      const uint32_t key = w.ekey[e];
      m = (int)(key >> 16) == si || (int)(key & 0xffffu) == si;
    }
    unsigned bits = __ballot_sync(FULL, m);
    while (bits) {
      const int ee = base + __ffs((int)bits) - 1;
      bits &= bits - 1;
      const uint32_t key = w.ekey[ee];
      const int sa = key >> 16, sb = key & 0xffffu;
      const double ax = sa == si ? x : w.x[sa], ay = sa == si ? y : w.y[sa];
      const double bx = sb == si ? x : w.x[sb], by = sb == si ? y : w.y[sb];
      if (intersects(w, ax, ay, bx, by)) return true;
    }
  }
  const int fe[3][2] = {{f0, f1}, {f1, f2}, {f0, f2}};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int sa = fe[q][0], sb = fe[q][1];
    const uint32_t key = (uint32_t)sa << 16 | (uint32_t)sb;
    const int known = find_edge(w, key);
    if (known >= 0 && w.e_in[known] && (sa == si || sb == si)) continue;   // tested above
    const double ax = sa == si ? x : w.x[sa], ay = sa == si ? y : w.y[sa];
    const double bx = sb == si ? x : w.x[sb], by = sb == si ? y : w.y[sb];
    if (intersects(w, ax, ay, bx, by)) return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------- add_face
// gpc_system.py:134-170 with its recursion (add_face -> update(update=False) -> add_face ...) on an explicit
// stack.  Frame phases per edge: 0 = register the face under the edge, 1 / 2 = try to close the edge's first /
// second mesh face, 3 = next edge.
__device__ void add_face(Wave& w, int f0, int f1, int f2) {
  const uint32_t fk = face_key(f0, f1, f2);
  int fidx = find_face(w, fk);
  if (fidx < 0) {
    if (w.nf >= w.FC) { w.flags |= GCB_GPC_OVERFLOW; return; }
    fidx = w.nf++;
    if (w.lane == 0) w.fkey[fidx] = fk;
    wsync();
  }
  const int floor_ns = w.ns;
  if (w.ns >= w.SC) { w.flags |= GCB_GPC_OVERFLOW; return; }
  if (w.lane == 0) w.stk[w.ns] = (uint32_t)fidx << 4;
  w.ns++;
  wsync();
  while (w.ns > floor_ns && !(w.flags & GCB_GPC_OVERFLOW)) {
    const uint32_t fr = w.stk[w.ns - 1];
    const int f = fr >> 4, ei = (fr >> 2) & 3, ph = fr & 3;
    const uint32_t key = w.fkey[f];
    const int a = key >> 20, b = (key >> 10) & 1023, c = key & 1023;
    const int sa = ei == 0 ? a : (ei == 1 ? b : a), sb = ei == 0 ? b : c;
    if (ph == 3) {
      wsync();
      if (ei == 2) { w.ns--; }
      else if (w.lane == 0) w.stk[w.ns - 1] = (uint32_t)f << 4 | (uint32_t)(ei + 1) << 2;
      wsync();
      continue;
    }
    wsync();
    if (w.lane == 0) w.stk[w.ns - 1] = (uint32_t)f << 4 | (uint32_t)ei << 2 | (uint32_t)(ph + 1);
    wsync();
    if (ph == 0) {
      const int e = get_edge(w, (uint32_t)sa << 16 | (uint32_t)sb);
      if (e < 0) return;
      const int n = w.e_nf[e];
      const bool have = (n > 0 && w.e_f[2 * e] == f) || (n > 1 && w.e_f[2 * e + 1] == f);
      if (!have) {
        if (n >= 2) { w.flags |= GCB_GPC_NONMANIFOLD; return; }
        if (w.lane == 0) { w.e_f[2 * e + n] = (uint16_t)f; w.e_nf[e] = (uint8_t)(n + 1); }
        wsync();
      }
      continue;
    }
    int k0, k1;
    mesh_faces_of_edge(w, w.vid[sa], w.vid[sb], k0, k1);
    const int gk = ph == 1 ? k0 : k1;
    if (gk < 0) continue;
    const int sk = find_vertex(w, gk);
    if (sk < 0 || isinf(w.rho[sk])) continue;
    int o0 = sa, o1 = sb, o2 = sk;
    sort3(w, o0, o1, o2);
    const uint32_t ok = face_key(o0, o1, o2);
    if (ok == key || find_face(w, ok) >= 0) continue;
    // update(o0, rho[o0], theta[o0], o1, [o2], update=False): the stored coordinates ARE rho * cos / sin (theta)
    if (any_crossing(w, o0, w.x[o0], w.y[o0], o0, o1, o2)) continue;
    add_edge(w, (uint32_t)o0 << 16 | (uint32_t)o1);
    add_edge(w, (uint32_t)o1 << 16 | (uint32_t)o2);
    add_edge(w, (uint32_t)o0 << 16 | (uint32_t)o2);
    if (w.nf >= w.FC || w.ns >= w.SC) { w.flags |= GCB_GPC_OVERFLOW; return; }
    const int nfi = w.nf++;
    if (w.lane == 0) { w.fkey[nfi] = ok; w.stk[w.ns] = (uint32_t)nfi << 4; }
    w.ns++;
    wsync();
  }
}

// gpc_system.py:172-254 with update=True
__device__ bool update(Wave& w, int gi, double rho_i, double th_i, int gj, int gk0, int gk1) {
  const int ks[2] = {gk0, gk1};
  for (int q = 0; q < 2; ++q) {
    if (ks[q] < 0) continue;
    const int sk = find_vertex(w, ks[q]);
    if (sk < 0 || isinf(w.rho[sk])) continue;
    const int si = get_vertex(w, gi), sj = find_vertex(w, gj);
    if (si < 0 || sj < 0) return false;
    int f0 = si, f1 = sj, f2 = sk;
    sort3(w, f0, f1, f2);
    const double x = __dmul_rn(rho_i, cos(th_i)), y = __dmul_rn(rho_i, sin(th_i));
    if (any_crossing(w, si, x, y, f0, f1, f2)) return false;
    wsync();
    if (w.lane == 0) { w.rho[si] = rho_i; w.th[si] = th_i; w.x[si] = x; w.y[si] = y; }
    wsync();
    add_edge(w, (uint32_t)f0 << 16 | (uint32_t)f1);
    add_edge(w, (uint32_t)f1 << 16 | (uint32_t)f2);
    add_edge(w, (uint32_t)f0 << 16 | (uint32_t)f2);
    add_face(w, f0, f1, f2);
    if (w.flags & (GCB_GPC_OVERFLOW | GCB_GPC_NONMANIFOLD)) return false;
  }
  return true;
}

// gpc_system_utils.py:150-212: coordinates vertex gi would get through edge {gi, gj}
__device__ bool candidate(Wave& w, int gi, int gj, double& u, double& th, int& k0, int& k1) {
  k0 = k1 = -1;
  const int si = find_vertex(w, gi), sj = find_vertex(w, gj);
  bool cached = false;
  if (si >= 0 && sj >= 0) {
    const int e = find_edge(w, edge_key(w, si, sj));
    if (e >= 0 && w.e_nf[e] > 0) {
      cached = true;
      for (int q = 0; q < w.e_nf[e]; ++q) {
        const uint32_t key = w.fkey[w.e_f[2 * e + q]];
        const int a = key >> 20, b = (key >> 10) & 1023, c = key & 1023;
        const int third = (a != si && a != sj) ? a : ((b != si && b != sj) ? b : c);
        (q == 0 ? k0 : k1) = w.vid[third];
      }
    }
  }
  if (!cached) mesh_faces_of_edge(w, gi, gj, k0, k1);
  if (k0 < 0 && k1 >= 0) { k0 = k1; k1 = -1; }
  bool found = false;
  int best_k = 0;
  const int ks[2] = {k0, k1};
  for (int q = 0; q < 2; ++q) {
    if (ks[q] < 0) continue;
    const int sk = find_vertex(w, ks[q]);
    if (sk < 0 || !(w.rho[sk] < INFINITY) || !(w.th[sk] >= 0.0)) continue;
    double uu, tt;
    dist_and_dir(w.pos + 3 * (size_t)gi, w.pos + 3 * (size_t)gj, w.pos + 3 * (size_t)ks[q], w.rho[sj], w.rho[sk],
                 w.th[sj], w.th[sk], uu, tt);
    // min() over (u, theta, k) tuples, the first of equal ones
    if (!found || uu < u || (uu == u && (tt < th || (tt == th && ks[q] < best_k)))) {
      u = uu; th = tt; best_k = ks[q]; found = true;
    }
  }
  return found;
}

__device__ void init_system(Wave& w, int src) {   // gpc_system.py:36-97
  const int beg = w.nptr[src], end = w.nptr[src + 1];
  const int val = end - beg;
  if (val + 1 > w.P) { w.flags |= GCB_GPC_OVERFLOW; return; }
  const double* p0 = w.pos + 3 * (size_t)src;
  double ref[3] = {0.0, 0.0, 0.0};
  if (val > 0) {
    const double* pr = w.pos + 3 * (size_t)w.nbr[beg];
    for (int c = 0; c < 3; ++c) ref[c] = __dsub_rn(pr[c], p0[c]);
  }
  if (w.lane == 0) { w.vid[0] = src; w.rho[0] = 0.0; w.th[0] = 0.0; w.x[0] = 0.0; w.y[0] = 0.0; }
  for (int n = w.lane; n < val; n += GCB_WARP) {
    const int g = w.nbr[beg + n];
    const double* p = w.pos + 3 * (size_t)g;
    double d[3];
    for (int c = 0; c < 3; ++c) d[c] = __dsub_rn(p[c], p0[c]);
    // np.linalg.norm(axis=-1): sqrt((d0^2 + d1^2) + d2^2)
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
    const double t = angle360(ref, d, w.nrm + 3 * (size_t)src);
    w.vid[1 + n] = g; w.rho[1 + n] = r; w.th[1 + n] = t;
    w.x[1 + n] = __dmul_rn(r, cos(t)); w.y[1 + n] = __dmul_rn(r, sin(t));
  }
  w.nv = val + 1;
  wsync();
  for (int n = 0; n < val && !(w.flags & (GCB_GPC_OVERFLOW | GCB_GPC_NONMANIFOLD)); ++n) {
    const int sn = 1 + n;
    add_edge(w, edge_key(w, 0, sn));
    int k0, k1;
    mesh_faces_of_edge(w, src, w.vid[sn], k0, k1);
    const int ks[2] = {k0, k1};
    for (int q = 0; q < 2; ++q) {
      if (ks[q] < 0) continue;
      int f0 = 0, f1 = sn, f2 = find_vertex(w, ks[q]);
      if (f2 < 0) continue;
      sort3(w, f0, f1, f2);
      add_face(w, f0, f1, f2);
    }
  }
}

__global__ void k_gpc_wavefront(const double* __restrict__ pos, const double* __restrict__ nrm,
                                const int32_t* __restrict__ nptr, const int32_t* __restrict__ nbr,
                                const int32_t* __restrict__ opp, const int32_t* __restrict__ sources, int n_sources,
                                double u_max, double eps, int P, int32_t* __restrict__ out_nv,
                                int32_t* __restrict__ out_vid, double* __restrict__ out_rho,
                                double* __restrict__ out_theta, int32_t* __restrict__ out_nf,
                                int32_t* __restrict__ out_tri, double* __restrict__ out_xy,
                                int32_t* __restrict__ out_flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warps_per_block = blockDim.x / GCB_WARP, wib = threadIdx.x / GCB_WARP;
  const int sys = blockIdx.x * warps_per_block + wib;
  if (sys >= n_sources) return;
  Wave w;
  w.P = P; w.EC = 4 * P; w.FC = 3 * P; w.HC = GCB_GPC_HEAP_MULT * P; w.SC = 3 * P; w.lane = threadIdx.x % GCB_WARP;
  w.pos = pos; w.nrm = nrm; w.nptr = nptr; w.nbr = nbr; w.opp = opp;
  const size_t per_warp = (size_t)gpc_smem_per_system(P);
  unsigned char* base = smem + per_warp * wib;
  w.rho = (double*)base; w.th = w.rho + P; w.x = w.th + P; w.y = w.x + P;
  w.hd = w.y + P;
  w.vid = (int32_t*)(w.hd + w.HC);
  w.hv = w.vid + P;
  w.ekey = (uint32_t*)(w.hv + w.HC);
  w.fkey = w.ekey + w.EC;
  w.stk = w.fkey + w.FC;
  w.e_f = (uint16_t*)(w.stk + w.SC);
  w.e_in = (uint8_t*)(w.e_f + 2 * w.EC);
  w.e_nf = w.e_in + w.EC;
  w.nv = w.ne = w.nf = w.nh = w.ns = 0;
  w.flags = 0;
  const int src = sources ? sources[sys] : sys;

  init_system(w, src);
  // min-"heap" over (radial coordinate, vertex id): gpc_system_group.py:85-91
  {
    const int beg = nptr[src], val = nptr[src + 1] - beg;
    if (val > w.HC) w.flags |= GCB_GPC_OVERFLOW;
    else {
      for (int n = w.lane; n < val; n += GCB_WARP) { w.hd[n] = w.rho[1 + n]; w.hv[n] = w.vid[1 + n]; }
      w.nh = val;
    }
    wsync();
  }
  int pops = 0;   // every pop follows a strict improvement of some vertex; the cap only bounds a runaway
  while (w.nh > 0 && !(w.flags & (GCB_GPC_OVERFLOW | GCB_GPC_NONMANIFOLD))) {
    if (++pops > 64 * P) { w.flags |= GCB_GPC_OVERFLOW; break; }
    // pop the smallest (distance, vertex) tuple
    double bd = INFINITY; int bv = 0x7fffffff, bi = -1;
    for (int h = w.lane; h < w.nh; h += GCB_WARP) {
      const double d = w.hd[h]; const int v = w.hv[h];
      if (bi < 0 || d < bd || (d == bd && v < bv)) { bd = d; bv = v; bi = h; }
    }
#pragma unroll
    for (int o = GCB_WARP / 2; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(FULL, bd, o); const int ov = __shfl_xor_sync(FULL, bv, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (oi >= 0 && (bi < 0 || od < bd || (od == bd && (ov < bv || (ov == bv && oi < bi))))) { bd = od; bv = ov; bi = oi; }
    }
    wsync();
    if (w.lane == 0) { w.hd[bi] = w.hd[w.nh - 1]; w.hv[bi] = w.hv[w.nh - 1]; }
    w.nh--;
    wsync();
    const int gj = bv;
    const int jb = nptr[gj], je = nptr[gj + 1];
    for (int p = jb; p < je && !(w.flags & (GCB_GPC_OVERFLOW | GCB_GPC_NONMANIFOLD)); ++p) {
      const int gi = nbr[p];
      if (gi == src) continue;
      double u, th; int k0, k1;
      if (!candidate(w, gi, gj, u, th, k0, k1)) continue;
      const int si = find_vertex(w, gi);
      const double old = si >= 0 ? w.rho[si] : INFINITY;
      if (u < u_max && __ddiv_rn(old, u) > __dadd_rn(1.0, eps)) {
        if (update(w, gi, u, th, gj, k0, k1)) {
          if (w.nh >= w.HC) { w.flags |= GCB_GPC_OVERFLOW; break; }
          wsync();
          if (w.lane == 0) { w.hd[w.nh] = u; w.hv[w.nh] = gi; }
          w.nh++;
          wsync();
        }
      }
    }
  }

  // results: vertex table, triangles (vertex ids + cartesian GPC coordinates) in insertion order
  const size_t vo = (size_t)sys * P, fo = (size_t)sys * w.FC;
  for (int s = w.lane; s < w.nv; s += GCB_WARP) {
    out_vid[vo + s] = w.vid[s]; out_rho[vo + s] = w.rho[s]; out_theta[vo + s] = w.th[s];
  }
  for (int f = w.lane; f < w.nf; f += GCB_WARP) {
    const uint32_t key = w.fkey[f];
    const int s3[3] = {(int)(key >> 20), (int)((key >> 10) & 1023), (int)(key & 1023)};
    for (int c = 0; c < 3; ++c) {
      out_tri[(fo + f) * 3 + c] = w.vid[s3[c]];
      out_xy[((fo + f) * 3 + c) * 2] = w.x[s3[c]];
      out_xy[((fo + f) * 3 + c) * 2 + 1] = w.y[s3[c]];
    }
  }
  if (w.lane == 0) { out_nv[sys] = w.nv; out_nf[sys] = w.nf; out_flags[sys] = w.flags; }
}

}  // namespace gcb

using namespace gcb;

extern "C" int64_t gcb_gpc_smem_per_system(int32_t max_patch) { return gpc_smem_per_system(max_patch); }

extern "C" int gcb_gpc_wavefront(const double* vertices, const double* normals, const int32_t* nbr_ptr,
                                 const int32_t* nbr, const int32_t* opposite, int32_t n_vertices,
                                 const int32_t* sources, int32_t n_sources, double u_max, double eps,
                                 int32_t max_patch, int32_t* out_nv, int32_t* out_vid, double* out_rho,
                                 double* out_theta, int32_t* out_nf, int32_t* out_tri, double* out_xy,
                                 int32_t* out_flags, void* stream) {
  GCB_REQUIRE(n_vertices >= 0 && n_sources >= 0, GCB_EINVAL, "gcb_gpc_wavefront: bad sizes");
  GCB_REQUIRE(max_patch >= 8 && max_patch <= 1023, GCB_EINVAL, "gcb_gpc_wavefront: max_patch must be in [8, 1023]");
  if (n_sources == 0) return GCB_OK;
  GCB_REQUIRE(vertices && normals && nbr_ptr && nbr && opposite, GCB_EINVAL, "gcb_gpc_wavefront: null mesh table");
  GCB_REQUIRE(out_nv && out_vid && out_rho && out_theta && out_nf && out_tri && out_xy && out_flags, GCB_EINVAL,
              "gcb_gpc_wavefront: null output");
  const int64_t per = gpc_smem_per_system(max_patch);
  const int64_t budget = 220 * 1024;
  GCB_REQUIRE(per <= budget, GCB_EUNSUPPORTED, "gcb_gpc_wavefront: max_patch %d needs %lld B of shared memory per system",
              (int)max_patch, (long long)per);
  int wpb = (int)(budget / per);
  if (wpb > 8) wpb = 8;
  // several resident blocks per SM beat one fat block: a block lives as long as its slowest system
  while (wpb > 1 && per * wpb > 56 * 1024) --wpb;
  const size_t smem = (size_t)per * wpb;
  GCB_CUDA_OK(cudaFuncSetAttribute(k_gpc_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gpc_wavefront<<<(unsigned)ceil_div64(n_sources, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(
      vertices, normals, nbr_ptr, nbr, opposite, sources, n_sources, u_max, eps, max_patch, out_nv, out_vid, out_rho,
      out_theta, out_nf, out_tri, out_xy, out_flags);
  GCB_LAUNCH_OK();
  return GCB_OK;
}
