// GPC systems -> barycentric coordinates on the GPU: the step right before the hot path.
// Replaces the triple Python loop of compute_barycentric_coordinates
// (/root/reference/src/geoconv/preprocessing/barycentric_coordinates.py:127-175) with
// interpolation (:46-69) and compute_barycentric (:7-43):
//   for every GPC system v and template vertex q = (r, a): the FIRST triangle (in the system's own face
//   order) that contains q gives the three vertex indices and barycentric weights; none -> zeros.
// One warp per (v, q): lanes test 32 triangles at a time in float64 with the reference's operation order
// (no FMA contraction: __dmul_rn / __dadd_rn), a ballot picks the lowest triangle index that is inside.
// HBM-bound integer/fp64 work: 56 B per triangle read once per template vertex from L2.
#include <float.h>

#include "gcb_common.cuh"

namespace gcb {

__global__ void k_bary_from_gpc(const double* __restrict__ tri_xy, const int32_t* __restrict__ tri_idx,
                                const int64_t* __restrict__ tri_ptr, int V, const double* __restrict__ tmpl_xy,
                                int RA, double* __restrict__ out) {
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (int64_t)V * RA) return;
  const int v = (int)(warp / RA), q = (int)(warp - (int64_t)v * RA);
  const double qx = tmpl_xy[2 * q], qy = tmpl_xy[2 * q + 1];
  const int64_t beg = tri_ptr[v], end = tri_ptr[v + 1];
  double* o = out + warp * 6;   // [3][2]: (index, weight) per triangle corner
  for (int64_t base = beg; base < end; base += 32) {
    const int64_t t = base + lane;
    bool inside = false;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    if (t < end) {
      const double* p = tri_xy + t * 6;
      const double ax = p[0], ay = p[1], bx = p[2], by = p[3], cx = p[4], cy = p[5];
      // v0 = C - A, v1 = B - A, v2 = Q - A   (barycentric_coordinates.py:26-28)
      const double v0x = __dsub_rn(cx, ax), v0y = __dsub_rn(cy, ay);
      const double v1x = __dsub_rn(bx, ax), v1y = __dsub_rn(by, ay);
      const double v2x = __dsub_rn(qx, ax), v2y = __dsub_rn(qy, ay);
      const double dot00 = __dadd_rn(__dmul_rn(v0x, v0x), __dmul_rn(v0y, v0y));
      const double dot01 = __dadd_rn(__dmul_rn(v0x, v1x), __dmul_rn(v0y, v1y));
      const double dot02 = __dadd_rn(__dmul_rn(v0x, v2x), __dmul_rn(v0y, v2y));
      const double dot11 = __dadd_rn(__dmul_rn(v1x, v1x), __dmul_rn(v1y, v1y));
      const double dot12 = __dadd_rn(__dmul_rn(v1x, v2x), __dmul_rn(v1y, v2y));
      double den = __dsub_rn(__dmul_rn(dot00, dot11), __dmul_rn(dot01, dot01));
      if (den == 0.0) den = __dadd_rn(den, DBL_MIN);   // sys.float_info.min (:34-35)
      w2 = __ddiv_rn(__dsub_rn(__dmul_rn(dot11, dot02), __dmul_rn(dot01, dot12)), den);
      w1 = __ddiv_rn(__dsub_rn(__dmul_rn(dot00, dot12), __dmul_rn(dot01, dot02)), den);
      w0 = __dsub_rn(__dsub_rn(1.0, w2), w1);
      inside = w2 > 0.0 && w1 > 0.0 && __dadd_rn(w2, w1) <= 1.0;   // (:40)
    }
    const unsigned hit = __ballot_sync(0xffffffffu, inside);
    if (hit) {
      if (lane == __ffs((int)hit) - 1) {
        const int32_t* ix = tri_idx + t * 3;
        o[0] = (double)ix[0]; o[1] = w0;
        o[2] = (double)ix[1]; o[3] = w1;
        o[4] = (double)ix[2]; o[5] = w2;
      }
      return;
    }
  }
  if (lane < 6) o[lane] = 0.0;   // no triangle contains the template vertex (:69)
}

}  // namespace gcb

using namespace gcb;

extern "C" int gcb_bary_from_gpc(const double* tri_xy, const int32_t* tri_idx, const int64_t* tri_ptr, int32_t V,
                                 const double* template_xy, int32_t RA, double* bary_out, void* stream) {
  GCB_REQUIRE(tri_ptr && template_xy && bary_out, GCB_EINVAL, "gcb_bary_from_gpc: null pointer");
  GCB_REQUIRE(V >= 0 && RA > 0, GCB_EINVAL, "gcb_bary_from_gpc: bad sizes");
  if (V == 0) return GCB_OK;
  GCB_REQUIRE(tri_xy && tri_idx, GCB_EINVAL, "gcb_bary_from_gpc: null triangle arrays");
  const int64_t warps = (int64_t)V * RA;
  const int warps_per_block = 8;
  k_bary_from_gpc<<<(unsigned)ceil_div64(warps, warps_per_block), warps_per_block * 32, 0, (cudaStream_t)stream>>>(
      tri_xy, tri_idx, tri_ptr, V, template_xy, RA, bary_out);
  GCB_LAUNCH_OK();
  return GCB_OK;
}
