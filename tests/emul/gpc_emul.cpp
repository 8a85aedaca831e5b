// TEST INFRASTRUCTURE.  Host build of the GPC wavefront kernel's algorithm text (see gpc_emul_shim.h).
// tests/test_gpc_wavefront.py::emul copies the DEVICE section of geoconv_b200/csrc/gcb_gpc.cu - everything between its
// #include "gcb_common.cuh" and the end of namespace gcb, verbatim - into _build/gcb_gpc_device_section.inc and compiles
//   g++ -O1 -shared -fPIC -I tests/emul -x c++ tests/emul/gpc_emul.cpp -o tests/emul/_build/libgpc_emul.so
// This file exports gpc_emul_run(), which calls the kernel body once per source vertex.
#include "gpc_emul_shim.h"
#include "_build/gcb_gpc_device_section.inc"

extern "C" int64_t gpc_emul_smem_per_system(int32_t P) { return gcb::gpc_smem_per_system(P); }

extern "C" int gpc_emul_run(const double* vertices, const double* normals, const int32_t* nbr_ptr, const int32_t* nbr,
                            const int32_t* opposite, const int32_t* sources, int32_t n_sources, double u_max, double eps,
                            int32_t max_patch, int32_t* out_nv, int32_t* out_vid, double* out_rho, double* out_theta,
                            int32_t* out_nf, int32_t* out_tri, double* out_xy, int32_t* out_flags) {
  if (gcb::gpc_smem_per_system(max_patch) > (int64_t)sizeof(gcb::smem)) return 1;
  blockDim.x = 1;
  threadIdx.x = 0;
  for (int s = 0; s < n_sources; ++s) {
    blockIdx.x = (unsigned)s;
    gcb::k_gpc_wavefront(vertices, normals, nbr_ptr, nbr, opposite, sources, n_sources, u_max, eps, max_patch, out_nv,
                         out_vid, out_rho, out_theta, out_nf, out_tri, out_xy, out_flags);
  }
  return 0;
}
