// Helpers shared by the tensor-core translation units (gcb_tc_fwd.cu defines them).
#pragma once
#include <cuda.h>

#include "gcb_common.cuh"

namespace gcb {

// 2-D fp32 tensor map: `inner` contiguous elements per row, `outer` rows of `row_stride_bytes`;
// box = box_inner x box_outer; swizzle span = box_inner * 4 bytes (128 / 64 / 32).
// atom32 = true selects the 128B swizzle with 32-byte atoms (the only layout UMMA accepts for
// MN-major tf32 operands).
int tc_make_tmap_2d(CUtensorMap* out, const float* base, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool atom32 = false);

// slot-major packed barycentric records [R*A][v_pad][2 x uint4], v_pad = multiple of 128
size_t tc_recs_bytes(int v_out, int R, int A);
int tc_pack_recs(const int32_t* idx, const float* w, int v_out, int R, int A, uint4* recs, cudaStream_t st);

// Wcat[t, 0:KRA] = Weff[t], Wcat[t, KRA:KRA+F] = Wself[t]: tf32 hi part (+ lo = v - hi when lo != NULL)
size_t tc_wcat_bytes(int R, int A, int F, int T);
int tc_build_wcat(const float* w_eff, const float* w_self, int T, int R, int A, int F, float* hi, float* lo,
                  cudaStream_t st);

// dSignal through G[v, rotated slot, t] and a dense tcgen05 GEMM (gcb_tc_dx.cu)
size_t tc_dx_workspace_bytes(int v_src, int R, int A, int T, bool x3);
int tc_dx_via_g(const float* h, const float* w, const int32_t* rot_sel, const int32_t* csr_row_ptr,
                const int32_t* csr_entries, const float* wcat_hi, const float* wcat_lo, int v_out, int v_src,
                int R, int A, int F, int T, bool x3, int accumulate, float* d_x, void* workspace,
                size_t workspace_bytes, cudaStream_t st);

// tensor-core backward (gcb_tc_bwd.cu)
bool tc_bwd_supported(int R, int A, int F, int T);
size_t tc_bwd_workspace_bytes(int v_out, int v_src, int R, int A, int F, int T, int precision);
int tc_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff, const float* w_self,
                const float* h, const int32_t* rot_sel, const int32_t* csr_row_ptr, const int32_t* csr_entries,
                int v_out, int v_src, int R, int A, int F, int T, int precision, int accumulate,
                const float* interp /* nullable: I saved by the forward */, float* d_x, float* d_w_eff,
                float* d_w_self, float* d_bias, void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace gcb
