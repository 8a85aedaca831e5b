// Helpers shared by the tensor-core translation units (gcb_tc_fwd.cu defines them).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "gcb_common.cuh"

namespace gcb {

// 2-D fp32 tensor map: `inner` contiguous elements per row, `outer` rows of `row_stride_bytes`;
// box = box_inner x box_outer; swizzle span = box_inner * 4 bytes (128 / 64 / 32).
// atom32 = true selects the 128B swizzle with 32-byte atoms (the only layout UMMA accepts for
// MN-major tf32 operands).
int tc_make_tmap_2d(CUtensorMap* out, const float* base, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool atom32 = false);

// MN-major operand view of a row-major matrix (3-D map: 32-column atoms), see gcb_tc_fwd.cu
int tc_make_tmap_mn(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                    uint32_t box_rows, uint32_t box_atoms);

// fp16 variants (GCB_PREC_F16X3).  2-D: box_inner halves per row, swizzle span = box_inner * 2 bytes.
// MN-major: row-major fp16 matrix [rows][cols_ld] seen as (64 columns, rows, cols_ld / 64); a box = `box_atoms`
// 64-column atoms x `box_rows` rows, each atom deposited as rows of 128 bytes with the standard 128-byte
// swizzle - the canonical MN-major layout of 16-bit UMMA operands (cols_ld % 64 == 0).
int tc_make_tmap_2d_h(CUtensorMap* out, const __half* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer);
int tc_make_tmap_mn_h(CUtensorMap* out, const __half* base, uint64_t rows, uint64_t cols_ld, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_atoms);
// Wcat * (*scale) = hi + lo in fp16, [T][R*A*F + F]
int tc_build_wcat_h(const float* w_eff, const float* w_self, int T, int R, int A, int F, const float* scale,
                    __half* hi, __half* lo, cudaStream_t st);
// fused operand maxima + fp16-mode scales (k_tc_maxima in gcb_tc_fwd.cu): mx = 8 zeroed words (mx[7] is the block
// ticket), segments of floats (max |a[i]|) or of a CSR row pointer (longest row), each reduced into mx[slot]
constexpr int TC_MAX_SEGS = 6;
struct TcMaxSeg { const void* ptr; long long n; int is_rowptr; int slot; };
// mode 0: maxima only; 1: forward scales; 2: backward scales.  ext_wsum (mode 1, nullable): max sum_j |w_j| of a cached
// record packing, used instead of mx[1]
struct TcMaxJob { TcMaxSeg seg[TC_MAX_SEGS]; int blk0[TC_MAX_SEGS + 1]; int n_seg, mode, have_csr; const unsigned* ext_wsum; };
int tc_maxima_scales(const TcMaxJob& job, unsigned* mx, float* scales, cudaStream_t st);
// max |a[i]| over up to two arrays as fp32 bits, atomicMax'ed into *out (zero it first)
int tc_absmax(const float* a, int64_t na, const float* b, int64_t nb, unsigned* out, cudaStream_t st);

// what a caller may prepare once and hand to every forward (gcb_fwd_pack_records / gcb_fwd_build_tables)
size_t tc_fwd_records_bytes(int v_out, int R, int A);
int tc_fwd_pack_records(const int32_t* idx, const float* w, int v_out, int R, int A, void* records, cudaStream_t st);
size_t tc_fwd_tables_bytes();
int tc_fwd_build_tables(const RotTable& rot, int n_rot, int v_out, int R, int A, int F, int T, int precision, void* tables,
                        cudaStream_t st);

// slot-major packed barycentric records [R*A][v_pad][2 x uint4], v_pad = multiple of 128
size_t tc_recs_bytes(int v_out, int R, int A);
int tc_pack_recs(const int32_t* idx, const float* w, int v_out, int R, int A, uint4* recs, cudaStream_t st);

// Wcat[t, 0:KRA] = Weff[t], Wcat[t, KRA:KRA+F] = Wself[t]: tf32 hi part (+ lo = v - hi when lo != NULL)
size_t tc_wcat_bytes(int R, int A, int F, int T);
int tc_build_wcat(const float* w_eff, const float* w_self, int T, int R, int A, int F, float* hi, float* lo,
                  cudaStream_t st);

// Backward through G[v, rotated slot, t] (gcb_tc_dx.cu): G is built once from the CSR inversion, then
//   dSignal = G x [Weff | Wself]   and   d[Weff | Wself] = G^T x X   are two dense tcgen05 GEMMs.
constexpr size_t TC_BUILD_G_SMEM_LIMIT = 200 * 1024;   // (R*A) x T fp32 accumulator of k_bw_build_g
struct TcGLayout {
  size_t g_bytes, xsplit_bytes, partial_bytes, total;
  int v_pad, kg, m_pad, bn, n_ntiles, vsplit, n_kch;
};
TcGLayout tc_g_layout(int v_src, int R, int A, int F, int T, bool x3, bool half = false);
// element-wise maximum of the buffer sizes over the tf32, fp16 and centre-only layouts (what the workspace provides)
TcGLayout tc_g_layout_max(int v_src, int R, int A, int F, int T, bool x3);
// fp16 route (half = true: `scales` = {sG, sW, sX, 1/(sG*sW), 1/(sG*sX)} on the device; g_hi / g_lo / wcat / x_hi /
// x_lo are then __half buffers with row strides kg_ld = align(kg, 64), x_ld = align(F, 64))
size_t tc_bwd_scales_bytes();
int tc_bwd_scales(const float* h, const float* w, const int32_t* csr_row_ptr, const float* w_eff, const float* w_self,
                  const float* x, int v_out, int v_src, int R, int A, int F, int T, void* scratch, cudaStream_t st,
                  int h_rows_per_vertex = 1);
// dense_n_rot > 0: h = [v_out][dense_n_rot][T] (a gradient for every rotation, offsets dense_offsets on the host),
// rot_sel unused
int tc_build_g(const float* h, const float* w, const int32_t* rot_sel, const int32_t* csr_row_ptr,
               const int32_t* csr_entries, int v_out, int v_src, int R, int A, int T, bool x3, float* g_hi,
               float* g_lo, cudaStream_t st, const float* scales = nullptr, bool ring_sorted = false,
               int dense_n_rot = 0, const int32_t* dense_offsets = nullptr);
int tc_dx_from_g(const float* g_hi, const float* g_lo, const float* wcat_hi, const float* wcat_lo, int v_src, int R,
                 int A, int F, int T, bool x3, int accumulate, float* d_x, cudaStream_t st,
                 const float* scales = nullptr);
int tc_dw_from_g(const float* g_hi, const float* g_lo, const float* x, int v_src, int R, int A, int F, int T,
                 bool x3, int accumulate, float* d_w_eff, float* d_w_self, float* x_hi, float* x_lo, float* partial,
                 size_t partial_bytes, cudaStream_t st, const float* scales = nullptr, bool x_is_split = false);

// tensor-core backward (gcb_tc_bwd.cu)
bool tc_bwd_supported(int R, int A, int F, int T);
// centre-only forward (ConvZero) as one tensor-core GEMM
bool tc_centre_fwd_supported(int F, int T);
size_t tc_centre_fwd_bytes(int v_out, int F, int T);
int tc_centre_fwd(const float* x_rows, const float* w_self, const float* bias, int n_rot, int v_out, int F, int T, int act,
                  int precision, float* y, float* z, int32_t* argmax, void* workspace, size_t workspace_bytes,
                  cudaStream_t st);
size_t tc_bwd_workspace_bytes(int v_out, int v_src, int R, int A, int F, int T, int precision);
bool tc_bwd_layout_fits(int v_out, int v_src, int R, int A, int F, int T, int precision);
// AMP-sparse entry: the backward starts from dZ, Z and the pooling decisions instead of h / rot_sel, which it then
// produces itself into `h` [v_out,T] and `rot_sel` [v_out] (caller-provided scratch)
struct TcBwdAmp {
  const float* grad_z;
  const float* z;
  const int32_t* argmax;
  RotTable rot;                 // offsets reduced to [0, A)
  int n_rot, act;
  const uint32_t* csr_stats;    // nullable: [1] max|w| bits, [2] longest CSR row (gcb_csr_build)
  const uint32_t* fwd_stats;    // nullable: [0] max|X| bits, [2] max|Wcat| bits (gcb_conv_fwd)
  float* h;
  int32_t* rot_sel;
};
int tc_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff, const float* w_self,
                const float* h, const int32_t* rot_sel, const int32_t* csr_row_ptr, const int32_t* csr_entries,
                int v_out, int v_src, int R, int A, int F, int T, int precision, int accumulate,
                const float* interp /* nullable: I saved by the forward */, float* d_x, float* d_w_eff,
                float* d_w_self, float* d_bias, void* workspace, size_t workspace_bytes, cudaStream_t st,
                const TcBwdAmp* amp = nullptr, bool csr_ring_grouped = false, int stages = 3, int dense_n_rot = 0,
                const int32_t* dense_offsets = nullptr);

}  // namespace gcb
