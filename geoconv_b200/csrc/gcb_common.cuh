// Shared device/host helpers for the geoconv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/geoconv_b200.h"

namespace gcb {

void set_error(const char* fmt, ...);

#define GCB_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      gcb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                        \
      return GCB_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

#define GCB_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      gcb::set_error(__VA_ARGS__);   \
      return (code);                 \
    }                                \
  } while (0)

// every kernel launch of the library goes through this: error check + launch accounting
void count_launch();
// gcb_yield_sms(): persistent GEMMs size their grid to the tile rounds they need instead of to the SM count
int yield_sms();
#define GCB_LAUNCH_OK()              \
  do {                               \
    gcb::count_launch();             \
    GCB_CUDA_OK(cudaGetLastError()); \
  } while (0)

// optional per-kernel timing (bench.py's roofline leg): events recorded around the dominant
// kernel on its own stream while profiling is enabled
void profile_begin(int which, cudaStream_t st);
void profile_end(int which, cudaStream_t st);
enum { GCB_PROF_FWD_CONV = 0, GCB_PROF_BWD_DW = 1, GCB_PROF_BWD_DI = 2, GCB_PROF_BWD_CSR = 3, GCB_PROF_N = 4 };

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- activations (ACTIVATIONS table, conv_intrinsic.py:10-17; torch default parameters) ----
#define GCB_SELU_ALPHA 1.6732632423543772848170429916717f
#define GCB_SELU_SCALE 1.0507009873554804934193349852946f

__device__ __forceinline__ float act_apply(int act, float x) {
  switch (act) {
    case GCB_ACT_RELU: return fmaxf(x, 0.f);
    case GCB_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case GCB_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
    case GCB_ACT_SELU: return GCB_SELU_SCALE * (x > 0.f ? x : GCB_SELU_ALPHA * expm1f(x));
    case GCB_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case GCB_ACT_TANH: return tanhf(x);
  }
  return x;
}

// derivative w.r.t. the pre-activation, written in terms of the activation output y
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
  switch (act) {
    case GCB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case GCB_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case GCB_ACT_LEAKY_RELU: return y > 0.f ? 1.f : 0.01f;
    case GCB_ACT_SELU: return y > 0.f ? GCB_SELU_SCALE : y + GCB_SELU_SCALE * GCB_SELU_ALPHA;
    case GCB_ACT_SIGMOID: return y * (1.f - y);
    case GCB_ACT_TANH: return 1.f - y * y;
  }
  return 1.f;
}

struct RotTable {
  int32_t off[GCB_MAX_ROT];
};

// ---- entry points implemented per translation unit ----------------------------------------
// fp32 CUDA-core path (gcb_ffma.cu)
size_t ffma_fwd_workspace_bytes(int v_out, int R, int A, int F);
int ffma_conv_fwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                  const float* w_self, const float* bias, const RotTable& rot, int n_rot,
                  int v_out, int v_src, int R, int A, int F, int T, int act, float* y,
                  void* workspace, cudaStream_t st, int row0 = 0);
size_t ffma_bwd_workspace_bytes(int v_out, int R, int A, int F);
int ffma_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                  const float* h, const int32_t* rot_sel, int v_out, int v_src, int R, int A,
                  int F, int T, int accumulate, float* d_w_eff, float* d_w_self, float* d_bias,
                  float* d_full /* out: [v_out,R,A,F] unrotated dI */, void* workspace,
                  cudaStream_t st);

// tcgen05 path (gcb_tc_fwd.cu / gcb_tc_bwd.cu)
bool tc_supported(int R, int A, int F, int T, int n_rot);
bool tc_half_supported(int R, int A, int F, int T, int n_rot);
size_t tc_fwd_workspace_bytes(int v_out, int R, int A, int F, int T, int n_rot, int precision);
int tc_conv_fwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                const float* w_self, const float* bias, const RotTable& rot, int n_rot, int v_out,
                int v_src, int R, int A, int F, int T, int act, int precision, float* y,
                float* z /* nullable: fused AngularMaxPooling */, int32_t* argmax,
                float* interp_out /* nullable: keep I[v_out,R,A,F] for the backward */,
                const void* records /* nullable: gcb_fwd_pack_records */, const void* tables /* nullable: gcb_fwd_build_tables */,
                uint32_t* stats_out /* nullable: 8 words of operand maxima for the backward */,
                int row0 /* output row k is vertex row0 + k of the signal (centre term) */, void* workspace,
                size_t workspace_bytes, cudaStream_t st);

// shared small kernels (gcb_pool.cu)
int amp_fwd(const float* y, int v, int n_rot, int T, float* z, int32_t* argmax, cudaStream_t st);
int csr_reduce_dx(const float* d_full, const float* w, const float* h, const float* w_self,
                  const int32_t* rot_sel, const int32_t* row_ptr, const int32_t* entries,
                  int v_out, int v_src, int R, int A, int F, int T, int accumulate, float* d_x,
                  cudaStream_t st);

}  // namespace gcb
