// AngularMaxPooling on a materialised Y, activation-gradient, and the deterministic CSR
// gather-reduce that produces dSignal.  All HBM/L2-bound row kernels: one warp or block per row,
// coalesced (float4 when F % 4 == 0), fixed summation order, no atomics.
#include "gcb_common.cuh"

namespace gcb {

// angular_max_pooling.py:27-29: argmax_i ||Y[k,i,:]||_2 (first maximum wins), Z[k] = Y[k,argmax].
__global__ void k_amp_fwd(const float* __restrict__ y, int v, int n_rot, int T,
                          float* __restrict__ z, int32_t* __restrict__ argmax) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= v) return;
  const float* row = y + (size_t)warp * n_rot * T;
  float best = -1.f;
  int arg = 0;
  for (int i = 0; i < n_rot; ++i) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) {
      float e = row[(size_t)i * T + t];
      s = fmaf(e, e, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float nrm = sqrtf(s);
    if (nrm > best) { best = nrm; arg = i; }
  }
  for (int t = lane; t < T; t += 32) z[(size_t)warp * T + t] = row[(size_t)arg * T + t];
  if (lane == 0) argmax[warp] = arg;
}

// AngularMinPooling / AngularAvgPooling on a materialised Y (mode GCB_POOL_MIN / GCB_POOL_AVG); one warp per vertex
__global__ void k_apool_fwd(const float* __restrict__ y, int v, int n_rot, int T, int mode,
                            float* __restrict__ z, int32_t* __restrict__ arg_out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= v) return;
  const float* row = y + (size_t)warp * n_rot * T;
  if (mode == GCB_POOL_AVG) {
    for (int t = lane; t < T; t += 32) {
      float acc = 0.f;
      for (int i = 0; i < n_rot; ++i) acc += row[(size_t)i * T + t];   // rotation order, like torch.mean's reduction
      z[(size_t)warp * T + t] = acc / (float)n_rot;
    }
    return;
  }
  float best = 0.f;
  int arg = 0;
  for (int i = 0; i < n_rot; ++i) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) {
      float e = row[(size_t)i * T + t];
      s = fmaf(e, e, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float nrm = sqrtf(s);
    const bool better = mode == GCB_POOL_MIN ? nrm < best : nrm > best;
    if (i == 0 || better) { best = nrm; arg = i; }
  }
  for (int t = lane; t < T; t += 32) z[(size_t)warp * T + t] = row[(size_t)arg * T + t];
  if (lane == 0 && arg_out) arg_out[warp] = arg;
}

__global__ void k_aavg_bwd(const float* __restrict__ gz, int64_t n /* v*n_rot*T */, int n_rot, int T,
                           float* __restrict__ gy) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  int t = (int)(e % T);
  int64_t k = e / T / n_rot;
  gy[e] = gz[k * T + t] / (float)n_rot;
}

// autograd of the row select: dY[k,i,:] = (i == argmax[k]) ? dZ[k,:] : 0
__global__ void k_amp_bwd(const float* __restrict__ gz, const int32_t* __restrict__ argmax,
                          int64_t n /* v*n_rot*T */, int n_rot, int T, float* __restrict__ gy) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  int t = (int)(e % T);
  int64_t ki = e / T;
  int i = (int)(ki % n_rot);
  int64_t k = ki / n_rot;
  gy[e] = (argmax[k] == i) ? gz[k * T + t] : 0.f;
}

__global__ void k_act_grad(const float* __restrict__ g, const float* __restrict__ out, int64_t n,
                           int act, float* __restrict__ h) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  h[e] = g[e] * act_grad_from_out(act, out[e]);
}

__global__ void k_select_rot(const int32_t* __restrict__ argmax, RotTable rot, int n_rot, int v,
                             int32_t* __restrict__ rot_sel) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= v) return;
  int a = argmax[k];
  rot_sel[k] = rot.off[a < 0 ? 0 : (a >= n_rot ? n_rot - 1 : a)];
}

// dX[v,:] (+)= sum_{slots s -> v} w[s] * D[k_s, x_s, (y_s + rot_sel[k_s]) % A, :]
//               + [v < v_out] sum_t h[v,t] * Wself[t,:]
// D is the unrotated dI [v_out,R,A,F]; the per-vertex rotation is applied as an index here.
// One block per source vertex.  The entry list is decoded cooperatively (one thread per entry:
// slot id -> row of D and weight) into shared memory, then every thread streams its float4 column
// over the rows with 8 independent loads in flight.  Fixed order => bit-reproducible.
constexpr int CSR_CHUNK = 256;

template <int VEC>
__global__ void k_csr_reduce_dx(const float* __restrict__ d_full, const float* __restrict__ w,
                                const float* __restrict__ h, const float* __restrict__ w_self,
                                const int32_t* __restrict__ rot_sel,
                                const int32_t* __restrict__ row_ptr,
                                const int32_t* __restrict__ entries, int v_out, int R, int A,
                                int F, int T, int accumulate, float* __restrict__ d_x) {
  __shared__ uint32_t s_row[CSR_CHUNK];
  __shared__ float s_w[CSR_CHUNK];
  const int v = blockIdx.x;
  const int beg = d_full ? row_ptr[v] : 0, end = d_full ? row_ptr[v + 1] : 0;  // d_full NULL: centre only
  const int RA = R * A;
  const int FV = F / VEC;
  const int fv = threadIdx.x;          // blockDim.x >= FV
  const bool active = fv < FV;
  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
  if (active && v < v_out) {
    for (int t = 0; t < T; ++t) {
      const float hv = h[(size_t)v * T + t];
      const float* ws = w_self + (size_t)t * F + (size_t)fv * VEC;
      if constexpr (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(ws));
        acc[0] = fmaf(hv, q.x, acc[0]); acc[1] = fmaf(hv, q.y, acc[1]);
        acc[2] = fmaf(hv, q.z, acc[2]); acc[3] = fmaf(hv, q.w, acc[3]);
      } else {
        acc[0] = fmaf(hv, __ldg(ws), acc[0]);
      }
    }
  }
  for (int base = beg; base < end; base += CSR_CHUNK) {
    const int n = min(CSR_CHUNK, end - base);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int s = entries[base + i];
      const int cell = s / 3;          // (k, x, y)
      const int k = cell / RA;
      const int slot = cell - k * RA;
      const int xr = slot / A, ya = slot - xr * A;
      int yr = ya + rot_sel[k];
      yr -= (yr >= A) ? A : 0;
      s_row[i] = (uint32_t)(k * RA + xr * A + yr);
      s_w[i] = w[s];
    }
    __syncthreads();
    if (active) {
      const float* col = d_full + (size_t)fv * VEC;
      int i = 0;
      for (; i + 8 <= n; i += 8) {
        if constexpr (VEC == 4) {
          float4 q[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) q[u] = __ldg(reinterpret_cast<const float4*>(col + (size_t)s_row[i + u] * F));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float ws = s_w[i + u];
            acc[0] = fmaf(ws, q[u].x, acc[0]); acc[1] = fmaf(ws, q[u].y, acc[1]);
            acc[2] = fmaf(ws, q[u].z, acc[2]); acc[3] = fmaf(ws, q[u].w, acc[3]);
          }
        } else {
          float q[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) q[u] = __ldg(col + (size_t)s_row[i + u] * F);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[0] = fmaf(s_w[i + u], q[u], acc[0]);
        }
      }
      for (; i < n; ++i) {
        const float ws = s_w[i];
        if constexpr (VEC == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(col + (size_t)s_row[i] * F));
          acc[0] = fmaf(ws, q.x, acc[0]); acc[1] = fmaf(ws, q.y, acc[1]);
          acc[2] = fmaf(ws, q.z, acc[2]); acc[3] = fmaf(ws, q.w, acc[3]);
        } else {
          acc[0] = fmaf(ws, __ldg(col + (size_t)s_row[i] * F), acc[0]);
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    float* dst = d_x + (size_t)v * F + (size_t)fv * VEC;
#pragma unroll
    for (int c = 0; c < VEC; ++c) dst[c] = accumulate ? dst[c] + acc[c] : acc[c];
  }
}

int amp_fwd(const float* y, int v, int n_rot, int T, float* z, int32_t* argmax, cudaStream_t st) {
  if (v <= 0) return GCB_OK;
  const int warps_per_block = 8;
  k_amp_fwd<<<(unsigned)ceil_div64(v, warps_per_block), warps_per_block * 32, 0, st>>>(y, v, n_rot, T, z, argmax);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int csr_reduce_dx(const float* d_full, const float* w, const float* h, const float* w_self,
                  const int32_t* rot_sel, const int32_t* row_ptr, const int32_t* entries,
                  int v_out, int v_src, int R, int A, int F, int T, int accumulate, float* d_x,
                  cudaStream_t st) {
  if (v_src <= 0) return GCB_OK;
  bool vec4 = (F % 4 == 0) && (((uintptr_t)d_full | (uintptr_t)w_self | (uintptr_t)d_x) % 16 == 0);
  int lanes = vec4 ? F / 4 : F;
  GCB_REQUIRE(lanes <= 1024, GCB_EUNSUPPORTED, "csr_reduce_dx: feature dimension %d too large", F);
  int threads = lanes < 32 ? 32 : (lanes + 31) / 32 * 32;
  if (vec4)
    k_csr_reduce_dx<4><<<v_src, threads, 0, st>>>(d_full, w, h, w_self, rot_sel, row_ptr, entries, v_out, R, A, F, T, accumulate, d_x);
  else
    k_csr_reduce_dx<1><<<v_src, threads, 0, st>>>(d_full, w, h, w_self, rot_sel, row_ptr, entries, v_out, R, A, F, T, accumulate, d_x);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int act_grad_launch(const float* grad, const float* out, int64_t n, int act, float* h, cudaStream_t st) {
  if (n <= 0) return GCB_OK;
  k_act_grad<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(grad, out, n, act, h);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int select_rot_launch(const int32_t* argmax, const RotTable& rot, int n_rot, int v, int32_t* rot_sel, cudaStream_t st) {
  if (v <= 0) return GCB_OK;
  k_select_rot<<<(unsigned)ceil_div64(v, 256), 256, 0, st>>>(argmax, rot, n_rot, v, rot_sel);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

}  // namespace gcb

using namespace gcb;

extern "C" int gcb_amp_fwd(const float* y, int32_t v, int32_t n_rot, int32_t T, float* z,
                           int32_t* argmax, void* stream) {
  GCB_REQUIRE(y && z && argmax, GCB_EINVAL, "gcb_amp_fwd: null pointer");
  GCB_REQUIRE(v >= 0 && n_rot > 0 && T > 0, GCB_EINVAL, "gcb_amp_fwd: bad sizes");
  return amp_fwd(y, v, n_rot, T, z, argmax, (cudaStream_t)stream);
}

extern "C" int gcb_amp_bwd(const float* grad_z, const int32_t* argmax, int32_t v, int32_t n_rot,
                           int32_t T, float* grad_y, void* stream) {
  GCB_REQUIRE(grad_z && argmax && grad_y, GCB_EINVAL, "gcb_amp_bwd: null pointer");
  int64_t n = (int64_t)v * n_rot * T;
  if (n <= 0) return GCB_OK;
  k_amp_bwd<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(grad_z, argmax, n, n_rot, T, grad_y);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_apool_fwd(const float* y, int32_t v, int32_t n_rot, int32_t T, int mode, float* z,
                             int32_t* arg, void* stream) {
  GCB_REQUIRE(y && z, GCB_EINVAL, "gcb_apool_fwd: null pointer");
  GCB_REQUIRE(v >= 0 && n_rot > 0 && T > 0, GCB_EINVAL, "gcb_apool_fwd: bad sizes");
  GCB_REQUIRE(mode == GCB_POOL_MAX || mode == GCB_POOL_MIN || mode == GCB_POOL_AVG, GCB_EINVAL,
              "gcb_apool_fwd: unknown pooling mode %d", mode);
  GCB_REQUIRE(mode == GCB_POOL_AVG || arg, GCB_EINVAL, "gcb_apool_fwd: max / min pooling needs the index output");
  if (v == 0) return GCB_OK;
  if (mode == GCB_POOL_MAX) return amp_fwd(y, v, n_rot, T, z, arg, (cudaStream_t)stream);
  const int warps_per_block = 8;
  k_apool_fwd<<<(unsigned)ceil_div64(v, warps_per_block), warps_per_block * 32, 0, (cudaStream_t)stream>>>(
      y, v, n_rot, T, mode, z, arg);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_apool_bwd(const float* grad_z, const int32_t* arg, int32_t v, int32_t n_rot, int32_t T,
                             int mode, float* grad_y, void* stream) {
  GCB_REQUIRE(grad_z && grad_y, GCB_EINVAL, "gcb_apool_bwd: null pointer");
  GCB_REQUIRE(mode == GCB_POOL_MAX || mode == GCB_POOL_MIN || mode == GCB_POOL_AVG, GCB_EINVAL,
              "gcb_apool_bwd: unknown pooling mode %d", mode);
  int64_t n = (int64_t)v * n_rot * T;
  if (n <= 0) return GCB_OK;
  if (mode == GCB_POOL_AVG) {
    k_aavg_bwd<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(grad_z, n, n_rot, T, grad_y);
  } else {
    GCB_REQUIRE(arg, GCB_EINVAL, "gcb_apool_bwd: max / min pooling needs the selected indices");
    k_amp_bwd<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(grad_z, arg, n, n_rot, T, grad_y);
  }
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_act_grad(const float* grad, const float* out, int64_t n, int act, float* h,
                            void* stream) {
  GCB_REQUIRE(grad && out && h, GCB_EINVAL, "gcb_act_grad: null pointer");
  GCB_REQUIRE(act >= 0 && act <= GCB_ACT_TANH, GCB_EINVAL, "gcb_act_grad: unknown activation %d", act);
  if (n <= 0) return GCB_OK;
  k_act_grad<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(grad, out, n, act, h);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_select_rot(const int32_t* argmax, const int32_t* rot_off_host, int32_t n_rot, int32_t A,
                              int32_t v, int32_t* rot_sel, void* stream) {
  GCB_REQUIRE(argmax && rot_off_host && rot_sel, GCB_EINVAL, "gcb_select_rot: null pointer");
  GCB_REQUIRE(n_rot > 0 && n_rot <= GCB_MAX_ROT, GCB_EINVAL, "gcb_select_rot: n_rot=%d out of range", n_rot);
  GCB_REQUIRE(A > 0, GCB_EINVAL, "gcb_select_rot: A must be positive");
  if (v <= 0) return GCB_OK;
  RotTable rot;
  // any integer shift is legal for torch.roll (`orientations=`): the backward kernels expect it reduced to [0, A)
  for (int i = 0; i < GCB_MAX_ROT; ++i) rot.off[i] = i < n_rot ? ((rot_off_host[i] % A) + A) % A : 0;
  k_select_rot<<<(unsigned)ceil_div64(v, 256), 256, 0, (cudaStream_t)stream>>>(argmax, rot, n_rot, v, rot_sel);
  GCB_LAUNCH_OK();
  return GCB_OK;
}
