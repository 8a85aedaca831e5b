// GCB_PREC_FP32: plain-fp32 CUDA-core (FFMA) implementation of the contraction stages.
// This is the exact-arithmetic path and the on-device cross-check for the tcgen05 kernels; it
// materialises the interpolations I[V,R,A,F] in the caller's workspace and runs a tiled SGEMM
// whose operand fetch applies the angular rotation as an index permutation.
#include "gcb_common.cuh"

namespace gcb {

// I[k,slot,:] = sum_j w[k,slot,j] * X[idx[k,slot,j],:]     (conv_intrinsic.py:226-240)
__global__ void k_interp(const float* __restrict__ x, const int32_t* __restrict__ idx,
                         const float* __restrict__ w, int64_t n_rows /* v_out*R*A */, int F,
                         float* __restrict__ out) {
  int64_t row = blockIdx.x;
  if (row >= n_rows) return;
  int i0 = idx[row * 3 + 0], i1 = idx[row * 3 + 1], i2 = idx[row * 3 + 2];
  float w0 = w[row * 3 + 0], w1 = w[row * 3 + 1], w2 = w[row * 3 + 2];
  const float* x0 = x + (size_t)i0 * F;
  const float* x1 = x + (size_t)i1 * F;
  const float* x2 = x + (size_t)i2 * F;
  float* o = out + (size_t)row * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) o[f] = x0[f] * w0 + x1[f] * w1 + x2[f] * w2;
}

// ---- generic tiled SGEMM: C[m][n] = sum_k A(m,k) * B(n,k) ------------------------------------
// Op supplies: int M,N,K;  float a(m,k), b(n,k) (only called in range);  void store(m,n,acc);
//              static bool kAContigK / kBContigK (which index is contiguous in memory).
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, NT = (BM / TM) * (BN / TN);

template <class Op>
__global__ void __launch_bounds__(NT) k_sgemm(Op op) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // two-level accumulation: every FLUSH k-blocks the running sums are folded into `tot`, so no
  // fp32 chain is longer than FLUSH*BK terms (keeps the error well below 1e-6 for K ~ 50k)
  constexpr int FLUSH = 16;
  float tot[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) tot[i][j] = 0.f;

  for (int k0 = 0, blk = 0; k0 < op.K; k0 += BK, ++blk) {
    if (blk == FLUSH) {
      blk = 0;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
    }
    for (int e = tid; e < BM * BK; e += NT) {
      int mm, kk;
      if (Op::kAContigK) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < op.M && k < op.K) ? op.a(m, k) : 0.f;
    }
    for (int e = tid; e < BN * BK; e += NT) {
      int nn, kk;
      if (Op::kBContigK) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < op.N && k < op.K) ? op.b(n, k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int m = m0 + ty * TM + i, n = n0 + tx * TN + j;
      if (m < op.M && n < op.N) op.store(m, n, tot[i][j] + acc[i][j]);
    }
}

template <class Op>
static int launch_sgemm(const Op& op, cudaStream_t st) {
  if (op.M <= 0 || op.N <= 0) return GCB_OK;
  dim3 grid((unsigned)ceil_div64(op.N, BN), (unsigned)ceil_div64(op.M, BM));
  k_sgemm<Op><<<grid, NT, 0, st>>>(op);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// Forward: rows = vertices, cols = (rotation, template), contraction over (x,y,f) then the
// centre slab (f).   Y[k,i,t] = act(sum Weff[t,x,(y+o_i)%A,f]*I[k,x,y,f] + X[k]·Wself[t] + b[t])
struct FwdOp {
  static constexpr bool kAContigK = true, kBContigK = true;
  int M, N, K;
  int KRA, A, F, T, act;
  int n_bcast;   // > 1: centre-only layer (all-zero prior) - every rotation gets the same row: N = T, written n_bcast times
  const float *interp, *x, *w_eff, *w_self, *bias;
  float* y;
  RotTable rot;
  __device__ float a(int m, int k) const {
    return k < KRA ? interp[(size_t)m * KRA + k] : x[(size_t)m * F + (k - KRA)];
  }
  __device__ float b(int n, int k) const {
    int i = n / T, t = n - i * T;
    if (k >= KRA) return w_self[(size_t)t * F + (k - KRA)];
    int slot = k / F, f = k - slot * F;
    int xr = slot / A, ya = slot - xr * A;
    int yw = ya + rot.off[i];
    yw -= (yw >= A) ? A : 0;
    return w_eff[(size_t)t * KRA + (size_t)(xr * A + yw) * F + f];
  }
  __device__ void store(int m, int n, float acc) const {
    int t = n % T;
    const float v = act_apply(act, acc + bias[t]);
    if (n_bcast <= 1) {
      y[(size_t)m * N + n] = v;
    } else {
      for (int i = 0; i < n_bcast; ++i) y[((size_t)m * n_bcast + i) * T + t] = v;
    }
  }
};

// dWeff / dWself / dbias: rows = templates, cols = (x,y',f), then the centre slab (f), then one
// all-ones column (its dot product with h is the bias gradient); contraction over vertices.
struct BwdWOp {
  static constexpr bool kAContigK = false, kBContigK = false;
  int M, N, K;
  int KRA, A, F, T, accumulate;
  const float *interp, *x, *h;
  const int32_t* rot_sel;
  float *d_w_eff, *d_w_self, *d_bias;
  __device__ float a(int m, int k) const { return h[(size_t)k * T + m]; }
  __device__ float b(int n, int k) const {
    if (n >= KRA + F) return 1.f;
    if (n >= KRA) return x[(size_t)k * F + (n - KRA)];
    int slot = n / F, f = n - slot * F;
    int xr = slot / A, yw = slot - xr * A;
    int ys = yw - rot_sel[k];
    ys += (ys < 0) ? A : 0;
    return interp[(size_t)k * KRA + (size_t)(xr * A + ys) * F + f];
  }
  __device__ void store(int m, int n, float acc) const {
    float* p = n < KRA ? d_w_eff + (size_t)m * KRA + n
               : (n < KRA + F ? d_w_self + (size_t)m * F + (n - KRA) : d_bias + m);
    *p = accumulate ? *p + acc : acc;
  }
};

// unrotated dI:  D[k,(x,y',f)] = sum_t h[k,t] * Weff[t,x,y',f]
struct BwdIOp {
  static constexpr bool kAContigK = true, kBContigK = false;
  int M, N, K;
  const float *h, *w_eff;
  float* d_full;
  __device__ float a(int m, int k) const { return h[(size_t)m * K + k]; }
  __device__ float b(int n, int k) const { return w_eff[(size_t)k * N + n]; }
  __device__ void store(int m, int n, float acc) const { d_full[(size_t)m * N + n] = acc; }
};

size_t ffma_fwd_workspace_bytes(int v_out, int R, int A, int F) {
  return align_up((size_t)v_out * R * A * F * sizeof(float), 256);
}

int ffma_conv_fwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                  const float* w_self, const float* bias, const RotTable& rot, int n_rot,
                  int v_out, int v_src, int R, int A, int F, int T, int act, float* y,
                  void* workspace, cudaStream_t st, int row0) {
  (void)v_src;
  float* interp = (float*)workspace;
  int64_t rows = (int64_t)v_out * R * A;
  if (w_eff) {  // w_eff == NULL: all-zero prior (ConvZero) => centre term only
    k_interp<<<(unsigned)rows, 128, 0, st>>>(x, idx, w, rows, F, interp);
    GCB_LAUNCH_OK();
  }
  FwdOp op;
  // ConvZero (reference: conv_zero.py:12-15 still pays the dense cost): the neighbour term vanishes, so all
  // rotations share one row - contract once (N = T) and write it n_rot times
  op.M = v_out; op.N = w_eff ? n_rot * T : T; op.KRA = w_eff ? R * A * F : 0; op.K = op.KRA + F;
  op.n_bcast = w_eff ? 1 : n_rot;
  op.A = A; op.F = F; op.T = T; op.act = act;
  op.interp = interp; op.x = x + (size_t)row0 * F /* centre term: output row k is vertex row0 + k */; op.w_eff = w_eff; op.w_self = w_self; op.bias = bias; op.y = y;
  op.rot = rot;
  return launch_sgemm(op, st);
}

size_t ffma_bwd_workspace_bytes(int v_out, int R, int A, int F) {
  return align_up((size_t)v_out * R * A * F * sizeof(float), 256);
}

int ffma_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                  const float* h, const int32_t* rot_sel, int v_out, int v_src, int R, int A,
                  int F, int T, int accumulate, float* d_w_eff, float* d_w_self, float* d_bias,
                  float* d_full, void* workspace, cudaStream_t st) {
  (void)v_src;
  float* interp = (float*)workspace;
  int64_t rows = (int64_t)v_out * R * A;
  if (w_eff) {
    k_interp<<<(unsigned)rows, 128, 0, st>>>(x, idx, w, rows, F, interp);
    GCB_LAUNCH_OK();
  } else if (!accumulate) {
    GCB_CUDA_OK(cudaMemsetAsync(d_w_eff, 0, (size_t)T * R * A * F * sizeof(float), st));
  }
  BwdWOp wop;
  wop.M = T; wop.KRA = w_eff ? R * A * F : 0; wop.N = wop.KRA + F + 1; wop.K = v_out;
  wop.A = A; wop.F = F; wop.T = T; wop.accumulate = accumulate;
  wop.interp = interp; wop.x = x; wop.h = h; wop.rot_sel = rot_sel;
  wop.d_w_eff = d_w_eff; wop.d_w_self = d_w_self; wop.d_bias = d_bias;
  int rc = launch_sgemm(wop, st);
  if (rc) return rc;
  if (!w_eff || !d_full) return GCB_OK;  // no neighbour term, or dSignal not requested
  BwdIOp iop;
  iop.M = v_out; iop.N = R * A * F; iop.K = T; iop.h = h; iop.w_eff = w_eff; iop.d_full = d_full;
  return launch_sgemm(iop, st);
}

}  // namespace gcb
