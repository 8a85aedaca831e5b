// The chain BETWEEN two convolutions of the reference's network block:
//   ... -> ISC -> AngularMaxPooling -> BatchNorm1d -> Dropout -> ISC -> ...
// (/root/reference/src/geoconv_examples/mpi_faust/pytorch/model.py:93-105 builds, :128-132 runs
//  do -> isc -> amp -> bn per block, i.e. bn of block l feeds the dropout of block l+1).
// Between the pooled output Z (V, T) of one convolution and the input of the next there are a column
// statistic and two elementwise maps; PyTorch runs them as ~4 kernels forward and ~5 backward, each a pass over
// (V, T).  Here:
//   forward   k_bn_stats        : per-feature mean / inverse std of Z (training mode), deterministic
//             k_bn_drop_fwd     : X' = dropout(gamma * (Z - mean) * invstd + beta)   - ONE pass, the mask comes
//                                 from a counter-based generator (Philox 4x32-10 on the element index), so it is
//                                 never stored: the backward regenerates it
//   backward  k_bn_drop_bwd_red : dY = mask * dX' / (1 - p);  d_beta = sum dY, d_gamma = sum dY * xhat   (one pass)
//             k_bn_drop_bwd     : dZ = gamma * invstd * (dY - d_beta / V - xhat * d_gamma / V)           (one pass)
// All four are HBM/L2-bound streaming kernels over V * T floats (2.6 MB at 6890 x 96); reductions add in a fixed
// order (double accumulators per thread, block partials summed in block order by the block that finishes last):
// bit-reproducible like the rest of the library.
//
// Why this is NOT folded into the convolution's gather producers (the obvious "fusion"): the affine map commutes
// with the barycentric interpolation (one FMA per element after the weighted sum), but the dropout mask does not -
// it is per signal ROW, and every output row gathers 3 * R * A other rows, so the producers would evaluate the
// generator 3 * R * A times per element instead of once (360 x at R5 x A8) inside the loop that bounds the kernel.
// One extra pass over 2.6 MB costs ~3 us next to a ~0.5 ms convolution.
#include "gcb_common.cuh"

namespace gcb {

// ---- Philox 4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3") -------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
    const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask of the four elements 4*g .. 4*g+3 (g = float4 index): element kept iff its 32 random bits < thresh
__device__ __forceinline__ uint4 keep4(uint64_t g, uint64_t seed, uint64_t stream) {
  return philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

constexpr int BN_BLOCKS = 48;    // target number of blocks of a column reduction
constexpr int BN_THREADS = 256;
constexpr int BN_COLS = 128;     // columns per column block (32 float4 lanes)

// Column sums of two per-element quantities.  Grid = (row blocks, column blocks of BN_COLS columns).  A block adds its
// `rows` rows: thread (row lane ty, float4 column tx) in row order (float), the row lanes through shared memory in lane
// order (double).  The block partials [column block][row block][2][BN_COLS] (double) of a column block are summed in
// row-block order by the block of that column block that finishes last, which hands every column's two sums to
// `fin`.  Fixed order everywhere: bit-reproducible.  Few row blocks per column block keep that last step short.
template <class Elem4, class Final>
__device__ void column_reduce(int V, int T, int rows, double* __restrict__ partial, unsigned* __restrict__ ticket,
                              Elem4 elem, Final fin) {
  __shared__ __align__(16) float s_red[(BN_THREADS / 32) * 2 * BN_COLS];   // [row lanes][2][cols] (>= 8 lanes x 128 columns)
  __shared__ unsigned last;
  const int col0 = blockIdx.y * BN_COLS, cols = min(BN_COLS, T - col0), c4n = cols >> 2;
  const int tpr = c4n < 32 ? c4n : 32;                // threads per row
  const int lanes_all = BN_THREADS / tpr;
  const int lanes = lanes_all < 8 ? lanes_all : 8;    // row lanes (shared-memory scratch holds 8)
  const int ty = threadIdx.x / tpr, tx = threadIdx.x - ty * tpr;
  const int r0 = blockIdx.x * rows, r1 = min(V, r0 + rows);
  const int T4 = T >> 2, cq0 = col0 >> 2;
  if (ty < lanes && tx < c4n) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    int r = r0 + ty;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {   // four rows in flight, added in row order
      float4 ea[4], eb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) elem((size_t)(r + u * lanes) * T4 + cq0 + tx, cq0 + tx, ea[u], eb[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a.x += ea[u].x; a.y += ea[u].y; a.z += ea[u].z; a.w += ea[u].w;
        b.x += eb[u].x; b.y += eb[u].y; b.z += eb[u].z; b.w += eb[u].w;
      }
    }
    for (; r < r1; r += lanes) {
      float4 ea, eb;
      elem((size_t)r * T4 + cq0 + tx, cq0 + tx, ea, eb);
      a.x += ea.x; a.y += ea.y; a.z += ea.z; a.w += ea.w;
      b.x += eb.x; b.y += eb.y; b.z += eb.z; b.w += eb.w;
    }
    *reinterpret_cast<float4*>(s_red + ((size_t)ty * 2 + 0) * BN_COLS + 4 * tx) = a;
    *reinterpret_cast<float4*>(s_red + ((size_t)ty * 2 + 1) * BN_COLS + 4 * tx) = b;
  }
  __syncthreads();
  double* mine = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * BN_COLS;
  for (int i = threadIdx.x; i < 2 * cols; i += BN_THREADS) {
    const int which = i / cols, c = i - which * cols;
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += (double)s_red[((size_t)l * 2 + which) * BN_COLS + c];
    mine[which * BN_COLS + c] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket + blockIdx.y, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block of this column block: thread (quantity, column) adds the row blocks' partials in block order,
  // eight loads in flight
  const double* base = partial + (size_t)blockIdx.y * gridDim.x * 2 * BN_COLS;
  const int nb = (int)gridDim.x;
  for (int i = threadIdx.x; i < 2 * cols; i += BN_THREADS) {
    const int which = i / cols, c = i - which * cols;
    const double* src = base + which * BN_COLS + c;
    double acc = 0.0;
    int blk = 0;
    for (; blk + 8 <= nb; blk += 8) {
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = __ldcg(src + (size_t)(blk + u) * 2 * BN_COLS);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += x[u];
    }
    for (; blk < nb; ++blk) acc += __ldcg(src + (size_t)blk * 2 * BN_COLS);
    reinterpret_cast<double*>(s_red)[which * BN_COLS + c] = acc;   // (every thread is past the row-lane scratch)
  }
  __syncthreads();
  const double* pair = reinterpret_cast<const double*>(s_red);
  for (int c = threadIdx.x; c < cols; c += BN_THREADS) fin(col0 + c, pair[c], pair[BN_COLS + c]);
  if (threadIdx.x == 0) ticket[blockIdx.y] = 0u;   // ready for the next call (graph replays included)
}

// training-mode statistics; running estimates updated in place when given (momentum as torch: new = (1-m) old + m x)
__global__ void __launch_bounds__(BN_THREADS) k_bn_stats(const float* __restrict__ z, int V, int T, int rows, float eps,
                                                         double* __restrict__ partial, unsigned* __restrict__ ticket,
                                                         float* __restrict__ mean, float* __restrict__ invstd,
                                                         float* __restrict__ running_mean,
                                                         float* __restrict__ running_var, float momentum) {
  const float4* z4 = reinterpret_cast<const float4*>(z);
  column_reduce(
      V, T, rows, partial, ticket,
      [&](size_t e4, int c4, float4& a, float4& b) {
        const float4 v = __ldg(z4 + e4);
        a = v;
        b = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
      },
      [&](int c, double s, double s2) {
        const double m = s / V;
        double var = s2 / V - m * m;
        var = var > 0.0 ? var : 0.0;
        mean[c] = (float)m;
        invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) {
          const double unbiased = V > 1 ? var * V / (V - 1) : var;
          running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
          running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        }
      });
}

__global__ void k_bn_drop_fwd(const float4* __restrict__ z, int64_t n4, int T4, const float4* __restrict__ mean,
                              const float4* __restrict__ invstd, const float4* __restrict__ gamma,
                              const float4* __restrict__ beta, float p, uint64_t seed, uint64_t stream,
                              float4* __restrict__ out) {
  const uint32_t thresh = p <= 0.f ? 0xffffffffu : (uint32_t)((1.0 - (double)p) * 4294967296.0);
  const float keep_scale = p <= 0.f ? 1.f : 1.f / (1.f - p);
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((uint32_t)g % (uint32_t)T4);   // (n4 < 2^31: checked on the host)
    const float4 v = __ldg(z + g), m = __ldg(mean + c), s = __ldg(invstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
    float4 y;
    y.x = fmaf((v.x - m.x) * s.x, ga.x, be.x); y.y = fmaf((v.y - m.y) * s.y, ga.y, be.y);
    y.z = fmaf((v.z - m.z) * s.z, ga.z, be.z); y.w = fmaf((v.w - m.w) * s.w, ga.w, be.w);
    if (p > 0.f) {
      const uint4 k = keep4((uint64_t)g, seed, stream);
      y.x = k.x < thresh ? y.x * keep_scale : 0.f; y.y = k.y < thresh ? y.y * keep_scale : 0.f;
      y.z = k.z < thresh ? y.z * keep_scale : 0.f; y.w = k.w < thresh ? y.w * keep_scale : 0.f;
    }
    out[g] = y;
  }
}

__global__ void __launch_bounds__(BN_THREADS)
k_bn_drop_bwd_red(const float* __restrict__ d_out, const float* __restrict__ z, int V, int T, int rows,
                  const float* __restrict__ mean, const float* __restrict__ invstd, float p, uint64_t seed,
                  uint64_t stream, double* __restrict__ partial, unsigned* __restrict__ ticket,
                  float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  const uint32_t thresh = p <= 0.f ? 0xffffffffu : (uint32_t)((1.0 - (double)p) * 4294967296.0);
  const float keep_scale = p <= 0.f ? 1.f : 1.f / (1.f - p);
  const float4* g4 = reinterpret_cast<const float4*>(d_out);
  const float4* z4 = reinterpret_cast<const float4*>(z);
  column_reduce(
      V, T, rows, partial, ticket,
      [&](size_t e4, int c4, float4& a, float4& b) {
        float4 dy = __ldg(g4 + e4);
        if (p > 0.f) {
          const uint4 k = keep4((uint64_t)e4, seed, stream);
          dy.x = k.x < thresh ? dy.x * keep_scale : 0.f; dy.y = k.y < thresh ? dy.y * keep_scale : 0.f;
          dy.z = k.z < thresh ? dy.z * keep_scale : 0.f; dy.w = k.w < thresh ? dy.w * keep_scale : 0.f;
        }
        const float4 v = __ldg(z4 + e4), m = __ldg(reinterpret_cast<const float4*>(mean) + c4),
                     s = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
        a = dy;
        b = make_float4(dy.x * ((v.x - m.x) * s.x), dy.y * ((v.y - m.y) * s.y), dy.z * ((v.z - m.z) * s.z),
                        dy.w * ((v.w - m.w) * s.w));
      },
      [&](int c, double s, double s2) { d_beta[c] = (float)s; d_gamma[c] = (float)s2; });
}

__global__ void k_bn_drop_bwd(const float4* __restrict__ d_out, const float4* __restrict__ z, int64_t n4, int T4,
                              int V, const float4* __restrict__ mean, const float4* __restrict__ invstd,
                              const float4* __restrict__ gamma, const float4* __restrict__ d_gamma,
                              const float4* __restrict__ d_beta, int training, float p, uint64_t seed, uint64_t stream,
                              float4* __restrict__ d_z) {
  const uint32_t thresh = p <= 0.f ? 0xffffffffu : (uint32_t)((1.0 - (double)p) * 4294967296.0);
  const float keep_scale = p <= 0.f ? 1.f : 1.f / (1.f - p);
  const float inv_v = 1.f / (float)V;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((uint32_t)g % (uint32_t)T4);
    float4 dy = __ldg(d_out + g);
    if (p > 0.f) {
      const uint4 k = keep4((uint64_t)g, seed, stream);
      dy.x = k.x < thresh ? dy.x * keep_scale : 0.f; dy.y = k.y < thresh ? dy.y * keep_scale : 0.f;
      dy.z = k.z < thresh ? dy.z * keep_scale : 0.f; dy.w = k.w < thresh ? dy.w * keep_scale : 0.f;
    }
    const float4 s = __ldg(invstd + c), ga = __ldg(gamma + c);
    float4 o;
    if (training) {
      const float4 v = __ldg(z + g), m = __ldg(mean + c), dg = __ldg(d_gamma + c), db = __ldg(d_beta + c);
      o.x = ga.x * s.x * (dy.x - db.x * inv_v - (v.x - m.x) * s.x * dg.x * inv_v);
      o.y = ga.y * s.y * (dy.y - db.y * inv_v - (v.y - m.y) * s.y * dg.y * inv_v);
      o.z = ga.z * s.z * (dy.z - db.z * inv_v - (v.z - m.z) * s.z * dg.z * inv_v);
      o.w = ga.w * s.w * (dy.w - db.w * inv_v - (v.w - m.w) * s.w * dg.w * inv_v);
    } else {   // running statistics are constants
      o.x = ga.x * s.x * dy.x; o.y = ga.y * s.y * dy.y; o.z = ga.z * s.z * dy.z; o.w = ga.w * s.w * dy.w;
    }
    d_z[g] = o;
  }
}

static int stream_grid(int64_t n4) {
  const int64_t blocks = ceil_div64(n4, 256);
  return (int)(blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8);
}

}  // namespace gcb

using namespace gcb;

static dim3 bn_grid(int V, int T) {   // (row blocks, column blocks)
  const int ncb = (int)ceil_div64(T, BN_COLS);
  int nrb = BN_BLOCKS / ncb;
  nrb = nrb < 1 ? 1 : nrb;
  const int rows = (int)ceil_div64(V, nrb);
  return dim3((unsigned)ceil_div64(V, rows), (unsigned)ncb);
}

extern "C" size_t gcb_bn_workspace_bytes(int32_t V, int32_t T) {
  (void)V;
  const size_t ncb = (size_t)ceil_div64(T > 0 ? T : 1, BN_COLS);
  return 256 + ncb * BN_BLOCKS * 2 * BN_COLS * sizeof(double);
}

static int bn_args_ok(const void* z, int V, int T, const void* ws, size_t ws_bytes, const char* who) {
  GCB_REQUIRE(z && ws, GCB_EINVAL, "%s: null pointer", who);
  GCB_REQUIRE((int64_t)V * T / 4 < (1ll << 31), GCB_EUNSUPPORTED, "%s: V * T too large", who);
  GCB_REQUIRE(V > 0 && T > 0 && T % 4 == 0, GCB_EINVAL, "%s: need V > 0 and T %% 4 == 0 (got %d, %d)", who, V, T);
  GCB_REQUIRE(ws_bytes >= gcb_bn_workspace_bytes(V, T), GCB_EWORKSPACE, "%s: workspace too small", who);
  GCB_REQUIRE((uintptr_t)z % 16 == 0 && (uintptr_t)ws % 16 == 0, GCB_EINVAL, "%s: pointers must be 16-byte aligned", who);
  return GCB_OK;
}

// The first 256 bytes of the workspace hold the ticket and must be ZERO before the first call (the kernels leave
// them zero again).
extern "C" int gcb_bn_stats(const float* z, int32_t V, int32_t T, float eps, float* mean, float* invstd,
                            float* running_mean, float* running_var, float momentum, void* workspace,
                            size_t workspace_bytes, void* stream) {
  int rc = bn_args_ok(z, V, T, workspace, workspace_bytes, "gcb_bn_stats");
  if (rc) return rc;
  GCB_REQUIRE(mean && invstd && (!running_mean == !running_var), GCB_EINVAL, "gcb_bn_stats: bad output pointers");
  unsigned* ticket = (unsigned*)workspace;
  double* partial = (double*)((char*)workspace + 256);
  const dim3 grid = bn_grid(V, T);
  const int rows = (int)ceil_div64(V, grid.x);
  k_bn_stats<<<grid, BN_THREADS, 0, (cudaStream_t)stream>>>(
      z, V, T, rows, eps, partial, ticket, mean, invstd, running_mean, running_var, momentum);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_bn_dropout_fwd(const float* z, int32_t V, int32_t T, const float* mean, const float* invstd,
                                  const float* gamma, const float* beta, float p, uint64_t seed, uint64_t stream_id,
                                  float* out, void* stream) {
  GCB_REQUIRE(z && mean && invstd && gamma && beta && out, GCB_EINVAL, "gcb_bn_dropout_fwd: null pointer");
  GCB_REQUIRE(V > 0 && T > 0 && T % 4 == 0 && p >= 0.f && p < 1.f, GCB_EINVAL, "gcb_bn_dropout_fwd: bad argument");
  GCB_REQUIRE((int64_t)V * T / 4 < (1ll << 31), GCB_EUNSUPPORTED, "gcb_bn_dropout_fwd: V * T too large");
  const int64_t n4 = (int64_t)V * T / 4;
  k_bn_drop_fwd<<<stream_grid(n4), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)z, n4, T / 4, (const float4*)mean, (const float4*)invstd, (const float4*)gamma, (const float4*)beta,
      p, seed, stream_id, (float4*)out);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_bn_dropout_bwd(const float* d_out, const float* z, int32_t V, int32_t T, const float* mean,
                                  const float* invstd, const float* gamma, int training, float p, uint64_t seed,
                                  uint64_t stream_id, float* d_z, float* d_gamma, float* d_beta, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  int rc = bn_args_ok(z, V, T, workspace, workspace_bytes, "gcb_bn_dropout_bwd");
  if (rc) return rc;
  GCB_REQUIRE(d_out && mean && invstd && gamma && d_z && d_gamma && d_beta && p >= 0.f && p < 1.f, GCB_EINVAL,
              "gcb_bn_dropout_bwd: bad argument");
  unsigned* ticket = (unsigned*)workspace;
  double* partial = (double*)((char*)workspace + 256);
  cudaStream_t st = (cudaStream_t)stream;
  GCB_REQUIRE((uintptr_t)d_out % 16 == 0 && (uintptr_t)mean % 16 == 0 && (uintptr_t)invstd % 16 == 0, GCB_EINVAL,
              "gcb_bn_dropout_bwd: pointers must be 16-byte aligned");
  const dim3 grid = bn_grid(V, T);
  const int rows = (int)ceil_div64(V, grid.x);
  k_bn_drop_bwd_red<<<grid, BN_THREADS, 0, st>>>(
      d_out, z, V, T, rows, mean, invstd, p, seed, stream_id, partial, ticket, d_gamma, d_beta);
  GCB_LAUNCH_OK();
  const int64_t n4 = (int64_t)V * T / 4;
  k_bn_drop_bwd<<<stream_grid(n4), 256, 0, st>>>((const float4*)d_out, (const float4*)z, n4, T / 4, V,
                                                (const float4*)mean, (const float4*)invstd, (const float4*)gamma,
                                                (const float4*)d_gamma, (const float4*)d_beta, training, p, seed,
                                                stream_id, (float4*)d_z);
  GCB_LAUNCH_OK();
  return GCB_OK;
}
