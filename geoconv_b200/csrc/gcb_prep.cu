// One-time / per-step preparation kernels: bary split, CSR inversion, prior folding, row packing.
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include "gcb_common.cuh"
#include "gcb_tc.cuh"

namespace gcb {

// ---- bary [n_slots,2] -> idx int32, w fp32  (conv_intrinsic.py:213-221) ---------------------
template <typename TIn>
__global__ void k_prep_bary(const TIn* __restrict__ bary, int64_t n_slots, int32_t v_src,
                            int32_t* __restrict__ idx, float* __restrict__ w,
                            int32_t* __restrict__ oob_count) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  // the reference casts the whole tensor to fp32 first, then truncates toward zero with .long()
  float fi = (float)bary[2 * s];
  float fw = (float)bary[2 * s + 1];
  long long li = (long long)fi;  // cvt.rzi
  int32_t i32 = (int32_t)li;
  if (li < 0 || li >= v_src) {
    atomicAdd(oob_count, 1);
    i32 = 0;
    fw = 0.f;
  }
  idx[s] = i32;
  w[s] = fw;
}

// ---- CSR inversion ---------------------------------------------------------------------------
// key = source vertex * rings + radial ring of the slot (rings = R when the caller gives the template size, else 1):
// inside a vertex's row the entries are then grouped by radial ring, ascending slot id inside a ring - what lets
// k_bw_build_g2 keep a ring's accumulators in registers (the rotation only moves a slot inside its ring)
__global__ void k_csr_keys(const int32_t* __restrict__ idx, const float* __restrict__ w,
                           int64_t n_slots, int32_t v_src, int32_t rings, int32_t A, int32_t* __restrict__ keys,
                           int32_t* __restrict__ vals) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  int32_t ring = 0;
  if (rings > 1) ring = (int32_t)(((s / 3) % ((int64_t)rings * A)) / A);
  keys[s] = (w[s] == 0.f) ? v_src * rings : idx[s] * rings + ring;  // zero-weight slots sort to the tail and are dropped
  vals[s] = (int32_t)s;
}

__global__ void k_csr_row_ptr(const int32_t* __restrict__ sorted_keys, int64_t n_slots,
                              int32_t v_src, int32_t rings, int32_t* __restrict__ row_ptr) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > v_src) return;
  // lower_bound(sorted_keys, v * rings)
  const int32_t key = v * rings;
  int64_t lo = 0, hi = n_slots;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  row_ptr[v] = (int32_t)lo;
}

static size_t cub_sort_temp_bytes(int64_t n_slots, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_slots, 0,
                                  end_bit);
  return bytes;
}

static int key_bits(int64_t max_key) {
  int bits = 1;
  while ((1ll << bits) <= (long long)max_key) ++bits;
  return bits;
}

// ---- prior folding ---------------------------------------------------------------------------
// out[t,q,f] = sum_p in[t,p,f] * (transpose ? prior[q,p] : prior[p,q]),  p,q in [0, R*A)
// One block = one template t and 32 consecutive features: the [RA x 32] input slab is staged in
// smem once; warp w produces the 8 output slots q = 8w .. 8w+7 in registers (lanes = features).  Per
// input slot p a thread reads its own value once and the 8 prior elements as two broadcast
// 128-bit loads, so the loop runs at 8 FMAs per 3 shared-memory instructions (the first version did
// 1 FMA per 2 and was shared-memory-issue bound at 31 us for 8.4 MB).
// FOLD_TB templates per block: the prior is staged once for all of them, their tiles fly together.  MAXT = the block
// size bound (4 * RAp threads): 256 covers R*A <= 64 and leaves the compiler 255 registers instead of the 64 a bound of
// 1024 allows (the 32 accumulators + 32 parked loads did not fit in 64).
template <int MAXT, int FOLD_TB>
__global__ void __launch_bounds__(MAXT) k_fold_prior(const float* __restrict__ in, const float* __restrict__ prior,
                                                    float* __restrict__ out, int T, int RA, int RAp, int F, int transpose) {
  extern __shared__ float4 s_mem4[];
  float* s_prior = reinterpret_cast<float*>(s_mem4);   // [RA][RAp]: s_prior[p*RAp + q] multiplies in[p] into out[q]
  float* s_in = s_prior + RA * RAp;                    // [FOLD_TB][RA][32]
  const int t0 = blockIdx.y * FOLD_TB, f0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // The input tiles come cold from HBM, the prior from L2: the loads of all FOLD_TB tiles are issued first and
  // parked in registers (blockDim = 4 * RAp threads, so a thread owns at most 8 of the RA * 32 elements of a tile),
  // the prior is staged while they fly - one exposed round trip per block (ncu, r01: half of the stall samples sat
  // on the shared-memory stores waiting for these loads; one template per block paid that round trip and the prior
  // staging 4x as often: 14 us per call).  blockDim = 4 * RAp also gives (row, column) of a prior element without
  // a division per element.
  float xin[FOLD_TB][8];
#pragma unroll
  for (int b = 0; b < FOLD_TB; ++b)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = threadIdx.x + j * blockDim.x;
      const int p = i >> 5, fl = i & 31;
      xin[b][j] = (i < RA * 32 && f0 + fl < F && t0 + b < T) ? __ldg(in + ((size_t)(t0 + b) * RA + p) * F + f0 + fl) : 0.f;
    }
  {
    const int p0 = threadIdx.x / RAp, q = threadIdx.x - p0 * RAp;
    for (int p = p0; p < RA; p += 4)
      s_prior[p * RAp + q] = q < RA ? __ldg(transpose ? prior + q * RA + p : prior + p * RA + q) : 0.f;
  }
#pragma unroll
  for (int b = 0; b < FOLD_TB; ++b)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = threadIdx.x + j * blockDim.x;
      if (i < RA * 32) s_in[b * RA * 32 + i] = xin[b][j];
    }
  __syncthreads();
  const int q0 = warp * 8;
  if (q0 >= RA || f0 + lane >= F) return;
  // p outermost: the eight prior values of this warp's columns are read once per p for all FOLD_TB templates
  // (32 FMAs per 2 + FOLD_TB shared-memory loads; one template at a time was 8 per 3 and load-issue bound)
  float acc[FOLD_TB][8];
#pragma unroll
  for (int b = 0; b < FOLD_TB; ++b)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[b][j] = 0.f;
  for (int p = 0; p < RA; ++p) {
    const float4 k0 = *reinterpret_cast<const float4*>(s_prior + p * RAp + q0);
    const float4 k1 = *reinterpret_cast<const float4*>(s_prior + p * RAp + q0 + 4);
#pragma unroll
    for (int b = 0; b < FOLD_TB; ++b) {
      const float x = s_in[(b * RA + p) * 32 + lane];
      acc[b][0] = fmaf(x, k0.x, acc[b][0]); acc[b][1] = fmaf(x, k0.y, acc[b][1]);
      acc[b][2] = fmaf(x, k0.z, acc[b][2]); acc[b][3] = fmaf(x, k0.w, acc[b][3]);
      acc[b][4] = fmaf(x, k1.x, acc[b][4]); acc[b][5] = fmaf(x, k1.y, acc[b][5]);
      acc[b][6] = fmaf(x, k1.z, acc[b][6]); acc[b][7] = fmaf(x, k1.w, acc[b][7]);
    }
  }
#pragma unroll
  for (int b = 0; b < FOLD_TB; ++b) {
    if (t0 + b >= T) break;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (q0 + j < RA) out[((size_t)(t0 + b) * RA + q0 + j) * F + f0 + lane] = acc[b][j];
  }
}

__global__ void k_pack_rows(const float* __restrict__ src, const int32_t* __restrict__ rows,
                            int n_rows, int F, float* __restrict__ out) {
  int r = blockIdx.x;
  if (r >= n_rows) return;
  const float* s = src + (size_t)rows[r] * F;
  float* d = out + (size_t)r * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) d[f] = s[f];
}

}  // namespace gcb

using namespace gcb;

extern "C" int gcb_prep_bary(const void* bary, int bary_is_f64, int64_t n_slots, int32_t v_src,
                             int32_t* idx, float* w, int32_t* oob_count, void* stream) {
  GCB_REQUIRE(bary && idx && w && oob_count, GCB_EINVAL, "gcb_prep_bary: null pointer");
  GCB_REQUIRE(n_slots >= 0 && v_src > 0, GCB_EINVAL, "gcb_prep_bary: bad sizes");
  if (n_slots == 0) return GCB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)ceil_div64(n_slots, 256);
  if (bary_is_f64)
    k_prep_bary<double><<<grid, 256, 0, st>>>((const double*)bary, n_slots, v_src, idx, w, oob_count);
  else
    k_prep_bary<float><<<grid, 256, 0, st>>>((const float*)bary, n_slots, v_src, idx, w, oob_count);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" size_t gcb_csr_workspace_bytes(int64_t n_slots, int32_t v_src) {
  size_t arr = align_up((size_t)n_slots * sizeof(int32_t), 256);
  return 3 * arr + align_up(cub_sort_temp_bytes(n_slots, 31), 256) + 256;   // (any key width)
}

extern "C" int gcb_csr_build(const int32_t* idx, const float* w, int64_t n_slots, int32_t v_src,
                             int32_t R, int32_t A, int32_t* row_ptr, int32_t* entries, uint32_t* stats,
                             void* workspace, size_t workspace_bytes, void* stream) {
  GCB_REQUIRE(idx && w && row_ptr && entries && workspace, GCB_EINVAL, "gcb_csr_build: null pointer");
  GCB_REQUIRE(R >= 0 && A >= 0 && (R == 0) == (A == 0), GCB_EINVAL, "gcb_csr_build: R and A go together");
  const int32_t rings = R > 0 ? R : 1;
  GCB_REQUIRE((int64_t)v_src * rings + rings < (1ll << 31), GCB_EUNSUPPORTED, "gcb_csr_build: V_src * R exceeds int32 keys");
  GCB_REQUIRE(R == 0 || n_slots % ((int64_t)R * A * 3) == 0, GCB_EINVAL, "gcb_csr_build: n_slots is not a multiple of R*A*3");
  GCB_REQUIRE(n_slots > 0 && n_slots < (1ll << 31) && v_src > 0, GCB_EINVAL, "gcb_csr_build: bad sizes");
  GCB_REQUIRE(workspace_bytes >= gcb_csr_workspace_bytes(n_slots, v_src), GCB_EWORKSPACE,
              "gcb_csr_build: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  size_t arr = align_up((size_t)n_slots * sizeof(int32_t), 256);
  char* base = (char*)workspace;
  int32_t* keys_in = (int32_t*)base;
  int32_t* keys_out = (int32_t*)(base + arr);
  int32_t* vals_in = (int32_t*)(base + 2 * arr);
  void* cub_tmp = base + 3 * arr;
  int bits = key_bits((int64_t)v_src * rings);
  size_t cub_bytes = cub_sort_temp_bytes(n_slots, bits);
  unsigned grid = (unsigned)ceil_div64(n_slots, 256);
  k_csr_keys<<<grid, 256, 0, st>>>(idx, w, n_slots, v_src, rings, A > 0 ? A : 1, keys_in, vals_in);
  GCB_LAUNCH_OK();
  // LSD radix sort is stable: slots of one source vertex stay in ascending slot order.
  GCB_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, vals_in,
                                              entries, (int)n_slots, 0, bits, st));
  k_csr_row_ptr<<<(unsigned)ceil_div64((int64_t)v_src + 1, 256), 256, 0, st>>>(keys_out, n_slots, v_src, rings, row_ptr);
  GCB_LAUNCH_OK();
  if (stats) {
    // what the fp16 backward needs to bound |G| (gcb_conv_amp_bwd): largest |weight| and longest row - properties of
    // the barycentric tensor, computed here once instead of in every backward
    GCB_CUDA_OK(cudaMemsetAsync(stats, 0, 32, st));
    TcMaxJob job{};
    job.mode = 0;
    job.n_seg = 2;
    job.seg[0] = {w, (long long)n_slots, 0, 1};
    job.seg[1] = {row_ptr, (long long)v_src, 1, 2};
    return tc_maxima_scales(job, stats, nullptr, st);
  }
  return GCB_OK;
}

extern "C" int gcb_fold_prior(const float* in, const float* prior, float* out, int32_t T, int32_t R,
                              int32_t A, int32_t F, int transpose, void* stream) {
  GCB_REQUIRE(in && prior && out, GCB_EINVAL, "gcb_fold_prior: null pointer");
  GCB_REQUIRE(T > 0 && R > 0 && A > 0 && F > 0, GCB_EINVAL, "gcb_fold_prior: bad sizes");
  int RA = R * A;
  const int RAp = (RA + 7) / 8 * 8;
  const int threads = RAp / 8 * 32;
  const int tb = 4;   // templates per block (two per block, i.e. twice the blocks: 25.4 us against 23.9 for fold + unfold)
  size_t smem = ((size_t)RA * RAp + (size_t)tb * RA * 32) * sizeof(float);
  GCB_REQUIRE(smem <= 160 * 1024 && RAp / 8 <= 32, GCB_EUNSUPPORTED, "gcb_fold_prior: R*A=%d too large", RA);
  dim3 grid((unsigned)ceil_div64(F, 32), (unsigned)ceil_div64(T, tb));
  auto launch = [&](auto kernel) -> int {
    if (smem > 48 * 1024)
      GCB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, (unsigned)threads, smem, (cudaStream_t)stream>>>(in, prior, out, T, RA, RAp, F, transpose);
    GCB_LAUNCH_OK();
    return GCB_OK;
  };
  return threads > 256 ? launch(k_fold_prior<1024, 4>) : launch(k_fold_prior<256, 4>);
}

extern "C" int gcb_pack_rows(const float* src, const int32_t* rows, int32_t n_rows, int32_t F,
                             float* out, void* stream) {
  GCB_REQUIRE(src && rows && out, GCB_EINVAL, "gcb_pack_rows: null pointer");
  if (n_rows <= 0) return GCB_OK;
  k_pack_rows<<<n_rows, 128, 0, (cudaStream_t)stream>>>(src, rows, n_rows, F, out);
  GCB_LAUNCH_OK();
  return GCB_OK;
}
