// tcgen05 backward of ConvIntrinsic for one rotation per vertex (the AMP-sparse case; a dense dY is
// back-propagated by calling this once per rotation).  Inputs: h = dZ * act' [V,T], rot_sel [V].
//
//   k_tc_dw   dWcat[t, kk'] = sum_k h[k,t] * Arow_k[kk'],  kk' over (x, y', f) then the centre slab f:
//             Arow_k = I[k, x, (y'-rot_sel[k])%A, f]  (gathered + interpolated on the fly)  |  X[k, f].
//             GEMM with M = 128 kk' values, N = templates, K = vertices.  The gathered operand is
//             written exactly like the forward's A tile (a vertex = one 128-byte row of 32 features)
//             and handed to the tensor core as an MN-major operand, so the contraction index
//             (vertex) is the row index and no transpose is ever materialised.  h^T tiles come by TMA.
//             Vertices are split over CTAs (fills the SMs, keeps accumulation chains short); partials
//             are summed in fp32 by k_tc_dw_finish in a fixed order.
//             This gather-fed kernel is the route taken when the caller supplies no CSR inversion of the
//             barycentric indices; with the CSR (what the torch layer always passes) both gradients go
//             through G (gcb_tc_dx.cu): d[Weff|Wself] = G^T X and dSignal = G [Weff|Wself], two dense GEMMs.
#include <cuda.h>

#include "gcb_common.cuh"
#include "gcb_ptx.cuh"
#include "gcb_tc.cuh"

namespace gcb {

constexpr int BW_SMEM_LIMIT = 227 * 1024;

// MN-major descriptor for tf32: the only layout UMMA accepts is SWIZZLE_128B_BASE32B (layout type 1):
// 32 fp32 along M/N are contiguous (a 128 B row), 4 K-rows form a 512 B atom whose four 32-byte chunks
// are XOR-swizzled with (row % 4), i.e. Swizzle<2,5,2> on byte addresses.  `lbo` = byte distance between
// atoms along M/N, `sbo` = between 4-row groups along K (one K=8 MMA reads two of them).
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__global__ void k_bw_transpose_h(const float* __restrict__ h, int v_out, int v_pad, int T,
                                 float* __restrict__ ht_hi, float* __restrict__ ht_lo,
                                 float* __restrict__ bias_part) {
  // ht[t][v] (v padded with zeros) for k_tc_dw;
  // bias_part[blockIdx.x][t] = sum of this block's 32 rows (first stage of d_bias)
  __shared__ float tile[32][33];
  const int v0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int v = v0 + i, t = t0 + threadIdx.x;
    float val = (v < v_out && t < T) ? h[(size_t)v * T + t] : 0.f;
    tile[i][threadIdx.x] = val;
  }
  __syncthreads();
  if (threadIdx.y == 0 && t0 + threadIdx.x < T) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += tile[i][threadIdx.x];
    bias_part[(size_t)blockIdx.x * T + t0 + threadIdx.x] = acc;
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, v = v0 + threadIdx.x;
    if (t < T && v < v_pad) {
      float val = tile[threadIdx.x][i];
      float hi = ptx::to_tf32(val);
      ht_hi[(size_t)t * v_pad + v] = hi;
      if (ht_lo) ht_lo[(size_t)t * v_pad + v] = val - hi;
    }
  }
}

// =================================================================================================
// k_tc_dw
// =================================================================================================
constexpr int DW_PRODUCER_WARPS = 8;
constexpr int DW_MMA_WARPS = 2;
constexpr int DW_THREADS = (DW_PRODUCER_WARPS + 1 + DW_MMA_WARPS) * 32;
constexpr int DW_VC = 32;  // vertices per stage
constexpr int DW_A_BYTES = 16 * 1024;  // 32 vertices x 128 features fp32

struct TcDwParams {
  const float* x;
  const uint4* recs;        // [R*A][v_pad][2]
  const int32_t* rot_sel;   // [v_out]
  const float* interp;      // optional [v_out][R*A*F] saved by the forward: read instead of re-gathering
  float* partial;           // [vsplit][T][kcat_pad]
  int v_out, v_pad, R, A, F, T, kcat, kcat_pad, n_t, vsplit, n_chunks, ns, tmem_cols;
};

template <bool X3>
__global__ void __launch_bounds__(DW_THREADS, 1)
k_tc_dw(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
        const TcDwParams p) {
  constexpr int NPROD = X3 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int b_bytes = p.n_t * 128;
  const int stage_bytes = NPROD * (DW_A_BYTES + b_bytes);
  const uint32_t stage_base = smem_base;
  const uint32_t rot_base = stage_base + (uint32_t)(p.ns * stage_bytes);   // int32 rot per vertex of my range
  const int chunk_begin = (int)(((int64_t)p.n_chunks * blockIdx.y) / p.vsplit);
  const int chunk_end = (int)(((int64_t)p.n_chunks * (blockIdx.y + 1)) / p.vsplit);
  const int n_local = chunk_end - chunk_begin;
  const int max_local = (p.n_chunks + p.vsplit - 1) / p.vsplit + 1;
  const uint32_t bar_base = (rot_base + (uint32_t)(max_local * DW_VC * 4) + 15u) & ~15u;
  const uint32_t full = bar_base, empty = full + 8u * p.ns, acc_full = empty + 8u * p.ns;
  const uint32_t tmem_slot = acc_full + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int t0 = blockIdx.z * p.n_t;
  const int RA = p.R * p.A;

  if (warp == DW_PRODUCER_WARPS && lane == 0) {
    for (int s = 0; s < p.ns; ++s) {
      ptx::mbar_init(full + 8u * s, DW_PRODUCER_WARPS + 1);
      ptx::mbar_init(empty + 8u * s, DW_MMA_WARPS);
    }
    ptx::mbar_init(acc_full, DW_MMA_WARPS);
    ptx::fence_mbar_init();
    ptx::prefetch_tensormap(&tmap_hi);
    if (X3) ptx::prefetch_tensormap(&tmap_lo);
  }
  if (warp == DW_PRODUCER_WARPS + 1) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  // rotation of every vertex in this CTA's range (zero beyond v_out)
  {
    int32_t* rot_s = reinterpret_cast<int32_t*>(smem_gen + (rot_base - smem_base));
    for (int i = threadIdx.x; i < n_local * DW_VC; i += blockDim.x) {
      int k = chunk_begin * DW_VC + i;
      rot_s[i] = k < p.v_out ? p.rot_sel[k] : 0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp < DW_PRODUCER_WARPS) {
    // ============ producers: every warp owns 4 vertices of the 32-vertex stage ============
    const int32_t* rot_s = reinterpret_cast<const int32_t*>(smem_gen + (rot_base - smem_base));
    const int c8 = lane & 7;
    const int vrow = warp * 4 + (lane >> 3);   // vertex row inside the stage
    // the four 32-feature blocks of this kk' tile: slot (x, y') or centre, feature offset
    int slot_j[4], f0_j[4];
    bool valid_j[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk0 = tile * 128 + 32 * j;
      valid_j[j] = kk0 < p.kcat;
      slot_j[j] = valid_j[j] ? kk0 / p.F : 0;
      f0_j[j] = valid_j[j] ? kk0 - slot_j[j] * p.F : 0;
    }
    uint4 ri[4], rw[4];
    auto load_recs = [&](int n) {
      const int k = (chunk_begin + n) * DW_VC + vrow;
      const int rot = rot_s[n * DW_VC + vrow];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (valid_j[j] && slot_j[j] < RA && p.interp) {
          // saved interpolation: one row segment of I[k, x, (y'-rot)%A, :] replaces three gathers
          const int xr = slot_j[j] / p.A;
          int ys = slot_j[j] - xr * p.A - rot;
          ys += ys < 0 ? p.A : 0;
          ri[j] = make_uint4((uint32_t)(xr * p.A + ys), 0u, 0u, 0u);
          rw[j] = make_uint4(__float_as_uint(k < p.v_out ? 1.f : 0.f), 0u, 0u, 0u);
          if (k >= p.v_out) ri[j].x = 0u;
        } else if (valid_j[j] && slot_j[j] < RA) {
          if (j > 0 && slot_j[j] == slot_j[j - 1]) {
            ri[j] = ri[j - 1];
            rw[j] = rw[j - 1];
          } else {
            const int xr = slot_j[j] / p.A;
            int ys = slot_j[j] - xr * p.A - rot;
            ys += ys < 0 ? p.A : 0;
            const uint4* rec = p.recs + ((size_t)(xr * p.A + ys) * p.v_pad + k) * 2;
            ri[j] = __ldg(rec);
            rw[j] = __ldg(rec + 1);
          }
        } else {
          // centre slab (or padding): the row itself with weight 1 (0 for padding rows)
          ri[j] = make_uint4((uint32_t)k, 0u, 0u, 0u);
          rw[j] = make_uint4(__float_as_uint((valid_j[j] && k < p.v_out) ? 1.f : 0.f), 0u, 0u, 0u);
          if (k >= p.v_out) ri[j].x = 0u;
        }
      }
    };
    if (n_local > 0) load_recs(0);
    int stage = 0;
    uint32_t empty_par = 1u;
    for (int n = 0; n < n_local; ++n) {
      float4 x0[4], x1[4], x2[4];
      float w0[4], w1[4], w2[4];
      const int kcur = (chunk_begin + n) * DW_VC + vrow;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* xcol = p.x + f0_j[j] + c8 * 4;
        w0[j] = __uint_as_float(rw[j].x); w1[j] = __uint_as_float(rw[j].y); w2[j] = __uint_as_float(rw[j].z);
        if (p.interp && valid_j[j] && slot_j[j] < RA) {
          const size_t krow = (size_t)(kcur < p.v_out ? kcur : 0) * RA + ri[j].x;
          x0[j] = __ldg(reinterpret_cast<const float4*>(p.interp + krow * p.F + f0_j[j] + c8 * 4));
          x1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          x2[j] = x1[j];
        } else {
          x0[j] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)ri[j].x * p.F));
          x1[j] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)ri[j].y * p.F));
          x2[j] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)ri[j].z * p.F));
        }
      }
      if (n + 1 < n_local) load_recs(n + 1);   // records of the next stage fly while X arrives
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j].x = fmaf(x2[j].x, w2[j], fmaf(x1[j].x, w1[j], x0[j].x * w0[j]));
        v[j].y = fmaf(x2[j].y, w2[j], fmaf(x1[j].y, w1[j], x0[j].y * w0[j]));
        v[j].z = fmaf(x2[j].z, w2[j], fmaf(x1[j].z, w1[j], x0[j].z * w0[j]));
        v[j].w = fmaf(x2[j].w, w2[j], fmaf(x1[j].w, w1[j], x0[j].w * w0[j]));
      }
      ptx::mbar_wait(empty + 8u * stage, empty_par);
      uint8_t* a_hi = smem_gen + (size_t)stage * stage_bytes;
      // MN-major tf32 tile: feature block j at j*4096, vertex row at vrow*128, 32-byte chunks
      // swizzled with (row % 4) (SWIZZLE_128B_BASE32B)
      const uint32_t off_row = (uint32_t)vrow * 128u + (uint32_t)((((c8 >> 1) ^ (vrow & 3)) << 5) | ((c8 & 1) << 4));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 hv;
        hv.x = ptx::to_tf32(v[j].x); hv.y = ptx::to_tf32(v[j].y);
        hv.z = ptx::to_tf32(v[j].z); hv.w = ptx::to_tf32(v[j].w);
        *reinterpret_cast<float4*>(a_hi + j * 4096 + off_row) = hv;
        if (X3) {
          float4 lv;
          lv.x = v[j].x - hv.x; lv.y = v[j].y - hv.y; lv.z = v[j].z - hv.z; lv.w = v[j].w - hv.w;
          *reinterpret_cast<float4*>(a_hi + DW_A_BYTES + j * 4096 + off_row) = lv;
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(full + 8u * stage);
      if (++stage == p.ns) { stage = 0; empty_par ^= 1u; }
    }
    // ============ epilogue: D0 + D1 -> partial[split][t][kk'] ============
    if (warp < 4) {
      ptx::mbar_wait(acc_full, 0);
      ptx::tc_fence_after();
      const int row = tile * 128 + 32 * warp + lane;   // kk' (< kcat_pad)
      const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * warp) << 16);
      float* out = p.partial + ((size_t)blockIdx.y * p.T + t0) * p.kcat_pad + row;
      for (int c0 = 0; c0 < p.n_t; c0 += 16) {
        uint32_t r0[16], r1[16];
        ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)c0, r0);
        ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)(p.n_t + c0), r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          out[(size_t)(c0 + q) * p.kcat_pad] = n_local > 0 ? __uint_as_float(r0[q]) + __uint_as_float(r1[q]) : 0.f;
      }
    }
  } else if (warp == DW_PRODUCER_WARPS) {
    // ============ TMA: h^T tiles [n_t templates x 32 vertices] ============
    if (lane == 0) {
      int stage = 0;
      uint32_t empty_par = 1u;
      for (int n = 0; n < n_local; ++n) {
        ptx::mbar_wait(empty + 8u * stage, empty_par);
        const uint32_t dst = stage_base + (uint32_t)(stage * stage_bytes + NPROD * DW_A_BYTES);
        ptx::mbar_arrive_expect_tx(full + 8u * stage, (uint32_t)(NPROD * b_bytes));
        ptx::tma_load_2d(dst, &tmap_hi, full + 8u * stage, (chunk_begin + n) * DW_VC, t0);
        if (X3) ptx::tma_load_2d(dst + b_bytes, &tmap_lo, full + 8u * stage, (chunk_begin + n) * DW_VC, t0);
        if (++stage == p.ns) { stage = 0; empty_par ^= 1u; }
      }
    }
  } else {
    // ============ MMA issuers: warp m owns k-blocks {2m, 2m+1} and accumulator D_m ============
    const int mw = warp - (DW_PRODUCER_WARPS + 1);
    if (lane == 0) {
      // idesc: f32 accum, tf32 x tf32, A MN-major (bit 15), B K-major, N = n_t, M = 128
      const uint32_t idesc = ptx::make_idesc_tf32(128, p.n_t) | (1u << 15);
      const uint32_t d = tmem_base + (uint32_t)(mw * p.n_t);
      int stage = 0;
      uint32_t full_par = 0, accum = 0;
      for (int n = 0; n < n_local; ++n) {
        ptx::mbar_wait(full + 8u * stage, full_par);
        ptx::tc_fence_after();
        const uint32_t a_hi = stage_base + (uint32_t)(stage * stage_bytes);
        const uint32_t b_hi = a_hi + NPROD * DW_A_BYTES;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int kb = mw * 2 + q;
          const uint64_t da = make_mnmajor_desc(a_hi + kb * 1024, 4096, 512);
          const uint64_t db = ptx::make_kmajor_desc<128>(b_hi + kb * 32);
          ptx::umma_tf32(d, da, db, idesc, accum);
          accum = 1u;
          if (X3) {
            ptx::umma_tf32(d, da, ptx::make_kmajor_desc<128>(b_hi + b_bytes + kb * 32), idesc, 1u);
            ptx::umma_tf32(d, make_mnmajor_desc(a_hi + DW_A_BYTES + kb * 1024, 4096, 512), db, idesc, 1u);
          }
        }
        ptx::umma_commit(empty + 8u * stage);
        if (++stage == p.ns) { stage = 0; full_par ^= 1u; }
      }
      ptx::umma_commit(acc_full);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == DW_PRODUCER_WARPS + 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// sums the vertex-split partials in a fixed order and scatters to d_w_eff / d_w_self
__global__ void k_tc_dw_finish(const float* __restrict__ partial, int vsplit, int T, int kcat, int kcat_pad,
                               int KRA, int F, int accumulate, float* __restrict__ d_w_eff,
                               float* __restrict__ d_w_self) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)T * kcat) return;
  const int t = (int)(e / kcat), kk = (int)(e - (int64_t)t * kcat);
  float acc = 0.f;
  for (int s = 0; s < vsplit; ++s) acc += partial[((size_t)s * T + t) * kcat_pad + kk];
  float* dst = kk < KRA ? (d_w_eff ? d_w_eff + (size_t)t * KRA + kk : nullptr) : d_w_self + (size_t)t * F + (kk - KRA);
  if (dst) *dst = accumulate ? *dst + acc : acc;
}

// bias_part[b][t] = sum of h[r][t] over the rows of strip b (first stage of d_bias, fixed order):
// 8 row lanes per column stride through the strip, then a fixed-order sum over the lanes
__global__ void k_bw_bias_partial(const float* __restrict__ h, int v_out, int T, int rows_per_block,
                                  float* __restrict__ bias_part) {
  __shared__ float s_part[8][33];
  const int t = blockIdx.y * 32 + threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(v_out, r0 + rows_per_block);
  float acc = 0.f;
  if (t < T)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += __ldg(h + (size_t)r * T + t);
  s_part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && t < T) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += s_part[i][threadIdx.x];
    bias_part[(size_t)blockIdx.x * T + t] = sum;
  }
}

// rows of h as a GEMM operand: tf32 hi (+ lo), zero rows from v_out up to v_pad
__global__ void k_bw_split_rows(const float* __restrict__ h, int v_out, int v_pad, int T, float* __restrict__ hi,
                                float* __restrict__ lo) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)v_pad * T) return;
  const float val = e < (int64_t)v_out * T ? h[e] : 0.f;
  const float hv = ptx::to_tf32(val);
  hi[e] = hv;
  if (lo) lo[e] = val - hv;
}

// d_bias[t] (+)= sum over the per-block partial sums (fixed order)
__global__ void k_bw_bias_final(const float* __restrict__ part, int nblocks, int T, int accumulate,
                                float* __restrict__ d_bias) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float acc = 0.f;
  for (int b = 0; b < nblocks; ++b) acc += part[(size_t)b * T + t];
  d_bias[t] = accumulate ? d_bias[t] + acc : acc;
}


// =================================================================================================
// k_bw_prepare: everything the AMP-sparse backward needs before its first big kernel, in ONE launch
//   role A (blocks [0, nb_h)):  h = dZ * act'(Z), rot_sel[k] = rot[argmax[k]], per-block column sums of h (first
//           stage of d_bias), max|h|; the block that finishes last (ticket) adds the column sums in block order
//           (fixed order: bit-reproducible) into d_bias and, in the fp16 mode, turns the operand maxima into the
//           five power-of-two scales {sG, sW, sX, 1/(sG sW), 1/(sG sX)}
//   role B (fp16 mode with the forward's maxima at hand):  X * sX = hi + lo (fp16)   - the B operand of G^T X
//   role C (same condition):  [Weff | Wself] * sW = hi + lo (fp16)                   - the B operand of G Wcat
// It replaces six launches (activation gradient, rotation select, maxima, two bias reductions, a memset) and runs
// the two operand splits, which depend on nothing computed here, beside them.
// =================================================================================================
struct BwPrepParams {
  const float* grad_z;
  const float* z;
  const int32_t* argmax;
  RotTable rot;              // offsets in [0, A)
  int n_rot, act, v_out, T;
  float* h;
  int32_t* rot_sel;
  float* bias_part;          // [nb_h][T]
  int rows_per_block, nb_h, accumulate;
  float* d_bias;
  unsigned* mx;              // 8 zeroed words: [0] max|h| bits, [6] ticket; [1..4] maxima when the stats are absent ([7]: k_tc_maxima's ticket)
  const unsigned* csr_stats; // nullable: [1] max|w| bits, [2] longest CSR row
  const unsigned* fwd_stats; // nullable: [0] max|X| bits, [2] max|Wcat| bits
  float* scales;             // nullable (tf32 modes): the five scales
  int have_csr;
  // roles B / C
  const float* x; long long x_rows; int F, x_ld; __half* xh; __half* xl; int nb_x;
  const float* w_eff; const float* w_self; int KRA; __half* wh; __half* wl; int nb_w;
};

__device__ __forceinline__ int bw_pow2_for(float bound) {   // bound * 2^e in [2^14, 2^15); 0 for a zero / non-finite bound
  if (!(bound > 0.f) || !isfinite(bound)) return 0;
  const int e = 14 - ilogbf(bound);
  return e < -60 ? -60 : (e > 60 ? 60 : e);
}

__device__ __forceinline__ void bw_split_half2(float a, float b, float s, uint32_t& hi, uint32_t& lo) {
  a *= s; b *= s;
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

constexpr int BWP_THREADS = 256;

__global__ void __launch_bounds__(BWP_THREADS) k_bw_prepare(const __grid_constant__ BwPrepParams p) {
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= p.nb_h) {
    // ---------------- roles B / C: fp16 operand splits (8 consecutive elements per thread) ----------------
    int b = (int)blockIdx.x - p.nb_h;
    if (b < p.nb_x) {
      const float s = ldexpf(1.f, bw_pow2_for(__uint_as_float(__ldg(p.fwd_stats + 0))));
      const int g8 = p.x_ld >> 3;
      const long long n = p.x_rows * g8;
      for (long long i8 = (long long)b * BWP_THREADS + tid; i8 < n; i8 += (long long)p.nb_x * BWP_THREADS) {
        const long long r = i8 / g8;
        const int c = (int)(i8 - r * g8) << 3;
        const float4* src = reinterpret_cast<const float4*>(p.x + r * p.F + c);
        const float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a = c + 4 <= p.F ? __ldg(src) : zz, bq = c + 8 <= p.F ? __ldg(src + 1) : zz;
        uint4 hi, lo;
        bw_split_half2(a.x, a.y, s, hi.x, lo.x); bw_split_half2(a.z, a.w, s, hi.y, lo.y);
        bw_split_half2(bq.x, bq.y, s, hi.z, lo.z); bw_split_half2(bq.z, bq.w, s, hi.w, lo.w);
        *reinterpret_cast<uint4*>(p.xh + i8 * 8) = hi;
        *reinterpret_cast<uint4*>(p.xl + i8 * 8) = lo;
      }
      return;
    }
    b -= p.nb_x;
    const float s = ldexpf(1.f, bw_pow2_for(__uint_as_float(__ldg(p.fwd_stats + 2))));
    const int kcat = p.KRA + p.F, k8n = kcat >> 3;
    const long long n = (long long)p.T * k8n;
    for (long long e8 = (long long)b * BWP_THREADS + tid; e8 < n; e8 += (long long)p.nb_w * BWP_THREADS) {
      const int t = (int)(e8 / k8n), k = (int)(e8 - (long long)t * k8n) << 3;
      const float* src = k < p.KRA ? p.w_eff + (size_t)t * p.KRA + k : p.w_self + (size_t)t * p.F + (k - p.KRA);
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), bq = __ldg(reinterpret_cast<const float4*>(src) + 1);
      uint4 hi, lo;
      bw_split_half2(a.x, a.y, s, hi.x, lo.x); bw_split_half2(a.z, a.w, s, hi.y, lo.y);
      bw_split_half2(bq.x, bq.y, s, hi.z, lo.z); bw_split_half2(bq.z, bq.w, s, hi.w, lo.w);
      const size_t o = (size_t)t * kcat + k;
      *reinterpret_cast<uint4*>(p.wh + o) = hi;
      *reinterpret_cast<uint4*>(p.wl + o) = lo;
    }
    return;
  }
  // ---------------- role A ----------------
  extern __shared__ float4 s_col[];            // [row lanes][T / 4]
  __shared__ unsigned s_max;
  __shared__ bool s_last;
  if (tid == 0) s_max = 0u;
  const int T4 = p.T >> 2;
  const int n_lanes = BWP_THREADS / T4;        // row lanes (>= 1: T <= 1024)
  const int c4 = tid % T4, rl = tid / T4;
  const int r0 = (int)blockIdx.x * p.rows_per_block, r1 = min(p.v_out, r0 + p.rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned m = 0;
  if (rl < n_lanes) {
    for (int r = r0 + rl; r < r1; r += n_lanes) {
      const size_t o = (size_t)r * T4 + c4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.grad_z) + o);
      const float4 zo = __ldg(reinterpret_cast<const float4*>(p.z) + o);
      float4 hv;
      hv.x = g.x * act_grad_from_out(p.act, zo.x); hv.y = g.y * act_grad_from_out(p.act, zo.y);
      hv.z = g.z * act_grad_from_out(p.act, zo.z); hv.w = g.w * act_grad_from_out(p.act, zo.w);
      reinterpret_cast<float4*>(p.h)[o] = hv;
      acc.x += hv.x; acc.y += hv.y; acc.z += hv.z; acc.w += hv.w;
      m = max(m, max(max(__float_as_uint(fabsf(hv.x)), __float_as_uint(fabsf(hv.y))),
                     max(__float_as_uint(fabsf(hv.z)), __float_as_uint(fabsf(hv.w)))));
      if (c4 == 0) {
        const int a = __ldg(p.argmax + r);
        p.rot_sel[r] = p.rot.off[a < 0 ? 0 : (a >= p.n_rot ? p.n_rot - 1 : a)];
      }
    }
    s_col[rl * T4 + c4] = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if ((tid & 31) == 0 && m) atomicMax(&s_max, m);
  if (tid < T4) {      // fixed-order sum over the row lanes
    float4 sum = s_col[tid];
    for (int l = 1; l < n_lanes; ++l) {
      const float4 v = s_col[l * T4 + tid];
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    reinterpret_cast<float4*>(p.bias_part)[(size_t)blockIdx.x * T4 + tid] = sum;
    __threadfence();   // visible to the block that finishes last before this block's ticket is
  }
  __syncthreads();
  if (tid == 0) {
    if (s_max) atomicMax(p.mx + 0, s_max);
    __threadfence();
    s_last = atomicAdd(p.mx + 6, 1u) == (unsigned)p.nb_h - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // d_bias: block partial sums added in a fixed order - row lane l takes blocks l, l + n_lanes, ... (independent
  // loads in flight instead of one dependent chain of nb_h L2 round trips), then the lanes in order
  if (rl < n_lanes) {
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = rl; b < p.nb_h; b += n_lanes) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p.bias_part) + (size_t)b * T4 + c4);
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    s_col[rl * T4 + c4] = sum;
  }
  __syncthreads();
  if (tid < T4) {
    float4 sum = s_col[tid];
    for (int l = 1; l < n_lanes; ++l) {
      const float4 v = s_col[l * T4 + tid];
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    float4* out = reinterpret_cast<float4*>(p.d_bias) + tid;
    if (p.accumulate) {
      const float4 old = *out;
      sum.x += old.x; sum.y += old.y; sum.z += old.z; sum.w += old.w;
    }
    *out = sum;
  }
  if (tid == 0 && p.scales) {
    volatile const unsigned* vm = p.mx;
    const float hm = __uint_as_float(vm[0]);
    const unsigned wbits = p.csr_stats ? __ldg(p.csr_stats + 1) : vm[1];
    const unsigned longest = p.csr_stats ? __ldg(p.csr_stats + 2) : vm[2];
    const unsigned wcbits = p.fwd_stats ? __ldg(p.fwd_stats + 2) : vm[3];
    const unsigned xbits = p.fwd_stats ? __ldg(p.fwd_stats + 0) : vm[4];
    const float gm = p.have_csr ? fmaxf(hm * __uint_as_float(wbits) * (float)longest, hm) : hm;
    const int eg = bw_pow2_for(gm), ew = bw_pow2_for(__uint_as_float(wcbits)), ex = bw_pow2_for(__uint_as_float(xbits));
    p.scales[0] = ldexpf(1.f, eg);
    p.scales[1] = ldexpf(1.f, ew);
    p.scales[2] = ldexpf(1.f, ex);
    p.scales[3] = ldexpf(1.f, -(eg + ew));
    p.scales[4] = ldexpf(1.f, -(eg + ex));
  }
}

// =================================================================================================
// host side
// =================================================================================================
bool tc_bwd_supported(int R, int A, int F, int T) {
  if (F % 32 != 0 || T % 32 != 0) return false;
  // k_bw_build_g keeps the (R*A) x T accumulator of one source vertex in shared memory
  if ((size_t)R * A * T * sizeof(float) > TC_BUILD_G_SMEM_LIMIT) return false;
  int n_t = T <= 256 ? T : 0;
  if (!n_t) {
    for (int c = 256; c >= 32; c -= 32)
      if (T % c == 0) { n_t = c; break; }
  }
  return n_t > 0;
}

static int dw_pick_nt(int T) {
  if (T <= 256) return T;
  for (int c = 256; c >= 32; c -= 32)
    if (T % c == 0) return c;
  return 0;
}

struct BwLayout {
  size_t recs, wcat, ht, partial, bias_part, via_g, total;
  int v_pad, kcat, kcat_pad, vsplit, n_chunks, n_t, bias_blocks, bias_rows;
};

static BwLayout bw_layout(int v_out, int v_src, int R, int A, int F, int T, bool x3) {
  BwLayout L{};
  L.v_pad = (int)ceil_div64(v_out, 128) * 128;
  L.kcat = R * A * F + F;
  L.kcat_pad = (int)ceil_div64(L.kcat, 128) * 128;
  L.n_chunks = L.v_pad / DW_VC;
  L.n_t = dw_pick_nt(T);
  const int64_t units = (int64_t)(L.kcat_pad / 128) * (T / L.n_t);
  // vertex splits: chains of at most ~512 accumulations per TMEM accumulator, then fill the SMs
  int s_min = x3 ? (int)ceil_div64((int64_t)L.n_chunks * 2 * 3, 512) : (int)ceil_div64((int64_t)L.n_chunks * 2, 1024);
  if (s_min < 1) s_min = 1;
  if (s_min > L.n_chunks) s_min = L.n_chunks;
  int best = s_min;
  double best_eff = -1.0;
  for (int s = s_min; s <= s_min + 8 && s <= L.n_chunks; ++s) {
    int64_t ctas = units * s;
    double eff = (double)ctas / (double)(ceil_div64(ctas, 148) * 148);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  L.vsplit = best;
  const int np = x3 ? 2 : 1;
  L.bias_rows = (int)ceil_div64(v_out, 48);
  if (L.bias_rows < 8) L.bias_rows = 8;
  L.bias_blocks = (int)ceil_div64(v_out, L.bias_rows);
  L.recs = tc_recs_bytes(v_out, R, A);
  L.wcat = np * tc_wcat_bytes(R, A, F, T);
  L.ht = align_up((size_t)np * T * L.v_pad * sizeof(float), 256);
  L.partial = align_up((size_t)L.vsplit * T * L.kcat_pad * sizeof(float), 256);
  const size_t bias_a = align_up((size_t)(L.v_pad / 32) * T * sizeof(float), 256);
  const size_t bias_b = align_up((size_t)L.bias_blocks * T * sizeof(float), 256);
  const size_t bias_c = align_up((size_t)296 * T * sizeof(float), 256);   // k_bw_prepare: at most 296 row blocks
  L.bias_part = bias_a > bias_b ? bias_a : bias_b;
  if (bias_c > L.bias_part) L.bias_part = bias_c;
  L.via_g = tc_g_layout_max(v_src, R, A, F, T, x3).total;   // G, the split signal, the GEMM partials (any route)
  const size_t gather_route = L.recs + L.ht + L.partial;
  L.total = L.bias_part + L.wcat + (gather_route > L.via_g ? gather_route : L.via_g) + tc_bwd_scales_bytes() + 1024;
  return L;
}

size_t tc_bwd_workspace_bytes(int v_out, int v_src, int R, int A, int F, int T, int precision) {
  return bw_layout(v_out, v_src, R, A, F, T, precision == GCB_PREC_TF32X3 || precision == GCB_PREC_F16X3).total;
}

// Every buffer layout a backward call can choose for this shape (tf32 or fp16 operands, with or without the
// neighbour term) against what the workspace query provides; the byte counts are recomputed here from the
// layouts' own tile / split counts, independently of tc_g_layout_max.
bool tc_bwd_layout_fits(int v_out, int v_src, int R, int A, int F, int T, int precision) {
  const bool x3 = precision == GCB_PREC_TF32X3 || precision == GCB_PREC_F16X3;
  const BwLayout L = bw_layout(v_out, v_src, R, A, F, T, x3);
  const TcGLayout GM = tc_g_layout_max(v_src, R, A, F, T, x3);
  const size_t fixed = L.bias_part + L.wcat + tc_bwd_scales_bytes() + 512 /* alignment of the base pointer */;
  if (fixed + GM.g_bytes + GM.xsplit_bytes + GM.partial_bytes > L.total) return false;
  for (int route = 0; route < 3; ++route) {
    const bool half = route == 1;
    if (half && !(precision == GCB_PREC_F16X3)) continue;
    const TcGLayout G = route == 2 ? tc_g_layout(v_src, 0, 1, F, T, x3, false) : tc_g_layout(v_src, R, A, F, T, x3, half);
    const int np = x3 ? 2 : 1;
    const size_t g_need = half ? (size_t)2 * G.v_pad * (size_t)(ceil_div64(G.kg, 64) * 64) * 2
                               : (size_t)np * G.v_pad * G.kg * 4;
    const size_t x_need = half ? (size_t)2 * v_src * (size_t)(ceil_div64(F, 64) * 64) * 2 : (size_t)np * v_src * F * 4;
    const size_t p_need = (size_t)G.vsplit * G.m_pad * F * 4;
    if (g_need > GM.g_bytes || x_need > GM.xsplit_bytes || p_need > GM.partial_bytes) return false;
  }
  return true;
}

template <bool X3>
static int launch_dw(const CUtensorMap& mh, const CUtensorMap& ml, const TcDwParams& p, dim3 grid, size_t smem,
                     cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_dw<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(GCB_PROF_BWD_DW, st);
  k_tc_dw<X3><<<grid, DW_THREADS, smem, st>>>(mh, ml, p);
  profile_end(GCB_PROF_BWD_DW, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// ---- centre-only forward (all-zero prior, ConvZero: conv_zero.py:12-15) on tensor cores ---------------------------
// Y[k, i, :] = act(X[k] Wself^T + b) for every rotation i; the pooling picks rotation 0.  The contraction is the
// persistent TMA-fed GEMM of the signal gradient (tc_dx_from_g) with the roles swapped: A = the signal rows (K-major,
// tf32 hi / lo), B = Wself^T [F x T] (MN-major).  The CUDA-core GEMM this replaces took 190 us at the FAUST shape.
__global__ void k_centre_wt(const float* __restrict__ w_self, int T, int F, float* __restrict__ hi, float* __restrict__ lo) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // over [F][T]
  if (e >= F * T) return;
  const int f = e / T, t = e - f * T;
  const float v = w_self[(size_t)t * F + f];
  const float hv = ptx::to_tf32(v);
  hi[e] = hv;
  if (lo) lo[e] = v - hv;
}

__global__ void k_centre_epilogue(const float4* __restrict__ acc, const float4* __restrict__ bias, int v_out, int T4,
                                  int n_rot, int act, float4* __restrict__ y, float4* __restrict__ z,
                                  int32_t* __restrict__ argmax) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)v_out * T4) return;
  const int k = (int)(e / T4), q = (int)(e - (int64_t)k * T4);
  const float4 a = acc[e], b = __ldg(bias + q);
  float4 o;
  o.x = act_apply(act, a.x + b.x); o.y = act_apply(act, a.y + b.y);
  o.z = act_apply(act, a.z + b.z); o.w = act_apply(act, a.w + b.w);
  if (y)
    for (int i = 0; i < n_rot; ++i) y[((size_t)k * n_rot + i) * T4 + q] = o;
  if (z) {
    z[e] = o;
    if (q == 0) argmax[k] = 0;   // equal norms: the first rotation (torch.argmax)
  }
}

size_t tc_centre_fwd_bytes(int v_out, int F, int T) {
  const size_t v_pad = (size_t)ceil_div64(v_out, 128) * 128;
  return align_up(2 * v_pad * F * sizeof(float), 256) + align_up((size_t)2 * F * T * sizeof(float), 256) +
         align_up((size_t)v_out * T * sizeof(float), 256) + 512;
}

bool tc_centre_fwd_supported(int F, int T) { return tc_bwd_supported(0, 1, T, F); }

int tc_centre_fwd(const float* x_rows, const float* w_self, const float* bias, int n_rot, int v_out, int F, int T, int act,
                  int precision, float* y, float* z, int32_t* argmax, void* workspace, size_t workspace_bytes,
                  cudaStream_t st) {
  GCB_REQUIRE(workspace_bytes >= tc_centre_fwd_bytes(v_out, F, T), GCB_EWORKSPACE, "tc_centre_fwd: workspace too small");
  GCB_REQUIRE(((uintptr_t)bias % 16 == 0) && (!y || (uintptr_t)y % 16 == 0) && (!z || (uintptr_t)z % 16 == 0), GCB_EINVAL,
              "tc_centre_fwd: bias, y and z must be 16-byte aligned");
  const bool x3 = precision != GCB_PREC_TF32;
  const int v_pad = (int)ceil_div64(v_out, 128) * 128;
  char* ws = (char*)align_up((size_t)workspace, 256);
  float* a_hi = (float*)ws;
  float* a_lo = x3 ? a_hi + (size_t)v_pad * F : nullptr;
  ws += align_up((size_t)2 * v_pad * F * sizeof(float), 256);
  float* b_hi = (float*)ws;
  float* b_lo = x3 ? b_hi + (size_t)F * T : nullptr;
  ws += align_up((size_t)2 * F * T * sizeof(float), 256);
  float* acc = (float*)ws;
  k_bw_split_rows<<<(unsigned)ceil_div64((int64_t)v_pad * F, 256), 256, 0, st>>>(x_rows, v_out, v_pad, F, a_hi, a_lo);
  GCB_LAUNCH_OK();
  k_centre_wt<<<(unsigned)ceil_div64((int64_t)F * T, 256), 256, 0, st>>>(w_self, T, F, b_hi, b_lo);
  GCB_LAUNCH_OK();
  // out[v, T] = A[v, F] x B[F, T]: the dSignal GEMM with (contraction, columns) = (F, T)
  int rc = tc_dx_from_g(a_hi, a_lo, b_hi, b_lo, v_out, 0, 1, /*columns*/ T, /*contraction*/ F, x3, 0, acc, st);
  if (rc) return rc;
  k_centre_epilogue<<<(unsigned)ceil_div64((int64_t)v_out * (T / 4), 256), 256, 0, st>>>(
      (const float4*)acc, (const float4*)bias, v_out, T / 4, n_rot, act, (float4*)y, (float4*)z, argmax);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int tc_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff, const float* w_self,
                const float* h, const int32_t* rot_sel, const int32_t* csr_row_ptr, const int32_t* csr_entries,
                int v_out, int v_src, int R, int A, int F, int T, int precision, int accumulate,
                const float* interp, float* d_x, float* d_w_eff, float* d_w_self, float* d_bias, void* workspace,
                size_t workspace_bytes, cudaStream_t st, const TcBwdAmp* amp, bool csr_ring_grouped, int stages,
                int dense_n_rot, const int32_t* dense_offsets) {
  // dense_n_rot > 0: a gradient for every rotation, h = [v_out][dense_n_rot][T]; only the route through G serves it
  GCB_REQUIRE(dense_n_rot == 0 || (!amp && w_eff && csr_row_ptr && csr_entries), GCB_EUNSUPPORTED,
              "tc_conv_bwd: the dense backward runs through G (neighbour weights and the CSR inversion required)");
  const int h_rows = dense_n_rot > 0 ? dense_n_rot : 1;   // rows of h per output vertex
  // fp16 operand splits (GCB_PREC_F16X3) serve the route through G; the centre-only and gather-fed routes run
  // the same precision request as 3xTF32
  static const bool no_half = getenv("GCB_BWD_NO_HALF") && atoi(getenv("GCB_BWD_NO_HALF")) != 0;   // A/B measurements
  const bool half = precision == GCB_PREC_F16X3 && w_eff && csr_row_ptr && csr_entries && !no_half;
  const bool x3 = precision == GCB_PREC_TF32X3 || precision == GCB_PREC_F16X3;
  BwLayout L = bw_layout(v_out, v_src, R, A, F, T, x3);
  GCB_REQUIRE(workspace_bytes >= L.total, GCB_EWORKSPACE, "tc_conv_bwd: workspace too small");
  GCB_REQUIRE((uintptr_t)x % 16 == 0, GCB_EINVAL, "tc_conv_bwd: x must be 16-byte aligned");
  const int KRA = R * A * F;
  char* ws = (char*)align_up((size_t)workspace, 256);
  float* bias_part = (float*)ws; ws += L.bias_part;
  float* wc_hi = (float*)ws;
  float* wc_lo = x3 ? (float*)(ws + L.wcat / 2) : nullptr;
  ws += L.wcat;
  void* scale_scratch = ws;
  ws += tc_bwd_scales_bytes();
  int rc;
  // AMP-sparse entry (gcb_conv_amp_bwd): h, rot_sel, d_bias, the fp16 scales and - when the forward's operand maxima
  // came along - both operand splits are produced by ONE launch (k_bw_prepare)
  bool bias_done = false, scales_done = false, x_split_done = false, w_split_done = false;
  const TcGLayout GMX = tc_g_layout_max(v_src, R, A, F, T, x3);
  // stages (route through G only): 1 = everything up to the weight gradients, 2 = the signal gradient from the G, the
  // scales and the weight split a stage-1 call left in the SAME workspace; 3 = both.  A caller that reduces weight
  // gradients over ranks starts its collective between the two calls, beside the dSignal GEMM.
  const bool can_split = w_eff && csr_row_ptr && csr_entries;
  const bool do_w = !can_split || (stages & 1), do_x = !can_split || (stages & 2);
  if (amp) {
    GCB_REQUIRE(T % 4 == 0 && T <= 1024 && ((uintptr_t)amp->grad_z % 16 == 0) && ((uintptr_t)amp->z % 16 == 0) &&
                    ((uintptr_t)d_bias % 16 == 0), GCB_EINVAL,
                "tc_conv_bwd: grad_z / z / d_bias must be 16-byte aligned and T a multiple of 4");
    unsigned* mx = (unsigned*)scale_scratch;
    if (do_w) GCB_CUDA_OK(cudaMemsetAsync(mx, 0, 32, st));
    const bool have_csr = csr_row_ptr && w_eff;
    const bool stats = half && amp->fwd_stats && (!have_csr || amp->csr_stats);
    if (half && !stats && do_w) {   // the maxima the caller did not bring along (mode 0: no scales yet)
      TcMaxJob job{};
      job.mode = 0;
      int n = 0;
      if (have_csr && !amp->csr_stats) {
        job.seg[n++] = {w, (long long)v_out * R * A * 3, 0, 1};
        job.seg[n++] = {csr_row_ptr, (long long)v_src, 1, 2};
      }
      if (!amp->fwd_stats) {
        if (w_eff) job.seg[n++] = {w_eff, (long long)T * R * A * F, 0, 3};
        job.seg[n++] = {w_self, (long long)T * F, 0, 3};
        job.seg[n++] = {x, (long long)v_src * F, 0, 4};
      }
      job.n_seg = n;
      if (n) {
        rc = tc_maxima_scales(job, mx, nullptr, st);   // (slots 1..4; its ticket is word 7, k_bw_prepare's word 6)
        if (rc) return rc;
      }
    }
    BwPrepParams pp{};
    pp.grad_z = amp->grad_z; pp.z = amp->z; pp.argmax = amp->argmax; pp.rot = amp->rot; pp.n_rot = amp->n_rot;
    pp.act = amp->act; pp.v_out = v_out; pp.T = T; pp.h = amp->h; pp.rot_sel = amp->rot_sel;
    pp.nb_h = (int)ceil_div64(v_out, 24);
    if (pp.nb_h > 296) pp.nb_h = 296;
    pp.rows_per_block = (int)ceil_div64(v_out, pp.nb_h);
    pp.nb_h = (int)ceil_div64(v_out, pp.rows_per_block);
    GCB_REQUIRE((size_t)pp.nb_h * T * sizeof(float) <= L.bias_part, GCB_EWORKSPACE, "tc_conv_bwd: bias partial sums do not fit");
    pp.bias_part = bias_part; pp.accumulate = accumulate; pp.d_bias = d_bias;
    pp.mx = mx;
    pp.have_csr = have_csr ? 1 : 0;
    pp.scales = half ? (float*)(mx + 8) : nullptr;
    pp.csr_stats = amp->csr_stats;   // (a null pointer: the maxima are in mx[1..4], from the job above)
    pp.fwd_stats = amp->fwd_stats;
    if (stats && F % 8 == 0) {
      const int x_ld = (int)ceil_div64(F, 64) * 64;
      char* wsx = ws + GMX.g_bytes;   // where the routes below place the split signal
      pp.x = x; pp.x_rows = v_src; pp.F = F; pp.x_ld = x_ld;
      pp.xh = (__half*)wsx; pp.xl = pp.xh + (size_t)v_src * x_ld;
      pp.nb_x = (int)ceil_div64((int64_t)v_src * (x_ld / 8), (int64_t)BWP_THREADS * 4);
      if (pp.nb_x > 148) pp.nb_x = 148;
      x_split_done = true;
      static const bool wsplit_here = !(getenv("GCB_BW_WSPLIT") && atoi(getenv("GCB_BW_WSPLIT")) == 0);   // A/B
      if (w_eff && d_x && wsplit_here) {
        pp.w_eff = w_eff; pp.w_self = w_self; pp.KRA = KRA;
        pp.wh = (__half*)wc_hi; pp.wl = (__half*)wc_lo;
        pp.nb_w = (int)ceil_div64((int64_t)T * ((KRA + F) / 8), (int64_t)BWP_THREADS * 4);
        if (pp.nb_w > 148) pp.nb_w = 148;
        w_split_done = true;
      }
    }
    const int T4 = T / 4;
    const size_t smem = (size_t)(BWP_THREADS / T4) * T4 * sizeof(float4);
    if (do_w) {
      k_bw_prepare<<<pp.nb_h + pp.nb_x + pp.nb_w, BWP_THREADS, smem, st>>>(pp);
      GCB_LAUNCH_OK();
    }
    h = amp->h;
    rot_sel = amp->rot_sel;
    bias_done = true;
    scales_done = half;
  }

  if (!w_eff) {
    // ---- centre-only layer (all-zero prior, ConvZero): G degenerates to its centre slab, h itself ----
    //   d_w_self = h^T X,  d_x = h Wself  (the same two tensor-core GEMMs with R*A = 0),  d_w_eff = 0
    const TcGLayout GL = tc_g_layout(v_src, 0, 1, F, T, x3);
    const TcGLayout GM = tc_g_layout_max(v_src, R, A, F, T, x3);
    float* g_hi = (float*)ws;
    float* g_lo = x3 ? g_hi + (size_t)GL.v_pad * GL.kg : nullptr;
    ws += GM.g_bytes;
    float* x_hi = (float*)ws;
    float* x_lo = x3 ? x_hi + (size_t)v_src * F : nullptr;
    ws += GM.xsplit_bytes;
    float* partial = (float*)ws;
    if (!bias_done) {
      k_bw_bias_partial<<<dim3((unsigned)L.bias_blocks, (unsigned)ceil_div64(T, 32)), dim3(32, 8), 0, st>>>(
          h, v_out, T, L.bias_rows, bias_part);
      GCB_LAUNCH_OK();
      k_bw_bias_final<<<(unsigned)ceil_div64(T, 128), 128, 0, st>>>(bias_part, L.bias_blocks, T, accumulate, d_bias);
      GCB_LAUNCH_OK();
    }
    if (!accumulate) GCB_CUDA_OK(cudaMemsetAsync(d_w_eff, 0, (size_t)T * R * A * F * sizeof(float), st));
    k_bw_split_rows<<<(unsigned)ceil_div64((int64_t)GL.v_pad * T, 256), 256, 0, st>>>(h, v_out, GL.v_pad, T, g_hi, g_lo);
    GCB_LAUNCH_OK();
    rc = tc_dw_from_g(g_hi, g_lo, x, v_src, 0, 1, F, T, x3, accumulate, nullptr, d_w_self, x_hi, x_lo, partial,
                      GM.partial_bytes, st);
    if (rc || !d_x) return rc;
    rc = tc_build_wcat(nullptr, w_self, T, 0, 1, F, wc_hi, wc_lo, st);
    if (rc) return rc;
    return tc_dx_from_g(g_hi, g_lo, wc_hi, wc_lo, v_src, 0, 1, F, T, x3, accumulate, d_x, st);
  }

  if (csr_row_ptr && csr_entries) {
    // ---- both gradients through G: dWcat = G^T X, dSignal = G Wcat (gcb_tc_dx.cu) ----
    const TcGLayout GL = tc_g_layout(v_src, R, A, F, T, x3);
    const TcGLayout GM = tc_g_layout_max(v_src, R, A, F, T, x3);
    float* g_hi = (float*)ws;
    float* g_lo = x3 ? g_hi + (size_t)GL.v_pad * GL.kg : nullptr;
    const float* scales = nullptr;
    if (half) {   // fp16 buffers: rows of kg_ld halves (they fit in the 3xTF32 sizes)
      g_lo = (float*)((__half*)g_hi + (size_t)GL.v_pad * (ceil_div64(GL.kg, 64) * 64));
      if (!scales_done && do_w) {
        rc = tc_bwd_scales(h, w, csr_row_ptr, w_eff, w_self, x, v_out, v_src, R, A, F, T, scale_scratch, st, h_rows);
        if (rc) return rc;
      }
      scales = (const float*)((const unsigned*)scale_scratch + 8);
    }
    ws += GM.g_bytes;
    float* x_hi = (float*)ws;
    float* x_lo = x3 ? x_hi + (size_t)v_src * F : nullptr;
    ws += GM.xsplit_bytes;
    float* partial = (float*)ws;
    if (do_w) {
      if (!bias_done) {   // (dense: the same strips, h_rows times as long)
        k_bw_bias_partial<<<dim3((unsigned)L.bias_blocks, (unsigned)ceil_div64(T, 32)), dim3(32, 8), 0, st>>>(
            h, v_out * h_rows, T, L.bias_rows * h_rows, bias_part);
        GCB_LAUNCH_OK();
        k_bw_bias_final<<<(unsigned)ceil_div64(T, 128), 128, 0, st>>>(bias_part, L.bias_blocks, T, accumulate, d_bias);
        GCB_LAUNCH_OK();
      }
      rc = tc_build_g(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, T, x3, g_hi, g_lo, st, scales, csr_ring_grouped,
                      dense_n_rot, dense_offsets);
      if (rc) return rc;
      rc = tc_dw_from_g(g_hi, g_lo, x, v_src, R, A, F, T, x3, accumulate, d_w_eff, d_w_self, x_hi, x_lo, partial,
                        GM.partial_bytes, st, scales, half && x_split_done);
      if (rc) return rc;
    }
    if (!d_x || !do_x) return GCB_OK;
    if (!(half && w_split_done)) {
      rc = half ? tc_build_wcat_h(w_eff, w_self, T, R, A, F, scales + 1, (__half*)wc_hi, (__half*)wc_lo, st)
                : tc_build_wcat(w_eff, w_self, T, R, A, F, wc_hi, wc_lo, st);
      if (rc) return rc;
    }
    return tc_dx_from_g(g_hi, g_lo, wc_hi, wc_lo, v_src, R, A, F, T, x3, accumulate, d_x, st, scales);
  }

  // ---- no CSR: weight gradients by the gather-fed contraction (dSignal needs the CSR) ----
  GCB_REQUIRE(!d_x, GCB_EINVAL, "tc_conv_bwd: d_x needs the CSR inversion of the barycentric indices");
  uint4* recs = (uint4*)ws; ws += L.recs;
  float* ht_hi = (float*)ws;
  float* ht_lo = x3 ? ht_hi + (size_t)T * L.v_pad : nullptr;
  ws += L.ht;
  float* partial = (float*)ws;
  rc = tc_pack_recs(idx, w, v_out, R, A, recs, st);
  if (rc) return rc;
  {
    dim3 grid((unsigned)(L.v_pad / 32), (unsigned)ceil_div64(T, 32));
    k_bw_transpose_h<<<grid, dim3(32, 8), 0, st>>>(h, v_out, L.v_pad, T, ht_hi, ht_lo, bias_part);
    GCB_LAUNCH_OK();
    if (!bias_done) {
      k_bw_bias_final<<<(unsigned)ceil_div64(T, 128), 128, 0, st>>>(bias_part, L.v_pad / 32, T, accumulate, d_bias);
      GCB_LAUNCH_OK();
    }
  }
  CUtensorMap mh, ml;
  rc = tc_make_tmap_2d(&mh, ht_hi, (uint64_t)L.v_pad, (uint64_t)T, (uint64_t)L.v_pad * 4, 32, (uint32_t)L.n_t);
  if (rc) return rc;
  rc = tc_make_tmap_2d(&ml, x3 ? ht_lo : ht_hi, (uint64_t)L.v_pad, (uint64_t)T, (uint64_t)L.v_pad * 4, 32,
                       (uint32_t)L.n_t);
  if (rc) return rc;
  TcDwParams p;
  p.x = x; p.recs = recs; p.rot_sel = rot_sel; p.partial = partial; p.interp = interp;
  p.v_out = v_out; p.v_pad = L.v_pad; p.R = R; p.A = A; p.F = F; p.T = T;
  p.kcat = L.kcat;
  p.kcat_pad = L.kcat_pad; p.n_t = L.n_t; p.vsplit = L.vsplit; p.n_chunks = L.n_chunks;
  int cols = 32;
  while (cols < 2 * L.n_t) cols <<= 1;
  p.tmem_cols = cols;
  const int np = x3 ? 2 : 1;
  const size_t stage_bytes = (size_t)np * (DW_A_BYTES + (size_t)L.n_t * 128);
  const size_t fixed = 1024 + ((size_t)((L.n_chunks + L.vsplit - 1) / L.vsplit + 1) * DW_VC * 4 + 16) + 512;
  int ns = (int)((BW_SMEM_LIMIT - fixed) / stage_bytes);
  if (ns > 8) ns = 8;
  GCB_REQUIRE(ns >= 2, GCB_EUNSUPPORTED, "tc_conv_bwd: dW stage does not fit in shared memory");
  p.ns = ns;
  const size_t smem = fixed + (size_t)ns * stage_bytes;
  dim3 grid((unsigned)(L.kcat_pad / 128), (unsigned)L.vsplit, (unsigned)(T / L.n_t));
  rc = x3 ? launch_dw<true>(mh, ml, p, grid, smem, st) : launch_dw<false>(mh, ml, p, grid, smem, st);
  if (rc) return rc;
  k_tc_dw_finish<<<(unsigned)ceil_div64((int64_t)T * L.kcat, 256), 256, 0, st>>>(
      partial, L.vsplit, T, L.kcat, L.kcat_pad, KRA, F, accumulate, d_w_eff, d_w_self);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

}  // namespace gcb
