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
//   k_tc_di   unrotated dI:  D[k, kk] = sum_t h[k,t] * Weff[t, kk]   (M = 128 vertices, N = 128 kk,
//             K = templates).  Persistent CTAs, TMA-fed (h K-major, Weff rows as an MN-major operand),
//             double-buffered TMEM accumulators, epilogue warps stream the 128x128 tile to HBM.  This
//             stage is bound by writing V*R*A*F floats; the per-vertex rotation is applied later as an
//             index by the CSR gather-reduce (gcb_pool.cu) that forms dSignal without atomics.
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
                                 float* __restrict__ h_hi, float* __restrict__ h_lo,
                                 float* __restrict__ bias_part) {
  // ht[t][v] (v padded with zeros) for k_tc_dw;  h_hi/h_lo [v_pad][T] for k_tc_di;
  // bias_part[blockIdx.x][t] = sum of this block's 32 rows (first stage of d_bias)
  __shared__ float tile[32][33];
  const int v0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int v = v0 + i, t = t0 + threadIdx.x;
    float val = (v < v_out && t < T) ? h[(size_t)v * T + t] : 0.f;
    tile[i][threadIdx.x] = val;
    if (t < T && v < v_pad) {
      float hi = ptx::to_tf32(val);
      h_hi[(size_t)v * T + t] = hi;
      if (h_lo) h_lo[(size_t)v * T + t] = val - hi;
    }
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

// d_bias[t] (+)= sum over the per-block partial sums written by k_bw_transpose_h (fixed order)
__global__ void k_bw_bias_final(const float* __restrict__ part, int nblocks, int T, int accumulate,
                                float* __restrict__ d_bias) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float acc = 0.f;
  for (int b = 0; b < nblocks; ++b) acc += part[(size_t)b * T + t];
  d_bias[t] = accumulate ? d_bias[t] + acc : acc;
}

// =================================================================================================
// k_tc_di
// =================================================================================================
constexpr int DI_EPI_WARPS = 4;
constexpr int DI_THREADS = (2 + DI_EPI_WARPS) * 32;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int DI_BN = 128;
constexpr int DI_STAGE_A = 16 * 1024;   // 128 vertices x 32 templates
constexpr int DI_STAGE_B = 16 * 1024;   // 32 templates x 128 kk  (4 MN-major atoms columns)

struct TcDiParams {
  float* d_full;    // [v_out][KRA]
  int v_out, KRA, T, n_mtiles, n_ntiles, ns;
};

template <bool X3>
__global__ void __launch_bounds__(DI_THREADS, 1)
k_tc_di(const __grid_constant__ CUtensorMap tmap_h_hi, const __grid_constant__ CUtensorMap tmap_h_lo,
        const __grid_constant__ CUtensorMap tmap_w_hi, const __grid_constant__ CUtensorMap tmap_w_lo,
        const TcDiParams p) {
  constexpr int NPROD = X3 ? 2 : 1;
  constexpr int STAGE_BYTES = NPROD * (DI_STAGE_A + DI_STAGE_B);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + (uint32_t)(p.ns * STAGE_BYTES);
  const uint32_t full = bar_base, empty = full + 8u * p.ns;
  const uint32_t tm_full = empty + 8u * p.ns, tm_empty = tm_full + 16u;
  const uint32_t tmem_slot = tm_empty + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_mtiles * p.n_ntiles;
  const int n_kc = p.T / 32;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.ns; ++s) {
      ptx::mbar_init(full + 8u * s, 1);
      ptx::mbar_init(empty + 8u * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tm_full + 8u * b, 1);
      ptx::mbar_init(tm_empty + 8u * b, DI_EPI_WARPS);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tensormap(&tmap_h_hi);
    ptx::prefetch_tensormap(&tmap_w_hi);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * DI_BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ============ TMA producer ============
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 1u;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
        for (int kc = 0; kc < n_kc; ++kc) {
          ptx::mbar_wait(empty + 8u * stage, par);
          const uint32_t a_dst = smem_base + (uint32_t)(stage * STAGE_BYTES);
          const uint32_t b_dst = a_dst + NPROD * DI_STAGE_A;
          ptx::mbar_arrive_expect_tx(full + 8u * stage, (uint32_t)STAGE_BYTES);
          ptx::tma_load_2d(a_dst, &tmap_h_hi, full + 8u * stage, kc * 32, mt * 128);
          if (X3) ptx::tma_load_2d(a_dst + DI_STAGE_A, &tmap_h_lo, full + 8u * stage, kc * 32, mt * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ptx::tma_load_2d(b_dst + j * 4096, &tmap_w_hi, full + 8u * stage, nt * DI_BN + 32 * j, kc * 32);
            if (X3)
              ptx::tma_load_2d(b_dst + DI_STAGE_B + j * 4096, &tmap_w_lo, full + 8u * stage, nt * DI_BN + 32 * j,
                               kc * 32);
          }
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============ MMA issuer ============
    if (lane == 0) {
      // A K-major (h), B MN-major (Weff rows: kk contiguous) -> bit 16
      const uint32_t idesc = ptx::make_idesc_tf32(128, DI_BN) | (1u << 16);
      int stage = 0;
      uint32_t par = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(tm_empty + 8u * buf, (((uint32_t)(it >> 1)) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * DI_BN);
        uint32_t accum = 0;
        for (int kc = 0; kc < n_kc; ++kc) {
          ptx::mbar_wait(full + 8u * stage, par);
          ptx::tc_fence_after();
          const uint32_t a_hi = smem_base + (uint32_t)(stage * STAGE_BYTES);
          const uint32_t b_hi = a_hi + NPROD * DI_STAGE_A;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t da = ptx::make_kmajor_desc<128>(a_hi + kb * 32);
            const uint64_t db = make_mnmajor_desc(b_hi + kb * 1024, 4096, 512);
            ptx::umma_tf32(d, da, db, idesc, accum);
            accum = 1u;
            if (X3) {
              ptx::umma_tf32(d, da, make_mnmajor_desc(b_hi + DI_STAGE_B + kb * 1024, 4096, 512), idesc, 1u);
              ptx::umma_tf32(d, ptx::make_kmajor_desc<128>(a_hi + DI_STAGE_A + kb * 32), db, idesc, 1u);
            }
          }
          ptx::umma_commit(empty + 8u * stage);
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
        ptx::umma_commit(tm_full + 8u * buf);
      }
    }
  } else {
    // ============ epilogue: TMEM -> HBM ============
    const int q = warp & 3;   // TMEM lane quadrant this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
      const int buf = it & 1;
      ptx::mbar_wait(tm_full + 8u * buf, ((uint32_t)(it >> 1)) & 1u);
      ptx::tc_fence_after();
      const int row = mt * 128 + 32 * q + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * DI_BN);
      float* out = p.d_full + (size_t)row * p.KRA + nt * DI_BN;
      for (int c0 = 0; c0 < DI_BN; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
        if (row < p.v_out && nt * DI_BN + c0 < p.KRA) {
#pragma unroll
          for (int u = 0; u < 16; u += 4)
            *reinterpret_cast<uint4*>(out + c0 + u) = make_uint4(r[u], r[u + 1], r[u + 2], r[u + 3]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tm_empty + 8u * buf);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 2 * DI_BN);
  }
}

// =================================================================================================
// host side
// =================================================================================================
bool tc_bwd_supported(int R, int A, int F, int T) {
  (void)R; (void)A;
  if (F % 32 != 0 || T % 32 != 0) return false;
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
  size_t recs, wcat, ht, hrow, partial, bias_part, d_full, total;
  int v_pad, v_pad32, kcat, kcat_pad, vsplit, n_chunks, n_t;
};

static BwLayout bw_layout(int v_out, int v_src, int R, int A, int F, int T, bool x3) {
  BwLayout L{};
  L.v_pad = (int)ceil_div64(v_out, 128) * 128;
  L.v_pad32 = L.v_pad;  // multiple of 128 is a multiple of 32
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
  L.recs = tc_recs_bytes(v_out, R, A);
  L.wcat = np * tc_wcat_bytes(R, A, F, T);
  L.ht = align_up((size_t)np * T * L.v_pad * sizeof(float), 256);
  L.hrow = align_up((size_t)np * L.v_pad * T * sizeof(float), 256);
  L.partial = align_up((size_t)L.vsplit * T * L.kcat_pad * sizeof(float), 256);
  L.bias_part = align_up((size_t)(L.v_pad / 32) * T * sizeof(float), 256);
  L.d_full = tc_dx_workspace_bytes(v_src, R, A, T, x3);   // the G operand of the dSignal GEMM
  L.total = L.recs + L.wcat + L.ht + L.hrow + L.partial + L.bias_part + L.d_full + 1024;
  return L;
}

size_t tc_bwd_workspace_bytes(int v_out, int v_src, int R, int A, int F, int T, int precision) {
  return bw_layout(v_out, v_src, R, A, F, T, precision == GCB_PREC_TF32X3).total;
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

template <bool X3>
static int launch_di(const CUtensorMap& hh, const CUtensorMap& hl, const CUtensorMap& wh, const CUtensorMap& wl,
                     const TcDiParams& p, int grid, size_t smem, cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_di<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(GCB_PROF_BWD_DI, st);
  k_tc_di<X3><<<grid, DI_THREADS, smem, st>>>(hh, hl, wh, wl, p);
  profile_end(GCB_PROF_BWD_DI, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int tc_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff, const float* w_self,
                const float* h, const int32_t* rot_sel, const int32_t* csr_row_ptr, const int32_t* csr_entries,
                int v_out, int v_src, int R, int A, int F, int T, int precision, int accumulate,
                const float* interp, float* d_x, float* d_w_eff, float* d_w_self, float* d_bias, void* workspace,
                size_t workspace_bytes, cudaStream_t st) {
  const bool x3 = precision == GCB_PREC_TF32X3;
  const int np = x3 ? 2 : 1;
  BwLayout L = bw_layout(v_out, v_src, R, A, F, T, x3);
  GCB_REQUIRE(workspace_bytes >= L.total, GCB_EWORKSPACE, "tc_conv_bwd: workspace too small");
  GCB_REQUIRE((uintptr_t)x % 16 == 0, GCB_EINVAL, "tc_conv_bwd: x must be 16-byte aligned");
  const int KRA = R * A * F;
  char* ws = (char*)align_up((size_t)workspace, 256);
  uint4* recs = (uint4*)ws; ws += L.recs;
  float* wc_hi = (float*)ws;
  float* wc_lo = x3 ? (float*)(ws + L.wcat / 2) : nullptr;
  ws += L.wcat;
  float* ht_hi = (float*)ws;
  float* ht_lo = x3 ? ht_hi + (size_t)T * L.v_pad : nullptr;
  ws += L.ht;
  float* hr_hi = (float*)ws;
  float* hr_lo = x3 ? hr_hi + (size_t)L.v_pad * T : nullptr;
  ws += L.hrow;
  float* partial = (float*)ws; ws += L.partial;
  float* bias_part = (float*)ws; ws += L.bias_part;
  float* d_full = (float*)ws;

  int rc = tc_pack_recs(idx, w, v_out, R, A, recs, st);
  if (rc) return rc;
  {
    dim3 grid((unsigned)(L.v_pad / 32), (unsigned)ceil_div64(T, 32));
    k_bw_transpose_h<<<grid, dim3(32, 8), 0, st>>>(h, v_out, L.v_pad, T, ht_hi, ht_lo, hr_hi, hr_lo, bias_part);
    GCB_LAUNCH_OK();
    k_bw_bias_final<<<(unsigned)ceil_div64(T, 128), 128, 0, st>>>(bias_part, L.v_pad / 32, T, accumulate, d_bias);
    GCB_LAUNCH_OK();
  }
  // ---- dW (neighbour + centre) ----
  {
    CUtensorMap mh, ml;
    rc = tc_make_tmap_2d(&mh, ht_hi, (uint64_t)L.v_pad, (uint64_t)T, (uint64_t)L.v_pad * 4, 32, (uint32_t)L.n_t);
    if (rc) return rc;
    rc = tc_make_tmap_2d(&ml, x3 ? ht_lo : ht_hi, (uint64_t)L.v_pad, (uint64_t)T, (uint64_t)L.v_pad * 4, 32,
                         (uint32_t)L.n_t);
    if (rc) return rc;
    TcDwParams p;
    p.x = x; p.recs = recs; p.rot_sel = rot_sel; p.partial = partial; p.interp = interp;
    p.v_out = v_out; p.v_pad = L.v_pad; p.R = R; p.A = A; p.F = F; p.T = T;
    p.kcat = w_eff ? L.kcat : L.kcat;  // centre-only layers never reach this path
    p.kcat_pad = L.kcat_pad; p.n_t = L.n_t; p.vsplit = L.vsplit; p.n_chunks = L.n_chunks;
    int cols = 32;
    while (cols < 2 * L.n_t) cols <<= 1;
    p.tmem_cols = cols;
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
  }
  if (!d_x) return GCB_OK;
  // ---- dSignal: G (entries regrouped by rotated slot) x weights on tensor cores (gcb_tc_dx.cu) ----
  rc = tc_build_wcat(w_eff, w_self, T, R, A, F, wc_hi, wc_lo, st);
  if (rc) return rc;
  return tc_dx_via_g(h, w, rot_sel, csr_row_ptr, csr_entries, wc_hi, wc_lo, v_out, v_src, R, A, F, T, x3, accumulate,
                     d_x, d_full, L.d_full, st);
}

}  // namespace gcb
