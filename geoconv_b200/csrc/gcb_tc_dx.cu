// dSignal on tensor cores without materialising dI.
//
//   dX[v,f] = sum_{slots s->v} w[s] * sum_t h[k_s,t] * Weff[t, x_s, (y_s+rot[k_s])%A, f]  +  h[v]·Wself[:,f]
// is regrouped by the ROTATED slot s' = (x, (y+rot)%A) of every entry:
//   G[v, s', t] = sum_{entries e of v with rotated slot s'} w_e * h[k_e, t]           (k_bw_build_g)
//   dX[v, f]    = sum_{s',t} G[v,s',t] * Weff[t,s',f] + sum_t h[v,t] * Wself[t,f]     (k_tc_gemm)
// G is V x (R*A+1)*T floats (109 MB at the FAUST shape) instead of the 600 MB dI that would have to be
// written once and re-read three times; the second line is a plain dense GEMM
// [V x (R*A+1)T] x [(R*A+1)T x F] whose B operand is the weight tensor itself read as MN-major tiles.
//   k_bw_build_g : one block per source vertex, CSR entry list decoded cooperatively, thread t owns
//                  column t of the (R*A) x T accumulator in shared memory; sequential over the
//                  vertex's entries (fixed order, no atomics); writes tf32 hi (+ lo) rows of G.
//   k_tc_gemm_tn : d[Weff | Wself] = G^T X.  The same regrouping makes the weight gradient a dense GEMM
//                  too:  dWeff[t,s',f] = sum_v G[v,s',t] * X[v,f]  (and the centre slab of G, h itself,
//                  gives dWself), M = (s',t), N = f, K = source vertices.  Both operands are row-major
//                  with the contraction index as the row, i.e. MN-major for the tensor core: 3-D TMA
//                  boxes deposit them in the SWIZZLE_128B_BASE32B layout, no transpose, no gather.
//                  It replaces a gather-fed contraction that re-read 1.8 GB of signal rows through L1.
//   k_tc_gemm    : persistent tcgen05 GEMM, 128 x 128 tiles, TMA-fed (A K-major, B MN-major
//                  SWIZZLE_128B_BASE32B), four TMEM accumulators per tile fed round-robin by the K
//                  chunks and summed in fp32 by the epilogue (the tensor core accumulates with
//                  truncation, so short chains are what keeps the 3xTF32 mode within 1e-5).
#include <cuda.h>

#include "gcb_common.cuh"
#include "gcb_ptx.cuh"
#include "gcb_tc.cuh"

namespace gcb {

constexpr int DX_SMEM_LIMIT = 227 * 1024;
constexpr int G_CHUNK = 128;

__device__ __forceinline__ uint64_t dx_mnmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

// ---- G[v, s', t] ----------------------------------------------------------------------------------
__global__ void k_bw_build_g(const float* __restrict__ h, const float* __restrict__ w,
                             const int32_t* __restrict__ rot_sel, const int32_t* __restrict__ row_ptr,
                             const int32_t* __restrict__ entries, int v_out, int v_src, int R, int A, int T,
                             int kg /* (R*A+1)*T */, float* __restrict__ g_hi, float* __restrict__ g_lo) {
  extern __shared__ float s_acc[];                       // [RA][T]
  __shared__ int s_k[G_CHUNK];
  __shared__ int s_slot[G_CHUNK];
  __shared__ float s_w[G_CHUNK];
  const int v = blockIdx.x;
  const int RA = R * A;
  const int t = threadIdx.x;
  float* out_hi = g_hi + (size_t)v * kg;
  float* out_lo = g_lo ? g_lo + (size_t)v * kg : nullptr;
  if (v >= v_src) {                                      // padding rows of the GEMM's A operand
    for (int i = threadIdx.x; i < kg; i += blockDim.x) {
      out_hi[i] = 0.f;
      if (out_lo) out_lo[i] = 0.f;
    }
    return;
  }
  for (int i = threadIdx.x; i < RA * T; i += blockDim.x) s_acc[i] = 0.f;
  const int beg = row_ptr[v], end = row_ptr[v + 1];
  for (int base = beg; base < end; base += G_CHUNK) {
    const int n = min(G_CHUNK, end - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int s = entries[base + i];
      const int cell = s / 3;
      const int k = cell / RA;
      const int slot = cell - k * RA;
      const int xr = slot / A, ya = slot - xr * A;
      int yr = ya + rot_sel[k];
      yr -= (yr >= A) ? A : 0;
      s_k[i] = k;
      s_slot[i] = xr * A + yr;
      s_w[i] = w[s];
    }
    __syncthreads();
    if (t < T) {
      int i = 0;
      for (; i + 8 <= n; i += 8) {
        float hv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) hv[u] = __ldg(h + (size_t)s_k[i + u] * T + t);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float* a = s_acc + s_slot[i + u] * T + t;
          *a = fmaf(s_w[i + u], hv[u], *a);
        }
      }
      for (; i < n; ++i) {
        float* a = s_acc + s_slot[i] * T + t;
        *a = fmaf(s_w[i], __ldg(h + (size_t)s_k[i] * T + t), *a);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RA * T; i += blockDim.x) {
    const float val = s_acc[i];
    const float hi = ptx::to_tf32(val);
    out_hi[i] = hi;
    if (out_lo) out_lo[i] = val - hi;
  }
  for (int i = threadIdx.x; i < T; i += blockDim.x) {   // centre slab: h[v] itself (zero for halo rows)
    const float val = v < v_out ? h[(size_t)v * T + i] : 0.f;
    const float hi = ptx::to_tf32(val);
    out_hi[RA * T + i] = hi;
    if (out_lo) out_lo[RA * T + i] = val - hi;
  }
}

// ---- persistent TMA-fed tcgen05 GEMM ---------------------------------------------------------------
constexpr int GM_EPI_WARPS = 4;
constexpr int GM_THREADS = (2 + GM_EPI_WARPS) * 32;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int GM_BN = 128;
constexpr int GM_STAGE_A = 16 * 1024;   // 128 rows x 32 k
constexpr int GM_STAGE_B = 16 * 1024;   // 32 k x 128 n (four MN-major column blocks)

struct TcGemmParams {
  float* out;
  int ld_out, rows_valid, cols_valid, accumulate;
  int n_mtiles, n_ntiles, ns;
  int n_kc;               // K chunks of 32
  int kc_per_group;       // B tile of chunk kc: rows (kc % kc_per_group)*32, cols (kc / kc_per_group)*group_stride + n0
  int group_stride;
};

template <bool X3>
__global__ void __launch_bounds__(GM_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
          const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
          const TcGemmParams p) {
  constexpr int NPROD = X3 ? 2 : 1;
  constexpr int STAGE_BYTES = NPROD * (GM_STAGE_A + GM_STAGE_B);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + (uint32_t)(p.ns * STAGE_BYTES);
  const uint32_t full = bar_base, empty = full + 8u * p.ns;
  const uint32_t tm_full = empty + 8u * p.ns, tm_empty = tm_full + 16u;
  const uint32_t tmem_slot = tm_empty + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_mtiles * p.n_ntiles;
  const int nacc = p.n_kc >= 4 ? 4 : 1;   // K chunks round-robin over this many accumulators

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.ns; ++s) {
      ptx::mbar_init(full + 8u * s, 1);
      ptx::mbar_init(empty + 8u * s, 1);
    }
    ptx::mbar_init(tm_full, 1);
    ptx::mbar_init(tm_empty, GM_EPI_WARPS);
    ptx::fence_mbar_init();
    ptx::prefetch_tensormap(&tmap_a_hi);
    ptx::prefetch_tensormap(&tmap_b_hi);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 4 * GM_BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 1u;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
        int grp = 0, in_grp = 0;
        for (int kc = 0; kc < p.n_kc; ++kc) {
          ptx::mbar_wait(empty + 8u * stage, par);
          const uint32_t a_dst = smem_base + (uint32_t)(stage * STAGE_BYTES);
          const uint32_t b_dst = a_dst + NPROD * GM_STAGE_A;
          const int b_row = in_grp * 32, b_col = grp * p.group_stride + nt * GM_BN;
          ptx::mbar_arrive_expect_tx(full + 8u * stage, (uint32_t)STAGE_BYTES);
          ptx::tma_load_2d(a_dst, &tmap_a_hi, full + 8u * stage, kc * 32, mt * 128);
          if (X3) ptx::tma_load_2d(a_dst + GM_STAGE_A, &tmap_a_lo, full + 8u * stage, kc * 32, mt * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ptx::tma_load_2d(b_dst + j * 4096, &tmap_b_hi, full + 8u * stage, b_col + 32 * j, b_row);
            if (X3) ptx::tma_load_2d(b_dst + GM_STAGE_B + j * 4096, &tmap_b_lo, full + 8u * stage, b_col + 32 * j, b_row);
          }
          if (++in_grp == p.kc_per_group) { in_grp = 0; ++grp; }
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_tf32(128, GM_BN) | (1u << 16);   // B MN-major
      int stage = 0;
      uint32_t par = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        ptx::mbar_wait(tm_empty, (((uint32_t)it) & 1u) ^ 1u);
        ptx::tc_fence_after();
        uint32_t started = 0;   // bit a: accumulator a already holds a partial sum
        for (int kc = 0; kc < p.n_kc; ++kc) {
          const int acc = nacc == 4 ? (kc & 3) : 0;
          const uint32_t d = tmem_base + (uint32_t)(acc * GM_BN);
          ptx::mbar_wait(full + 8u * stage, par);
          ptx::tc_fence_after();
          const uint32_t a_hi = smem_base + (uint32_t)(stage * STAGE_BYTES);
          const uint32_t b_hi = a_hi + NPROD * GM_STAGE_A;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t da = ptx::make_kmajor_desc<128>(a_hi + kb * 32);
            const uint64_t db = dx_mnmajor_desc(b_hi + kb * 1024, 4096, 512);
            ptx::umma_tf32(d, da, db, idesc, (started >> acc) & 1u);
            started |= 1u << acc;
            if (X3) {
              ptx::umma_tf32(d, da, dx_mnmajor_desc(b_hi + GM_STAGE_B + kb * 1024, 4096, 512), idesc, 1u);
              ptx::umma_tf32(d, ptx::make_kmajor_desc<128>(a_hi + GM_STAGE_A + kb * 32), db, idesc, 1u);
            }
          }
          ptx::umma_commit(empty + 8u * stage);
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
        ptx::umma_commit(tm_full);
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
      ptx::mbar_wait(tm_full, ((uint32_t)it) & 1u);
      ptx::tc_fence_after();
      const int row = mt * 128 + 32 * q + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
      float* out = p.out + (size_t)row * p.ld_out + nt * GM_BN;
      for (int c0 = 0; c0 < GM_BN; c0 += 16) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r0);
        if (nacc == 4) {
          ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(GM_BN + c0), r1);
          ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(2 * GM_BN + c0), r2);
          ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(3 * GM_BN + c0), r3);
        }
        ptx::tmem_ld_wait();
        if (row < p.rows_valid && nt * GM_BN + c0 < p.cols_valid) {
#pragma unroll
          for (int u = 0; u < 16; u += 4) {
            float4 o;
            o.x = __uint_as_float(r0[u]); o.y = __uint_as_float(r0[u + 1]);
            o.z = __uint_as_float(r0[u + 2]); o.w = __uint_as_float(r0[u + 3]);
            if (nacc == 4) {
              o.x = (o.x + __uint_as_float(r1[u])) + (__uint_as_float(r2[u]) + __uint_as_float(r3[u]));
              o.y = (o.y + __uint_as_float(r1[u + 1])) + (__uint_as_float(r2[u + 1]) + __uint_as_float(r3[u + 1]));
              o.z = (o.z + __uint_as_float(r1[u + 2])) + (__uint_as_float(r2[u + 2]) + __uint_as_float(r3[u + 2]));
              o.w = (o.w + __uint_as_float(r1[u + 3])) + (__uint_as_float(r2[u + 3]) + __uint_as_float(r3[u + 3]));
            }
            float4* dst = reinterpret_cast<float4*>(out + c0 + u);
            if (p.accumulate) {
              const float4 old = *dst;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *dst = o;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tm_empty);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 4 * GM_BN);
  }
}

template <bool X3>
static int launch_gemm(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                       const TcGemmParams& p, int grid, size_t smem, cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(GCB_PROF_BWD_DI, st);
  k_tc_gemm<X3><<<grid, GM_THREADS, smem, st>>>(ah, al, bh, bl, p);
  profile_end(GCB_PROF_BWD_DI, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// ---- dWcat = G^T X : both operands MN-major ---------------------------------------------------------
constexpr int TN_KCH = 16;                     // contraction rows (source vertices) per stage
constexpr int TN_STAGE_A = 128 * TN_KCH * 4;   // 8 KB: four 32-row atoms of M
constexpr int TN_THREADS = (2 + GM_EPI_WARPS) * 32;

struct TcGemmTnParams {
  float* partial;    // [vsplit][m_pad][ld]
  int ld, n_cols;    // row length of the output (= columns of X)
  int m_pad, n_mtiles, n_ntiles, vsplit;
  int bn;            // columns per N tile (multiple of 32, <= 256); the last tile may be narrower
  int n_kch;         // K chunks of TN_KCH rows in total
  int ns, nacc;
};

template <bool X3>
__global__ void __launch_bounds__(TN_THREADS, 1)
k_tc_gemm_tn(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
             const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
             const TcGemmTnParams p) {
  constexpr int NPROD = X3 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int stage_b = p.bn * TN_KCH * 4;
  const int stage_bytes = NPROD * (TN_STAGE_A + stage_b);
  const uint32_t bar_base = smem_base + (uint32_t)(p.ns * stage_bytes);
  const uint32_t full = bar_base, empty = full + 8u * p.ns;
  const uint32_t tm_full = empty + 8u * p.ns, tm_empty = tm_full + 8u;
  const uint32_t tmem_slot = tm_empty + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_mtiles * p.vsplit * p.n_ntiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.ns; ++s) {
      ptx::mbar_init(full + 8u * s, 1);
      ptx::mbar_init(empty + 8u * s, 1);
    }
    ptx::mbar_init(tm_full, 1);
    ptx::mbar_init(tm_empty, GM_EPI_WARPS);
    ptx::fence_mbar_init();
    ptx::prefetch_tensormap(&tmap_a_hi);
    ptx::prefetch_tensormap(&tmap_b_hi);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // tile -> (M tile, vertex split, N tile): the N tiles of one (M tile, split) are neighbours, so the
  // CTAs that run them at the same time share that panel of G in L2 (G hi + lo exceeds the L2)
  auto decode = [&](int tile, int& mt, int& vs, int& nt) {
    nt = tile % p.n_ntiles;
    const int rest = tile / p.n_ntiles;
    vs = rest % p.vsplit;
    mt = rest / p.vsplit;
  };
  auto k_range = [&](int vs, int& kb, int& ke) {
    kb = (int)(((int64_t)p.n_kch * vs) / p.vsplit);
    ke = (int)(((int64_t)p.n_kch * (vs + 1)) / p.vsplit);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 1u;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mt, vs, nt, kb, ke;
        decode(tile, mt, vs, nt);
        k_range(vs, kb, ke);
        for (int kc = kb; kc < ke; ++kc) {
          ptx::mbar_wait(empty + 8u * stage, par);
          const uint32_t a_dst = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_dst = a_dst + NPROD * TN_STAGE_A;
          ptx::mbar_arrive_expect_tx(full + 8u * stage, (uint32_t)stage_bytes);
          ptx::tma_load_3d(a_dst, &tmap_a_hi, full + 8u * stage, 0, kc * TN_KCH, mt * 4);
          if (X3) ptx::tma_load_3d(a_dst + TN_STAGE_A, &tmap_a_lo, full + 8u * stage, 0, kc * TN_KCH, mt * 4);
          ptx::tma_load_3d(b_dst, &tmap_b_hi, full + 8u * stage, 0, kc * TN_KCH, nt * (p.bn / 32));
          if (X3) ptx::tma_load_3d(b_dst + stage_b, &tmap_b_lo, full + 8u * stage, 0, kc * TN_KCH, nt * (p.bn / 32));
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        int mt, vs, nt, kb, ke;
        decode(tile, mt, vs, nt);
        k_range(vs, kb, ke);
        const int n_valid = min(p.bn, p.n_cols - nt * p.bn);
        const uint32_t idesc = ptx::make_idesc_tf32(128, n_valid) | (1u << 15) | (1u << 16);   // A and B MN-major
        ptx::mbar_wait(tm_empty, (((uint32_t)it) & 1u) ^ 1u);
        ptx::tc_fence_after();
        uint32_t started = 0;   // bit a: accumulator a already holds a partial sum
        for (int kc = kb; kc < ke; ++kc) {
          const int acc = (kc - kb) & (p.nacc - 1);
          const uint32_t d = tmem_base + (uint32_t)(acc * p.bn);
          ptx::mbar_wait(full + 8u * stage, par);
          ptx::tc_fence_after();
          const uint32_t a_hi = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_hi = a_hi + NPROD * TN_STAGE_A;
#pragma unroll
          for (int k8 = 0; k8 < TN_KCH / 8; ++k8) {
            const uint64_t da = dx_mnmajor_desc(a_hi + k8 * 1024, TN_KCH * 128, 512);
            const uint64_t db = dx_mnmajor_desc(b_hi + k8 * 1024, TN_KCH * 128, 512);
            ptx::umma_tf32(d, da, db, idesc, (started >> acc) & 1u);
            started |= 1u << acc;
            if (X3) {
              ptx::umma_tf32(d, da, dx_mnmajor_desc(b_hi + stage_b + k8 * 1024, TN_KCH * 128, 512), idesc, 1u);
              ptx::umma_tf32(d, dx_mnmajor_desc(a_hi + TN_STAGE_A + k8 * 1024, TN_KCH * 128, 512), db, idesc, 1u);
            }
          }
          ptx::umma_commit(empty + 8u * stage);
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
        ptx::umma_commit(tm_full);
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      int mt, vs, nt, kb, ke;
      decode(tile, mt, vs, nt);
      k_range(vs, kb, ke);
      const int n_valid = min(p.bn, p.n_cols - nt * p.bn);
      const int used = min(p.nacc, ke - kb);   // accumulators that received at least one chunk
      ptx::mbar_wait(tm_full, ((uint32_t)it) & 1u);
      ptx::tc_fence_after();
      const int row = mt * 128 + 32 * q + lane;   // < m_pad: the partial buffer is padded
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
      float* out = p.partial + ((size_t)vs * p.m_pad + row) * p.ld + nt * p.bn;
      for (int c0 = 0; c0 < n_valid; c0 += 16) {
        uint32_t r[16];
        float o[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) o[u] = 0.f;
        for (int a = 0; a < used; ++a) {
          ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(a * p.bn + c0), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] += __uint_as_float(r[u]);
        }
#pragma unroll
        for (int u = 0; u < 16; u += 4)
          *reinterpret_cast<float4*>(out + c0 + u) = make_float4(o[u], o[u + 1], o[u + 2], o[u + 3]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tm_empty);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// tf32 hi (+ lo = v - hi) parts of a row-major matrix
__global__ void k_split_tf32(const float4* __restrict__ src, int64_t n4, float4* __restrict__ hi,
                             float4* __restrict__ lo) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(src + i);
  float4 h;
  h.x = ptx::to_tf32(v.x); h.y = ptx::to_tf32(v.y); h.z = ptx::to_tf32(v.z); h.w = ptx::to_tf32(v.w);
  hi[i] = h;
  if (lo) lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// sums the vertex-split partials [vs][(s',t)][f] in a fixed order and scatters to d_w_eff[t][s'][f] / d_w_self[t][f]
__global__ void k_tc_dwg_finish(const float* __restrict__ partial, int vsplit, int m_pad, int T, int RA, int F,
                                int accumulate, float* __restrict__ d_w_eff, float* __restrict__ d_w_self) {
  const int F4 = F >> 2;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)T * (RA + 1) * F4) return;
  const int f4 = (int)(e % F4);
  const int ts = (int)(e / F4);
  const int sl = ts % (RA + 1), t = ts / (RA + 1);
  const int m = sl * T + t;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < vsplit; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)s * m_pad + m) * F) + f4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float4* dst = sl < RA ? (d_w_eff ? reinterpret_cast<float4*>(d_w_eff + ((size_t)t * RA + sl) * F) + f4 : nullptr)
                        : reinterpret_cast<float4*>(d_w_self + (size_t)t * F) + f4;
  if (!dst) return;
  if (accumulate) {
    const float4 old = *dst;
    acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
  }
  *dst = acc;
}

template <bool X3>
static int launch_gemm_tn(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                          const TcGemmTnParams& p, int grid, size_t smem, cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm_tn<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(GCB_PROF_BWD_DW, st);
  k_tc_gemm_tn<X3><<<grid, TN_THREADS, smem, st>>>(ah, al, bh, bl, p);
  profile_end(GCB_PROF_BWD_DW, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

TcGLayout tc_g_layout(int v_src, int R, int A, int F, int T, bool x3) {
  TcGLayout L{};
  const int np = x3 ? 2 : 1;
  L.v_pad = (int)ceil_div64(v_src, 128) * 128;
  L.kg = (R * A + 1) * T;
  L.m_pad = (int)ceil_div64(L.kg, 128) * 128;
  L.n_ntiles = (int)ceil_div64(F, 256);
  L.bn = (int)ceil_div64(ceil_div64(F, L.n_ntiles), 32) * 32;
  L.n_ntiles = (int)ceil_div64(F, L.bn);
  L.n_kch = L.v_pad / TN_KCH;
  const int nacc = 512 / L.bn >= 4 ? 4 : (512 / L.bn >= 2 ? 2 : 1);
  // vertex splits: chains of at most ~512 (3xTF32) / ~1024 accumulations per TMEM accumulator, then fill the SMs
  const int64_t per_chunk = (int64_t)(TN_KCH / 8) * (x3 ? 3 : 1);
  int s_min = (int)ceil_div64((int64_t)L.n_kch * per_chunk, (int64_t)nacc * (x3 ? 512 : 1024));
  if (s_min < 1) s_min = 1;
  if (s_min > L.n_kch) s_min = L.n_kch;
  const int64_t units = (int64_t)(L.m_pad / 128) * L.n_ntiles;
  int best = s_min;
  double best_eff = -1.0;
  for (int s = s_min; s <= s_min + 8 && s <= L.n_kch; ++s) {
    const int64_t tiles = units * s;
    const double eff = (double)tiles / (double)(ceil_div64(tiles, 148) * 148);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  L.vsplit = best;
  L.g_bytes = align_up((size_t)np * L.v_pad * L.kg * sizeof(float), 256);
  L.xsplit_bytes = align_up((size_t)np * v_src * F * sizeof(float), 256);
  L.partial_bytes = align_up((size_t)L.vsplit * L.m_pad * F * sizeof(float), 256);
  L.total = L.g_bytes + L.xsplit_bytes + L.partial_bytes + 256;
  return L;
}

int tc_build_g(const float* h, const float* w, const int32_t* rot_sel, const int32_t* csr_row_ptr,
               const int32_t* csr_entries, int v_out, int v_src, int R, int A, int T, bool x3, float* g_hi,
               float* g_lo, cudaStream_t st) {
  const int RA = R * A;
  const int kg = (RA + 1) * T;
  const int v_pad = (int)ceil_div64(v_src, 128) * 128;
  const size_t smem = (size_t)RA * T * sizeof(float);
  GCB_REQUIRE(smem <= 200 * 1024, GCB_EUNSUPPORTED, "tc_build_g: R*A*T too large for the G accumulator");
  if (smem > 48 * 1024)
    GCB_CUDA_OK(cudaFuncSetAttribute(k_bw_build_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = (T + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  profile_begin(GCB_PROF_BWD_CSR, st);
  k_bw_build_g<<<v_pad, threads, smem, st>>>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, T, kg, g_hi,
                                             x3 ? g_lo : nullptr);
  profile_end(GCB_PROF_BWD_CSR, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// d_x[v_src, F] (+)= G x Wcat, the weight tensor maps over Wcat [T x (R*A*F + F)] (hi / lo)
int tc_dx_from_g(const float* g_hi, const float* g_lo, const float* wcat_hi, const float* wcat_lo, int v_src, int R,
                 int A, int F, int T, bool x3, int accumulate, float* d_x, cudaStream_t st) {
  GCB_REQUIRE((uintptr_t)d_x % 16 == 0 && F % 4 == 0, GCB_EINVAL, "tc_dx_from_g: d_x must be 16-byte aligned");
  const int RA = R * A, KRA = RA * F, kcat = KRA + F;
  const int kg = (RA + 1) * T;
  const int v_pad = (int)ceil_div64(v_src, 128) * 128;
  CUtensorMap ah, al, bh, bl;
  int rc = tc_make_tmap_2d(&ah, g_hi, (uint64_t)kg, (uint64_t)v_pad, (uint64_t)kg * 4, 32, 128);
  if (rc) return rc;
  rc = tc_make_tmap_2d(&al, x3 ? g_lo : g_hi, (uint64_t)kg, (uint64_t)v_pad, (uint64_t)kg * 4, 32, 128);
  if (rc) return rc;
  rc = tc_make_tmap_2d(&bh, wcat_hi, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * 4, 32, 32, true);
  if (rc) return rc;
  rc = tc_make_tmap_2d(&bl, x3 ? wcat_lo : wcat_hi, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * 4, 32, 32, true);
  if (rc) return rc;
  TcGemmParams p;
  p.out = d_x; p.ld_out = F; p.rows_valid = v_src; p.cols_valid = F; p.accumulate = accumulate;
  p.n_mtiles = v_pad / 128;
  p.n_ntiles = (int)ceil_div64(F, GM_BN);
  p.n_kc = kg / 32;
  p.kc_per_group = T / 32;
  p.group_stride = F;
  const size_t stage_bytes = (size_t)(x3 ? 2 : 1) * (GM_STAGE_A + GM_STAGE_B);
  int ns = (int)((DX_SMEM_LIMIT - 2048) / stage_bytes);
  if (ns > 6) ns = 6;
  p.ns = ns;
  const size_t smem = 2048 + (size_t)ns * stage_bytes;
  const int n_tiles = p.n_mtiles * p.n_ntiles;
  const int grid = n_tiles < 148 ? n_tiles : 148;
  return x3 ? launch_gemm<true>(ah, al, bh, bl, p, grid, smem, st) : launch_gemm<false>(ah, al, bh, bl, p, grid, smem, st);
}

// d_w_eff[T,R,A,F], d_w_self[T,F] (+)= G^T x X
int tc_dw_from_g(const float* g_hi, const float* g_lo, const float* x, int v_src, int R, int A, int F, int T,
                 bool x3, int accumulate, float* d_w_eff, float* d_w_self, float* x_hi, float* x_lo, float* partial,
                 cudaStream_t st) {
  GCB_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)d_w_self % 16 == 0 && (uintptr_t)d_w_eff % 16 == 0 && F % 32 == 0 &&
                  T % 32 == 0,
              GCB_EINVAL, "tc_dw_from_g: operands must be 16-byte aligned, F and T multiples of 32");
  const TcGLayout L = tc_g_layout(v_src, R, A, F, T, x3);
  {
    const int64_t n4 = (int64_t)v_src * F / 4;
    k_split_tf32<<<(unsigned)ceil_div64(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n4,
                                                                reinterpret_cast<float4*>(x_hi),
                                                                x3 ? reinterpret_cast<float4*>(x_lo) : nullptr);
    GCB_LAUNCH_OK();
  }
  CUtensorMap ah, al, bh, bl;
  int rc = tc_make_tmap_mn(&ah, g_hi, (uint64_t)L.v_pad, (uint64_t)L.kg, (uint64_t)L.kg * 4, TN_KCH, 4);
  if (rc) return rc;
  rc = tc_make_tmap_mn(&al, x3 ? g_lo : g_hi, (uint64_t)L.v_pad, (uint64_t)L.kg, (uint64_t)L.kg * 4, TN_KCH, 4);
  if (rc) return rc;
  rc = tc_make_tmap_mn(&bh, x_hi, (uint64_t)v_src, (uint64_t)F, (uint64_t)F * 4, TN_KCH, (uint32_t)(L.bn / 32));
  if (rc) return rc;
  rc = tc_make_tmap_mn(&bl, x3 ? x_lo : x_hi, (uint64_t)v_src, (uint64_t)F, (uint64_t)F * 4, TN_KCH,
                       (uint32_t)(L.bn / 32));
  if (rc) return rc;
  TcGemmTnParams p;
  p.partial = partial; p.ld = F; p.n_cols = F;
  p.m_pad = L.m_pad; p.n_mtiles = L.m_pad / 128; p.n_ntiles = L.n_ntiles; p.vsplit = L.vsplit;
  p.bn = L.bn; p.n_kch = L.n_kch;
  p.nacc = 512 / L.bn >= 4 ? 4 : (512 / L.bn >= 2 ? 2 : 1);
  const size_t stage_bytes = (size_t)(x3 ? 2 : 1) * (TN_STAGE_A + (size_t)L.bn * TN_KCH * 4);
  int ns = (int)((DX_SMEM_LIMIT - 2048) / stage_bytes);
  if (ns > 10) ns = 10;
  GCB_REQUIRE(ns >= 2, GCB_EUNSUPPORTED, "tc_dw_from_g: stage does not fit in shared memory");
  p.ns = ns;
  const size_t smem = 2048 + (size_t)ns * stage_bytes;
  const int n_tiles = p.n_mtiles * p.vsplit * p.n_ntiles;
  const int grid = n_tiles < 148 ? n_tiles : 148;
  rc = x3 ? launch_gemm_tn<true>(ah, al, bh, bl, p, grid, smem, st) : launch_gemm_tn<false>(ah, al, bh, bl, p, grid, smem, st);
  if (rc) return rc;
  const int64_t n_out = (int64_t)T * (R * A + 1) * (F / 4);
  k_tc_dwg_finish<<<(unsigned)ceil_div64(n_out, 256), 256, 0, st>>>(partial, L.vsplit, L.m_pad, T, R * A, F, accumulate,
                                                                    d_w_eff, d_w_self);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

}  // namespace gcb
