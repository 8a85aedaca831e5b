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
// GCB_PREC_F16X3 (template flag H): G, X and the weights are written as power-of-two scaled fp16 hi / lo pairs
// (k_tc_maxima -> scales), both GEMMs issue kind::f16 (K = 16 per instruction: half the tensor time), the
// MN-major operands use 64-column atoms with the standard 128-byte swizzle, the epilogues undo the scales.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

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

// fp16 MN-major operand, standard 128-byte swizzle: atoms of 64 (M or N) elements x 8 k-rows of 128 bytes;
// lbo = byte distance between atoms along M/N, sbo = between 8-row groups along K
__device__ __forceinline__ uint64_t dx_mnmajor_desc_h(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// v * s = hi + lo (fp16)
__device__ __forceinline__ void split_half(float v, float s, __half& hi, __half& lo) {
  v *= s;
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// ---- G[v, s', t] ----------------------------------------------------------------------------------
// H: G * scales[0] as fp16 hi / lo rows of kg_ld (>= kg, padded with zeros) halves
template <bool H>
__global__ void k_bw_build_g(const float* __restrict__ h, const float* __restrict__ w,
                             const int32_t* __restrict__ rot_sel, const int32_t* __restrict__ row_ptr,
                             const int32_t* __restrict__ entries, int v_out, int v_src, int R, int A, int T,
                             int kg /* (R*A+1)*T */, int kg_ld, const float* __restrict__ scales,
                             float* __restrict__ g_hi, float* __restrict__ g_lo) {
  extern __shared__ float s_acc[];                       // [RA][T]
  // per staged entry {row offset of h[k], row offset of its rotated slot in s_acc, weight bits, -}: ONE broadcast
  // 16-byte read per entry in the loop below (the kernel runs at the load/store-unit wavefront rate: with three
  // separate arrays half of its shared-memory wavefronts were these metadata reads)
  __shared__ int4 s_meta[G_CHUNK];
  const int v = blockIdx.x;
  const int RA = R * A;
  const int t = threadIdx.x;
  float* out_hi = g_hi + (size_t)v * kg;
  float* out_lo = g_lo ? g_lo + (size_t)v * kg : nullptr;
  __half* hh = reinterpret_cast<__half*>(g_hi) + (size_t)v * kg_ld;
  __half* hl = reinterpret_cast<__half*>(g_lo) + (size_t)v * kg_ld;
  const float sg = H ? __ldg(scales) : 1.f;
  if (H) {
    for (int i = kg + threadIdx.x; i < kg_ld; i += blockDim.x) { hh[i] = __float2half_rn(0.f); hl[i] = __float2half_rn(0.f); }
  }
  if (v >= v_src) {                                      // padding rows of the GEMM's A operand
    for (int i = threadIdx.x; i < kg; i += blockDim.x) {
      if (H) { hh[i] = __float2half_rn(0.f); hl[i] = __float2half_rn(0.f); continue; }
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
      s_meta[i] = make_int4(k * T, (xr * A + yr) * T, __float_as_int(w[s]), 0);
    }
    __syncthreads();
    if (t < T) {
      int i = 0;
      for (; i + 8 <= n; i += 8) {
        float hv[8];
        int4 m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          m[u] = s_meta[i + u];
          hv[u] = __ldg(h + (size_t)m[u].x + t);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float* a = s_acc + m[u].y + t;
          *a = fmaf(__int_as_float(m[u].z), hv[u], *a);
        }
      }
      for (; i < n; ++i) {
        const int4 m = s_meta[i];
        float* a = s_acc + m.y + t;
        *a = fmaf(__int_as_float(m.z), __ldg(h + (size_t)m.x + t), *a);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RA * T; i += blockDim.x) {
    const float val = s_acc[i];
    if (H) { split_half(val, sg, hh[i], hl[i]); continue; }
    const float hi = ptx::to_tf32(val);
    out_hi[i] = hi;
    if (out_lo) out_lo[i] = val - hi;
  }
  for (int i = threadIdx.x; i < T; i += blockDim.x) {   // centre slab: h[v] itself (zero for halo rows)
    const float val = v < v_out ? h[(size_t)v * T + i] : 0.f;
    if (H) { split_half(val, sg, hh[RA * T + i], hl[RA * T + i]); continue; }
    const float hi = ptx::to_tf32(val);
    out_hi[RA * T + i] = hi;
    if (out_lo) out_lo[RA * T + i] = val - hi;
  }
}


// ---- G[v, s', t], ring-grouped CSR: one register accumulator per lane and template --------------------
// The rotation moves a template vertex only INSIDE its radial ring: s' = (x, (y + rot_k) % A).  With the CSR
// entries of a source vertex grouped by ring x (gcb_csr_build with the template size) the kernel needs no
// accumulator array at all:
//   one warp per source vertex, lane l owns templates l, l + 32, ... (TPL per pass, passes over T for wide layers);
//   per 32 entries the lanes decode one entry each ({h row, ring, rotated angular index y', weight}) into a per-warp
//   staging row and keep (ring, y') in registers;
//   for every ring segment of the chunk and every y' a ballot gives the segment's entries with that y'; their h
//   rows are summed (CSR order: the order of k_bw_build_g, same bits) into TPL registers and stored - converted -
//   straight to G.  Rings without entries are written as zeros.
//   A ring that straddles a 32-entry chunk parks its running sums in a per-warp shared-memory tile and continues
//   from them (rare: a ring has ~24 entries at the FAUST shape).
// Against k_bw_build_g (an (R*A) x T accumulator in shared memory, read-modify-written per entry: 4 load/store-unit
// wavefronts per (entry, warp), 84 us at the FAUST shape) an entry costs one broadcast read and TPL loads.
// Two earlier variants are recorded in profiles/r02_notes.md: A accumulators per template selected by a switch on y'
// (187 us: three indirect branches per entry) and the same with 8 entries' loads in flight (178 us).
constexpr int G2_WARPS = 8;

// (a, b, c, d) * s = hi + lo as two packed fp16 quadruples
__device__ __forceinline__ void g2_split4(const float4 v, float s, uint2& hi, uint2& lo) {
  const float a = v.x * s, b = v.y * s, c = v.z * s, d = v.w * s;
  const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(a - f0.x, b - f0.y), l1 = __floats2half2_rn(c - f1.x, d - f1.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// Lane l owns the four templates t_base + 4 l .. + 3 of a 128-template pass: one 16-byte load of an h row, four
// FMAs and no per-template address arithmetic per entry (the variant with one template per lane and load was
// instruction-issue bound at 94 us: ~25 instructions per entry and ~56 per (ring, y') group).
// DENSE (a gradient for EVERY rotation, h = [v_out][n_rot][T]): an entry with angular index y feeds the rotated slot
// y' from rotation i with offset o_i = (y' - y) mod A (`inv` maps an offset to its rotation, -1: not in the list),
// so every entry of a ring takes part in every y' group - n_rot times the gather of the sparse case, but the two
// GEMMs behind G stay the size of ONE rotation instead of running n_rot times.
struct RotInv { int8_t of_offset[GCB_MAX_ROT]; };

template <bool H, bool DENSE>
__global__ void __launch_bounds__(G2_WARPS * 32)
k_bw_build_g2(const float* __restrict__ h, const float* __restrict__ w, const int32_t* __restrict__ rot_sel,
              const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ entries, int v_out, int v_src, int v_pad,
              int R, int A, int T, int kg, int kg_ld, const float* __restrict__ scales, float* __restrict__ g_hi,
              float* __restrict__ g_lo, int n_rot, const RotInv inv) {
  extern __shared__ float4 s_carry_all[];              // [G2_WARPS][A][32]: running sums of a straddling ring
  __shared__ int2 s_meta[G2_WARPS][32];                // {h row offset, weight bits}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x * G2_WARPS + warp;
  if (v >= v_pad) return;
  const int RA = R * A;
  float4* s_carry = s_carry_all + (size_t)warp * A * 32;
  float* out_hi = g_hi + (size_t)v * kg;
  float* out_lo = g_lo ? g_lo + (size_t)v * kg : nullptr;
  __half* hh = reinterpret_cast<__half*>(g_hi) + (size_t)v * kg_ld;
  __half* hl = reinterpret_cast<__half*>(g_lo) + (size_t)v * kg_ld;
  const float sg = H ? __ldg(scales) : 1.f;
  auto put4 = [&](int col, const float4 val) {          // col % 4 == 0
    if (H) {
      uint2 hi, lo;
      g2_split4(val, sg, hi, lo);
      *reinterpret_cast<uint2*>(hh + col) = hi;
      *reinterpret_cast<uint2*>(hl + col) = lo;
    } else {
      float4 hi;
      hi.x = ptx::to_tf32(val.x); hi.y = ptx::to_tf32(val.y); hi.z = ptx::to_tf32(val.z); hi.w = ptx::to_tf32(val.w);
      *reinterpret_cast<float4*>(out_hi + col) = hi;
      if (out_lo)
        *reinterpret_cast<float4*>(out_lo + col) = make_float4(val.x - hi.x, val.y - hi.y, val.z - hi.z, val.w - hi.w);
    }
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (H)
    for (int i = kg + lane; i < kg_ld; i += 32) { hh[i] = __float2half_rn(0.f); hl[i] = __float2half_rn(0.f); }
  if (v >= v_src) {                                      // padding rows of the GEMM's A operand
    for (int i = 4 * lane; i < kg; i += 128) put4(i, zero4);
    return;
  }
  const int beg = row_ptr[v], end = row_ptr[v + 1];
  for (int t_base = 0; t_base < T; t_base += 128) {
    const int t = t_base + 4 * lane;
    const bool t_on = t < T;
    int next_x = 0;           // rings below it are written
    bool open = false;        // ring `next_x` has running sums parked in s_carry (it may continue in the next chunk)
    auto zero_rings = [&](int x0, int x1) {
      if (t_on)
        for (int x = x0; x < x1; ++x)
          for (int y = 0; y < A; ++y) put4((x * A + y) * T + t, zero4);
    };
    auto close_ring = [&]() {   // the parked ring is complete: convert and store it
      if (t_on)
        for (int y = 0; y < A; ++y) put4((next_x * A + y) * T + t, s_carry[y * 32 + lane]);
      open = false;
      ++next_x;
    };
    for (int base = beg; base < end; base += 32) {
      const int n = min(32, end - base);
      __syncwarp();
      int my_x = -1, my_y = -1;
      if (lane < n) {
        const int s = __ldg(entries + base + lane);
        const int cell = s / 3;
        const int k = cell / RA;
        const int slot = cell - k * RA;
        my_x = slot / A;
        my_y = slot - my_x * A;
        if (!DENSE) {
          my_y += __ldg(rot_sel + k);
          my_y -= (my_y >= A) ? A : 0;
        }
        s_meta[warp][lane] = make_int2(DENSE ? k * n_rot * T : k * T, __float_as_int(__ldg(w + s)));
      }
      __syncwarp();
      int pos = 0;
      while (pos < n) {
        const int x = __shfl_sync(0xffffffffu, my_x, pos);
        const unsigned seg = __ballot_sync(0xffffffffu, my_x == x);          // (entries of a ring are adjacent)
        const int seg_end = pos + __popc(seg);
        if (open && x != next_x) close_ring();
        const bool cont = open;                                               // this segment continues the parked ring
        if (!cont) { zero_rings(next_x, x); next_x = x; }
        const bool more = seg_end == n && base + n < end;                     // the ring may go on in the next chunk
        for (int y = 0; y < A; ++y) {
          int my_i = 0;   // DENSE: the rotation through which this lane's entry reaches slot y
          if (DENSE) {
            int d = y - my_y;
            d += d < 0 ? A : 0;
            my_i = my_y >= 0 ? inv.of_offset[d] : -1;
          }
          unsigned bits = seg & __ballot_sync(0xffffffffu, DENSE ? my_i >= 0 : my_y == y);
          float4 acc = cont ? s_carry[y * 32 + lane] : zero4;
          while (bits) {
            // up to four entries' rows are requested before the first is added (CSR order kept)
            float wg[4];
            float4 hv[4];
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (bits) {
                const int e = __ffs((int)bits) - 1;
                bits &= bits - 1;
                const int2 m = s_meta[warp][e];
                wg[u] = __int_as_float(m.y);
                const int roff = DENSE ? __shfl_sync(0xffffffffu, my_i, e) * T : 0;
                hv[u] = t_on ? __ldg(reinterpret_cast<const float4*>(h + (size_t)m.x + roff + t)) : zero4;
                ++cnt;
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (u < cnt) {
                acc.x = fmaf(wg[u], hv[u].x, acc.x); acc.y = fmaf(wg[u], hv[u].y, acc.y);
                acc.z = fmaf(wg[u], hv[u].z, acc.z); acc.w = fmaf(wg[u], hv[u].w, acc.w);
              }
          }
          if (cont || more) s_carry[y * 32 + lane] = acc;
          else if (t_on) put4((x * A + y) * T + t, acc);
        }
        if (cont || more) {
          open = true;
          if (!more) close_ring();
        } else {
          ++next_x;
        }
        pos = seg_end;
      }
    }
    if (open) close_ring();
    zero_rings(next_x, R);                 // empty rings after the last one that had entries
  }
  for (int i = 4 * lane; i < T; i += 128) {  // centre slab: h[v] itself (zero for halo rows); DENSE: summed over rotations
    float4 c = zero4;
    if (v < v_out) {
      if (DENSE) {
        for (int r = 0; r < n_rot; ++r) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(h + ((size_t)v * n_rot + r) * T + i));
          c.x += a.x; c.y += a.y; c.z += a.z; c.w += a.w;
        }
      } else {
        c = __ldg(reinterpret_cast<const float4*>(h + (size_t)v * T + i));
      }
    }
    put4(RA * T + i, c);
  }
}

// ---- persistent TMA-fed tcgen05 GEMM ---------------------------------------------------------------
constexpr int GM_EPI_WARPS = 4;
constexpr int GM_THREADS = (2 + GM_EPI_WARPS) * 32;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int GM_BN = 128;
constexpr int GM_STAGE_A = 16 * 1024;   // 128 rows x 32 k
constexpr int GM_STAGE_B = 16 * 1024;   // 32 k x 128 n (four MN-major column blocks)

struct TcGemmParams {
  const float* inv_scale;  // fp16 route: the accumulators carry the operand scales; *inv_scale undoes them
  uint32_t h_lbo, h_sbo, h_kstep;   // fp16 MN-major descriptor fields (bytes)
  float* out;
  int ld_out, rows_valid, cols_valid, accumulate;
  int n_mtiles, n_ntiles, ns;
  int n_kc;               // K chunks of 32
  int kparts;             // passes over K per tile: every pass ends in an fp32 (RN) add into the output, which keeps
                          // the truncating accumulation chains of the tensor core short for long contractions
  int kc_per_group;       // B tile of chunk kc: rows (kc % kc_per_group)*32, cols (kc / kc_per_group)*group_stride + n0
  int group_stride;
};

// H (fp16 operands): a stage holds two sub-chunks of 32 k, each an (A 128 x 64 B K-major SWIZZLE_64B, B 32 k x 128 n
// as two 64-column MN-major atoms) pair laid out like the tf32 stage at half the size.
// MP (multi-pass): K is walked in p.kparts passes; MP = false compiles the single-pass loops of the common case
// (with the pass loops around them the one-pass kernel measured 93 us against 84 at the FAUST shape)
template <bool X3, bool H, bool MP>
__global__ void __launch_bounds__(GM_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
          const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
          const TcGemmParams p) {
  constexpr int NPROD = X3 ? 2 : 1;
  constexpr int NSUB = H ? 2 : 1;                       // 32-k sub-chunks per stage
  constexpr int SUB_A = GM_STAGE_A / NSUB, SUB_B = GM_STAGE_B / NSUB;
  constexpr int SUB_BYTES = NPROD * (SUB_A + SUB_B);
  constexpr int STAGE_BYTES = NSUB * SUB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + (uint32_t)(p.ns * STAGE_BYTES);
  const uint32_t full = bar_base, empty = full + 8u * p.ns;
  const uint32_t tm_full = empty + 8u * p.ns, tm_empty = tm_full + 16u;
  const int n_st = (p.n_kc + NSUB - 1) / NSUB;         // stages per tile
  const uint32_t tmem_slot = tm_empty + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_mtiles * p.n_ntiles;
  const int nacc = p.n_kc >= 4 ? 4 : 1;   // K chunks round-robin over this many accumulators
  const int kparts = MP ? p.kparts : 1;

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
        for (int si = 0; si < n_st; ++si) {
          ptx::mbar_wait(empty + 8u * stage, par);
          const int n_sub = min(NSUB, p.n_kc - si * NSUB);
          ptx::mbar_arrive_expect_tx(full + 8u * stage, (uint32_t)(n_sub * SUB_BYTES));
          for (int sub = 0; sub < n_sub; ++sub) {
            const int kc = si * NSUB + sub;
            const uint32_t a_dst = smem_base + (uint32_t)(stage * STAGE_BYTES + sub * SUB_BYTES);
            const uint32_t b_dst = a_dst + NPROD * SUB_A;
            const int b_row = in_grp * 32, b_col = grp * p.group_stride + nt * GM_BN;
            ptx::tma_load_2d(a_dst, &tmap_a_hi, full + 8u * stage, kc * 32, mt * 128);
            if (X3) ptx::tma_load_2d(a_dst + SUB_A, &tmap_a_lo, full + 8u * stage, kc * 32, mt * 128);
            constexpr int BOX_N = H ? 64 : 32;            // columns per MN-major atom (128 bytes)
#pragma unroll
            for (int j = 0; j < GM_BN / BOX_N; ++j) {
              ptx::tma_load_2d(b_dst + j * 4096, &tmap_b_hi, full + 8u * stage, b_col + BOX_N * j, b_row);
              if (X3) ptx::tma_load_2d(b_dst + SUB_B + j * 4096, &tmap_b_lo, full + 8u * stage, b_col + BOX_N * j, b_row);
            }
            if (++in_grp == p.kc_per_group) { in_grp = 0; ++grp; }
          }
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc<H>(128, GM_BN) | (1u << 16);   // B MN-major
      constexpr int A_ROW = H ? 64 : 128;                 // bytes of one K-major A row (32 k)
      constexpr int KB = H ? 2 : 4;                       // MMAs per 32-k sub-chunk (K = 16 halves / 8 tf32)
      int stage = 0;
      uint32_t par = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
       for (int kp = 0; kp < kparts; ++kp, ++it) {
        ptx::mbar_wait(tm_empty, (((uint32_t)it) & 1u) ^ 1u);
        ptx::tc_fence_after();
        uint32_t started = 0;   // bit a: accumulator a already holds a partial sum
        const int si_b = MP ? (int)(((int64_t)n_st * kp) / kparts) : 0, si_e = MP ? (int)(((int64_t)n_st * (kp + 1)) / kparts) : n_st;
        for (int si = si_b; si < si_e; ++si) {
          ptx::mbar_wait(full + 8u * stage, par);
          ptx::tc_fence_after();
          const int n_sub = min(NSUB, p.n_kc - si * NSUB);
          for (int sub = 0; sub < n_sub; ++sub) {
            const int kc = si * NSUB + sub;
            const int acc = nacc == 4 ? (kc & 3) : 0;
            const uint32_t d = tmem_base + (uint32_t)(acc * GM_BN);
            const uint32_t a_hi = smem_base + (uint32_t)(stage * STAGE_BYTES + sub * SUB_BYTES);
            const uint32_t b_hi = a_hi + NPROD * SUB_A;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const uint64_t da = ptx::make_kmajor_desc<A_ROW>(a_hi + kb * 32);
              const uint64_t da_lo = ptx::make_kmajor_desc<A_ROW>(a_hi + SUB_A + kb * 32);
              const uint64_t db = H ? dx_mnmajor_desc_h(b_hi + kb * p.h_kstep, p.h_lbo, p.h_sbo)
                                    : dx_mnmajor_desc(b_hi + kb * 1024, 4096, 512);
              const uint64_t db_lo = H ? dx_mnmajor_desc_h(b_hi + SUB_B + kb * p.h_kstep, p.h_lbo, p.h_sbo)
                                       : dx_mnmajor_desc(b_hi + SUB_B + kb * 1024, 4096, 512);
              ptx::umma<H>(d, da, db, idesc, (started >> acc) & 1u);
              started |= 1u << acc;
              if (X3) {
                ptx::umma<H>(d, da, db_lo, idesc, 1u);
                ptx::umma<H>(d, da_lo, db, idesc, 1u);
              }
            }
          }
          ptx::umma_commit(empty + 8u * stage);
          if (++stage == p.ns) { stage = 0; par ^= 1u; }
        }
        ptx::umma_commit(tm_full);
       }
      }
    }
  } else {
    const int q = warp & 3;
    const float inv = H ? __ldg(p.inv_scale) : 1.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
     for (int kp = 0; kp < kparts; ++kp, ++it) {
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
            if (H) { o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv; }
            float4* dst = reinterpret_cast<float4*>(out + c0 + u);
            if (p.accumulate || (MP && kp > 0)) {   // (kp > 0: this thread's own earlier store to the same address)
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
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 4 * GM_BN);
  }
}

template <bool X3, bool H, bool MP>
static int launch_gemm_mp(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                          const TcGemmParams& p, int grid, size_t smem, cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm<X3, H, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(GCB_PROF_BWD_DI, st);
  k_tc_gemm<X3, H, MP><<<grid, GM_THREADS, smem, st>>>(ah, al, bh, bl, p);
  profile_end(GCB_PROF_BWD_DI, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

template <bool X3, bool H = false>
static int launch_gemm(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                       const TcGemmParams& p, int grid, size_t smem, cudaStream_t st) {
  return p.kparts > 1 ? launch_gemm_mp<X3, H, true>(ah, al, bh, bl, p, grid, smem, st)
                      : launch_gemm_mp<X3, H, false>(ah, al, bh, bl, p, grid, smem, st);
}

// ---- dWcat = G^T X : both operands MN-major ---------------------------------------------------------
constexpr int TN_KCH = 16;                     // contraction rows (source vertices) per stage
constexpr int TN_STAGE_A = 128 * TN_KCH * 4;   // 8 KB: four 32-row atoms of M
constexpr int TN_THREADS = (2 + GM_EPI_WARPS) * 32;

struct TcGemmTnParams {
  const float* inv_scale;           // fp16 route: undoes the operand scales
  uint32_t h_lbo, h_sbo, h_kstep;   // fp16 MN-major descriptor fields (bytes)
  float* partial;    // [vsplit][m_pad][ld]
  int ld, n_cols;    // row length of the output (= columns of X)
  int m_pad, n_mtiles, n_ntiles, vsplit;
  int bn;            // columns per N tile (multiple of 32, <= 256); the last tile may be narrower
  int n_kch;         // K chunks of TN_KCH rows in total
  int ns, nacc;
};

// H (fp16 operands): 64-column atoms (3-D boxes 64 x 32 rows x atoms, standard 128-byte swizzle), 32 contraction
// rows per stage = two K = 16 steps.
template <bool X3, bool H>
__global__ void __launch_bounds__(TN_THREADS, 1)
k_tc_gemm_tn(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
             const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
             const TcGemmTnParams p) {
  constexpr int NPROD = X3 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  constexpr int KCH = H ? 32 : TN_KCH;          // contraction rows per stage
  constexpr int ATOM = H ? 64 : 32;             // operand columns per 128-byte atom row
  const int stage_b = p.bn * TN_KCH * 4;        // (= bn * 32 rows * 2 bytes with fp16 operands)
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
          ptx::tma_load_3d(a_dst, &tmap_a_hi, full + 8u * stage, 0, kc * KCH, mt * (128 / ATOM));
          if (X3) ptx::tma_load_3d(a_dst + TN_STAGE_A, &tmap_a_lo, full + 8u * stage, 0, kc * KCH, mt * (128 / ATOM));
          ptx::tma_load_3d(b_dst, &tmap_b_hi, full + 8u * stage, 0, kc * KCH, nt * (p.bn / ATOM));
          if (X3) ptx::tma_load_3d(b_dst + stage_b, &tmap_b_lo, full + 8u * stage, 0, kc * KCH, nt * (p.bn / ATOM));
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
        const uint32_t idesc = ptx::make_idesc<H>(128, n_valid) | (1u << 15) | (1u << 16);   // A and B MN-major
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
          for (int k8 = 0; k8 < 2; ++k8) {   // two K steps per stage (8 tf32 rows / 16 fp16 rows each)
            const uint64_t da = H ? dx_mnmajor_desc_h(a_hi + k8 * p.h_kstep, p.h_lbo, p.h_sbo)
                                  : dx_mnmajor_desc(a_hi + k8 * 1024, TN_KCH * 128, 512);
            const uint64_t db = H ? dx_mnmajor_desc_h(b_hi + k8 * p.h_kstep, p.h_lbo, p.h_sbo)
                                  : dx_mnmajor_desc(b_hi + k8 * 1024, TN_KCH * 128, 512);
            ptx::umma<H>(d, da, db, idesc, (started >> acc) & 1u);
            started |= 1u << acc;
            if (X3) {
              const uint64_t db_lo = H ? dx_mnmajor_desc_h(b_hi + stage_b + k8 * p.h_kstep, p.h_lbo, p.h_sbo)
                                       : dx_mnmajor_desc(b_hi + stage_b + k8 * 1024, TN_KCH * 128, 512);
              const uint64_t da_lo = H ? dx_mnmajor_desc_h(a_hi + TN_STAGE_A + k8 * p.h_kstep, p.h_lbo, p.h_sbo)
                                       : dx_mnmajor_desc(a_hi + TN_STAGE_A + k8 * 1024, TN_KCH * 128, 512);
              ptx::umma<H>(d, da, db_lo, idesc, 1u);
              ptx::umma<H>(d, da_lo, db, idesc, 1u);
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
    const float inv = H ? __ldg(p.inv_scale) : 1.f;
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
        if (H) {
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] *= inv;
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

// src[rows][cols] * (*scale) = hi + lo in fp16, rows of ld (>= cols, zero padded) halves; ld % 8 == 0 and
// cols % 4 == 0, a thread converts 8 consecutive columns
__global__ void k_split_half(const float* __restrict__ src, int64_t rows, int cols, int ld,
                             const float* __restrict__ scale, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int64_t i8 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int g = ld >> 3;
  if (i8 >= rows * g) return;
  const int64_t r = i8 / g;
  const int c = (int)(i8 - r * g) << 3;
  const float4* p = reinterpret_cast<const float4*>(src + r * cols + c);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = c + 4 <= cols ? __ldg(p) : z, b = c + 8 <= cols ? __ldg(p + 1) : z;
  const float s = __ldg(scale);
  __half h[8], l[8];
  split_half(a.x, s, h[0], l[0]); split_half(a.y, s, h[1], l[1]);
  split_half(a.z, s, h[2], l[2]); split_half(a.w, s, h[3], l[3]);
  split_half(b.x, s, h[4], l[4]); split_half(b.y, s, h[5], l[5]);
  split_half(b.z, s, h[6], l[6]); split_half(b.w, s, h[7], l[7]);
  *reinterpret_cast<uint4*>(hi + i8 * 8) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo + i8 * 8) = *reinterpret_cast<const uint4*>(l);
}

// Power-of-two operand scales of the fp16 backward (k_tc_maxima, mode 2): |G| <= max|h| * max|w| * longest CSR
// row - a loose bound only costs absolute resolution far below fp32's (2^-25 of 2^14 / looseness).
size_t tc_bwd_scales_bytes() { return 256; }

// scratch: 8 unsigned maxima, then the 5 scales (the caller passes scratch + 8 words as `scales`)
int tc_bwd_scales(const float* h, const float* w, const int32_t* csr_row_ptr, const float* w_eff, const float* w_self,
                  const float* x, int v_out, int v_src, int R, int A, int F, int T, void* scratch, cudaStream_t st,
                  int h_rows_per_vertex) {
  unsigned* mx = (unsigned*)scratch;
  GCB_CUDA_OK(cudaMemsetAsync(mx, 0, 32, st));
  const bool have_csr = csr_row_ptr && w_eff;
  TcMaxJob job{};
  job.mode = 2; job.have_csr = have_csr ? 1 : 0;
  int n = 0;
  job.seg[n++] = {h, (long long)v_out * h_rows_per_vertex * T, 0, 0};
  if (have_csr) {
    job.seg[n++] = {w, (long long)v_out * R * A * 3, 0, 1};
    job.seg[n++] = {csr_row_ptr, (long long)v_src, 1, 2};
    job.seg[n++] = {w_eff, (long long)T * R * A * F, 0, 3};
  }
  job.seg[n++] = {w_self, (long long)T * F, 0, 3};
  job.seg[n++] = {x, (long long)v_src * F, 0, 4};
  job.n_seg = n;
  return tc_maxima_scales(job, mx, (float*)(mx + 8), st);
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

template <bool X3, bool H = false>
static int launch_gemm_tn(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                          const TcGemmTnParams& p, int grid, size_t smem, cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm_tn<X3, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(GCB_PROF_BWD_DW, st);
  k_tc_gemm_tn<X3, H><<<grid, TN_THREADS, smem, st>>>(ah, al, bh, bl, p);
  profile_end(GCB_PROF_BWD_DW, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// (half: fp16 operands - 64-column atoms, 32 contraction rows per stage; the buffer sizes stay those of the
// 3xTF32 route, which bound the fp16 ones)
TcGLayout tc_g_layout(int v_src, int R, int A, int F, int T, bool x3, bool half) {
  TcGLayout L{};
  const int np = x3 ? 2 : 1;
  const int atom = half ? 64 : 32;
  L.v_pad = (int)ceil_div64(v_src, 128) * 128;
  L.kg = (R * A + 1) * T;
  L.m_pad = (int)ceil_div64(L.kg, 128) * 128;
  L.n_ntiles = (int)ceil_div64(F, 256);
  L.bn = (int)ceil_div64(ceil_div64(F, L.n_ntiles), atom) * atom;
  L.n_ntiles = (int)ceil_div64(F, L.bn);
  L.n_kch = L.v_pad / (half ? 32 : TN_KCH);
  const int nacc = 512 / L.bn >= 4 ? 4 : (512 / L.bn >= 2 ? 2 : 1);
  // vertex splits: chains of at most ~512 (3xTF32) / ~1024 accumulations per TMEM accumulator, then fill the SMs
  const int64_t per_chunk = (int64_t)2 * (x3 ? 3 : 1);   // two K steps per stage
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

// Buffer sizes that serve every route tc_conv_bwd can take for this shape: the tf32 and the fp16 layouts differ in
// their K chunking and N tile, hence in the vertex-split count of the G^T X partial sums, and the centre-only route
// (R*A = 0) has a layout of its own.  The workspace is sized and carved with the element-wise maximum.
TcGLayout tc_g_layout_max(int v_src, int R, int A, int F, int T, bool x3) {
  TcGLayout L = tc_g_layout(v_src, R, A, F, T, x3, false);
  auto widen = [&](const TcGLayout& o) {
    if (o.g_bytes > L.g_bytes) L.g_bytes = o.g_bytes;
    if (o.xsplit_bytes > L.xsplit_bytes) L.xsplit_bytes = o.xsplit_bytes;
    if (o.partial_bytes > L.partial_bytes) L.partial_bytes = o.partial_bytes;
  };
  if (x3) {
    TcGLayout H = tc_g_layout(v_src, R, A, F, T, true, true);
    // fp16 G / X rows are padded to multiples of 64 halves
    const size_t kg_ld = (size_t)ceil_div64(H.kg, 64) * 64, x_ld = (size_t)ceil_div64(F, 64) * 64;
    H.g_bytes = align_up((size_t)2 * H.v_pad * kg_ld * sizeof(__half), 256);
    H.xsplit_bytes = align_up((size_t)2 * v_src * x_ld * sizeof(__half), 256);
    widen(H);
  }
  widen(tc_g_layout(v_src, 0, 1, F, T, x3, false));
  L.total = L.g_bytes + L.xsplit_bytes + L.partial_bytes + 256;
  return L;
}

template <bool H, bool DENSE>
static int launch_build_g2(const float* h, const float* w, const int32_t* rot_sel, const int32_t* row_ptr,
                           const int32_t* entries, int v_out, int v_src, int v_pad, int R, int A, int T, int kg, int kg_ld,
                           const float* scales, float* g_hi, float* g_lo, cudaStream_t st, int n_rot, const RotInv& inv) {
  const size_t smem = (size_t)G2_WARPS * A * 32 * sizeof(float4);
  if (smem > 40 * 1024)
    GCB_CUDA_OK(cudaFuncSetAttribute(k_bw_build_g2<H, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bw_build_g2<H, DENSE><<<(unsigned)ceil_div64(v_pad, G2_WARPS), G2_WARPS * 32, smem, st>>>(
      h, w, rot_sel, row_ptr, entries, v_out, v_src, v_pad, R, A, T, kg, kg_ld, scales, g_hi, g_lo, n_rot, inv);
  return GCB_OK;
}

int tc_build_g(const float* h, const float* w, const int32_t* rot_sel, const int32_t* csr_row_ptr,
               const int32_t* csr_entries, int v_out, int v_src, int R, int A, int T, bool x3, float* g_hi,
               float* g_lo, cudaStream_t st, const float* scales, bool ring_sorted, int dense_n_rot,
               const int32_t* dense_offsets) {
  const int RA = R * A;
  RotInv inv;
  for (int i = 0; i < GCB_MAX_ROT; ++i) inv.of_offset[i] = -1;
  if (dense_n_rot > 0) {
    GCB_REQUIRE(ring_sorted && A <= GCB_MAX_ROT && T % 32 == 0 && ((uintptr_t)h % 16 == 0) && dense_offsets, GCB_EUNSUPPORTED,
                "tc_build_g: the dense route needs a ring-grouped CSR, A <= %d and T %% 32 == 0", GCB_MAX_ROT);
    for (int i = 0; i < dense_n_rot; ++i) {
      const int o = ((dense_offsets[i] % A) + A) % A;
      GCB_REQUIRE(inv.of_offset[o] < 0, GCB_EUNSUPPORTED, "tc_build_g: the dense route needs distinct rotation offsets");
      inv.of_offset[o] = (int8_t)i;
    }
  }
  const int kg = (RA + 1) * T;
  const int v_pad = (int)ceil_div64(v_src, 128) * 128;
  // GCB_BUILD_G2=0 (read per call): the shared-memory kernel k_bw_build_g, for A/B measurements
  const bool g2_off = getenv("GCB_BUILD_G2") && atoi(getenv("GCB_BUILD_G2")) == 0;
  if (ring_sorted && A <= GCB_MAX_ROT && T % 32 == 0 && !g2_off && ((uintptr_t)h % 16 == 0)) {
    const bool half = scales != nullptr;
    const int kg_ld = half ? (int)ceil_div64(kg, 64) * 64 : kg;
    float* lo = half || x3 ? g_lo : nullptr;
    int rc;
    profile_begin(GCB_PROF_BWD_CSR, st);
    if (dense_n_rot > 0)
      rc = half ? launch_build_g2<true, true>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, v_pad, R, A, T, kg, kg_ld, scales, g_hi, lo, st, dense_n_rot, inv)
                : launch_build_g2<false, true>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, v_pad, R, A, T, kg, kg_ld, scales, g_hi, lo, st, dense_n_rot, inv);
    else
      rc = half ? launch_build_g2<true, false>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, v_pad, R, A, T, kg, kg_ld, scales, g_hi, lo, st, 0, inv)
                : launch_build_g2<false, false>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, v_pad, R, A, T, kg, kg_ld, scales, g_hi, lo, st, 0, inv);
    profile_end(GCB_PROF_BWD_CSR, st);
    if (rc) return rc;
    GCB_LAUNCH_OK();
    return GCB_OK;
  }
  GCB_REQUIRE(dense_n_rot == 0, GCB_EUNSUPPORTED, "tc_build_g: the dense route is served by k_bw_build_g2 only");
  const size_t smem = (size_t)RA * T * sizeof(float);
  GCB_REQUIRE(smem <= TC_BUILD_G_SMEM_LIMIT, GCB_EUNSUPPORTED, "tc_build_g: R*A*T too large for the G accumulator");
  const bool half = scales != nullptr;
  const int kg_ld = (int)ceil_div64(kg, 64) * 64;
  if (smem > 48 * 1024) {
    GCB_CUDA_OK(cudaFuncSetAttribute(k_bw_build_g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GCB_CUDA_OK(cudaFuncSetAttribute(k_bw_build_g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int threads = (T + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  profile_begin(GCB_PROF_BWD_CSR, st);
  if (half)
    k_bw_build_g<true><<<v_pad, threads, smem, st>>>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, T, kg,
                                                     kg_ld, scales, g_hi, g_lo);
  else
    k_bw_build_g<false><<<v_pad, threads, smem, st>>>(h, w, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, T,
                                                      kg, kg, nullptr, g_hi, x3 ? g_lo : nullptr);
  profile_end(GCB_PROF_BWD_CSR, st);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// d_x[v_src, F] (+)= G x Wcat, the weight tensor maps over Wcat [T x (R*A*F + F)] (hi / lo)
int tc_dx_from_g(const float* g_hi, const float* g_lo, const float* wcat_hi, const float* wcat_lo, int v_src, int R,
                 int A, int F, int T, bool x3, int accumulate, float* d_x, cudaStream_t st, const float* scales) {
  GCB_REQUIRE((uintptr_t)d_x % 16 == 0 && F % 4 == 0, GCB_EINVAL, "tc_dx_from_g: d_x must be 16-byte aligned");
  const int RA = R * A, KRA = RA * F, kcat = KRA + F;
  const int kg = (RA + 1) * T;
  const int v_pad = (int)ceil_div64(v_src, 128) * 128;
  const bool half = scales != nullptr;
  CUtensorMap ah, al, bh, bl;
  int rc;
  if (half) {
    const uint64_t kg_ld = (uint64_t)ceil_div64(kg, 64) * 64;
    rc = tc_make_tmap_2d_h(&ah, (const __half*)g_hi, (uint64_t)kg, (uint64_t)v_pad, kg_ld * 2, 32, 128);
    if (rc) return rc;
    rc = tc_make_tmap_2d_h(&al, (const __half*)g_lo, (uint64_t)kg, (uint64_t)v_pad, kg_ld * 2, 32, 128);
    if (rc) return rc;
    // B: 32 template rows x 64 feature columns of Wcat per box = one MN-major atom column (4 KB)
    rc = tc_make_tmap_2d_h(&bh, (const __half*)wcat_hi, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * 2, 64, 32);
    if (rc) return rc;
    rc = tc_make_tmap_2d_h(&bl, (const __half*)wcat_lo, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * 2, 64, 32);
    if (rc) return rc;
  } else {
    rc = tc_make_tmap_2d(&ah, g_hi, (uint64_t)kg, (uint64_t)v_pad, (uint64_t)kg * 4, 32, 128);
    if (rc) return rc;
    rc = tc_make_tmap_2d(&al, x3 ? g_lo : g_hi, (uint64_t)kg, (uint64_t)v_pad, (uint64_t)kg * 4, 32, 128);
    if (rc) return rc;
    rc = tc_make_tmap_2d(&bh, wcat_hi, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * 4, 32, 32, true);
    if (rc) return rc;
    rc = tc_make_tmap_2d(&bl, x3 ? wcat_lo : wcat_hi, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * 4, 32, 32, true);
    if (rc) return rc;
  }
  TcGemmParams p;
  p.inv_scale = half ? scales + 3 : nullptr;
  {
    static const int lbo_env = getenv("GCB_H_LBO") ? atoi(getenv("GCB_H_LBO")) : 0;   // experiments
    static const int sbo_env = getenv("GCB_H_SBO") ? atoi(getenv("GCB_H_SBO")) : 0;
    static const int ks_env = getenv("GCB_H_KSTEP") ? atoi(getenv("GCB_H_KSTEP")) : 0;
    p.h_lbo = lbo_env > 0 ? lbo_env : 4096;   // between 64-column atoms (32 rows x 128 bytes)
    p.h_sbo = sbo_env > 0 ? sbo_env : 1024;   // between 8-row groups
    p.h_kstep = ks_env > 0 ? ks_env : 2048;   // 16 rows
  }
  p.out = d_x; p.ld_out = F; p.rows_valid = v_src; p.cols_valid = F; p.accumulate = accumulate;
  p.n_mtiles = v_pad / 128;
  p.n_ntiles = (int)ceil_div64(F, GM_BN);
  p.n_kc = kg / 32;
  {
    // accumulation steps per TMEM accumulator and pass: chunks / 4 accumulators x (4 tf32 | 2 fp16) k-steps; measured
    // (tests/test_gpu_configs.py, randn signal): 3xTF32 stays inside 1e-5 at 164 steps (T = 128) and leaves it at
    // 328 (T = 256), so passes are cut at <= 192 steps
    const int steps = half ? p.n_kc / 2 : p.n_kc;
    int kparts = x3 ? (steps + 191) / 192 : 1;
    if (kparts > p.n_kc / 8) kparts = p.n_kc / 8;   // every pass feeds all four accumulators
    p.kparts = kparts < 1 ? 1 : kparts;
  }
  p.kc_per_group = T / 32;
  p.group_stride = F;
  const size_t stage_bytes = (size_t)(x3 ? 2 : 1) * (GM_STAGE_A + GM_STAGE_B);
  int ns = (int)((DX_SMEM_LIMIT - 2048) / stage_bytes);
  if (ns > 6) ns = 6;
  p.ns = ns;
  const size_t smem = 2048 + (size_t)ns * stage_bytes;
  const int n_tiles = p.n_mtiles * p.n_ntiles;
  int grid = n_tiles < 148 ? n_tiles : 148;
  if (yield_sms() && n_tiles > 148) {
    // a collective runs beside this GEMM (gcb_yield_sms): 270 tiles take two rounds on 148 CTAs and on 135 - the
    // smaller grid costs nothing and leaves whole SMs (227 KB of shared memory each) to the collective's CTAs
    const int rounds = (n_tiles + 147) / 148;
    grid = (n_tiles + rounds - 1) / rounds;
  }
  if (half) return launch_gemm<true, true>(ah, al, bh, bl, p, grid, smem, st);
  return x3 ? launch_gemm<true>(ah, al, bh, bl, p, grid, smem, st) : launch_gemm<false>(ah, al, bh, bl, p, grid, smem, st);
}

// d_w_eff[T,R,A,F], d_w_self[T,F] (+)= G^T x X
int tc_dw_from_g(const float* g_hi, const float* g_lo, const float* x, int v_src, int R, int A, int F, int T,
                 bool x3, int accumulate, float* d_w_eff, float* d_w_self, float* x_hi, float* x_lo, float* partial,
                 size_t partial_bytes, cudaStream_t st, const float* scales, bool x_is_split) {
  const bool half = scales != nullptr;
  GCB_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)d_w_self % 16 == 0 && (uintptr_t)d_w_eff % 16 == 0 && F % 32 == 0 &&
                  T % 32 == 0,
              GCB_EINVAL, "tc_dw_from_g: operands must be 16-byte aligned, F and T multiples of 32");
  const TcGLayout L = tc_g_layout(v_src, R, A, F, T, x3, half);
  GCB_REQUIRE((size_t)L.vsplit * L.m_pad * F * sizeof(float) <= partial_bytes, GCB_EWORKSPACE,
              "tc_dw_from_g: %d vertex splits of the partial sums do not fit the %zu bytes provided", L.vsplit,
              partial_bytes);
  CUtensorMap ah, al, bh, bl;
  int rc;
  if (half) {
    const int x_ld = (int)ceil_div64(F, 64) * 64, kg_ld = (int)ceil_div64(L.kg, 64) * 64;
    __half* xh = (__half*)x_hi;
    __half* xl = xh + (size_t)v_src * x_ld;
    if (!x_is_split) {   // (k_bw_prepare has already written the split signal in this layout)
      const int64_t n = (int64_t)v_src * (x_ld / 8);
      k_split_half<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, v_src, F, x_ld, scales + 2, xh, xl);
      GCB_LAUNCH_OK();
    }
    rc = tc_make_tmap_mn_h(&ah, (const __half*)g_hi, (uint64_t)L.v_pad, (uint64_t)kg_ld, (uint64_t)kg_ld * 2, 32, 2);
    if (rc) return rc;
    rc = tc_make_tmap_mn_h(&al, (const __half*)g_lo, (uint64_t)L.v_pad, (uint64_t)kg_ld, (uint64_t)kg_ld * 2, 32, 2);
    if (rc) return rc;
    rc = tc_make_tmap_mn_h(&bh, xh, (uint64_t)v_src, (uint64_t)x_ld, (uint64_t)x_ld * 2, 32, (uint32_t)(L.bn / 64));
    if (rc) return rc;
    rc = tc_make_tmap_mn_h(&bl, xl, (uint64_t)v_src, (uint64_t)x_ld, (uint64_t)x_ld * 2, 32, (uint32_t)(L.bn / 64));
    if (rc) return rc;
  } else {
    const int64_t n4 = (int64_t)v_src * F / 4;
    k_split_tf32<<<(unsigned)ceil_div64(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n4,
                                                                reinterpret_cast<float4*>(x_hi),
                                                                x3 ? reinterpret_cast<float4*>(x_lo) : nullptr);
    GCB_LAUNCH_OK();
    rc = tc_make_tmap_mn(&ah, g_hi, (uint64_t)L.v_pad, (uint64_t)L.kg, (uint64_t)L.kg * 4, TN_KCH, 4);
    if (rc) return rc;
    rc = tc_make_tmap_mn(&al, x3 ? g_lo : g_hi, (uint64_t)L.v_pad, (uint64_t)L.kg, (uint64_t)L.kg * 4, TN_KCH, 4);
    if (rc) return rc;
    rc = tc_make_tmap_mn(&bh, x_hi, (uint64_t)v_src, (uint64_t)F, (uint64_t)F * 4, TN_KCH, (uint32_t)(L.bn / 32));
    if (rc) return rc;
    rc = tc_make_tmap_mn(&bl, x3 ? x_lo : x_hi, (uint64_t)v_src, (uint64_t)F, (uint64_t)F * 4, TN_KCH,
                         (uint32_t)(L.bn / 32));
    if (rc) return rc;
  }
  TcGemmTnParams p;
  p.inv_scale = half ? scales + 4 : nullptr;
  {
    static const int lbo_env = getenv("GCB_H_LBO") ? atoi(getenv("GCB_H_LBO")) : 0;   // experiments
    static const int sbo_env = getenv("GCB_H_SBO") ? atoi(getenv("GCB_H_SBO")) : 0;
    static const int ks_env = getenv("GCB_H_KSTEP") ? atoi(getenv("GCB_H_KSTEP")) : 0;
    p.h_lbo = lbo_env > 0 ? lbo_env : 4096;
    p.h_sbo = sbo_env > 0 ? sbo_env : 1024;
    p.h_kstep = ks_env > 0 ? ks_env : 2048;
  }
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
  if (half) rc = launch_gemm_tn<true, true>(ah, al, bh, bl, p, grid, smem, st);
  else rc = x3 ? launch_gemm_tn<true>(ah, al, bh, bl, p, grid, smem, st) : launch_gemm_tn<false>(ah, al, bh, bl, p, grid, smem, st);
  if (rc) return rc;
  const int64_t n_out = (int64_t)T * (R * A + 1) * (F / 4);
  k_tc_dwg_finish<<<(unsigned)ceil_div64(n_out, 256), 256, 0, st>>>(partial, L.vsplit, L.m_pad, T, R * A, F, accumulate,
                                                                    d_w_eff, d_w_self);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

}  // namespace gcb
