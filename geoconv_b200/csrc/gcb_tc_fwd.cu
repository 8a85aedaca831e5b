// tcgen05 forward of ConvIntrinsic (GCB_PREC_TF32 / GCB_PREC_TF32X3), fused with AngularMaxPooling.
//
//   Y[k,i,t] = act( sum_{x,y,f} Weff[t,x,(y+o_i)%A,f] * I[k,x,y,f] + X[k]·Wself[t] + b[t] )
// is a GEMM with M = vertices, N = (rotation, template), K = (x, y, f) plus one centre slab (f).
// The rotation never moves data: for a fixed (x, f-chunk) "super-step" all A angular weight tiles
// sit in a shared-memory ring and rotation i pairs the gathered tile y with weight tile (y+o_i)%A,
// so every operand tile is staged once and used g times.
//
// Kernel 1  k_tc_conv: one CTA = 128 vertices x g rotations x n_t templates x one K split.
//   warp 0     TMA: weight tiles (cp.async.bulk.tensor, hardware swizzle) into the ring.
//   warps 1-2  MMA: one thread per warp issues tcgen05.mma kind::tf32 into the TMEM accumulators of the
//              units it owns (g*n_t columns in total) and releases operand slots with tcgen05.commit.
//              Two issuers because the tensor pipe queues only ~2 MMAs: while one warp is in its
//              barrier waits / bookkeeping the other keeps the pipe fed.
//   warps 3-10 producers.  Every warp fills its rows of the tiles its group owns: barycentric 3-neighbour gather
//              of signal rows (float4 loads of the L2-resident signal; the packed records of the
//              current radial ring are cached in smem), weighted sum in registers, rounded to tf32
//              (RN), written straight into the swizzled K-major UMMA layout.  The loads of tile
//              n+D are issued before tile n is combined and stored (register double/triple
//              buffering), so a warp always has gathers in flight; measured before this change a
//              warp's load -> combine -> store -> arrive chain (~2300 cycles) set the tile rate.
//              At the end the producers drain TMEM to the split's partial-sum buffer.
//   Merged rotations: with N = n_t = 96 an MMA reads 4 KB of A and 3 KB of B from shared memory in
//   48 tensor cycles - 149 B/cycle against the 128 B/cycle an SM can deliver - so the unmerged
//   kernel was shared-memory bound (3xTF32: 3060 cycles per tile measured vs 3062 predicted from
//   bytes).  Rotations whose offsets step uniformly (o_{i+1} = o_i + d) read weight tiles that
//   are d apart on the ring; the ring is therefore laid out cycle by cycle of "+d", with the head of
//   each cycle duplicated behind its tail, so that m consecutive rotations always find their m
//   tiles ADJACENT and are issued as ONE MMA of N = m*n_t <= 256: the A tile is read once per m
//   rotations (104 B/cycle at m = 2, n_t = 96).
//   The contraction is split over CTAs (split-K) for two reasons: 108 tiles do not fill 148 SMs at
//   the FAUST shape, and the tensor core accumulates with truncation, so a chain of thousands of
//   accumulations loses ~1e-5 relative; short chains + an fp32 (RN) sum of the partials do not.
// Kernel 2  k_tc_finish: per vertex, sums the K-split partials in a fixed order, adds bias,
//   applies the activation, writes Y, and in the same pass does the angular max-pooling
//   (norm over templates, first-max argmax over rotations, row select -> Z, argmax).
//
// TF32X3 (fp32-faithful): operands are split v = hi + lo (both tf32) and D += Ahi*Bhi + Ahi*Blo +
// Alo*Bhi - three MMAs per k-step, relative error ~2^-21 per product.
//
// F16X3 (fp32-faithful at twice the tensor rate): fp16 has the same 11-bit significand as tf32 and
// kind::f16 contracts 16 elements per instruction where kind::tf32 contracts 8.  Both operands are
// multiplied by a power of two chosen per call from max|X| * max(sum_j |w_j|) and max|Wcat| so that the
// largest element sits in [2^14, 2^15) (no overflow; what falls into the fp16 subnormal range keeps an
// ABSOLUTE error of 2^-25, i.e. 2^-39 of the largest element), split v*s = hi + lo in fp16, contracted
// with the same three products, and the finishing kernel multiplies the sum by the exact inverse scale.
// Shared-memory layouts are identical byte for byte (a row of KC 4-byte units now holds 2*KC features).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include <type_traits>

#include "gcb_common.cuh"
#include "gcb_ptx.cuh"
#include "gcb_tc.cuh"

namespace gcb {

constexpr int TC_M = 128;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_MMA_WARPS = 2;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 2 + TC_MMA_WARPS) * 32;
// Warp roles.  The two control warps come FIRST: measured (clock64 traces), the warp scheduler favours
// the oldest (lowest-numbered) ready warp, and with the producers in front the TMA warp needed
// ~5000 cycles for ~60 instructions whenever the producers of its scheduler were busy.
constexpr int TC_WARP_TMA = 0;
constexpr int TC_WARP_MMA = 1;
constexpr int TC_WARP_SCOUT = TC_WARP_MMA + TC_MMA_WARPS;   // collects the landing barriers of a tile into one
constexpr int TC_WARP_PROD0 = TC_WARP_SCOUT + 1;
constexpr int TC_READY_RING = 16;   // tile-ready barriers (> the A ring depth, so the scout cannot lap a waiter)
constexpr int TC_WARP_ROWS = TC_M / TC_PRODUCER_WARPS;   // tile rows one producer warp fills
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_G_MAX = 8;              // rotations per CTA (accumulator groups) upper bound
constexpr int TC_OWN = TC_G_MAX;   // MMA units one issuing warp can own (all of them when it works alone)
constexpr int TC_REC_BYTES = TC_M * 32;  // one (x,y) slot of packed records for a 128-vertex tile
// TF32X3: contraction length per accumulator chain.  The tensor core accumulates in fp32 with truncation, so
// the error grows with the chain: at the FAUST shape with a randn signal (heavy cancellation) the forward's
// normalised error against float64 is 1.27e-5 / 9.4e-6 / 8.5e-6 / 5.8e-6 for 8 / 10 / 12 / 16 K splits
// (tools/prec_check.py); 1920 gives 12 splits there, inside the 1e-5 bar, for 2 % of forward time.
constexpr int TC_X3_CHAIN_K = 1920;
// F16X3: kind::f16 sums 16 products per instruction, so a chain of the same element count has half the
// accumulations; measured at the FAUST shape (tile-balanced splits): 8 splits 5.9e-6 (randn) / 2.0e-6 (SHOT-shaped),
// 12 splits 4.3e-6 / 1.3e-6, 6 splits 9.2e-6 / 2.6e-6.
constexpr int TC_H3_CHAIN_K = 2880;
constexpr uint8_t TC_NO_DUP = 0xFF;

// Where the A angular weight tiles of a super-step live in the shared-memory ring.
//   m      rotations merged into one MMA (1 = every rotation issues its own N = n_t MMAs)
//   dstep  offset step between consecutive rotations (mod A) when m > 1
//   pos[s] ring position of angular tile s; dup[s] a second position that receives a copy of it
//   tile_of[q]  the angular tile whose data sits at position q (main position or copy)
//   shift  odd super-steps use the same layout `shift` positions further on (n_pos + shift positions in
//          total): the entries of the next super-step then land in positions that were vacated several
//          steps earlier instead of the ones their namesakes still occupy
// All tile indices here are relative to the CTA's first rotation (tile s of the CTA = (s - rel0) mod A).
struct TcRing {
  int m, dstep, n_pos, shift;
  uint8_t pos[GCB_MAX_ROT], dup[GCB_MAX_ROT], tile_of[2 * GCB_MAX_ROT];
};

struct TcFwdParams {
  const float* x;       // [v_src, F]
  const float* x_centre;  // row 0 of the centre term: output row k multiplies Wself with x_centre[k] (= x + row0 * F)
  const uint4* recs;    // [R*A, v_pad, 2]: {i0,i1,i2,-} {w0,w1,w2,-}
  float* partial;       // [ksplit, v_pad, n_rot, T] raw accumulators
  float* interp_out;    // optional [v_out, R*A*F]: the interpolations, kept for the backward's dW
  const float* scales;  // F16X3: {scale of the gathered operand, scale of the weights, 1 / (product)}
  int v_out, v_pad, R, A, F, T, n_rot;
  int nfc;              // F / (features per tile row: KC, or 2*KC with fp16 operands)
  int n_t;              // templates per CTA (columns of one rotation's accumulator)
  int g;                // rotations per CTA
  int ns_a;             // A-tile ring depth
  int tmem_cols;        // power of two >= g*n_t
  int ksplit;           // K splits (grid.z = t_chunks * ksplit)
  int phase_shift;      // angular visiting order of super-step k starts at (k * phase_shift) % A
  int phase_alt;        // 1: the phase alternates 0, phase_shift, 0, ... with the two ring layouts instead of accumulating
  int bar_bytes;        // size of the mbarrier block in shared memory
  int tab_bytes;        // size of one rotation group's ring schedules (k_tc_build_tabs)
  const uint8_t* tabs;  // [n_groups][tab_bytes]
  const int* splits;    // [ksplit + 1] first super-step of every K split (tc_split_table)
  int drain_cw;         // accumulator columns staged per bulk store in the drain (0: registers -> global directly)
  int split_reverse;    // 1: K splits are dispatched last-to-first
  // fused finish (fuse = 1): the K splits of a (vertex tile, template chunk, rotation group) unit are neighbours in
  // dispatch order; the split that finishes last sums the unit's partial sums (still L2-resident) in split order and
  // applies bias / activation, the unit that finishes a vertex tile last does the angular max-pooling
  int fuse;              // 1: fused finish; 2: the same dispatch order, reduction left to k_tc_finish (A/B of the order)
  unsigned* tickets;     // [m_tiles * t_chunks * n_groups] arrivals per unit, then [m_tiles] finished units per tile (zeroed)
  float* nrm2;           // [v_pad][n_rot][t_chunks] squared norms of the activated rows, per template chunk
  const float* bias;
  float* y;              // [v_out, n_rot, T]
  float* z;              // nullable [v_out, T]
  int32_t* argmax;       // nullable [v_out]
  int act;
  int discard;           // 1: partial-sum lines are discarded from L2 once summed (never written back to DRAM)
  long long* trace;     // GCB_TC_TRACE (investigation builds): clock64 stamps of CTA (1,0,0), [role][tile][8]
  long long* cta_log;   // GCB_TC_TRACE=2: per CTA {globaltimer at entry, at exit, clock64 at entry, at exit, SM id}
  int debug;            // GCB_TC_DEBUG bits (timing experiments only): 1 = all gathers read row 0, 2 = no MMAs
  TcRing ring;
  RotTable rot;         // offsets in [0, A)
};

struct TcConfig {
  int kc, n_t, g, n_groups, ns_a, tmem_cols;
  TcRing ring;
  size_t smem_bytes;
  bool ok;
};

static int gcd_int(int a, int b) {
  while (b) { int t = a % b; a = b; b = t; }
  return a;
}

// Ring layout for rotation lists that step uniformly by d (mod A): the cycles of "+d" on Z_A are
// laid out one after the other, each followed by copies of its first m-1 tiles, so that the tiles
// s, s+d, ..., s+(m-1)d of m consecutive rotations are always at adjacent positions.
static TcRing tc_ring_layout(int A, int dstep, int m) {
  TcRing r{};
  r.m = m; r.dstep = m > 1 ? dstep : 0;
  for (int s = 0; s < GCB_MAX_ROT; ++s) { r.pos[s] = (uint8_t)(s < A ? s : 0); r.dup[s] = TC_NO_DUP; }
  r.n_pos = A;
  for (int s = 0; s < 2 * GCB_MAX_ROT; ++s) r.tile_of[s] = (uint8_t)(s < A ? s : 0);
  if (m <= 1) return r;
  const int gc = gcd_int(dstep, A), L = A / gc;
  int at = 0;
  for (int c = 0; c < gc; ++c) {
    for (int k = 0; k < L; ++k) { r.tile_of[at] = (uint8_t)((c + k * dstep) % A); r.pos[(c + k * dstep) % A] = (uint8_t)at++; }
    for (int k = 0; k < m - 1; ++k) { r.tile_of[at] = (uint8_t)((c + k * dstep) % A); r.dup[(c + k * dstep) % A] = (uint8_t)at++; }
  }
  r.n_pos = at;
  return r;
}

// uniform step of a rotation list (0 = not uniform / single rotation)
static int tc_uniform_step(const RotTable& rot, int n_rot, int A) {
  if (n_rot < 2) return 0;
  const int d = ((rot.off[1] - rot.off[0]) % A + A) % A;
  if (d == 0) return 0;
  for (int i = 1; i + 1 < n_rot; ++i)
    if (((rot.off[i + 1] - rot.off[i]) % A + A) % A != d) return 0;
  return d;
}

// per-phase weight-tile schedules in shared memory: need / retire masks (uint32) and first / last steps (uint8)
static bool tc_use_pairs() {
  static const bool use_pairs = !(getenv("GCB_TC_CLUSTER") && atoi(getenv("GCB_TC_CLUSTER")) == 0);
  return use_pairs;
}

// schedules of the weight ring in shared memory, per phase (and super-step parity), in ring-POSITION space
static size_t tc_tab_bytes(int A, int n_pos) {
  return align_up((size_t)A * A * 8 + 8 * A + (size_t)A * n_pos * 2 + (size_t)(2 * A + 1) * n_pos + (size_t)2 * A * n_pos + 64, 16);
}
// mbarriers: A-stage full/empty (<= 8 stages), per (parity, position) full, per absolute position empty,
// accumulators; TMEM slot
static size_t tc_bar_bytes(int n_pos, int n_ring) {
  return align_up((size_t)8 * (16 + 2 * n_pos + n_ring + 1 + TC_READY_RING) + 16 + 16 /* fused finish */, 64);
}

// Refill slack of the weight ring for a visiting order that advances `phase_shift` steps per super-step and a
// ring that alternates between two layouts `ring.shift` positions apart: the smallest number of whole steps
// between the last reader of a position in super-step k and the first reader of the entry that replaces it in
// super-step k+1, over every rotation group and both parities.  A refill (commit -> TMA -> barrier) takes
// ~2000 cycles, i.e. two to three steps; with one layout no visiting order leaves more than one step when a
// CTA owns A/2 rotations.  Entries are ring positions (a tile's copy is loaded on its own).
static int tc_refill_slack(const RotTable& rot, int n_rot, int A, int g, int n_groups, bool pairs, const TcRing& ring,
                           int phase_shift, bool alt) {
  int slack = 4 * A;
  for (int grp = 0; grp < n_groups; ++grp) {
    const int rot0 = grp * g;
    const int g_cnt = g < n_rot - rot0 ? g : n_rot - rot0;
    const int anchor = rot.off[(pairs ? (grp & ~1) : grp) * g];
    const int rel0 = ((rot.off[rot0] - anchor) % A + A) % A;
    int first[32], last_prev[32], last_cur[32];
    auto reads = [&](int phase, int* fo, int* lo) {
      for (int q = 0; q < 32; ++q) { fo[q] = A; lo[q] = -1; }
      for (int y = 0; y < A; ++y)
        for (int u = 0; u * ring.m < g_cnt; ++u) {
          const int rel = rot.off[rot0 + u * ring.m] - anchor;
          const int s0 = ((y + phase + rel - rel0) % A + A) % A;
          const int wd = ring.m < g_cnt - u * ring.m ? ring.m : g_cnt - u * ring.m;
          for (int i = 0; i < wd; ++i) {
            const int q = ring.pos[s0] + i;
            if (fo[q] > y) fo[q] = y;
            lo[q] = y;
          }
        }
    };
    // walk the sequence of super-steps (parity alternates; the phase accumulates or alternates) until it repeats
    int phase = 0;
    reads(phase, first, last_prev);
    for (int k = 0; k < 2 * A; ++k) {
      const int par = k & 1;                     // parity of super-step k; k + 1 has the other
      const int phase_n = alt ? (par ? 0 : phase_shift % A) : (phase + phase_shift) % A;
      reads(phase_n, first, last_cur);
      for (int q = 0; q < ring.n_pos; ++q) {     // entry q of super-step k + 1
        if (last_cur[q] < 0) continue;           // never read: never loaded
        const int qp = q + (par ^ 1) * ring.shift - par * ring.shift;   // the same position in k's numbering
        const int ret = (qp >= 0 && qp < ring.n_pos) ? last_prev[qp] : -1;
        const int free_steps = (ret < 0 ? A : A - 1 - ret) + first[q];
        slack = free_steps < slack ? free_steps : slack;
      }
      for (int q = 0; q < 32; ++q) last_prev[q] = last_cur[q];
      phase = phase_n;
    }
  }
  return slack;
}

// best visiting-order advance for a given ring (ties: the smallest)
static int tc_pick_phase_shift(const RotTable& rot, int n_rot, int A, int g, int n_groups, bool pairs,
                               const TcRing& ring, bool alt, int* slack_out) {
  int best = 0, best_slack = -1;
  for (int shift = 0; shift < A; ++shift) {
    const int sl = tc_refill_slack(rot, n_rot, A, g, n_groups, pairs, ring, shift, alt);
    if (sl > best_slack) { best_slack = sl; best = shift; }
  }
  if (slack_out) *slack_out = best_slack;
  return best;
}

// dstep = 0: every rotation on its own (works for any rotation list)
// One candidate (template chunk n_t, tile row width kc): the rotation grouping, the widest rotation merge whose
// weight ring fits, the number of A stages.  ok = false if nothing fits in shared memory.
static TcConfig tc_try_config(int A, int T, int n_rot, bool x3, int dstep, int ring_shift, int n_t, int kc) {
  TcConfig c{};
  c.ok = false;
  const int nprod = x3 ? 2 : 1;
  int g_max = 512 / n_t;
  if (g_max > TC_G_MAX) g_max = TC_G_MAX;
  if (g_max > n_rot) g_max = n_rot;
  const int n_groups = (n_rot + g_max - 1) / g_max;
  const int g = (n_rot + n_groups - 1) / n_groups;
  int m_top = 1;
  if (dstep > 0) {
    const int L = A / gcd_int(dstep, A);
    while (m_top * 2 <= g && m_top * 2 * n_t <= 256 && m_top * 2 - 1 < L) m_top *= 2;
  }
  for (int m = m_top; m >= 1; m >>= 1) {
    TcRing ring = tc_ring_layout(A, dstep, m);
    if (ring.n_pos > 32) continue;   // position masks are 32-bit
    ring.shift = ring_shift;
    const bool pairs = tc_use_pairs() && n_groups % 2 == 0 && kc >= 16;   // a pair member stages half of the records
    size_t b_bytes = (size_t)(ring.n_pos + ring_shift) * nprod * n_t * kc * 4;
    size_t a_stage = (size_t)nprod * TC_M * kc * 4;
    size_t fixed = b_bytes + (size_t)A * TC_REC_BYTES / (pairs ? 2 : 1) + tc_bar_bytes(ring.n_pos, ring.n_pos + ring_shift) +
                   tc_tab_bytes(A, ring.n_pos) + 1024 /* alignment slack */;
    if (fixed + 4 * a_stage > (size_t)TC_SMEM_LIMIT) continue;
    int ns_a = (int)(((size_t)TC_SMEM_LIMIT - fixed) / a_stage);
    if (ns_a > 8) ns_a = 8;
    static const int ns_cap = getenv("GCB_TC_NSA") ? atoi(getenv("GCB_TC_NSA")) : 0;   // experiments
    if (ns_cap >= 3 && ns_a > ns_cap) ns_a = ns_cap;
    int cols = 32;
    while (cols < g * n_t) cols <<= 1;
    c.kc = kc; c.n_t = n_t; c.g = g; c.n_groups = n_groups; c.ns_a = ns_a; c.tmem_cols = cols;
    c.ring = ring;
    c.smem_bytes = fixed + (size_t)ns_a * a_stage;
    c.ok = true;
    return c;
  }
  return c;
}

// Tile row width FIRST, template chunk second: a wide layer (T = 256, 384) does not fit its weight ring with the widest
// rows AND the largest template chunk, and giving up the row width costs far more than giving up chunk width -
// measured at V = 6890, 256 -> 384: 1428 us with kc = 8 (n_t as large as possible) against 925 us with kc = 16 in
// the fp16-split mode, 963 against 714 us (kc = 16 / 32) in the TF32 mode (profiles/r02_notes.md, section 14).  Narrow
// rows halve the work per tile, so twice as many barrier rounds, ring entries and staging fences per FLOP.
// The same holds for A = 12 at T = 96 (n_t = 48 with kc = 16: forward 1356 us, n_t = 96 with kc = 8: 1635 us at R = 5).
// Chunks below 48 templates are only taken when nothing else fits.
static TcConfig tc_pick_config(int A, int F, int T, int n_rot, bool x3, bool half, int dstep, int ring_shift = 0) {
  TcConfig c{};
  c.ok = false;
  if (T % 16 != 0 || F % 8 != 0 || A > GCB_MAX_ROT || n_rot > GCB_MAX_ROT) return c;
  static const int kc_env = getenv("GCB_TC_KC") ? atoi(getenv("GCB_TC_KC")) : 0;   // experiments
  static const bool old_order = getenv("GCB_TC_NT_FIRST") && atoi(getenv("GCB_TC_NT_FIRST")) != 0;   // A/B: round-1 order
  const int nt_top = T < 256 ? T : 256;
  // kc = 4-byte units per tile row; fp16 rows hold 2 * kc features and stay <= 64 bytes so that a producer
  // lane keeps the same number of gathered float4 in flight as in the tf32 layouts
  const int kc_top = half ? 16 : 32;
  if (!old_order) {
    const int nt_floor = T < 48 ? T : 48;
    for (int kc = kc_top; kc >= 8; kc >>= 1) {
      if (F % (half ? 2 * kc : kc) != 0) continue;
      if (kc_env > 0 && kc != kc_env) continue;
      for (int n_t = nt_top; n_t >= nt_floor; n_t -= 16) {
        if (T % n_t != 0) continue;
        c = tc_try_config(A, T, n_rot, x3, dstep, ring_shift, n_t, kc);
        if (c.ok) return c;
      }
    }
  }
  for (int n_t = nt_top; n_t >= 16; n_t -= 16) {
    if (T % n_t != 0) continue;
    for (int kc = kc_top; kc >= 8; kc >>= 1) {
      if (F % (half ? 2 * kc : kc) != 0) continue;
      if (kc_env > 0 && kc != kc_env) continue;
      c = tc_try_config(A, T, n_rot, x3, dstep, ring_shift, n_t, kc);
      if (c.ok) return c;
    }
  }
  return c;
}

// K-split count: enough to keep 3xTF32 accumulation chains short, then the split that minimises a
// simple time model: whole waves of 148 CTAs, each paying its share of the tile MMAs plus a fixed
// prologue/drain, plus the partial-sum traffic every extra split adds (written by this kernel, read back
// by k_tc_finish).
static int tc_pick_ksplit(const TcConfig& cfg, int v_out, int R, int A, int F, int T, bool x3, bool half, int n_rot) {
  const int kf = half ? 2 * cfg.kc : cfg.kc;   // features per tile row
  const int n_ss = (R + 1) * (F / kf);
  const int64_t units = ceil_div64(v_out, TC_M) * cfg.n_groups * (T / cfg.n_t);
  static const int ks_env = getenv("GCB_TC_KSPLIT") ? atoi(getenv("GCB_TC_KSPLIT")) : 0;   // experiments
  if (ks_env > 0) return ks_env < n_ss ? ks_env : n_ss;
  int s_min = 1;
  if (x3) s_min = (int)ceil_div64((int64_t)(R * A + 1) * F, half ? TC_H3_CHAIN_K : TC_X3_CHAIN_K);
  if (s_min > n_ss) s_min = n_ss;
  if (units >= 148 * 6) return s_min;
  const double tile_cycles = (double)(cfg.kc / 8) * (x3 ? 3 : 1) * cfg.g * cfg.n_t / 2.0;   // MMA time of one tile
  const double work = ((double)R * A + 1.0) * (F / kf) * tile_cycles * 1.1;                 // whole K range of one unit
  const double cta_fixed = 22000.0;   // prologue + pipeline fill + drain + exit (traced: ~25 k cycles, part of it overlapped)
  const double split_cost = 2.0 * (double)ceil_div64(v_out, TC_M) * TC_M * n_rot * T * 4.0 / 6.5e12 * 1.7e9;
  int best = s_min;
  double best_t = 0.0;
  for (int s = s_min; s <= s_min + 12 && s <= n_ss; ++s) {
    const double waves = (double)ceil_div64(units * s, 148);
    const double t = waves * (work / s + cta_fixed) + s * split_cost;
    if (s == s_min || t < best_t) { best_t = t; best = s; }
  }
  return best;
}

// (a, b) * s = hi + lo with hi, lo fp16 pairs
__device__ __forceinline__ void split_half2(float a, float b, float s, uint32_t& hi, uint32_t& lo) {
  a *= s; b *= s;
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- preparation kernels ---------------------------------------------------------------------
// slot-major packed barycentric records, padded to a multiple of 128 vertices
// (wsum_max, optional: max over slots of |w0| + |w1| + |w2| as fp32 bits - bounds |interpolation| / max|X|)
__global__ void k_tc_pack_recs(const int32_t* __restrict__ idx, const float* __restrict__ w, int v_out,
                               int v_pad, int RA, uint4* __restrict__ recs, unsigned* __restrict__ wsum_max) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t total = (int64_t)RA * v_pad;
  float ws = 0.f;
  if (e < total) {
    int slot = (int)(e / v_pad), v = (int)(e - (int64_t)slot * v_pad);
    uint4 ri = make_uint4(0, 0, 0, 0), rw = make_uint4(0, 0, 0, 0);
    if (v < v_out) {
      size_t base = ((size_t)v * RA + slot) * 3;
      ri = make_uint4((uint32_t)idx[base], (uint32_t)idx[base + 1], (uint32_t)idx[base + 2], 0);
      rw = make_uint4(__float_as_uint(w[base]), __float_as_uint(w[base + 1]), __float_as_uint(w[base + 2]), 0);
      ws = fabsf(w[base]) + fabsf(w[base + 1]) + fabsf(w[base + 2]);
    }
    recs[e * 2] = ri;
    recs[e * 2 + 1] = rw;
  }
  if (wsum_max) {   // (max of non-negative floats = max of their bit patterns; order-independent)
    unsigned b = __float_as_uint(ws);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(wsum_max, b);
  }
}

// max |a[i]| over two arrays, as fp32 bits (a NaN gives a pattern above +inf: the caller sees "not finite")
__global__ void k_tc_absmax(const float* __restrict__ a, int64_t na, const float* __restrict__ b2, int64_t nb,
                            unsigned* __restrict__ out) {
  unsigned m = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < na; i += stride)
    m = max(m, __float_as_uint(fabsf(__ldg(a + i))));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += stride)
    m = max(m, __float_as_uint(fabsf(__ldg(b2 + i))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__device__ __forceinline__ int tc_pow2_for(float bound) {   // bound * 2^e in [2^14, 2^15); 0 for a zero / non-finite bound
  if (!(bound > 0.f) || !isfinite(bound)) return 0;
  const int e = 14 - ilogbf(bound);
  return e < -60 ? -60 : (e > 60 ? 60 : e);
}

// 0x7FC00000 if any component is a NaN (fmaxf would silently drop it), else 0
__device__ __forceinline__ unsigned tc_nan_bits(const float4 v) {
  return (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) ? 0x7FC00000u : 0u;
}

// Operand maxima and the power-of-two scales of the fp16 modes in ONE launch: every block strides over every
// segment (max |a[i]|, or the longest CSR row for a row-pointer segment) and atomicMax'es its results into mx[slot];
// the block that finishes last (ticket counter mx[7]) turns the maxima into scales.
//   mode 1 (forward):  mx = {max|X|, max sum_j|w_j| (from k_tc_pack_recs), max|Wcat|}
//                      scales = {sA, sB, 1 / (sA sB)},  sA from max|X| * max(sum|w|, 1)
//   mode 2 (backward): mx = {max|h|, max|w|, longest CSR row, max|Wcat|, max|X|}
//                      scales = {sG, sW, sX, 1 / (sG sW), 1 / (sG sX)},  |G| <= max|h| * max|w| * longest row
// (maxima of non-negative floats are compared as bit patterns; a NaN compares above +inf and gives scale 1)
__global__ void k_tc_maxima(const TcMaxJob job, unsigned* __restrict__ mx, float* __restrict__ scales) {
  __shared__ unsigned s_max;
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_max = 0u;
  __syncthreads();
  // blocks [blk0[sg], blk0[sg + 1]) work on segment sg (shares proportional to its size, set by the host): the
  // inputs come cold from HBM, so every thread keeps four 16-byte loads in flight
  int sg = 0;
#pragma unroll
  for (int i = 1; i < TC_MAX_SEGS; ++i) sg += (i < job.n_seg && (int)blockIdx.x >= job.blk0[i]) ? 1 : 0;
  const int nb = job.blk0[sg + 1] - job.blk0[sg], lb = (int)blockIdx.x - job.blk0[sg];
  const int64_t n = job.seg[sg].n, stride = (int64_t)nb * blockDim.x, i0 = (int64_t)lb * blockDim.x + threadIdx.x;
  unsigned m = 0;
  if (job.seg[sg].is_rowptr) {
    const int32_t* rp = static_cast<const int32_t*>(job.seg[sg].ptr);
    for (int64_t i = i0; i < n; i += stride) m = max(m, (unsigned)(rp[i + 1] - rp[i]));
  } else {
    const float* a = static_cast<const float*>(job.seg[sg].ptr);
    int64_t done = 0;
    if ((reinterpret_cast<uintptr_t>(a) & 15) == 0) {
      const float4* a4 = reinterpret_cast<const float4*>(a);
      const int64_t n4 = n >> 2;
      int64_t i = i0;
      for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 v0 = __ldg(a4 + i), v1 = __ldg(a4 + i + stride), v2 = __ldg(a4 + i + 2 * stride),
                     v3 = __ldg(a4 + i + 3 * stride);
        const float q0 = fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w)));
        const float q1 = fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)));
        const float q2 = fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w)));
        const float q3 = fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w)));
        // (bit patterns, so that a NaN is not dropped the way fmaxf drops it)
        m = max(m, max(max(__float_as_uint(fabsf(v0.x)), __float_as_uint(q0)), __float_as_uint(q1)));
        m = max(m, max(__float_as_uint(q2), __float_as_uint(q3)));
        m = max(m, tc_nan_bits(v0) | tc_nan_bits(v1) | tc_nan_bits(v2) | tc_nan_bits(v3));
      }
      for (; i < n4; i += stride) {
        const float4 v = __ldg(a4 + i);
        m = max(m, max(max(__float_as_uint(fabsf(v.x)), __float_as_uint(fabsf(v.y))),
                       max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
      }
      done = n4 << 2;
    }
    for (int64_t i = done + i0; i < n; i += stride) m = max(m, __float_as_uint(fabsf(__ldg(a + i))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_max, m);
  __syncthreads();
  if (threadIdx.x == 0 && s_max) {
    atomicMax(mx + job.seg[sg].slot, s_max);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(mx + 7, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0 || job.mode == 0) return;   // mode 0: maxima only
  __threadfence();
  volatile const unsigned* vm = mx;
  if (job.mode == 1) {
    if (job.ext_wsum) mx[1] = __ldg(job.ext_wsum);   // cached record packing: its max sum_j |w_j| was computed then
    const float xm = __uint_as_float(vm[0]), wm = fmaxf(__uint_as_float(vm[1]), 1.f), bm = __uint_as_float(vm[2]);
    const int ea = tc_pow2_for(xm * wm), eb = tc_pow2_for(bm);
    scales[0] = ldexpf(1.f, ea);
    scales[1] = ldexpf(1.f, eb);
    scales[2] = ldexpf(1.f, -(ea + eb));
  } else {
    const float hm = __uint_as_float(vm[0]);
    const float gm = job.have_csr ? fmaxf(hm * __uint_as_float(vm[1]) * (float)vm[2], hm) : hm;
    const int eg = tc_pow2_for(gm), ew = tc_pow2_for(__uint_as_float(vm[3])), ex = tc_pow2_for(__uint_as_float(vm[4]));
    scales[0] = ldexpf(1.f, eg);
    scales[1] = ldexpf(1.f, ew);
    scales[2] = ldexpf(1.f, ex);
    scales[3] = ldexpf(1.f, -(eg + ew));
    scales[4] = ldexpf(1.f, -(eg + ex));
  }
}

// fp16 variant of k_tc_wcat: Wcat * (*scales) = hi + lo (both fp16); a thread converts 8 consecutive elements of a
// row (R*A*F and F are multiples of 8, so a group never straddles [Weff | Wself]) with 16-byte loads and stores
__global__ void k_tc_wcat_h(const float* __restrict__ w_eff, const float* __restrict__ w_self, int T, int KRA,
                            int F, const float* __restrict__ scales, __half* __restrict__ hi,
                            __half* __restrict__ lo) {
  const int64_t e8 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int kcat = KRA + F, k8n = kcat >> 3;
  if (e8 >= (int64_t)T * k8n) return;
  const int t = (int)(e8 / k8n), k = (int)(e8 - (int64_t)t * k8n) << 3;
  const float* src = k < KRA ? w_eff + (size_t)t * KRA + k : w_self + (size_t)t * F + (k - KRA);
  const float s = __ldg(scales);
  float v[8];
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(src + j);
  }
  uint4 h, l;
  split_half2(v[0], v[1], s, h.x, l.x);
  split_half2(v[2], v[3], s, h.y, l.y);
  split_half2(v[4], v[5], s, h.z, l.z);
  split_half2(v[6], v[7], s, h.w, l.w);
  const size_t o = (size_t)t * kcat + k;   // multiple of 8 halves: 16-byte aligned
  *reinterpret_cast<uint4*>(hi + o) = h;
  *reinterpret_cast<uint4*>(lo + o) = l;
}

// Wcat[t, 0:KRA] = Weff[t], Wcat[t, KRA:KRA+F] = Wself[t], split into tf32 hi (+ lo for X3); a thread converts 4
// consecutive elements of a row (R*A*F and F are multiples of 4 on the tensor path, so a group never straddles
// [Weff | Wself]) with 16-byte accesses (one element per thread and a division per element took 9 us per call)
__global__ void k_tc_wcat(const float* __restrict__ w_eff, const float* __restrict__ w_self, int T, int KRA,
                          int F, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t e4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int kcat = KRA + F, k4n = kcat >> 2;
  if (e4 >= (int64_t)T * k4n) return;
  const int t = (int)(e4 / k4n), k = (int)(e4 - (int64_t)t * k4n) << 2;
  const float* src = k < KRA ? w_eff + (size_t)t * KRA + k : w_self + (size_t)t * F + (k - KRA);
  float4 v;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) v = __ldg(reinterpret_cast<const float4*>(src));
  else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
  float4 h;
  h.x = ptx::to_tf32(v.x); h.y = ptx::to_tf32(v.y); h.z = ptx::to_tf32(v.z); h.w = ptx::to_tf32(v.w);
  const size_t o = (size_t)t * kcat + k;   // multiple of 4 floats: 16-byte aligned
  *reinterpret_cast<float4*>(hi + o) = h;
  if (lo) *reinterpret_cast<float4*>(lo + o) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// ---- kernel 1: the contraction ---------------------------------------------------------------
template <int KC>
__device__ __forceinline__ uint32_t swz_chunk(int row, int c) {
  if (KC == 32) return (uint32_t)(c ^ (row & 7));
  if (KC == 16) return (uint32_t)(c ^ ((row >> 1) & 3));
  return (uint32_t)(c ^ ((row >> 2) & 1));
}

// phase of the super-step after one with phase `phi`
__device__ __forceinline__ int tc_next_phase(const TcFwdParams& p, int phi) {
  if (p.phase_alt) return phi == 0 ? p.phase_shift : 0;
  const int n = phi + p.phase_shift;
  return n >= p.A ? n - p.A : n;
}

constexpr int TC_CTRL_THREADS = TC_WARP_PROD0 * 32;   // the control warps come first
__device__ __forceinline__ void ctrl_bar_sync() {
  asm volatile("bar.sync 2, %0;" ::"n"(TC_CTRL_THREADS) : "memory");
}
__device__ __forceinline__ void producer_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCER_WARPS * 32) : "memory");
}

// K-split boundaries: splits are balanced by TILES (a radial super-step has A tiles, a centre super-step one),
// cut at super-step boundaries, every split at least one super-step long (an empty split would drain an
// accumulator nobody wrote).  n_nc / n_c = radial / centre super-steps; out[0..ksplit], ksplit <= n_nc + n_c.
__device__ inline void tc_split_table(int n_nc, int n_c, int A, int ksplit, int* out) {
  const long long total = (long long)n_nc * A + n_c;
  const int n_ss = n_nc + n_c;
  int b = 0;
  out[0] = 0;
  for (int j = 1; j < ksplit; ++j) {
    const long long t = (total * j + ksplit / 2) / ksplit;   // target tile count before split j
    int cand = t >= (long long)n_nc * A ? n_nc + (int)(t - (long long)n_nc * A) : (int)((t + A / 2) / A);
    b = cand > b + 1 ? cand : b + 1;
    const int hi = n_ss - (ksplit - j);
    b = b < hi ? b : hi;
    out[j] = b;
  }
  out[ksplit] = n_ss;
}

// position in the CTA's tile sequence: super-step ss = (radial ring xr, feature chunk fc) or the
// centre slab, angular step yy inside it
struct TcTileIt {
  int ss, yy, xr, fc, n_y;
  bool centre;
  __device__ __forceinline__ void decode(const TcFwdParams& p) {
    centre = ss >= p.R * p.nfc;
    xr = centre ? p.R : ss / p.nfc;
    fc = ss - xr * p.nfc;
    n_y = centre ? 1 : p.A;
  }
  __device__ __forceinline__ void start(const TcFwdParams& p, int ss0) { ss = ss0; yy = 0; decode(p); }
  __device__ __forceinline__ void next(const TcFwdParams& p) {
    if (++yy == n_y) { yy = 0; ++ss; decode(p); }
  }
};

// Barrier / operand addresses the MMA warp works with
struct TcMmaCtx {
  uint32_t a_full, a_empty, b_full, b_empty, acc_full, ready;   // mbarrier addresses (arrays indexed * 8)
  uint32_t a_lo0, a_stage_lo;   // descriptor low word of A stage 0, distance between stages (>> 4)
  uint32_t b_lo0, b_tile_lo, b_lo_ring;   // same for ring position 0, one tile, hi ring -> lo ring
  uint32_t tmem_base;
  int rot0, g_cnt, n_units, off_anchor, rel0;
  int mw, n_mma;                // this issuing warp and how many are active: warp mw owns units mw, mw + n_mma, ...
  int n_nc, n_c;                // non-centre super-steps / centre tiles of this CTA
};

// The MMA-issuing warp.  Measured on B200 (tools/microbench/mma_issue.cu): ONE thread issues
// back-to-back tcgen05.mma at the full tensor rate (N = 192: 96.0 cycles per MMA, commits free), and a
// successful mbarrier try_wait costs ~80 cycles (sync_cost.cu).  What bounded every earlier version
// of this kernel was the issuing warp's own instruction stream: ~300-2000 cycles of integer / branch
// / shuffle bookkeeping per tile around 768 cycles of MMAs (a single warp retires a dependent
// instruction every ~8 cycles).  So this loop is kept minimal: NU (units) is a template parameter,
// everything is warp-uniform (the whole warp runs converged, only the tcgen05 instructions sit under
// elect.sync, so values live in uniform registers) and centre tiles have their own loop.
//
// Two measured stalls shape the loop (clock64 traces, FAUST shape):
//  * the tensor pipe queues only ~2 MMAs behind the one executing, so the ~300 cycles the thread
//    spent between the last MMA of a tile and the first of the next (commits, loop, two barrier
//    waits) left the pipe idle ~160-300 cycles per tile.  The barrier waits of tile n+1 now sit in
//    the MIDDLE of tile n's MMA list: they run while the first half executes, the second half and
//    the next tile's first MMAs then issue back to back.
//  * a weight tile can only be reloaded after its last reader; with the visiting order fixed, the
//    tiles read at the last step of a super-step (rotation windows that wrap) were also the first
//    ones needed by the next super-step: ~2000 cycles of exposed commit -> TMA -> barrier latency
//    per super-step.  The visiting order now starts `phase_shift` steps later every super-step
//    (chosen on the host so that every slot stays free for at least one step before its next use);
//    the per-phase first/last-use tables are built once per CTA in shared memory.
template <int KC, bool X3, bool H, int NU, int U0, int U1, int J0, int J1>
__device__ __forceinline__ void tc_issue_range(const uint64_t desc_hi, const uint32_t a_lo, const uint32_t (&b_lo)[NU],
                                               const uint32_t (&d_u)[NU], const uint32_t (&idesc_u)[NU],
                                               const bool (&on)[NU], const uint32_t b_lo_ring, const uint32_t accum) {
  constexpr int ROW_BYTES = KC * 4;
  constexpr uint32_t A_LO_HALF = (uint32_t)(TC_M * ROW_BYTES) >> 4;   // hi tile -> lo tile of an A stage
#pragma unroll
  for (int u = U0; u < U1; ++u) {
    if (on[u]) {
#pragma unroll
      for (int j = J0; j < J1; ++j) {
        const uint64_t da = desc_hi | (uint64_t)(a_lo + 2u * j);
        const uint64_t db = desc_hi | (uint64_t)(b_lo[u] + 2u * j);
        ptx::umma<H>(d_u[u], da, db, idesc_u[u], j == 0 ? accum : 1u);
        if (X3) {
          ptx::umma<H>(d_u[u], da, desc_hi | (uint64_t)(b_lo[u] + b_lo_ring + 2u * j), idesc_u[u], 1u);
          ptx::umma<H>(d_u[u], desc_hi | (uint64_t)(a_lo + A_LO_HALF + 2u * j), db, idesc_u[u], 1u);
        }
      }
    }
  }
}

// wait for the ring entries in `bits` (barrier set of one super-step parity); `seen` holds, per entry, the
// parity of the barrier phase the next wait is for
__device__ __forceinline__ void tc_wait_entries(uint32_t bar0, uint32_t bits, uint32_t& seen) {
  while (bits) {
    const uint32_t q = (uint32_t)(__ffs((int)bits) - 1);
    ptx::mbar_wait(bar0 + 8u * q, (seen >> q) & 1u);
    seen ^= 1u << q;
    bits &= bits - 1;
  }
}

template <int KC, bool X3, bool H, int CL, int NU>
__device__ __forceinline__ void tc_mma_role(const TcFwdParams& p, const TcMmaCtx& c, int lane,
                                            const uint32_t* need_tab, const uint32_t* retire_tab,
                                            const uint32_t* seam_tab) {
  constexpr int ROW_BYTES = KC * 4;
  constexpr int KS = KC / 8;
  constexpr uint32_t A_LO_HALF = (uint32_t)(TC_M * ROW_BYTES) >> 4;
  // the MMA list of a tile is issued in two halves around the next tile's barrier waits
  constexpr int U_SPLIT = NU >= 2 ? NU / 2 : NU;
  constexpr int J_SPLIT = NU >= 2 ? KS : (KS >= 2 ? KS / 2 : KS);
  const uint64_t desc_hi = ptx::make_kmajor_desc<ROW_BYTES>(0) & 0xFFFFFFFF00000000ull;
  const int A_ = p.A, n_t_ = p.n_t, ns_a_ = p.ns_a, m_ = p.ring.m, shift_ = p.phase_shift;
  const bool no_mma = (p.debug & 2) != 0;
  int sl[NU];              // angular weight tile the unit's first rotation reads at the current step
  uint32_t idesc_u[NU], d_u[NU];
  bool on[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int gu = c.mw + c.n_mma * u;   // unit index in the CTA
    on[u] = gu < c.n_units;
    const int wd = on[u] ? min(m_, c.g_cnt - gu * m_) : 1;
    int rel = on[u] ? p.rot.off[c.rot0 + gu * m_] - c.off_anchor - c.rel0 : 0;   // relative to the CTA's first rotation
    rel += rel < 0 ? A_ : 0;
    sl[u] = rel < 0 ? rel + A_ : rel;
    idesc_u[u] = ptx::make_idesc<H>(TC_M, wd * n_t_);
    d_u[u] = c.tmem_base + (uint32_t)(gu * m_ * n_t_);
  }
  const bool trace_on = p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && c.mw == 0;
  int stage = 0;
  uint32_t par = 0, accum = 0;
  int n = 0, phi = 0;
  // The per-step schedule of a super-step lives in the LANES of this warp: lane y holds what step y needs
  // (weight tiles to wait for, tiles to release, descriptor address words of the units' weight tiles), read
  // from shared / constant memory once per super-step; a tile then costs a few shuffles instead of dependent
  // memory reads in the issuing thread's critical path.
  uint32_t retire_l = 0, blo_l[NU];
  auto load_schedule = [&](int ph, uint32_t ring_off, const int (&s0)[NU], uint32_t& retire_o, uint32_t (&blo_o)[NU]) {
    const int yy = lane < A_ ? lane : 0;
    retire_o = retire_tab[ph * A_ + yy];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      int sg = s0[u] + yy;
      sg -= sg >= A_ ? A_ : 0;
      blo_o[u] = c.b_lo0 + ((uint32_t)p.ring.pos[sg] + ring_off) * c.b_tile_lo;
    }
  };
  // wait until the scout has seen everything tile `t` reads (one barrier per tile, ring of TC_READY_RING)
  auto wait_ready = [&](int t) {
    ptx::mbar_wait(c.ready + 8u * (uint32_t)(t & (TC_READY_RING - 1)), (uint32_t)(t / TC_READY_RING) & 1u);
    ptx::tc_fence_after();
  };
  uint32_t ring_cur = 0u, ring_nxt = (uint32_t)p.ring.shift;   // the two ring layouts alternate per super-step
  load_schedule(0, ring_cur, sl, retire_l, blo_l);
  uint32_t b_lo[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) b_lo[u] = __shfl_sync(0xffffffffu, blo_l[u], 0);
  const int n_tiles = c.n_nc * A_ + c.n_c;
  if (n_tiles > 0) wait_ready(0);   // every later tile is waited for one tile ahead
  for (int k = 0; k < c.n_nc; ++k) {
    const int phi_n = tc_next_phase(p, phi);
    // the next super-step's first tile may need an entry whose ring position this super-step only vacates with
    // the commit of its last tile: then that tile must be committed before the wait (else: deadlock)
    const bool seam_dep = (seam_tab[(k & 1) * A_ + phi] & need_tab[phi_n * A_]) != 0u;
    int sl_n[NU];            // the units' first weight tiles at step 0 of the next super-step
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      sl_n[u] = sl[u] + (phi_n - phi);   // the units' tiles move with the phase
      sl_n[u] += sl_n[u] < 0 ? A_ : 0;
      sl_n[u] -= sl_n[u] >= A_ ? A_ : 0;
    }
    uint32_t retire_nl = 0, blo_nl[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) blo_nl[u] = 0;
    for (int y = 0; y < A_; ++y, ++n) {
      long long* trp = trace_on && n < 128 ? p.trace + ((size_t)2 * 128 + n) * 8 : nullptr;
      if (trp) trp[0] = clock64();
      const uint32_t a_lo = c.a_lo0 + (uint32_t)stage * c.a_stage_lo;
      if (!no_mma && ptx::elect_one())
        tc_issue_range<KC, X3, H, NU, 0, U_SPLIT, 0, J_SPLIT>(desc_hi, a_lo, b_lo, d_u, idesc_u, on, c.b_lo_ring, accum);
      __syncwarp();
      // ---- while the first half executes: this step's releases, the next tile's addresses and readiness ----
      const uint32_t rm = __shfl_sync(0xffffffffu, retire_l, y);
      const bool last_y = y + 1 == A_;
      uint32_t b_lo_n[NU];
      if (!last_y) {
#pragma unroll
        for (int u = 0; u < NU; ++u) b_lo_n[u] = __shfl_sync(0xffffffffu, blo_l[u], y + 1);
      } else {
        if (k + 1 < c.n_nc) load_schedule(phi_n, ring_nxt, sl_n, retire_nl, blo_nl);
#pragma unroll
        for (int u = 0; u < NU; ++u) b_lo_n[u] = __shfl_sync(0xffffffffu, blo_nl[u], 0);
      }
      const bool wait_late = last_y && (seam_dep || k + 1 == c.n_nc);   // (centre tiles wait at their own top)
      if (n + 1 < n_tiles && !wait_late) wait_ready(n + 1);
      if (trp) trp[1] = clock64();
      if (ptx::elect_one()) {
        if (!no_mma) {
          if (U_SPLIT < NU)
            tc_issue_range<KC, X3, H, NU, U_SPLIT, NU, 0, KS>(desc_hi, a_lo, b_lo, d_u, idesc_u, on, c.b_lo_ring, accum);
          else if (J_SPLIT < KS)
            tc_issue_range<KC, X3, H, NU, 0, NU, J_SPLIT, KS>(desc_hi, a_lo, b_lo, d_u, idesc_u, on, c.b_lo_ring, accum);
        }
        uint32_t r = rm;
        while (r) {   // vacate the ring positions read for the last time
          ptx::umma_commit(c.b_empty + 8u * ((uint32_t)(__ffs((int)r) - 1) + ring_cur));
          r &= r - 1;
        }
        if (CL == 2) ptx::umma_commit_mc(c.a_empty + 8u * stage, (uint16_t)3);
        else ptx::umma_commit(c.a_empty + 8u * stage);
      }
      __syncwarp();
      if (n + 1 < n_tiles && wait_late && k + 1 < c.n_nc) wait_ready(n + 1);
      if (trp) trp[2] = clock64();
      accum = 1u;
#pragma unroll
      for (int u = 0; u < NU; ++u) b_lo[u] = b_lo_n[u];
      if (++stage == ns_a_) { stage = 0; par ^= 1u; }
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      sl[u] = sl_n[u];
      blo_l[u] = blo_nl[u];
    }
    retire_l = retire_nl;
    phi = phi_n;
    { const uint32_t t = ring_cur; ring_cur = ring_nxt; ring_nxt = t; }
  }
  // centre tiles: X[k] * Wself - one weight tile (the ring position of tile 0 in the current layout) serves
  // every rotation, one N = n_t MMA chain each; the layouts keep alternating
  const uint32_t idesc1 = ptx::make_idesc<H>(TC_M, n_t_);
  const uint32_t q_c = (uint32_t)p.ring.pos[0];
  for (int k = 0; k < c.n_c; ++k, ++n) {
    const uint32_t pos_c = q_c + ring_cur;
    const uint32_t b_lo_c = c.b_lo0 + pos_c * c.b_tile_lo;
    if (n > 0) wait_ready(n);
    const uint32_t a_lo = c.a_lo0 + (uint32_t)stage * c.a_stage_lo;
    if (ptx::elect_one()) {
      if (!no_mma) {
        for (int r = 0; r < c.g_cnt; ++r) {
          // a rotation's accumulator belongs to the warp that owns its unit in the main loop: MMAs of two
          // threads into the same accumulator would execute in an order that varies from run to run
          if ((r / m_) % c.n_mma != c.mw) continue;
          const uint32_t d = c.tmem_base + (uint32_t)(r * n_t_);
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const uint64_t da = desc_hi | (uint64_t)(a_lo + 2u * j);
            const uint64_t db = desc_hi | (uint64_t)(b_lo_c + 2u * j);
            ptx::umma<H>(d, da, db, idesc1, j == 0 ? accum : 1u);
            if (X3) {
              ptx::umma<H>(d, da, desc_hi | (uint64_t)(b_lo_c + c.b_lo_ring + 2u * j), idesc1, 1u);
              ptx::umma<H>(d, desc_hi | (uint64_t)(a_lo + A_LO_HALF + 2u * j), db, idesc1, 1u);
            }
          }
        }
      }
      ptx::umma_commit(c.b_empty + 8u * pos_c);
      if (CL == 2) ptx::umma_commit_mc(c.a_empty + 8u * stage, (uint16_t)3);
      else ptx::umma_commit(c.a_empty + 8u * stage);
    }
    __syncwarp();
    accum = 1u;
    if (++stage == ns_a_) { stage = 0; par ^= 1u; }
    { const uint32_t t = ring_cur; ring_cur = ring_nxt; ring_nxt = t; }
  }
  if (ptx::elect_one()) ptx::umma_commit(c.acc_full);   // accumulators complete
  __syncwarp();
}

// Weight-ring schedules of rotation group `grp` (they do not depend on the vertex tile or the K split): built
// once per launch by k_tc_build_tabs into global memory; every CTA copies its group's ~1 KB.  Layout:
//   need_tab / retire_tab [phase][step] u32: ring entries (positions of the even layout) first / last read at that step
//   seam_tab [parity][phase] u32: entries of the NEXT super-step whose position is only vacated by the last step
//   firstp / lastp [phase][entry] u8: first / last step at which the entry is read (0xFF: never - never loaded)
//   order_tab [parity * A + phase][i] u8: the i-th entry the TMA warp (re)loads (row 2A: the first super-step)
//   occ_tab [parity][phase][entry] u8: 1 + last step at which the previous super-step still reads the position
//   last 16 bytes: n_ent = entries that are read at all
__global__ void k_tc_build_tabs(const __grid_constant__ TcFwdParams p, int pairs, int tab_bytes,
                                uint8_t* __restrict__ tabs, int* __restrict__ splits) {
  extern __shared__ __align__(16) uint8_t tab_smem[];
  const int tid = threadIdx.x, nthr = blockDim.x, grp = blockIdx.x;
  if (grp == 0 && tid == nthr - 1) tc_split_table(p.R * p.nfc, p.nfc, p.A, p.ksplit, splits);
  const int NP = p.ring.n_pos, AA = p.A * p.A;
  uint32_t* need_tab = reinterpret_cast<uint32_t*>(tab_smem);
  uint32_t* retire_tab = need_tab + AA;
  uint32_t* seam_tab = retire_tab + AA;
  uint8_t* firstp_tab = reinterpret_cast<uint8_t*>(seam_tab + 2 * p.A);
  uint8_t* lastp_tab = firstp_tab + p.A * NP;
  uint8_t* order_tab = lastp_tab + p.A * NP;
  uint8_t* occ_tab = order_tab + (2 * p.A + 1) * NP;
  const int rot0 = grp * p.g;
  const int g_cnt = min(p.g, p.n_rot - rot0);
  const int n_units = (g_cnt + p.ring.m - 1) / p.ring.m;
  const int off_anchor = p.rot.off[(pairs ? (grp & ~1) : grp) * p.g];
  int rel0 = p.rot.off[rot0] - off_anchor;
  rel0 += rel0 < 0 ? p.A : 0;
  for (int i = tid; i < tab_bytes / 4; i += nthr) reinterpret_cast<uint32_t*>(tab_smem)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < p.A * NP; i += nthr) {
    // first / last reader of ring entry q under phase phi: a unit reads m adjacent positions starting at the
    // position of its first tile (tiles relabelled by this CTA's first rotation)
    const int phi = i / NP, q = i - phi * NP;
    int first = 0xFF, last = 0xFF;
    for (int y = 0; y < p.A; ++y)
      for (int u = 0; u < n_units; ++u) {
        int s0 = (y + phi + p.rot.off[rot0 + u * p.ring.m] - off_anchor - rel0) % p.A;
        s0 += s0 < 0 ? p.A : 0;
        const int wd = min(p.ring.m, g_cnt - u * p.ring.m);
        const int q0 = p.ring.pos[s0];
        if (q >= q0 && q < q0 + wd) {
          if (first == 0xFF) first = y;
          last = y;
        }
      }
    firstp_tab[i] = (uint8_t)first;
    lastp_tab[i] = (uint8_t)last;
  }
  __syncthreads();
  for (int i = tid; i < AA; i += nthr) {
    const int phi = i / p.A, yy = i - phi * p.A;
    uint32_t need = 0, retire = 0;
    for (int q = 0; q < NP; ++q) {
      need |= firstp_tab[phi * NP + q] == yy ? 1u << q : 0u;
      retire |= lastp_tab[phi * NP + q] == yy ? 1u << q : 0u;
    }
    need_tab[i] = need;
    retire_tab[i] = retire;
  }
  for (int i = tid; i < 2 * p.A * NP; i += nthr) {
    const int par = i / (p.A * NP), phi = (i / NP) % p.A, q = i % NP;   // entry q of a super-step (par, phi)
    int phi_prev = phi - p.phase_shift;            // accumulating phases; alternating: the other of {0, phase_shift}
    phi_prev += phi_prev < 0 ? p.A : 0;
    if (p.phase_alt) phi_prev = phi == 0 ? p.phase_shift : 0;
    const int qp = q + par * p.ring.shift - (par ^ 1) * p.ring.shift;  // its position in the previous layout's numbering
    int ret = 0;
    if (qp >= 0 && qp < NP && lastp_tab[phi_prev * NP + qp] != 0xFF) ret = 1 + (int)lastp_tab[phi_prev * NP + qp];
    occ_tab[i] = (uint8_t)ret;
  }
  __syncthreads();
  // Reload order: the entries of the next super-step are loaded in the order in which their positions are
  // vacated (under the PREVIOUS super-step's phase), ties in the order the next super-step needs them; entries
  // nobody reads come last and are skipped.  Loading in index order would park the TMA thread on a position
  // that is vacated late while others wait for their refill.
  for (int i = tid; i < (2 * p.A + 1) * NP; i += nthr) {
    const int row = i / NP, q = i - row * NP;
    const bool steady = row < 2 * p.A;
    const int phi = steady ? row % p.A : 0;
    auto key_of = [&](int e) {
      const int fr = firstp_tab[phi * NP + e];
      return fr == 0xFF ? (1 << 30) + e : ((steady ? occ_tab[row * NP + e] : 0) * (p.A + 1) + fr) * NP + e;
    };
    const int my_key = key_of(q);
    int rank = 0;
    for (int o = 0; o < NP; ++o) rank += key_of(o) < my_key ? 1 : 0;
    order_tab[row * NP + rank] = (uint8_t)q;
  }
  for (int i = tid; i < 2 * p.A; i += nthr) {
    const int par = i / p.A, phi = i - par * p.A;
    const int phi_n = tc_next_phase(p, phi);
    uint32_t seam = 0;
    for (int q = 0; q < NP; ++q) seam |= occ_tab[((par ^ 1) * p.A + phi_n) * NP + q] == p.A ? 1u << q : 0u;
    seam_tab[i] = seam;
  }
  __syncthreads();
  __syncthreads();
  if (tid == 0) {
    int n_ent = 0;
    for (int q = 0; q < NP; ++q) n_ent += firstp_tab[q] != 0xFF ? 1 : 0;
    *reinterpret_cast<int*>(tab_smem + tab_bytes - 16) = n_ent;
  }
  __syncthreads();
  for (int i = tid; i < tab_bytes / 16; i += nthr)
    reinterpret_cast<uint4*>(tabs + (size_t)grp * tab_bytes)[i] = reinterpret_cast<const uint4*>(tab_smem)[i];
}

// CL = 2: the two CTAs of a (1,2,1) cluster work on the SAME 128-vertex tile and K range with
// different rotation groups.  Each gathers and interpolates only half of the rows and hands them to
// the peer's A stage with the bulk-copy engine (DSMEM), so the gather traffic of a tile - the
// largest L2 -> SM stream of the kernel - is paid once, not once per rotation group.  Stage
// barriers span the pair: a_full collects both CTAs' producers (remote arrive + complete_tx),
// a_empty both CTAs' MMA warps (tcgen05.commit multicast).
template <int KC, bool X3, bool H, int CL, bool KEEP>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_conv(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
          const __grid_constant__ TcFwdParams p) {
  constexpr int ROW_BYTES = KC * 4;
  constexpr int NPROD = X3 ? 2 : 1;
  constexpr int A_TILE_BYTES = TC_M * ROW_BYTES;
  constexpr int CPR = KC / 4;                // 16-byte chunks per row
  constexpr int ROWS_PER_PASS = 32 / CPR;
  // producer warps form NGRP groups that take tiles round-robin; a group covers this CTA's share of
  // the tile rows (128 rows alone, 64 rows in a cluster pair), so a CTA keeps 2 (4) tiles in flight
  constexpr int NGRP = 2 * CL;
  constexpr int GRP_WARPS = TC_PRODUCER_WARPS / NGRP;
  constexpr int WARP_ROWS = (TC_M / CL) / GRP_WARPS;   // tile rows one producer warp fills (32)
  constexpr int ITEMS = WARP_ROWS / ROWS_PER_PASS;     // 16-byte tile chunks per lane per tile
  constexpr int NV = H ? 2 : 1;                        // float4 of features behind one chunk (8 halves / 4 tf32)
  constexpr int KF = KC * NV;                          // features per tile row
  static_assert(!H || X3, "fp16 operands are only used in the three-product mode");

  extern __shared__ uint8_t smem_raw[];
  // life-cycle stamps of one CTA (GCB_TC_TRACE): entry, prologue done, accumulators complete, drained, exit
  long long* const life = p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 &&
                                  threadIdx.x == TC_WARP_PROD0 * 32 ? p.trace + ((size_t)3 * 128 + 127) * 8 : nullptr;
  if (life) life[0] = clock64();
  long long* const clog = p.cta_log && threadIdx.x == TC_WARP_PROD0 * 32
                              ? p.cta_log + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (clog) {
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    clog[0] = (long long)gt; clog[2] = clock64(); clog[4] = sm;
  }
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_tile_bytes = p.n_t * ROW_BYTES;
  const uint32_t b_base = smem_base;                                        // [NPROD][n_pos][b_tile]
  const int NP = p.ring.n_pos;                                                // ring entries (positions of one layout)
  const int n_ring = NP + p.ring.shift;                                       // ring positions in shared memory
  const uint32_t a_base = b_base + (uint32_t)(n_ring * NPROD * b_tile_bytes);        // [ns_a][NPROD][A_TILE]
  const uint32_t rec_base = a_base + (uint32_t)(p.ns_a * NPROD * A_TILE_BYTES);      // [A][rows of this CTA][32 B]
  const uint32_t bar_base = rec_base + (uint32_t)(p.A * (TC_REC_BYTES / CL));
  const uint32_t a_full = bar_base;
  const uint32_t a_empty = a_full + 8u * p.ns_a;
  // per (super-step parity, ring entry): its data has landed.  Two sets because the two ring layouts let the
  // TMA warp fill an entry of super-step k+1 before the entry of the same index of super-step k has been
  // consumed: one barrier would then run two phases ahead of its waiter.
  const uint32_t b_full = a_empty + 8u * p.ns_a;
  const uint32_t b_empty = b_full + 16u * NP;      // per ring POSITION: every reader of its occupant is done
  const uint32_t acc_full = b_empty + 8u * n_ring;
  // tile n is ready (its A stage and every weight entry it reads have landed): the scout warp collects the
  // individual barriers, so an issuing warp pays ONE barrier wait per tile (each costs the issuing thread
  // ~250 cycles under load; 2-5 of them per tile boundary drained the tensor pipe, traced)
  const uint32_t ready = acc_full + 8u;
  const uint32_t tmem_slot = ready + 8u * TC_READY_RING;
  const uint32_t tab_base = bar_base + (uint32_t)p.bar_bytes;   // weight-ring schedules, built below
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));     // generic view of smem_base
  const int AA = p.A * p.A;
  // [phase][step]: ring entries (positions of the even layout) first / last read at that step
  uint32_t* need_tab = reinterpret_cast<uint32_t*>(smem_gen + (tab_base - smem_base));
  uint32_t* retire_tab = need_tab + AA;
  // [parity][phase]: entries of the NEXT super-step whose position is only vacated by the last step of this one
  uint32_t* seam_tab = retire_tab + AA;
  // [phase][entry]: first / last step at which the entry is read (0xFF: never - then it is never loaded)
  uint8_t* firstp_tab = reinterpret_cast<uint8_t*>(seam_tab + 2 * p.A);
  uint8_t* lastp_tab = firstp_tab + p.A * NP;
  // [parity * A + phase][i]: the i-th entry the TMA warp (re)loads for such a super-step (row 2A: the CTA's
  // first super-step, when every position is vacant)
  uint8_t* order_tab = lastp_tab + p.A * NP;           // [2 * A + 1][NP]
  // [parity][phase][entry]: 1 + last step at which the previous super-step still reads the entry's position
  // (0 = vacant for the whole previous super-step)
  uint8_t* occ_tab = order_tab + (2 * p.A + 1) * NP;   // [2][A][NP]

  // the shuffle tells the compiler the warp index is warp-uniform (role dispatch, uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int grp = blockIdx.y;
  const int t_chunks = p.T / p.n_t;
  int tile_m, tz, sk;
  if (p.fuse) {
    // blockIdx.x = (vertex tile, template chunk, K split) with the split fastest: the splits of a unit are dispatched
    // together, finish together, and the one that finishes last finds the others' partial sums in L2
    const int bx = (int)blockIdx.x;
    sk = bx % p.ksplit;
    const int rest = bx / p.ksplit;
    tz = rest % t_chunks;
    tile_m = rest / t_chunks;
  } else {
    // CTAs are dispatched in blockIdx order: the split that holds the centre super-steps (one tile each, so one
    // weight-ring reload per tile) goes FIRST instead of forming the tail wave (GCB_TC_SPLIT_ORDER=0: index order)
    tile_m = (int)blockIdx.x;
    tz = blockIdx.z / p.ksplit;
    sk = p.split_reverse ? p.ksplit - 1 - (int)(blockIdx.z - tz * p.ksplit) : (int)(blockIdx.z - tz * p.ksplit);
  }
  const int m0 = tile_m * TC_M;
  const int t0 = tz * p.n_t;
  const int rot0 = grp * p.g;
  const int g_cnt = min(p.g, p.n_rot - rot0);
  const int RA = p.R * p.A;
  const int n_ss = (p.R + 1) * p.nfc;  // super-steps: (x, f-chunk) for x < R, then the centre slab
  const int ss_begin = __ldg(p.splits + sk), ss_end = __ldg(p.splits + sk + 1);   // tc_split_table
  const int n_units = (g_cnt + p.ring.m - 1) / p.ring.m;   // MMA units: groups of <= m merged rotations
  // The angular visiting order is common to the cluster pair (they share the gathered tiles): it is
  // anchored at the first rotation of the pair's first CTA.  rel0 = this CTA's first offset
  // relative to that anchor = the first weight tile this CTA needs in every super-step.
  const int off_anchor = p.rot.off[(CL == 2 ? (grp & ~1) : grp) * p.g];
  int rel0 = p.rot.off[rot0] - off_anchor;
  rel0 += rel0 < 0 ? p.A : 0;

  const uint32_t crank = CL == 2 ? ptx::cluster_ctarank() : 0u;
  // issuing warps: units are dealt round-robin to up to TC_MMA_WARPS warps (separate accumulators, so
  // the order in which the tensor pipe interleaves their MMAs does not change any result)
  const int n_mma = n_units < TC_MMA_WARPS ? n_units : TC_MMA_WARPS;
  int mma_arrivals = n_mma;      // commits that release an A stage: every issuing warp of the cluster
  if (CL == 2) {
    const int rot0_peer = (grp ^ 1) * p.g;
    const int units_peer = (min(p.g, p.n_rot - rot0_peer) + p.ring.m - 1) / p.ring.m;
    mma_arrivals += units_peer < TC_MMA_WARPS ? units_peer : TC_MMA_WARPS;
  }
  if (warp == TC_WARP_TMA && lane == 0) {
    for (int s = 0; s < p.ns_a; ++s) {
      ptx::mbar_init(a_full + 8u * s, GRP_WARPS * CL);
      ptx::mbar_init(a_empty + 8u * s, (uint32_t)mma_arrivals);
    }
    for (int s = 0; s < 2 * NP; ++s) ptx::mbar_init(b_full + 8u * s, 1);
    for (int s = 0; s < n_ring; ++s) ptx::mbar_init(b_empty + 8u * s, (uint32_t)n_mma);
    ptx::mbar_init(acc_full, (uint32_t)n_mma);
    for (int s = 0; s < TC_READY_RING; ++s) ptx::mbar_init(ready + 8u * s, 1);
    ptx::mbar_init(tmem_slot + 16u, 1);   // fused finish: two staging buffers
    ptx::mbar_init(tmem_slot + 24u, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tensormap(&tmap_hi);
    if (X3) ptx::prefetch_tensormap(&tmap_lo);
  }
  if (warp == TC_WARP_MMA) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp == TC_WARP_SCOUT) {   // this group's ring schedules (k_tc_build_tabs), ~1 KB from L2
    const uint4* src = reinterpret_cast<const uint4*>(p.tabs + (size_t)grp * p.tab_bytes);
    uint4* dst = reinterpret_cast<uint4*>(smem_gen + (tab_base - smem_base));
    for (int i = lane; i < p.tab_bytes / 16; i += 32) dst[i] = __ldg(src + i);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync();   // the peer's barriers exist before anyone arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(
      0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base)), 0);

  if (life) life[1] = clock64();
  if (warp >= TC_WARP_PROD0) {
    // ===================== producers: gather + interpolate -> swizzled A tiles =====================
    const int pw = warp - TC_WARP_PROD0, ptid = (int)threadIdx.x - TC_WARP_PROD0 * 32;
    const int pg = pw / GRP_WARPS, wq = pw % GRP_WARPS;
    constexpr int ROWS_CTA = TC_M / CL;                                   // tile rows this CTA gathers
    const int row_cta0 = CL == 2 ? ROWS_CTA * (int)crank : 0;
    const int row_w = row_cta0 + WARP_ROWS * wq;                          // first tile row of this warp
    const int c = lane % CPR;
    const int row_l = lane / CPR;
    const uint4* rec_smem = reinterpret_cast<const uint4*>(smem_gen + (rec_base - smem_base));
    const bool keep_interp = KEEP && tz == 0 && (CL == 2 ? (grp >> 1) == 0 : grp == 0);
    const float scale_a = H ? __ldg(p.scales) : 1.f;
    int ring = -1;
    int n = 0;
    int phi = 0;               // visiting-order phase of the current super-step
    int stage = 0;             // n % ns_a, kept as a rolling counter (no integer division per tile)
    uint32_t empty_par = 1u;   // parity to wait for on a_empty[stage]: ((n / ns_a) & 1) ^ 1
    for (int ss = ss_begin; ss < ss_end; ++ss) {
      const bool centre = ss >= p.R * p.nfc;
      const int xr = centre ? p.R : ss / p.nfc;
      const int fc = ss - xr * p.nfc;
      const int n_y = centre ? 1 : p.A;
      if (!centre && xr != ring) {
        // stage this radial ring's packed records (A slots x 128 vertices x 32 B) in smem
        producer_bar_sync();
        // (a cluster-pair member only gathers - and stages - its own half of the tile rows)
        uint4* dst = reinterpret_cast<uint4*>(smem_gen + (rec_base - smem_base));
        if (ptid < ROWS_CTA * 2)
          for (int y = 0; y < p.A; ++y)
            dst[y * (ROWS_CTA * 2) + ptid] =
                __ldg(p.recs + ((size_t)(xr * p.A + y) * p.v_pad + m0 + row_cta0) * 2 + ptid);
        producer_bar_sync();
        ring = xr;
      }
      for (int yy = 0; yy < n_y; ++yy, ++n) {
        // Tiles are visited in the order y = yy + phase - off[anchor rotation]: rotation j then reads
        // weight tile (yy + phase + off_j - off_anchor) % A, so a CTA whose first rotation has relative
        // offset rel0 needs the weight tiles in the order phase+rel0, phase+rel0+1, ...
        int y = yy;
        if (!centre) {
          y = yy + phi - off_anchor;
          y += y < 0 ? p.A : 0;
          y -= y >= p.A ? p.A : 0;
        }
        const int my_stage = stage;
        const uint32_t my_par = empty_par;
        if (++stage == p.ns_a) { stage = 0; empty_par ^= 1u; }
        if ((n & (NGRP - 1)) != pg) continue;
        const bool tr = p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && wq == 0 &&
                        pg < 1 && n < 128;
        long long* trp = tr ? p.trace + ((size_t)pg * 128 + n) * 8 : nullptr;
        if (tr) trp[0] = clock64();
        const float* xcol = p.x + fc * KF + c * (4 * NV);
        float4 v[ITEMS * NV];
        if (p.debug & 4) {
#pragma unroll
          for (int j = 0; j < ITEMS * NV; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (!centre) {
          const uint4* rec = rec_smem + (size_t)(y * ROWS_CTA + (row_w - row_cta0) + row_l) * 2;
          float4 x0[ITEMS * NV], x1[ITEMS * NV], x2[ITEMS * NV];
          float w0[ITEMS], w1[ITEMS], w2[ITEMS];
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            uint4 ri = rec[j * ROWS_PER_PASS * 2];
            const uint4 rw = rec[j * ROWS_PER_PASS * 2 + 1];
            if (p.debug & 1) { ri.x = 0; ri.y = 0; ri.z = 0; }
            w0[j] = __uint_as_float(rw.x); w1[j] = __uint_as_float(rw.y); w2[j] = __uint_as_float(rw.z);
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              x0[j * NV + q] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)ri.x * p.F) + q);
              x1[j * NV + q] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)ri.y * p.F) + q);
              x2[j * NV + q] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)ri.z * p.F) + q);
            }
          }
          if (tr) trp[1] = clock64();
#pragma unroll
          for (int i = 0; i < ITEMS * NV; ++i) {
            const int j = i / NV;
            v[i].x = fmaf(x2[i].x, w2[j], fmaf(x1[i].x, w1[j], x0[i].x * w0[j]));
            v[i].y = fmaf(x2[i].y, w2[j], fmaf(x1[i].y, w1[j], x0[i].y * w0[j]));
            v[i].z = fmaf(x2[i].z, w2[j], fmaf(x1[i].z, w1[j], x0[i].z * w0[j]));
            v[i].w = fmaf(x2[i].w, w2[j], fmaf(x1[i].w, w1[j], x0[i].w * w0[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            const int row = m0 + row_w + row_l + j * ROWS_PER_PASS;
#pragma unroll
            for (int q = 0; q < NV; ++q)
              v[j * NV + q] = row < p.v_out ? __ldg(reinterpret_cast<const float4*>(xcol + (p.x_centre - p.x) + (size_t)row * p.F) + q)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (KEEP && keep_interp && !centre) {
          // every (vertex, slot, feature) value is produced exactly once across the K splits: keep it
          float* irow = p.interp_out + (size_t)(xr * p.A + y) * p.F + fc * KF + c * (4 * NV);
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            const int row = m0 + row_w + row_l + j * ROWS_PER_PASS;
#pragma unroll
            for (int q = 0; q < NV; ++q)
              if (row < p.v_out) *(reinterpret_cast<float4*>(irow + (size_t)row * RA * p.F) + q) = v[j * NV + q];
          }
        }
        if (tr) trp[2] = clock64() + (__float_as_int(v[0].x) & 0);   // after the loads have landed
        ptx::mbar_wait(a_empty + 8u * my_stage, my_par);
        if (tr) trp[3] = clock64();
        const uint32_t a_stage_addr = a_base + (uint32_t)(my_stage * NPROD * A_TILE_BYTES);
        uint8_t* a_hi = smem_gen + (a_stage_addr - smem_base);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          if (p.debug & 12) break;   // experiments: no stores
          const int row = row_w + row_l + j * ROWS_PER_PASS;
          const uint32_t off = (uint32_t)row * ROW_BYTES + swz_chunk<KC>(row, c) * 16u;
          if (H) {
            // 8 features -> 8 fp16 hi + 8 fp16 lo of the scaled value
            const float4 va = v[j * NV], vb = v[j * NV + NV - 1];
            uint4 h, l;
            split_half2(va.x, va.y, scale_a, h.x, l.x);
            split_half2(va.z, va.w, scale_a, h.y, l.y);
            split_half2(vb.x, vb.y, scale_a, h.z, l.z);
            split_half2(vb.z, vb.w, scale_a, h.w, l.w);
            *reinterpret_cast<uint4*>(a_hi + off) = h;
            *reinterpret_cast<uint4*>(a_hi + A_TILE_BYTES + off) = l;
          } else {
            float4 h;
            h.x = ptx::to_tf32(v[j].x); h.y = ptx::to_tf32(v[j].y);
            h.z = ptx::to_tf32(v[j].z); h.w = ptx::to_tf32(v[j].w);
            *reinterpret_cast<float4*>(a_hi + off) = h;
            if (X3) {
              float4 l;
              l.x = v[j].x - h.x; l.y = v[j].y - h.y; l.z = v[j].z - h.z; l.w = v[j].w - h.w;
              *reinterpret_cast<float4*>(a_hi + A_TILE_BYTES + off) = l;
            }
          }
        }
        if (tr) trp[4] = clock64();
        ptx::fence_proxy_async_smem();   // generic-proxy stores -> async proxy (tcgen05.mma, bulk copy)
        __syncwarp();
        if (tr) trp[5] = clock64();
        if (lane == 0) {
          ptx::mbar_arrive(a_full + 8u * my_stage);
          if (CL == 2) {
            // hand this warp's 32-row slab (hi, then lo) to the peer CTA with the bulk-copy engine;
            // the copy completes on the peer's a_full barrier, which also gets this warp's arrival
            const uint32_t a_peer = ptx::mapa(a_stage_addr, crank ^ 1u);
            const uint32_t slab = (uint32_t)row_w * ROW_BYTES, slab_bytes = WARP_ROWS * ROW_BYTES;
            const uint32_t bar_peer = ptx::mapa(a_full + 8u * my_stage, crank ^ 1u);
            ptx::mbar_arrive_expect_tx_cluster(bar_peer, NPROD * slab_bytes);
            ptx::bulk_copy_to_peer(a_peer + slab, a_stage_addr + slab, slab_bytes, bar_peer);
            if (X3) ptx::bulk_copy_to_peer(a_peer + A_TILE_BYTES + slab, a_stage_addr + A_TILE_BYTES + slab,
                                           slab_bytes, bar_peer);
          }
        }
        if (tr) trp[6] = clock64();
      }
      if (!centre) phi = tc_next_phase(p, phi);
    }
    // ===================== drain: TMEM -> this split's partial sums =====================
    if (life) life[2] = clock64();
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    if (life) life[3] = clock64();
    const int dq = warp & 3, dh = pw >> 2;   // TMEM lane quadrant of this warp (hardware: warp id % 4), and which half of the rotations
    const int row = m0 + 32 * dq + lane;       // < v_pad: the partial buffer is padded
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * dq) << 16);
    float* prow = p.partial + (((size_t)sk * p.v_pad + row) * p.n_rot + rot0) * p.T + t0;
    if (p.drain_cw > 0) {
      // A lane holds one accumulator ROW, and rows are n_rot * T floats apart in the partial buffer: storing from
      // registers made every warp store touch 32 sectors (traced: 20.7 k cycles per CTA for 196 KB).  Instead
      // the lane parks up to drain_cw columns of its row in the (now idle) weight ring - row stride = an odd
      // number of 16-byte units, so the 128-bit stores are conflict-free - and hands the row to the bulk-copy
      // engine as one contiguous piece.
      const uint32_t row_bytes = (uint32_t)p.drain_cw * 4u + 16u;
      const uint32_t my_row = b_base + ((uint32_t)pw * 32u + (uint32_t)lane) * row_bytes;
      for (int il = dh; il < g_cnt; il += TC_PRODUCER_WARPS / 4) {
        for (int c0 = 0; c0 < p.n_t; c0 += p.drain_cw) {
          const int cw = min(p.drain_cw, p.n_t - c0);
          ptx::bulk_wait_read_all();   // the previous piece has left this lane's staging row
          // (TMEM loads of up to 64 columns are issued back to back and waited for once: one wait per 16 columns
          // left the load latency exposed six times per 96-column piece)
          int cc = 0;
          for (; cc + 64 <= cw; cc += 64) {
            uint32_t ra[32], rb[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(il * p.n_t + c0 + cc), ra);
            ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(il * p.n_t + c0 + cc + 32), rb);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              ptx::st_shared_v4(my_row + (uint32_t)(cc + q) * 4u, ra[q], ra[q + 1], ra[q + 2], ra[q + 3]);
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              ptx::st_shared_v4(my_row + (uint32_t)(cc + 32 + q) * 4u, rb[q], rb[q + 1], rb[q + 2], rb[q + 3]);
          }
          for (; cc + 32 <= cw; cc += 32) {
            uint32_t ra[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(il * p.n_t + c0 + cc), ra);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              ptx::st_shared_v4(my_row + (uint32_t)(cc + q) * 4u, ra[q], ra[q + 1], ra[q + 2], ra[q + 3]);
          }
          for (; cc < cw; cc += 16) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)(il * p.n_t + c0 + cc), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; q += 4)
              ptx::st_shared_v4(my_row + (uint32_t)(cc + q) * 4u, r[q], r[q + 1], r[q + 2], r[q + 3]);
          }
          ptx::fence_proxy_async_smem();
          ptx::bulk_store_global(prow + (size_t)il * p.T + c0, my_row, (uint32_t)cw * 4u);
          ptx::bulk_commit_group();
        }
      }
      if (p.fuse == 1) ptx::bulk_wait_all();   // ... and the rows must have reached global memory before the ticket is taken
      else ptx::bulk_wait_read_all();     // shared memory must outlive the copies' reads
    } else {
      for (int il = dh; il < g_cnt; il += TC_PRODUCER_WARPS / 4) {
        for (int c0 = 0; c0 < p.n_t; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)(il * p.n_t + c0), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            *reinterpret_cast<uint4*>(prow + (size_t)il * p.T + c0 + q) = make_uint4(r[q], r[q + 1], r[q + 2], r[q + 3]);
        }
      }
    }
    if (life) life[4] = clock64();
    if (p.fuse == 1) {
      // ===================== fused finish: split reduction, bias, activation, angular max-pooling =====================
      // Every producer thread has completed its own partial-sum stores (bulk copies waited for / plain stores);
      // publish them (fence, barrier, one ticket per CTA).  Nobody waits for anybody: the CTA whose ticket is the
      // last one of its unit does the unit's reduction, every other CTA is done.
      volatile uint32_t* const flag = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base)) + 1;
      ptx::fence_proxy_async_all();
      __threadfence();
      producer_bar_sync();
      const int unit = (tile_m * t_chunks + tz) * (int)gridDim.y + grp;
      if (clog) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        clog[5] = (long long)gt;
      }
      if (ptid == 0) {
        const unsigned t = atomicAdd(p.tickets + unit, 1u);
        __threadfence();
        *flag = t == (unsigned)p.ksplit - 1u ? 1u : 0u;
      }
      producer_bar_sync();
      if (*flag) {
        // The unit's partial sums - ksplit slabs x (rows of this tile) x (g_cnt rotations x n_t templates) - are
        // fetched by the bulk-copy engine into the now idle operand rings, a batch of rows at a time, double
        // buffered: plain loads (one row per warp, a few 16-byte loads per lane in flight) left this CTA waiting on
        // L2 latency for ~100 us; the copy engine keeps ~100 KB in flight.
        const float inv = H ? __ldg(p.scales + 2) : 1.f;
        const int T4c = p.n_t >> 2, T4 = p.T >> 2;
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias) + (t0 >> 2);
        const int rows_valid = min(TC_M, p.v_out - m0);
        const uint32_t seg_floats = (uint32_t)(g_cnt * p.n_t);              // one (row, split): g_cnt rotations x n_t
        const uint32_t row_bytes = (uint32_t)p.ksplit * seg_floats * 4u;    // one row: every split
        const uint32_t region = rec_base - b_base;
        int nb = (int)(region / (2u * row_bytes));
        nb = nb > TC_PRODUCER_WARPS ? TC_PRODUCER_WARPS : nb;               // one row per warp and batch
        const bool contiguous = t_chunks == 1;                              // the g_cnt rotations of a row are adjacent
        const int pieces = contiguous ? 1 : g_cnt;                          // copies per (row, split)
        const uint32_t piece_bytes = (contiguous ? seg_floats : (uint32_t)p.n_t) * 4u;
        const uint32_t fin_bar = tmem_slot + 16u;
        const int n_batches = nb > 0 ? (rows_valid + nb - 1) / nb : 0;
        ptx::fence_proxy_async_all();
        auto issue = [&](int b, int buf) {
          const int r_lo = b * nb, r_n = min(nb, rows_valid - r_lo);
          if (p.debug & 16) {   // timing experiment: no copies (the sums are garbage)
            if (ptid == 0) ptx::mbar_arrive(fin_bar + 8u * (uint32_t)buf);
            return;
          }
          if (ptid == 0) ptx::mbar_arrive_expect_tx(fin_bar + 8u * (uint32_t)buf, (uint32_t)r_n * row_bytes);
          for (int i = ptid; i < r_n * p.ksplit * pieces; i += TC_PRODUCER_WARPS * 32) {
            const int pc = i % pieces, sp = (i / pieces) % p.ksplit, rr = i / (pieces * p.ksplit);
            const float* src = p.partial + (((size_t)sp * p.v_pad + m0 + r_lo + rr) * p.n_rot + rot0 + pc) * p.T + t0;
            const uint32_t dst = b_base + (uint32_t)buf * (uint32_t)nb * row_bytes +
                                 ((uint32_t)(rr * p.ksplit + sp) * seg_floats + (uint32_t)(pc * p.n_t)) * 4u;
            ptx::bulk_load_global(dst, src, piece_bytes, fin_bar + 8u * (uint32_t)buf);
          }
        };
        if (n_batches > 0) issue(0, 0);
        if (n_batches > 1) issue(1, 1);
        for (int b = 0; b < n_batches; ++b) {
          const int buf = b & 1;
          ptx::mbar_wait(fin_bar + 8u * (uint32_t)buf, (uint32_t)(b >> 1) & 1u);
          const int rr = pw, vrow = m0 + b * nb + rr;
          if (rr < nb && b * nb + rr < rows_valid && !(p.debug & 32)) {   // (32: timing experiment, copies only)
            const float4* sm = reinterpret_cast<const float4*>(
                smem_gen + (b_base - smem_base) + (size_t)buf * nb * row_bytes + (size_t)rr * row_bytes);
            const uint32_t seg4 = seg_floats >> 2;
            for (int il = 0; il < g_cnt; ++il) {
              float4* dst = reinterpret_cast<float4*>(p.y) + ((size_t)vrow * p.n_rot + rot0 + il) * T4 + (t0 >> 2);
              float sq = 0.f;
              for (int q = lane; q < T4c; q += 32) {
                const float4 bv = __ldg(bias4 + q);
                float4 acc = H ? make_float4(0.f, 0.f, 0.f, 0.f) : bv;
                for (int sp = 0; sp < p.ksplit; ++sp) {          // split order: the order k_tc_finish adds in
                  const float4 pv = sm[(size_t)sp * seg4 + (size_t)il * T4c + q];
                  acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
                }
                if (H) {
                  acc.x = fmaf(acc.x, inv, bv.x); acc.y = fmaf(acc.y, inv, bv.y);
                  acc.z = fmaf(acc.z, inv, bv.z); acc.w = fmaf(acc.w, inv, bv.w);
                }
                float4 o;
                o.x = act_apply(p.act, acc.x); o.y = act_apply(p.act, acc.y);
                o.z = act_apply(p.act, acc.z); o.w = act_apply(p.act, acc.w);
                dst[q] = o;
                sq = fmaf(o.x, o.x, sq); sq = fmaf(o.y, o.y, sq); sq = fmaf(o.z, o.z, sq); sq = fmaf(o.w, o.w, sq);
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
              if (lane == 0 && p.z) p.nrm2[((size_t)vrow * p.n_rot + rot0 + il) * t_chunks + tz] = sq;
            }
            if (p.discard && contiguous && (p.T & 31) == 0) {
              // the row's partial sums are dead: drop their (dirty) L2 lines instead of writing them back to DRAM
              const int lines = (int)(seg_floats >> 5);           // 128-byte lines per (row, split)
              for (int i = lane; i < p.ksplit * lines; i += 32) {
                const int sp = i / lines, ln = i - sp * lines;
                const float* a = p.partial + (((size_t)sp * p.v_pad + vrow) * p.n_rot + rot0) * p.T + t0 + ln * 32;
                if ((reinterpret_cast<uintptr_t>(a) & 127) == 0) asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
              }
            }
          }
          producer_bar_sync();                                   // everybody is done with this buffer
          if (b + 2 < n_batches) issue(b + 2, buf);
        }
        if (n_batches == 0) {
          // (a row does not fit the staging area twice: plain loads)
          const size_t split_stride4 = (size_t)p.v_pad * p.n_rot * T4;
          for (int rr = pw; rr < rows_valid; rr += TC_PRODUCER_WARPS) {
            const int vrow = m0 + rr;
            for (int il = 0; il < g_cnt; ++il) {
              const size_t row4 = ((size_t)vrow * p.n_rot + rot0 + il) * T4 + (t0 >> 2);
              const float4* src = reinterpret_cast<const float4*>(p.partial) + row4;
              float4* dst = reinterpret_cast<float4*>(p.y) + row4;
              float sq = 0.f;
              for (int q = lane; q < T4c; q += 32) {
                const float4 bv = __ldg(bias4 + q);
                float4 acc = H ? make_float4(0.f, 0.f, 0.f, 0.f) : bv;
                for (int sp = 0; sp < p.ksplit; ++sp) {
                  const float4 pv = __ldcg(src + (size_t)sp * split_stride4 + q);
                  acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
                }
                if (H) {
                  acc.x = fmaf(acc.x, inv, bv.x); acc.y = fmaf(acc.y, inv, bv.y);
                  acc.z = fmaf(acc.z, inv, bv.z); acc.w = fmaf(acc.w, inv, bv.w);
                }
                float4 o;
                o.x = act_apply(p.act, acc.x); o.y = act_apply(p.act, acc.y);
                o.z = act_apply(p.act, acc.z); o.w = act_apply(p.act, acc.w);
                dst[q] = o;
                sq = fmaf(o.x, o.x, sq); sq = fmaf(o.y, o.y, sq); sq = fmaf(o.z, o.z, sq); sq = fmaf(o.w, o.w, sq);
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
              if (lane == 0 && p.z) p.nrm2[((size_t)vrow * p.n_rot + rot0 + il) * t_chunks + tz] = sq;
            }
          }
        }
        if (clog) {
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          clog[6] = 1; clog[7] = (long long)gt;
        }
        if (p.z) {
          // the unit that completes this vertex tile pools it: norms over all template chunks (chunk order), first
          // maximum over the rotations (a NaN norm counts as the maximum, like torch.argmax), row select
          __threadfence();
          producer_bar_sync();
          if (ptid == 0) {
            const unsigned t = atomicAdd(p.tickets + (size_t)gridDim.x / p.ksplit * gridDim.y + tile_m, 1u);
            __threadfence();
            *flag = t == (unsigned)(t_chunks * (int)gridDim.y) - 1u ? 1u : 0u;
          }
          producer_bar_sync();
          if (*flag) {
            for (int rr = pw; rr < TC_M; rr += TC_PRODUCER_WARPS) {
              const int vrow = m0 + rr;
              if (vrow >= p.v_out) break;
              float nrm = -1.f;
              if (lane < p.n_rot) {
                float sq = 0.f;
                for (int c = 0; c < t_chunks; ++c) sq += __ldcg(p.nrm2 + ((size_t)vrow * p.n_rot + lane) * t_chunks + c);
                nrm = sqrtf(sq);
              }
              const unsigned nan_mask = __ballot_sync(0xffffffffu, nrm != nrm);
              float best = nrm == nrm ? nrm : -1.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
              const unsigned max_mask = __ballot_sync(0xffffffffu, lane < p.n_rot && nrm == best);
              const int arg = nan_mask ? __ffs((int)nan_mask) - 1 : (max_mask ? __ffs((int)max_mask) - 1 : 0);
              const float4* sel = reinterpret_cast<const float4*>(p.y) + ((size_t)vrow * p.n_rot + arg) * T4;
              float4* zr = reinterpret_cast<float4*>(p.z) + (size_t)vrow * T4;
              for (int q = lane; q < T4; q += 32) zr[q] = __ldcg(sel + q);
              if (lane == 0) p.argmax[vrow] = arg;
            }
          }
        }
      }
    }
  } else if (warp == TC_WARP_TMA) {
    // ===================== TMA: weight tiles -> ring =====================
    // (reload order: order_tab, built in the prologue; an entry = one ring position of the current layout)
    uint64_t fills = 0;  // bit P = parity of the number of fills of ring position P
    const uint32_t lo_ring = (uint32_t)(n_ring * b_tile_bytes);
    const int n_ent = *reinterpret_cast<const int*>(smem_gen + (tab_base - smem_base) + p.tab_bytes - 16);
    int load_no = 0;
    int phi = 0, par = 0;
    bool first_ss = true;
    for (int ss = ss_begin; ss < ss_end; ++ss) {
      const bool centre = ss >= p.R * p.nfc;
      const int xr = centre ? p.R : ss / p.nfc;
      const int fc = ss - xr * p.nfc;
      const int n_load = centre ? 1 : n_ent;
      const uint8_t* order = order_tab + (first_ss ? 2 * p.A : par * p.A + phi) * NP;
      if (lane == 0) {
        for (int i = 0; i < n_load; ++i) {
          const int q = centre ? p.ring.pos[0] : order[i];          // entry; the centre slab uses tile 0's
          const uint32_t pos = (uint32_t)(q + par * p.ring.shift);  // ring position in shared memory
          int sg = p.ring.tile_of[q] + rel0;                        // angular tile (absolute)
          sg -= sg >= p.A ? p.A : 0;
          const bool tr = p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && load_no < 128;
          long long* trp = tr ? p.trace + ((size_t)3 * 128 + load_no) * 8 : nullptr;
          ++load_no;
          if (tr) trp[0] = clock64();
          ptx::mbar_wait(b_empty + 8u * pos, ((uint32_t)(fills >> pos) & 1u) ^ 1u);
          if (tr) trp[1] = clock64();
          const int col = (centre ? RA * p.F : (xr * p.A + sg) * p.F) + fc * KF;
          const uint32_t bar = b_full + 8u * (uint32_t)(par * NP + q);
          ptx::mbar_arrive_expect_tx(bar, (uint32_t)(NPROD * b_tile_bytes));
          const uint32_t dst = b_base + pos * (uint32_t)b_tile_bytes;
          ptx::tma_load_2d(dst, &tmap_hi, bar, col, t0);
          if (X3) ptx::tma_load_2d(dst + lo_ring, &tmap_lo, bar, col, t0);
          fills ^= 1ull << pos;
          if (tr) trp[2] = clock64();
        }
      }
      par ^= 1;   // centre super-steps alternate layouts too
      if (!centre) {
        first_ss = false;
        phi = tc_next_phase(p, phi);
      }
    }
  } else if (warp == TC_WARP_SCOUT) {
    // ===================== scout: per tile, wait for its A stage and weight entries, then publish =====================
    if (lane == 0) {
      const int ss_nc = p.R * p.nfc;
      const int n_nc = min(ss_end, ss_nc) - min(ss_begin, ss_nc);
      const int n_c = ss_end > max(ss_begin, ss_nc) ? ss_end - max(ss_begin, ss_nc) : 0;
      uint32_t seen_cur = 0u, seen_nxt = 0u;
      uint32_t bf_cur = b_full, bf_nxt = b_full + 8u * (uint32_t)NP;
      int stage = 0, n = 0, phi = 0;
      uint32_t par = 0;
      auto publish = [&]() {
        ptx::mbar_arrive(ready + 8u * (uint32_t)(n & (TC_READY_RING - 1)));
        ++n;
        if (++stage == p.ns_a) { stage = 0; par ^= 1u; }
      };
      for (int k = 0; k < n_nc + n_c; ++k) {
        const bool centre = k >= n_nc;
        const int n_y = centre ? 1 : p.A;
        for (int y = 0; y < n_y; ++y) {
          long long* trp = p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && n < 128
                               ? p.trace + ((size_t)1 * 128 + n) * 8 : nullptr;
          if (trp) trp[0] = clock64();
          // weight entries first: they landed steps ago (refill slack), so their waits (~350 cycles each,
          // four at the first step of a super-step) overlap the producers' latency for this tile's A stage
          tc_wait_entries(bf_cur, centre ? 1u << p.ring.pos[0] : need_tab[phi * p.A + y], seen_cur);
          if (trp) trp[1] = clock64();
          ptx::mbar_wait(a_full + 8u * stage, par);
          if (trp) trp[2] = clock64();
          publish();
        }
        if (!centre) phi = tc_next_phase(p, phi);
        { const uint32_t t = seen_cur; seen_cur = seen_nxt; seen_nxt = t; }
        { const uint32_t t = bf_cur; bf_cur = bf_nxt; bf_nxt = t; }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    const int mw = warp - TC_WARP_MMA;
    if (mw < n_mma) {
      const int ss_nc = p.R * p.nfc;
      const uint32_t lo_flags = (uint32_t)(ptx::make_kmajor_desc<ROW_BYTES>(0) & 0xFFFF0000ull);
      TcMmaCtx c;
      c.a_full = a_full; c.a_empty = a_empty; c.b_full = b_full; c.b_empty = b_empty; c.acc_full = acc_full;
      c.ready = ready;
      c.a_lo0 = (a_base >> 4) | lo_flags;
      c.a_stage_lo = (uint32_t)(NPROD * A_TILE_BYTES) >> 4;
      c.b_lo0 = (b_base >> 4) | lo_flags;
      c.b_tile_lo = (uint32_t)b_tile_bytes >> 4;
      c.b_lo_ring = (uint32_t)(n_ring * b_tile_bytes) >> 4;
      c.tmem_base = tmem_base;
      c.rot0 = rot0; c.g_cnt = g_cnt; c.n_units = n_units; c.off_anchor = off_anchor; c.rel0 = rel0;
      c.n_nc = min(ss_end, ss_nc) - min(ss_begin, ss_nc);
      c.n_c = ss_end > max(ss_begin, ss_nc) ? ss_end - max(ss_begin, ss_nc) : 0;
      c.mw = mw; c.n_mma = n_mma;
      const int n_own = (n_units - mw + n_mma - 1) / n_mma;
      if (n_own == 1) tc_mma_role<KC, X3, H, CL, 1>(p, c, lane, need_tab, retire_tab, seam_tab);
      else if (n_own == 2) tc_mma_role<KC, X3, H, CL, 2>(p, c, lane, need_tab, retire_tab, seam_tab);
      else if (n_own <= 4) tc_mma_role<KC, X3, H, CL, 4>(p, c, lane, need_tab, retire_tab, seam_tab);
      else tc_mma_role<KC, X3, H, CL, 8>(p, c, lane, need_tab, retire_tab, seam_tab);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (CL == 2) ptx::cluster_sync();   // neither CTA leaves while the peer may still write / arrive here
  if (life) life[5] = clock64();
  if (clog) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    clog[1] = (long long)gt; clog[3] = clock64();
  }
}

// ---- kernel 2: split reduction + bias + activation + angular max-pooling -------------------------
// One warp per vertex; lanes own float4 column groups.  Summation order over splits is fixed.
// (scales != NULL: the partial sums carry the fp16 mode's operand scales; their sum is multiplied by the exact
// inverse, scales[2], before the bias is added)
__global__ void k_tc_finish(const float* __restrict__ partial, const float* __restrict__ bias, int ksplit,
                            int v_out, int v_pad, int n_rot, int T, int act, const float* __restrict__ scales,
                            float* __restrict__ y, float* __restrict__ z, int32_t* __restrict__ argmax) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= v_out) return;
  const size_t split_stride = (size_t)v_pad * n_rot * T;
  const int T4 = T >> 2;
  float best = -1.f;
  int arg = 0;
  const float inv = scales ? __ldg(scales + 2) : 1.f;
  for (int i = 0; i < n_rot; ++i) {
    const float* src = partial + ((size_t)k * n_rot + i) * T;
    float* dst = y + ((size_t)k * n_rot + i) * T;
    float s = 0.f;
    for (int q = lane; q < T4; q += 32) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + q);
      float4 acc = scales ? make_float4(0.f, 0.f, 0.f, 0.f) : bv;
      for (int sp = 0; sp < ksplit; ++sp) {
        const float4 pv = __ldcs(reinterpret_cast<const float4*>(src + sp * split_stride) + q);   // read once: streaming
        acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
      }
      if (scales) {
        acc.x = fmaf(acc.x, inv, bv.x); acc.y = fmaf(acc.y, inv, bv.y);
        acc.z = fmaf(acc.z, inv, bv.z); acc.w = fmaf(acc.w, inv, bv.w);
      }
      float4 o;
      o.x = act_apply(act, acc.x); o.y = act_apply(act, acc.y);
      o.z = act_apply(act, acc.z); o.w = act_apply(act, acc.w);
      reinterpret_cast<float4*>(dst)[q] = o;
      s = fmaf(o.x, o.x, s); s = fmaf(o.y, o.y, s); s = fmaf(o.z, o.z, s); s = fmaf(o.w, o.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = sqrtf(s);
    if (nrm > best) { best = nrm; arg = i; }
  }
  if (z) {
    __syncwarp();
    const float4* sel = reinterpret_cast<const float4*>(y + ((size_t)k * n_rot + arg) * T);
    for (int q = lane; q < T4; q += 32) reinterpret_cast<float4*>(z + (size_t)k * T)[q] = sel[q];
    if (lane == 0) argmax[k] = arg;
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

int tc_make_tmap_2d(CUtensorMap* out, const float* base, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool atom32) {
  PFN_encodeTiled enc = get_encode_fn();
  GCB_REQUIRE(enc, GCB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (box_inner == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  if (atom32) sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCB_REQUIRE(r == CUDA_SUCCESS, GCB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GCB_OK;
}

// Row-major fp32 matrix [rows][cols] seen as (32 columns, rows, cols / 32): one box = `box_atoms`
// 32-column atoms x `box_rows` rows, written to shared memory atom after atom - the MN-major UMMA
// operand layout (SWIZZLE_128B with 32-byte atoms) for `box_atoms * 32` rows/columns of M or N.
int tc_make_tmap_mn(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                    uint32_t box_rows, uint32_t box_atoms) {
  PFN_encodeTiled enc = get_encode_fn();
  GCB_REQUIRE(enc, GCB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  GCB_REQUIRE(cols % 32 == 0, GCB_EINVAL, "tc_make_tmap_mn: cols must be a multiple of 32");
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_bytes, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_atoms};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCB_REQUIRE(r == CUDA_SUCCESS, GCB_ECUDA, "cuTensorMapEncodeTiled (3-D, MN-major) failed with CUresult %d", (int)r);
  return GCB_OK;
}

static int make_weight_tmap(CUtensorMap* out, const float* base, int T, int kcat, int kc, int n_t) {
  return tc_make_tmap_2d(out, base, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * sizeof(float), (uint32_t)kc,
                         (uint32_t)n_t);
}

int tc_make_tmap_2d_h(CUtensorMap* out, const __half* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  GCB_REQUIRE(enc, GCB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                : (box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCB_REQUIRE(r == CUDA_SUCCESS, GCB_ECUDA, "cuTensorMapEncodeTiled (fp16) failed with CUresult %d", (int)r);
  return GCB_OK;
}

int tc_make_tmap_mn_h(CUtensorMap* out, const __half* base, uint64_t rows, uint64_t cols_ld, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_atoms) {
  PFN_encodeTiled enc = get_encode_fn();
  GCB_REQUIRE(enc, GCB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  GCB_REQUIRE(cols_ld % 64 == 0, GCB_EINVAL, "tc_make_tmap_mn_h: the row length must be a multiple of 64");
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols_ld / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_bytes, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_atoms};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCB_REQUIRE(r == CUDA_SUCCESS, GCB_ECUDA, "cuTensorMapEncodeTiled (fp16, MN-major) failed with CUresult %d", (int)r);
  return GCB_OK;
}

// fp16 weights [T][kcat]: box = 2 * kc halves (kc 4-byte units) x n_t rows, swizzle span = the box row
static int make_weight_tmap_h(CUtensorMap* out, const __half* base, int T, int kcat, int kc, int n_t) {
  return tc_make_tmap_2d_h(out, base, (uint64_t)kcat, (uint64_t)T, (uint64_t)kcat * sizeof(__half), (uint32_t)(2 * kc),
                           (uint32_t)n_t);
}

int tc_build_wcat_h(const float* w_eff, const float* w_self, int T, int R, int A, int F, const float* scale,
                    __half* hi, __half* lo, cudaStream_t st) {
  const int KRA = R * A * F;
  GCB_REQUIRE(F % 8 == 0, GCB_EINVAL, "tc_build_wcat_h: F must be a multiple of 8");
  const int64_t wt = (int64_t)T * ((KRA + F) / 8);
  k_tc_wcat_h<<<(unsigned)ceil_div64(wt, 256), 256, 0, st>>>(w_eff, w_self, T, KRA, F, scale, hi, lo);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int tc_absmax(const float* a, int64_t na, const float* b, int64_t nb, unsigned* out, cudaStream_t st) {
  k_tc_absmax<<<148 * 2, 256, 0, st>>>(a, na, b, nb, out);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int tc_maxima_scales(const TcMaxJob& job_in, unsigned* mx, float* scales, cudaStream_t st) {
  // blocks per segment: proportional to its bytes, at least one; empty segments are dropped
  TcMaxJob job = job_in;
  int n = 0;
  for (int i = 0; i < job_in.n_seg; ++i)
    if (job_in.seg[i].ptr && job_in.seg[i].n > 0) job.seg[n++] = job_in.seg[i];
  job.n_seg = n;
  GCB_REQUIRE(n > 0 && n <= TC_MAX_SEGS, GCB_EINVAL, "tc_maxima_scales: no input");
  long long total = 0;
  for (int i = 0; i < n; ++i) total += job.seg[i].n;
  const int budget = 148 * 4;
  int at = 0;
  for (int i = 0; i < n; ++i) {
    job.blk0[i] = at;
    long long want = (job.seg[i].n * budget + total - 1) / total;
    const long long cap = (job.seg[i].n + 1023) / 1024;   // no more blocks than 1024-element pieces
    if (want > cap) want = cap;
    at += want < 1 ? 1 : (int)want;
  }
  for (int i = n; i <= TC_MAX_SEGS; ++i) job.blk0[i] = at;
  k_tc_maxima<<<at, 256, 0, st>>>(job, mx, scales);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

static int tc_pack_recs_max(const int32_t* idx, const float* w, int v_out, int R, int A, uint4* recs,
                            unsigned* wsum_max, cudaStream_t st) {
  const int v_pad = (int)ceil_div64(v_out, TC_M) * TC_M;
  int64_t total = (int64_t)R * A * v_pad;
  k_tc_pack_recs<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(idx, w, v_out, v_pad, R * A, recs, wsum_max);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

int tc_pack_recs(const int32_t* idx, const float* w, int v_out, int R, int A, uint4* recs, cudaStream_t st) {
  return tc_pack_recs_max(idx, w, v_out, R, A, recs, nullptr, st);
}

int tc_build_wcat(const float* w_eff, const float* w_self, int T, int R, int A, int F, float* hi, float* lo,
                  cudaStream_t st) {
  const int KRA = R * A * F;
  GCB_REQUIRE(F % 4 == 0, GCB_EINVAL, "tc_build_wcat: F must be a multiple of 4");
  int64_t wt = (int64_t)T * ((KRA + F) / 4);
  k_tc_wcat<<<(unsigned)ceil_div64(wt, 256), 256, 0, st>>>(w_eff, w_self, T, KRA, F, hi, lo);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

bool tc_supported(int R, int A, int F, int T, int n_rot) {
  (void)R;
  return tc_pick_config(A, F, T, n_rot, true, false, 0).ok && tc_pick_config(A, F, T, n_rot, false, false, 0).ok;
}

size_t tc_recs_bytes(int v_out, int R, int A) {
  size_t v_pad = (size_t)ceil_div64(v_out, TC_M) * TC_M;
  return align_up((size_t)R * A * v_pad * 32, 256);
}
size_t tc_wcat_bytes(int R, int A, int F, int T) {
  return align_up((size_t)T * ((size_t)R * A * F + F) * sizeof(float), 256);
}
static size_t tc_partial_bytes(int v_out, int n_rot, int T, int ksplit) {
  size_t v_pad = (size_t)ceil_div64(v_out, TC_M) * TC_M;
  return align_up((size_t)ksplit * v_pad * n_rot * T * sizeof(float), 256);
}

// The tiling (and with it the K-split count) depends on the rotation list, which the size query does
// not see: take the largest requirement over every uniform step and the unmerged layout.
// fp16 operands need 16 features per row at least (a 32-byte swizzle span)
bool tc_half_supported(int R, int A, int F, int T, int n_rot) {
  (void)R;
  return tc_pick_config(A, F, T, n_rot, true, true, 0).ok;
}

// fused finish: one arrival counter per (vertex tile, template chunk, rotation group) unit and one per vertex tile
// (bounds: at most n_rot groups and T / 16 template chunks), squared row norms per template chunk
static size_t tc_ticket_bytes(int v_out, int n_rot, int T) {
  return align_up((size_t)ceil_div64(v_out, TC_M) * ((size_t)n_rot * (T / 16) + 1) * sizeof(unsigned), 256);
}
static size_t tc_nrm2_bytes(int v_out, int n_rot, int T) {
  return align_up((size_t)ceil_div64(v_out, TC_M) * TC_M * n_rot * (T / 16) * sizeof(float), 256);
}

constexpr size_t TC_TABS_BYTES = (size_t)GCB_MAX_ROT * 16384;   // ring schedules of every rotation group
constexpr size_t TC_SPLITS_BYTES = 16384;   // K-split boundaries
constexpr size_t TC_SCALES_BYTES = 256;   // {max|X|, max sum|w|, max|Wcat|} as bits, then the three scales

size_t tc_fwd_workspace_bytes(int v_out, int R, int A, int F, int T, int n_rot, int precision) {
  const bool half = precision == GCB_PREC_F16X3;
  const bool x3 = half || precision == GCB_PREC_TF32X3;
  int ksplit = 0;
  for (int dstep = 0; dstep < A; ++dstep) {
    TcConfig cfg = tc_pick_config(A, F, T, n_rot, x3, half, dstep);
    if (!cfg.ok) return 0;
    int s = tc_pick_ksplit(cfg, v_out, R, A, F, T, x3, half, n_rot);
    if (s > ksplit) ksplit = s;
  }
  return tc_recs_bytes(v_out, R, A) + (x3 ? 2 : 1) * tc_wcat_bytes(R, A, F, T) +
         tc_partial_bytes(v_out, n_rot, T, ksplit) + TC_SCALES_BYTES + TC_TABS_BYTES + TC_SPLITS_BYTES +
         tc_ticket_bytes(v_out, n_rot, T) + tc_nrm2_bytes(v_out, n_rot, T) + 1024;
}

template <int KC, bool X3, bool H, int CL, bool KEEP>
static int tc_launch_k(const CUtensorMap& mh, const CUtensorMap& ml, const TcFwdParams& p, const TcConfig& cfg,
                       dim3 grid, cudaStream_t st) {
  GCB_CUDA_OK(cudaFuncSetAttribute(k_tc_conv<KC, X3, H, CL, KEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)cfg.smem_bytes));
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = dim3(TC_THREADS);
  lc.dynamicSmemBytes = cfg.smem_bytes;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CL;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  profile_begin(GCB_PROF_FWD_CONV, st);
  cudaError_t e = cudaLaunchKernelEx(&lc, k_tc_conv<KC, X3, H, CL, KEEP>, mh, ml, p);
  profile_end(GCB_PROF_FWD_CONV, st);
  GCB_CUDA_OK(e);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

// Pairs of rotation groups share their gathered tiles through a 2-CTA cluster when the group count
// is even (GCB_TC_CLUSTER=0 switches the pairing off for A/B measurements).
template <int KC, bool X3, bool H = false>
static int tc_launch(const CUtensorMap& mh, const CUtensorMap& ml, const TcFwdParams& p, const TcConfig& cfg,
                     dim3 grid, cudaStream_t st) {
  if (tc_use_pairs() && cfg.n_groups % 2 == 0 && KC >= 16)
    return p.interp_out ? tc_launch_k<KC, X3, H, 2, true>(mh, ml, p, cfg, grid, st)
                        : tc_launch_k<KC, X3, H, 2, false>(mh, ml, p, cfg, grid, st);
  return p.interp_out ? tc_launch_k<KC, X3, H, 1, true>(mh, ml, p, cfg, grid, st)
                      : tc_launch_k<KC, X3, H, 1, false>(mh, ml, p, cfg, grid, st);
}

// Everything the host decides for one forward launch; a function of the shape, the rotation list and the precision only
struct TcFwdPlan {
  TcConfig cfg;
  int phase_shift, ksplit, drain_cw, split_reverse;
  bool pairs, half, x3;
};

static int tc_fwd_plan_compute(const RotTable& rot, int n_rot, int v_out, int R, int A, int F, int T, int precision, TcFwdPlan* out);

// The plan is searched on the host (tilings, ring layouts, the visiting-order advance with the most refill slack:
// 0.1-0.3 ms of host time at A = 12) and depends on nothing but its arguments, so the last few are remembered: a
// training loop that reads its loss every step starts every step with an idle GPU and pays the host time in full.
static int tc_fwd_plan(const RotTable& rot, int n_rot, int v_out, int R, int A, int F, int T, int precision, TcFwdPlan* out) {
  struct Entry { int key[8 + GCB_MAX_ROT]; TcFwdPlan plan; bool used; };
  static Entry cache[16];
  static int next = 0;
  static std::mutex mu;
  int key[8 + GCB_MAX_ROT] = {n_rot, v_out, R, A, F, T, precision, 0};
  for (int i = 0; i < n_rot && i < GCB_MAX_ROT; ++i) key[8 + i] = rot.off[i];
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
      if (e.used && memcmp(e.key, key, sizeof(key)) == 0) { *out = e.plan; return GCB_OK; }
  }
  const int rc = tc_fwd_plan_compute(rot, n_rot, v_out, R, A, F, T, precision, out);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(mu);
  Entry& e = cache[next];
  next = (next + 1) % 16;
  memcpy(e.key, key, sizeof(key));
  e.plan = *out;
  e.used = true;
  return GCB_OK;
}

static int tc_fwd_plan_compute(const RotTable& rot, int n_rot, int v_out, int R, int A, int F, int T, int precision, TcFwdPlan* out) {
  const bool half = precision == GCB_PREC_F16X3;
  const bool x3 = half || precision == GCB_PREC_TF32X3;
  static const bool no_merge = getenv("GCB_TC_NOMERGE") && atoi(getenv("GCB_TC_NOMERGE")) != 0;
  const int dstep = no_merge ? 0 : tc_uniform_step(rot, n_rot, A);
  TcConfig cfg = tc_pick_config(A, F, T, n_rot, x3, half, dstep);
  GCB_REQUIRE(cfg.ok, GCB_EUNSUPPORTED, "tc_conv_fwd: no tcgen05 tiling for A=%d F=%d T=%d", A, F, T);
  // Weight ring: spend spare shared memory on a second, shifted layout when that buys refill slack (3 steps
  // cover the commit -> TMA -> barrier latency) without changing the tiling or starving the A ring.
  int phase_shift = 0;
  static const int shift_env = getenv("GCB_TC_RINGSHIFT") ? atoi(getenv("GCB_TC_RINGSHIFT")) : -1;   // experiments
  static const int phase_env = getenv("GCB_TC_PHASE") ? atoi(getenv("GCB_TC_PHASE")) : -1;
  const bool pairs = tc_use_pairs() && cfg.n_groups % 2 == 0 && cfg.kc >= 16;
  int best_slack = -1;
  phase_shift = tc_pick_phase_shift(rot, n_rot, A, cfg.g, cfg.n_groups, pairs, cfg.ring, false, &best_slack);
  for (int rs = 1; rs <= 4; ++rs) {
    if (shift_env >= 0 && rs != shift_env) continue;
    if (shift_env < 0 && best_slack >= 3) break;
    TcConfig c2 = tc_pick_config(A, F, T, n_rot, x3, half, dstep, rs);
    if (!c2.ok || c2.kc != cfg.kc || c2.ring.m != cfg.ring.m || c2.n_t != cfg.n_t || c2.g != cfg.g || c2.ns_a < 4) break;
    int sl2 = -1;
    const int ph2 = tc_pick_phase_shift(rot, n_rot, A, c2.g, c2.n_groups, pairs, c2.ring, true, &sl2);   // two layouts: two phases
    if (sl2 > best_slack || shift_env >= 0) { best_slack = sl2; phase_shift = ph2; cfg = c2; }
  }
  if (phase_env >= 0) phase_shift = phase_env % A;
  static const bool verbose = getenv("GCB_TC_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "[gcb] tc forward: kc %d n_t %d g %d groups %d m %d ring %d+%d positions, phase shift %d, refill slack %d steps, A stages %d\n",
            cfg.kc, cfg.n_t, cfg.g, cfg.n_groups, cfg.ring.m, cfg.ring.n_pos, cfg.ring.shift, phase_shift,
            tc_refill_slack(rot, n_rot, A, cfg.g, cfg.n_groups, pairs, cfg.ring, phase_shift, cfg.ring.shift > 0), cfg.ns_a);
  out->cfg = cfg;
  out->phase_shift = phase_shift;
  out->pairs = pairs;
  out->half = half;
  out->x3 = x3;
  out->ksplit = tc_pick_ksplit(cfg, v_out, R, A, F, T, x3, half, n_rot);
  {
    static const int order_env = getenv("GCB_TC_SPLIT_ORDER") ? atoi(getenv("GCB_TC_SPLIT_ORDER")) : 1;   // experiments
    out->split_reverse = (order_env && x3) ? 1 : 0;   // (measured: +1 % with three products per k-step, -0.4 % in tf32 mode)
  }
  {
    // drain staging: 8 warps x 32 rows x (cw * 4 + 16) bytes inside the weight ring
    static const int drain_env = getenv("GCB_TC_DRAIN") ? atoi(getenv("GCB_TC_DRAIN")) : -1;   // experiments (0: direct)
    const size_t ring_bytes = (size_t)(cfg.ring.n_pos + cfg.ring.shift) * (x3 ? 2 : 1) * cfg.n_t * cfg.kc * 4;
    int cw = cfg.n_t < 128 ? cfg.n_t : 128;
    while (cw >= 16 && (size_t)TC_PRODUCER_WARPS * 32 * (cw * 4 + 16) > ring_bytes) cw -= 16;
    out->drain_cw = cw >= 16 ? cw : 0;
    if (drain_env >= 0) out->drain_cw = drain_env < out->drain_cw ? (drain_env / 16) * 16 : out->drain_cw;
  }
  GCB_REQUIRE((size_t)(out->ksplit + 1) * sizeof(int) <= TC_SPLITS_BYTES, GCB_EUNSUPPORTED, "tc_conv_fwd: too many K splits");
  GCB_REQUIRE((size_t)cfg.n_groups * tc_tab_bytes(A, cfg.ring.n_pos) <= TC_TABS_BYTES, GCB_EUNSUPPORTED,
              "tc_conv_fwd: ring schedules exceed their workspace");
  return GCB_OK;
}

// the launch parameters that depend on the plan only (no data pointers)
static void tc_fwd_plan_params(const TcFwdPlan& plan, const RotTable& rot, int n_rot, int v_out, int R, int A, int F, int T,
                               TcFwdParams* p) {
  const TcConfig& cfg = plan.cfg;
  const int m_tiles = (int)ceil_div64(v_out, TC_M);
  p->v_out = v_out; p->v_pad = m_tiles * TC_M; p->R = R; p->A = A; p->F = F; p->T = T; p->n_rot = n_rot;
  p->nfc = F / (plan.half ? 2 * cfg.kc : cfg.kc); p->n_t = cfg.n_t; p->g = cfg.g; p->ns_a = cfg.ns_a; p->tmem_cols = cfg.tmem_cols;
  p->ksplit = plan.ksplit;
  p->phase_shift = plan.phase_shift;
  p->phase_alt = cfg.ring.shift > 0 ? 1 : 0;
  p->bar_bytes = (int)tc_bar_bytes(cfg.ring.n_pos, cfg.ring.n_pos + cfg.ring.shift);
  p->split_reverse = plan.split_reverse;
  p->tab_bytes = (int)tc_tab_bytes(A, cfg.ring.n_pos);
  p->drain_cw = plan.drain_cw;
  p->ring = cfg.ring;
  p->rot = rot;
}

// Ring schedules + K-split boundaries of one (shape, rotation list, precision): `tables` = TC_TABS_BYTES of schedules
// followed by TC_SPLITS_BYTES of split boundaries.  They depend on nothing else, so a caller may build them once
// (gcb_fwd_build_tables) and hand them to every forward of that shape.
size_t tc_fwd_tables_bytes() { return TC_TABS_BYTES + TC_SPLITS_BYTES; }

int tc_fwd_build_tables(const RotTable& rot, int n_rot, int v_out, int R, int A, int F, int T, int precision, void* tables,
                        cudaStream_t st) {
  TcFwdPlan plan;
  int rc = tc_fwd_plan(rot, n_rot, v_out, R, A, F, T, precision, &plan);
  if (rc) return rc;
  TcFwdParams p{};
  tc_fwd_plan_params(plan, rot, n_rot, v_out, R, A, F, T, &p);
  k_tc_build_tabs<<<plan.cfg.n_groups, 128, p.tab_bytes, st>>>(p, plan.pairs ? 1 : 0, p.tab_bytes, (uint8_t*)tables,
                                                                (int*)((uint8_t*)tables + TC_TABS_BYTES));
  GCB_LAUNCH_OK();
  return GCB_OK;
}

size_t tc_fwd_records_bytes(int v_out, int R, int A) { return tc_recs_bytes(v_out, R, A) + 256; }

// records = the slot-major packed records, then 256 bytes whose first word is max_slots(|w0|+|w1|+|w2|) as fp32 bits
int tc_fwd_pack_records(const int32_t* idx, const float* w, int v_out, int R, int A, void* records, cudaStream_t st) {
  unsigned* tail = (unsigned*)((char*)records + tc_recs_bytes(v_out, R, A));
  GCB_CUDA_OK(cudaMemsetAsync(tail, 0, 256, st));
  return tc_pack_recs_max(idx, w, v_out, R, A, (uint4*)records, tail, st);
}

int tc_conv_fwd(const float* x, const int32_t* idx, const float* w, const float* w_eff, const float* w_self,
                const float* bias, const RotTable& rot, int n_rot, int v_out, int v_src, int R, int A, int F,
                int T, int act, int precision, float* y, float* z, int32_t* argmax, float* interp_out,
                const void* records, const void* tables, uint32_t* stats_out, int row0,
                void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TcFwdPlan plan;
  int rc = tc_fwd_plan(rot, n_rot, v_out, R, A, F, T, precision, &plan);
  if (rc) return rc;
  const bool half = plan.half, x3 = plan.x3;
  const TcConfig cfg = plan.cfg;
  GCB_REQUIRE(workspace_bytes >= tc_fwd_workspace_bytes(v_out, R, A, F, T, n_rot, precision), GCB_EWORKSPACE,
              "tc_conv_fwd: workspace too small");
  GCB_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)bias % 16 == 0) &&
                  ((uintptr_t)z % 16 == 0),
              GCB_EINVAL, "tc_conv_fwd: x, y, z and bias must be 16-byte aligned");
  const int ksplit = plan.ksplit;
  const int RA = R * A, KRA = RA * F, kcat = KRA + F;
  const int m_tiles = (int)ceil_div64(v_out, TC_M);
  const int v_pad = m_tiles * TC_M;
  char* ws = (char*)align_up((size_t)workspace, 256);
  uint4* recs = (uint4*)ws;
  ws += tc_recs_bytes(v_out, R, A);
  float* wc_hi = (float*)ws;
  ws += tc_wcat_bytes(R, A, F, T);
  float* wc_lo = nullptr;
  if (x3) {
    wc_lo = (float*)ws;
    ws += tc_wcat_bytes(R, A, F, T);
  }
  float* partial = (float*)ws;
  ws += tc_partial_bytes(v_out, n_rot, T, ksplit);
  float* nrm2 = (float*)ws;
  ws += tc_nrm2_bytes(v_out, n_rot, T);
  unsigned* tickets = (unsigned*)ws;            // zeroed together with the operand maxima that follow them
  const size_t ticket_bytes = tc_ticket_bytes(v_out, n_rot, T);
  ws += ticket_bytes;
  unsigned* mx = (unsigned*)ws;                 // fp16 mode: operand maxima, then the scales
  float* scales = half ? (float*)(mx + 8) : nullptr;
  // GCB_TC_FUSE (read per call, so one process can run both): 0 (default) = k_tc_finish reduces the K splits in a
  // launch of its own; 1 = the reduction, bias, activation and pooling run inside k_tc_conv (last-arriving split of a
  // unit); 2 = the fused mode's dispatch order with the separate finish.  Measured at the FAUST shape (profiles/
  // r02_notes.md): 1 removes the launch and 4/5 of the DRAM traffic but a finishing CTA needs ~75 us for its 1.5 MB
  // of partial sums (one SM pulls ~13 B/cycle from L2) while its SM issues no MMAs: 640 us against 499 + 38.
  const int fuse_env = getenv("GCB_TC_FUSE") ? atoi(getenv("GCB_TC_FUSE")) : 0;
  const long long grid_x = (long long)ksplit * m_tiles * (T / cfg.n_t);
  const bool fuse = fuse_env != 0 && grid_x < (1ll << 31) && T % 4 == 0;
  const unsigned* ext_wsum = nullptr;
  if (records) {   // cached packing of this barycentric tensor (gcb_fwd_pack_records)
    recs = (uint4*)records;
    ext_wsum = (const unsigned*)((const char*)records + tc_recs_bytes(v_out, R, A));
  }

  CUtensorMap mh, ml;
  if (fuse) GCB_CUDA_OK(cudaMemsetAsync(tickets, 0, ticket_bytes + (half ? 32 : 0), st));
  else if (half) GCB_CUDA_OK(cudaMemsetAsync(mx, 0, 32, st));
  if (half) {
    if (!records) {
      rc = tc_pack_recs_max(idx, w, v_out, R, A, recs, mx + 1, st);
      if (rc) return rc;
    }
    TcMaxJob job{};
    job.n_seg = 3; job.mode = 1;
    job.ext_wsum = ext_wsum;
    job.seg[0] = {x, (long long)v_src * F, 0, 0};
    job.seg[1] = {w_eff, (long long)T * RA * F, 0, 2};
    job.seg[2] = {w_self, (long long)T * F, 0, 2};
    rc = tc_maxima_scales(job, mx, scales, st);
    if (rc) return rc;
    rc = tc_build_wcat_h(w_eff, w_self, T, R, A, F, scales + 1, (__half*)wc_hi, (__half*)wc_lo, st);
    if (rc) return rc;
    rc = make_weight_tmap_h(&mh, (const __half*)wc_hi, T, kcat, cfg.kc, cfg.n_t);
    if (rc) return rc;
    rc = make_weight_tmap_h(&ml, (const __half*)wc_lo, T, kcat, cfg.kc, cfg.n_t);
    if (rc) return rc;
    if (stats_out)   // {max|X|, max sum|w|, max|Wcat|, ...} for the backward of the same step (word 7 = ticket != 0: valid)
      GCB_CUDA_OK(cudaMemcpyAsync(stats_out, mx, 32, cudaMemcpyDeviceToDevice, st));
  } else {
    if (!records) {
      rc = tc_pack_recs(idx, w, v_out, R, A, recs, st);
      if (rc) return rc;
    }
    rc = tc_build_wcat(w_eff, w_self, T, R, A, F, wc_hi, wc_lo, st);
    if (rc) return rc;
    rc = make_weight_tmap(&mh, wc_hi, T, kcat, cfg.kc, cfg.n_t);
    if (rc) return rc;
    rc = make_weight_tmap(&ml, x3 ? wc_lo : wc_hi, T, kcat, cfg.kc, cfg.n_t);
    if (rc) return rc;
    if (stats_out) GCB_CUDA_OK(cudaMemsetAsync(stats_out, 0, 32, st));
  }

  TcFwdParams p;
  tc_fwd_plan_params(plan, rot, n_rot, v_out, R, A, F, T, &p);
  p.x = x; p.recs = recs; p.partial = partial; p.interp_out = interp_out; p.scales = scales;
  p.x_centre = x + (size_t)row0 * F;
  p.fuse = fuse ? (fuse_env == 2 ? 2 : 1) : 0;
  p.tickets = tickets; p.nrm2 = nrm2; p.bias = bias; p.y = y; p.z = z; p.argmax = argmax; p.act = act;
  {
    static const int discard_env = getenv("GCB_TC_DISCARD") ? atoi(getenv("GCB_TC_DISCARD")) : 1;   // A/B
    p.discard = discard_env ? 1 : 0;
  }
  if (tables) {
    p.tabs = (const uint8_t*)tables;
  } else {
    p.tabs = (const uint8_t*)mx + TC_SCALES_BYTES;
  }
  p.splits = (const int*)(p.tabs + TC_TABS_BYTES);
  static const int debug = getenv("GCB_TC_DEBUG") ? atoi(getenv("GCB_TC_DEBUG")) : 0;
  p.debug = debug;
  p.trace = nullptr;
  p.cta_log = nullptr;
  static long long* trace_dev = nullptr;
  static long long* clog_dev = nullptr;
  const size_t n_ctas = (size_t)m_tiles * cfg.n_groups * (T / cfg.n_t) * ksplit;
  if (getenv("GCB_TC_TRACE") && atoi(getenv("GCB_TC_TRACE")) == 2) {
    if (!clog_dev) GCB_CUDA_OK(cudaMalloc(&clog_dev, 16384 * 8 * sizeof(long long)));
    if (n_ctas <= 16384) {
      GCB_CUDA_OK(cudaMemsetAsync(clog_dev, 0, n_ctas * 8 * sizeof(long long), st));
      p.cta_log = clog_dev;
    }
  }
  if (getenv("GCB_TC_TRACE")) {   // investigation aid only: the one place the library touches cudaMalloc
    if (!trace_dev) GCB_CUDA_OK(cudaMalloc(&trace_dev, 4 * 128 * 8 * sizeof(long long)));
    GCB_CUDA_OK(cudaMemsetAsync(trace_dev, 0, 4 * 128 * 8 * sizeof(long long), st));
    p.trace = trace_dev;
  }
  if (!tables) {
    k_tc_build_tabs<<<cfg.n_groups, 128, p.tab_bytes, st>>>(p, plan.pairs ? 1 : 0, p.tab_bytes, (uint8_t*)p.tabs,
                                                             (int*)p.splits);
    GCB_LAUNCH_OK();
  }
  dim3 grid((unsigned)m_tiles, (unsigned)cfg.n_groups, (unsigned)((T / cfg.n_t) * ksplit));
  if (fuse) grid = dim3((unsigned)grid_x, (unsigned)cfg.n_groups, 1u);
  if (half) rc = cfg.kc == 16 ? tc_launch<16, true, true>(mh, ml, p, cfg, grid, st) : tc_launch<8, true, true>(mh, ml, p, cfg, grid, st);
  else if (cfg.kc == 32) rc = x3 ? tc_launch<32, true>(mh, ml, p, cfg, grid, st) : tc_launch<32, false>(mh, ml, p, cfg, grid, st);
  else if (cfg.kc == 16) rc = x3 ? tc_launch<16, true>(mh, ml, p, cfg, grid, st) : tc_launch<16, false>(mh, ml, p, cfg, grid, st);
  else rc = x3 ? tc_launch<8, true>(mh, ml, p, cfg, grid, st) : tc_launch<8, false>(mh, ml, p, cfg, grid, st);
  if (rc) return rc;
  if (p.fuse != 1) {
    const int warps_per_block = 8;
    k_tc_finish<<<(unsigned)ceil_div64(v_out, warps_per_block), warps_per_block * 32, 0, st>>>(
        partial, bias, ksplit, v_out, v_pad, n_rot, T, act, scales, y, z, argmax);
    GCB_LAUNCH_OK();
  }
  if (p.cta_log) {
    // per-CTA life times: printed by dispatch order in groups of 148 (one "wave" if all CTAs took equally long)
    static long long hl[16384 * 8];
    GCB_CUDA_OK(cudaMemcpyAsync(hl, p.cta_log, n_ctas * 8 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    GCB_CUDA_OK(cudaStreamSynchronize(st));
    long long t_min = hl[0], t_max = 0;
    for (size_t i = 0; i < n_ctas; ++i) { if (hl[i * 8] < t_min) t_min = hl[i * 8]; if (hl[i * 8 + 1] > t_max) t_max = hl[i * 8 + 1]; }
    fprintf(stderr, "[cta log] %zu CTAs, kernel span %.1f us\n", n_ctas, (t_max - t_min) * 1e-3);
    {
      // dispatch order = ascending start time
      static size_t order[16384];
      for (size_t i = 0; i < n_ctas; ++i) order[i] = i;
      for (size_t i = 1; i < n_ctas; ++i) {   // insertion sort (nearly sorted already)
        const size_t o = order[i];
        size_t j = i;
        while (j > 0 && hl[order[j - 1] * 8] > hl[o * 8]) { order[j] = order[j - 1]; --j; }
        order[j] = o;
      }
      for (size_t g0 = 0; g0 < n_ctas; g0 += 148) {
        const size_t g1 = g0 + 148 < n_ctas ? g0 + 148 : n_ctas;
        double s_ns = 0, s_cy = 0, st0 = 0, en = 0, s_main = 0, s_fin = 0; long long mx_ns = 0; int n_fin = 0;
        for (size_t q = g0; q < g1; ++q) {
          const size_t i = order[q];
          const long long ns = hl[i * 8 + 1] - hl[i * 8], cy = hl[i * 8 + 3] - hl[i * 8 + 2];
          s_ns += ns; s_cy += cy; st0 += hl[i * 8] - t_min; en += hl[i * 8 + 1] - t_min;
          if (hl[i * 8 + 5]) s_main += hl[i * 8 + 5] - hl[i * 8];
          if (hl[i * 8 + 6]) { ++n_fin; s_fin += hl[i * 8 + 7] - hl[i * 8 + 5]; }
          if (ns > mx_ns) mx_ns = ns;
        }
        const double n = (double)(g1 - g0);
        fprintf(stderr, "  CTAs %5zu-%5zu (by start): mean start %7.1f us, mean end %7.1f us, life %6.1f us (max %6.1f) = %7.0f cycles -> %.0f MHz; to ticket %6.1f us; %d finishers, %6.1f us each\n",
                g0, g1 - 1, st0 / n * 1e-3, en / n * 1e-3, s_ns / n * 1e-3, mx_ns * 1e-3, s_cy / n, s_cy / s_ns * 1e3,
                s_main / n * 1e-3, n_fin, n_fin ? s_fin / n_fin * 1e-3 : 0.0);
      }
    }
  }
  if (p.trace) {
    static long long host[4 * 128 * 8];
    GCB_CUDA_OK(cudaMemcpyAsync(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost, st));
    GCB_CUDA_OK(cudaStreamSynchronize(st));
    long long t0 = host[0];
    const char* names[4] = {"prod g0: start issued landed empty_ok stored fenced arrived", "scout: start entries a_full",
                            "mma: start waited done a_full_next", "tma: start slot_free issued ss_top"};
    {
      long long* e = host + ((size_t)3 * 128 + 127) * 8;
      fprintf(stderr, "[trace life: entry %lld, prologue done %lld, producers done %lld, accumulators complete %lld, drained %lld, exit %lld]\n",
              e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0);
    }
    for (int role = 0; role < 4; ++role) {
      fprintf(stderr, "[trace %s]\n", names[role]);
      for (int n = 0; n < 48; ++n) {
        long long* e = host + ((size_t)role * 128 + n) * 8;
        if (!e[0]) continue;
        fprintf(stderr, "  %3d:", n);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %8lld", e[k] ? e[k] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return GCB_OK;
}

}  // namespace gcb
