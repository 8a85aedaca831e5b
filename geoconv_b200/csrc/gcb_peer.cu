// Data-parallel weight-gradient all-reduce over NVLink peer memory, written for the one place the hot path has an
// exchange step (batches of meshes): after the weight-gradient stage of the backward every rank holds
// d[Weff | Wself | bias] (8.6 MB at the FAUST shape) and needs the sum over the ranks while its dSignal GEMM is still
// running on 135 of the 148 SMs.  NCCL's all-reduce with the 8 CTAs that fit beside that GEMM needs ~145 us at 8
// ranks (86 us of GEMM to hide behind: 60 us exposed, bench N = 8).  This kernel does the same job as a two-shot
// all-reduce on directly mapped peer buffers (torch symmetric memory: every rank's buffer is mapped into every
// rank's address space; NVSwitch gives every pair full bandwidth):
//
//   phase 0   copy the local gradients into this rank's staging area S_r (peer-readable); the CTA that finishes last
//             tells every peer "S_r is ready" (flag = epoch, release at system scope);
//   phase 1   wait for all peers' flags; rank r owns slice r of the flattened gradient: it adds S_0 .. S_{W-1} over
//             that slice in rank order (fixed order: every rank ends up with the same bits, run after run) and
//             writes the sums into every rank's result area OUT_q, its own gradients directly in place;
//   phase 2   the CTA that finishes last tells every peer "slice r delivered"; when all W slices have arrived the
//             other ranks' slices are copied from OUT_r into the local gradient tensors.
//
// Nobody on THIS GPU waits for another kernel of this GPU (the spins are on flags written by other GPUs, whose
// kernels run on their own hardware); a bounded spin turns a missing peer into an error flag instead of a hang.
// The epoch lives in device memory and is advanced by the kernel itself, so the launch can sit in a CUDA graph.
#include "gcb_common.cuh"

namespace gcb {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_THREADS = 512;
constexpr long long PEER_SPIN_LIMIT = 4000000000ll;   // clock64 ticks (~2 s): then give up and raise the error word

struct PeerParams {
  float* seg_ptr[3];          // the local gradient tensors (d_w_eff, d_w_self, d_bias), reduced in place
  long long seg_end[3];       // cumulative element counts: segment i holds flat elements [seg_end[i-1], seg_end[i])
  long long n;                // total elements (multiple of 4; every segment 16-byte aligned and a multiple of 4 long)
  const unsigned long long* bufs;   // device array [world]: base address of every rank's symmetric buffer
  int rank, world;
  long long out_off;          // float offset of OUT inside a symmetric buffer (S starts at 0)
  long long flag_off;         // byte offset of the flag block: uint32 ready[world], delivered[world], then
                              // {epoch, ticket0, ticket1, ticket2, error} (local words, never written by peers)
  int staged;                 // 1: the caller has already written its gradients into S_r and takes the sums from OUT_r
                              // (no copies in this kernel: the backward writes its weight gradients straight into S_r)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float4* peer_seg_addr(const PeerParams& p, long long e, float* const (&seg)[3]) {
  // flat element e (multiple of 4) -> address inside the local gradient tensors
  if (e < p.seg_end[0]) return reinterpret_cast<float4*>(seg[0] + e);
  if (e < p.seg_end[1]) return reinterpret_cast<float4*>(seg[1] + (e - p.seg_end[0]));
  return reinterpret_cast<float4*>(seg[2] + (e - p.seg_end[1]));
}

// wait until every word of flags[0 .. world) has reached `epoch` (flags only grow)
__device__ __forceinline__ bool peer_wait_all(const unsigned* flags, int world, unsigned epoch, unsigned* error) {
  if (threadIdx.x < (unsigned)world) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > PEER_SPIN_LIMIT) { atomicExch(error, 1u); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  return true;
}

// WC = compile-time cap of the world size; U = 16 / WC slice elements per thread and iteration, so that 16 loads of
// 16 bytes are in flight per thread whatever the world size (NVLink round trips are ~2 us: with one element per
// iteration the first version needed 187 us for 8.6 MB on 12 CTAs at 2 ranks)
template <int WC>
__global__ void __launch_bounds__(PEER_THREADS) k_peer_allreduce(const PeerParams p) {
  constexpr int U = 16 / WC;
  float* const seg[3] = {p.seg_ptr[0], p.seg_ptr[1], p.seg_ptr[2]};
  char* const mine = reinterpret_cast<char*>(p.bufs[p.rank]);
  unsigned* const ready = reinterpret_cast<unsigned*>(mine + p.flag_off);
  unsigned* const delivered = ready + p.world;
  unsigned* const state = delivered + p.world;             // {epoch, ticket0, ticket1, ticket2, error}
  __shared__ bool s_last;
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(state) + 1u;   // this call's epoch (state[0] = last finished)
  const long long n4 = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x, tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  // ---- phase 0: local gradients -> S_r (skipped when the caller staged them itself) ----
  float4* const s_me = reinterpret_cast<float4*>(mine);
  if (!p.staged) {
    for (long long i0 = tid0; i0 < n4; i0 += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * stride < n4) v[u] = *peer_seg_addr(p, (i0 + u * stride) << 2, seg);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * stride < n4) s_me[i0 + u * stride] = v[u];
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(state + 1, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < (unsigned)p.world)
      st_release_sys(reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p.bufs[threadIdx.x]) + p.flag_off) + p.rank, epoch);
    if (threadIdx.x == 0) state[1] = 0u;
  }

  // ---- phase 1: reduce this rank's slice over all ranks, deliver it to everybody ----
  peer_wait_all(ready, p.world, epoch, state + 4);
  const long long lo4 = n4 * p.rank / p.world, hi4 = n4 * (p.rank + 1) / p.world;
  for (long long i0 = lo4 + tid0; i0 < hi4; i0 += U * stride) {
    float4 v[U][WC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < WC; ++q)
        if (q < p.world && i0 + u * stride < hi4) v[u][q] = __ldcg(reinterpret_cast<const float4*>(p.bufs[q]) + i0 + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi4) break;
      float4 acc = v[u][0];
#pragma unroll
      for (int q = 1; q < WC; ++q)
        if (q < p.world) { acc.x += v[u][q].x; acc.y += v[u][q].y; acc.z += v[u][q].z; acc.w += v[u][q].w; }
#pragma unroll
      for (int q = 0; q < WC; ++q)
        if (q < p.world && (p.staged || q != p.rank)) (reinterpret_cast<float4*>(p.bufs[q]) + (p.out_off >> 2))[i] = acc;
      if (!p.staged) *peer_seg_addr(p, i << 2, seg) = acc;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(state + 2, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < (unsigned)p.world)
      st_release_sys(reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p.bufs[threadIdx.x]) + p.flag_off) + p.world + p.rank, epoch);
    if (threadIdx.x == 0) state[2] = 0u;
  }

  // ---- phase 2: all slices have arrived in OUT_r; unless staged, copy the other ranks' slices into the local tensors ----
  peer_wait_all(delivered, p.world, epoch, state + 4);
  if (!p.staged) {
    const float4* const out_me = reinterpret_cast<const float4*>(mine) + (p.out_off >> 2);
    for (long long i0 = tid0; i0 < n4; i0 += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4 && (i < lo4 || i >= hi4)) v[u] = __ldcg(out_me + i);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4 && (i < lo4 || i >= hi4)) *peer_seg_addr(p, i << 2, seg) = v[u];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 3, 1u) == gridDim.x - 1) {   // everybody has read the epoch and is done: advance it
      state[3] = 0u;
      __threadfence();
      state[0] = epoch;
    }
  }
}

}  // namespace gcb

using namespace gcb;

extern "C" size_t gcb_peer_allreduce_bytes(int64_t n_elements, int32_t world) {
  const size_t n = (size_t)((n_elements + 3) / 4 * 4);
  return align_up(2 * n * sizeof(float), 256) + align_up((size_t)(2 * world + 8) * sizeof(unsigned), 256);
}

static int peer_launch(const PeerParams& p, int n_ctas, cudaStream_t st) {
  if (n_ctas < 1) n_ctas = 12;
  if (n_ctas > 64) n_ctas = 64;   // all CTAs must be resident together: the phases hand over through tickets and flags
  if (p.world <= 2) k_peer_allreduce<2><<<n_ctas, PEER_THREADS, 0, st>>>(p);
  else if (p.world <= 4) k_peer_allreduce<4><<<n_ctas, PEER_THREADS, 0, st>>>(p);
  else if (p.world <= 8) k_peer_allreduce<8><<<n_ctas, PEER_THREADS, 0, st>>>(p);
  else k_peer_allreduce<16><<<n_ctas, PEER_THREADS, 0, st>>>(p);
  GCB_LAUNCH_OK();
  return GCB_OK;
}

extern "C" int gcb_peer_allreduce_staged(int64_t n_elements, const uint64_t* peer_bufs_dev, int32_t rank, int32_t world,
                                         int32_t n_ctas, void* stream) {
  GCB_REQUIRE(n_elements > 0 && n_elements % 4 == 0 && peer_bufs_dev, GCB_EINVAL, "gcb_peer_allreduce_staged: bad argument");
  GCB_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, GCB_EINVAL,
              "gcb_peer_allreduce_staged: rank %d / world %d out of range (at most %d ranks)", rank, world, PEER_MAX_WORLD);
  PeerParams p{};
  p.n = n_elements;
  p.seg_end[0] = p.seg_end[1] = p.seg_end[2] = n_elements;
  p.bufs = reinterpret_cast<const unsigned long long*>(peer_bufs_dev);
  p.rank = rank; p.world = world;
  p.out_off = p.n;
  p.flag_off = (long long)align_up(2 * (size_t)p.n * sizeof(float), 256);
  p.staged = 1;
  return peer_launch(p, n_ctas, (cudaStream_t)stream);
}

extern "C" int gcb_peer_allreduce(float* g0, int64_t n0, float* g1, int64_t n1, float* g2, int64_t n2,
                                  const uint64_t* peer_bufs_dev, int32_t rank, int32_t world, int32_t n_ctas,
                                  void* stream) {
  GCB_REQUIRE(g0 && n0 > 0 && peer_bufs_dev, GCB_EINVAL, "gcb_peer_allreduce: null pointer");
  GCB_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, GCB_EINVAL,
              "gcb_peer_allreduce: rank %d / world %d out of range (at most %d ranks)", rank, world, PEER_MAX_WORLD);
  GCB_REQUIRE(n0 % 4 == 0 && n1 % 4 == 0 && n2 % 4 == 0 && (uintptr_t)g0 % 16 == 0 && (!g1 || (uintptr_t)g1 % 16 == 0) &&
                  (!g2 || (uintptr_t)g2 % 16 == 0),
              GCB_EINVAL, "gcb_peer_allreduce: every tensor must be 16-byte aligned and a multiple of 4 elements long");
  PeerParams p{};
  p.seg_ptr[0] = g0; p.seg_ptr[1] = g1 ? g1 : g0; p.seg_ptr[2] = g2 ? g2 : g0;
  p.seg_end[0] = n0; p.seg_end[1] = n0 + (g1 ? n1 : 0); p.seg_end[2] = p.seg_end[1] + (g2 ? n2 : 0);
  p.n = p.seg_end[2];
  p.bufs = reinterpret_cast<const unsigned long long*>(peer_bufs_dev);
  p.rank = rank; p.world = world;
  p.out_off = p.n;
  p.flag_off = (long long)align_up(2 * (size_t)p.n * sizeof(float), 256);
  p.staged = 0;
  return peer_launch(p, n_ctas, (cudaStream_t)stream);
}
