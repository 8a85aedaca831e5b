// Microbenchmark: tcgen05.mma kind::tf32 issue / execution rate from ONE thread, and from W warps in
// parallel, for several N.  Operands are whatever is in shared memory (zeros); no producer traffic.
//   cycles/MMA (back-to-back, one thread)  vs  the nominal max(M,128)*N/256 = N/2 cycles
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geoconv_b200/csrc/gcb_ptx.cuh"
using namespace gcb;

// mode 0: all MMAs into the same accumulator; 1: alternate between two accumulators
__global__ void __launch_bounds__(160, 1) k(long long* out, int n_mma, int N, int warps, int mode, int d_stride, int commits) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t sb = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(&tmem_slot), 0);
  if (warp < warps) {
    const uint64_t desc_hi = ptx::make_kmajor_desc<128>(0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_flags = (uint32_t)(ptx::make_kmajor_desc<128>(0) & 0xFFFF0000ull);
    const uint32_t a_lo = (sb >> 4) | lo_flags;                  // A tile 128 x 128 B = 16 KB
    const uint32_t b_lo = ((sb + 16384u) >> 4) | lo_flags;       // B tile up to 256 x 128 B = 32 KB
    const uint32_t idesc = ptx::make_idesc_tf32(128, N);
    const uint32_t d0 = tmem_base + (uint32_t)(warp * d_stride);
    const uint32_t bar = ptx::smem_u32(&bars[warp]);
    long long t0 = clock64();
    if (ptx::elect_one()) {
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = mode == 1 ? d0 + (uint32_t)((j & 1) * N) : d0;
          ptx::umma_tf32(d, desc_hi | (uint64_t)(a_lo + 2u * j), desc_hi | (uint64_t)(b_lo + 2u * j), idesc, 1u);
        }
        for (int c = 0; c < commits; ++c) ptx::umma_commit(ptx::smem_u32(&bars[4 + (c & 3)]));   // count-1 barriers, never waited on
      }
      ptx::umma_commit(bar);
    }
    __syncwarp();
    long long t1 = clock64();
    ptx::mbar_wait(bar, 0);
    long long t2 = clock64();
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 512;
  printf("%-6s %-6s %-5s %-9s %-7s | issue cyc/MMA  total cyc/MMA (per warp)   nominal N/2\n", "N", "warps", "mode", "d_stride", "commits");
  int Ns[] = {96, 192};
  for (int N : Ns) {
    for (int warps = 1; warps <= 2; warps *= 2) {
     for (int commits = 0; commits <= 3; ++commits) {
      for (int mode = 0; mode < 1; ++mode) {
        int d_stride = mode == 1 ? 2 * N : N;
        if (warps * d_stride > 512) continue;
        cudaMemset(d, 0, 16 * sizeof(long long));
        k<<<1, 160, 64 * 1024>>>(d, n_mma, N, warps, mode, d_stride, commits);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-6d %-6d %-5d %-9d %-7d | %8.1f %8.1f   aggregate %6.1f cyc/MMA   %d\n", N, warps, mode, d_stride, commits,
               (double)h[0] / n_mma, (double)h[1] / n_mma, (double)h[1] / n_mma / warps, N / 2);
      }
     }
    }
  }
  return 0;
}
