// Microbenchmark: cycle cost of the synchronisation primitives the tensor-core kernels use per tile.
// One CTA, one warp measured; prints average cycles per operation.  nvcc -arch=sm_100a sync_cost.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geoconv_b200/csrc/gcb_ptx.cuh"
using namespace gcb;

__global__ void k(long long* out, int iters) {
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t b0 = ptx::smem_u32(&bars[0]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) ptx::mbar_init(b0 + 8 * i, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 32); ptx::tmem_relinquish(); }
  __syncthreads();
  if (warp != 0) return;
  // complete phase 0 of barrier 0 so that waits on parity 0 succeed immediately
  if (lane == 0) ptx::mbar_arrive(b0);
  __syncwarp();
  long long t0, t1;
  int r = 0;
  // (a) try_wait already complete, all 32 lanes
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { ptx::mbar_wait(b0, 0); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (b) try_wait already complete, one lane + syncwarp
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { if (lane == 0) ptx::mbar_wait(b0, 0); __syncwarp(); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (c) elect_one + syncwarp
  t0 = clock64();
  int acc = 0;
  for (int i = 0; i < iters; ++i) { if (ptx::elect_one()) acc += i; __syncwarp(); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (d) tcgen05.commit by elected lane onto barrier 1 (count 1): completes a phase each time
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { if (ptx::elect_one()) ptx::umma_commit(b0 + 8); __syncwarp(); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (e) mbarrier.arrive by one lane on barrier 2
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { if (lane == 0) ptx::mbar_arrive(b0 + 16); __syncwarp(); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (f) fence.proxy.async (all lanes)
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { ptx::fence_proxy_async_smem(); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (g) tcgen05.fence::after_thread_sync
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { ptx::tc_fence_after(); }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (h) commit + wait for it (round trip): commit on barrier 3 then wait its phase
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (ptx::elect_one()) ptx::umma_commit(b0 + 24);
    __syncwarp();
    ptx::mbar_wait(b0 + 24, i & 1);
  }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (i) arrive + wait round trip (one lane arrives, all wait)
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (lane == 0) ptx::mbar_arrive(b0 + 32);
    __syncwarp();
    ptx::mbar_wait(b0 + 32, i & 1);
  }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (j) __shfl_sync x8
  t0 = clock64();
  int v = lane;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) v += __shfl_sync(0xffffffffu, v, (i + q) & 31);
  }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / iters; ++r;
  // (k) try_wait NOT complete -> how long one failing try_wait blocks (barrier 5 never completes)
  t0 = clock64();
  int fails = 0;
  for (int i = 0; i < 64; ++i) { fails += ptx::mbar_try_wait(b0 + 40, 0) ? 0 : 1; }
  t1 = clock64(); if (lane == 0) out[r] = (t1 - t0) / 64; ++r;
  if (lane == 0) out[15] = acc + v + fails;
  __syncwarp();
  ptx::tmem_dealloc(tmem_slot, 32);
}

int main() {
  long long* d; cudaMalloc(&d, 16 * sizeof(long long)); cudaMemset(d, 0, 16 * sizeof(long long));
  k<<<1, 64>>>(d, 1000);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"try_wait complete, 32 lanes", "try_wait complete, 1 lane + syncwarp", "elect + syncwarp",
                         "tcgen05.commit (elected) + syncwarp", "mbarrier.arrive (1 lane) + syncwarp",
                         "fence.proxy.async", "tcgen05.fence::after", "commit -> wait round trip",
                         "arrive -> wait round trip", "8 x shfl", "failing try_wait (time limit)"};
  for (int i = 0; i < 11; ++i) printf("%-42s %6lld cycles\n", names[i], h[i]);
  return 0;
}
