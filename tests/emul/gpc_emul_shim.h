// TEST INFRASTRUCTURE.  Lets g++ compile geoconv_b200/csrc/gcb_gpc.cu's algorithm text for the host with a
// one-lane "warp" (GCB_WARP = 1): every warp-wide primitive degenerates to its single-lane meaning.  Used by
// tests/test_gpc_wavefront.py (CPU suite) to run the kernel's logic against the reference fixtures without a GPU.  The harness (test_gpc_wavefront.py::emul) places the
// device section of the product source behind this header, unchanged.
#pragma once
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/geoconv_b200.h"

#define GCB_WARP 1
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__
#define __align__(n)

struct EmulDim { unsigned x, y, z; };
static EmulDim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1};
namespace gcb { alignas(16) unsigned char smem[1 << 20]; }

static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline void __syncwarp() {}
static inline int __ffs(int v) { return __builtin_ffs(v); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
// IEEE operations; g++ for baseline x86-64 does not contract them
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
