// Shared device/host helpers for the bdpose sm_100a kernels.
// Everything here is internal to libbdpose.so; the public surface is include/bdpose.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bdpose.h"

#define BDP_EPS 1e-6f          // helperFunctions.py:20  (clamp / small-angle threshold)
#define BDP_NORM_EPS 1e-12f    // torch.nn.functional.normalize default eps

#define BDP_FULL_MASK 0xffffffffu

// ---- error plumbing -------------------------------------------------------------------------
// No exceptions cross the C ABI: every entry point returns a status and parks a message in a
// thread-local buffer that bdp_last_error() hands back.
void bdp_set_error(const char* fmt, ...);

#define BDP_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      bdp_set_error(__VA_ARGS__);               \
      return BDP_ERR_ARG;                       \
    }                                           \
  } while (0)

#define BDP_CUDA_CHECK_LAUNCH(what)                                              \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      bdp_set_error("%s: launch failed: %s", what, cudaGetErrorString(e__));     \
      return BDP_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define BDP_CUDA_CALL(expr)                                                      \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      bdp_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));            \
      return BDP_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

int bdp_num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device

// ---- warp helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BDP_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BDP_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(BDP_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BDP_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BDP_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BDP_FULL_MASK, v, o);
  return v;
}

// (value, index) max with lowest-index tie-break: the torch.max / np.argmax "first maximal
// value" rule that the reference relies on (learnGeodesicBDModel.py:175, 217).
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(BDP_FULL_MASK, v, o);
    int oi = __shfl_xor_sync(BDP_FULL_MASK, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// streaming 128-bit global accesses: the big per-row operands are touched exactly once.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
