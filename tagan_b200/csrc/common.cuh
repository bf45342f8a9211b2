// Shared device helpers for libtagan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/tagan_b200.h"

#define TAGAN_API extern "C" __attribute__((visibility("default")))

#define FULL_MASK 0xffffffffu

static inline int tagan_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

static inline cudaStream_t as_stream(tagan_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// The opt-in to more than 48 KB of dynamic shared memory is a per-DEVICE attribute of a kernel: remember the largest size
// set per device ordinal (one instance per kernel / call site), so a process that drives several GPUs opts in on each.
struct SmemOptIn {
  size_t bytes[64] = {};
  template <typename Kern>
  cudaError_t ensure(Kern kern, size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
    if (dev >= 0 && bytes[dev] >= smem) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && dev >= 0) bytes[dev] = smem;
    return e;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
// Sum over aligned groups of `group` lanes (group a power of two <= 32).
__device__ __forceinline__ float group_sum(float v, int group) {
  for (int o = group >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// 128-bit read-only loads.  Gathered rows are reused across warps through L2, so the default
// policy is kept for them; pure streams use the no-allocate flavour.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

template <int VEC> struct VecIO;
template <> struct VecIO<1> {
  static __device__ __forceinline__ void load(float* d, const float* p) { d[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float* s) { p[0] = s[0]; }
};
template <> struct VecIO<2> {
  static __device__ __forceinline__ void load(float* d, const float* p) { float2 t = ldg2(p); d[0] = t.x; d[1] = t.y; }
  static __device__ __forceinline__ void store(float* p, const float* s) { *reinterpret_cast<float2*>(p) = make_float2(s[0], s[1]); }
};
template <> struct VecIO<4> {
  static __device__ __forceinline__ void load(float* d, const float* p) { float4 t = ldg4(p); d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float* s) { *reinterpret_cast<float4*>(p) = make_float4(s[0], s[1], s[2], s[3]); }
};

static inline int ceil_div_i64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
