// Shared helpers for the audiogan_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/audiogan_b200.h"

namespace ag {

void set_error(const char* fmt, ...);
// diagnostics of the recurrent dispatch (abi.cu): why a fast path declined / which family ran (ag_lstm_last_path)
void set_decline(const char* fmt, ...);
void clear_decline();
void set_path(const char* family);

#define AG_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ag::set_error(__VA_ARGS__);               \
      return AG_EINVAL;                         \
    }                                           \
  } while (0)

#define AG_CUDA(call)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ag::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return AG_ECUDA;                                                                \
    }                                                                                 \
  } while (0)

#define AG_LAUNCH_CHECK() AG_CUDA(cudaGetLastError())

int sm_count();
int smem_optin();

__device__ __forceinline__ float ld_any(const void* p, int64_t i, int dtype) {
  return dtype ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
               : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, int64_t i, float v, int dtype) {
  if (dtype) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
// 4 consecutive elements at element index i (16-byte aligned for fp32, 8-byte for bf16), read-only path / plain store
__device__ __forceinline__ float4 ldg4_any(const void* p, int64_t i, int dtype) {
  if (dtype == 0) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ float4 ld4_plain_any(const void* p, int64_t i, int dtype) {
  if (dtype == 0) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
  const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ void st4_any(void* p, int64_t i, float4 v, int dtype) {
  if (dtype == 0) { *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v; return; }
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum; `red` is >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  if (threadIdx.x == 0) red[0] = v;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace ag
