// Shared helpers for the fpmatch sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define FPM_OK 0
#define FPM_ERR_ARG -1
#define FPM_ERR_UNSUPPORTED -2

extern "C" void fpm_set_error(const char* msg);

#define FPM_CHECK_ARG(cond, msg)                 \
  do {                                           \
    if (!(cond)) {                               \
      fpm_set_error(msg);                        \
      return FPM_ERR_ARG;                        \
    }                                            \
  } while (0)

#define FPM_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) {                            \
      fpm_set_error(cudaGetErrorString(e__));            \
      return (int)e__;                                   \
    }                                                    \
  } while (0)

#define FPM_LAUNCH_CHECK()                               \
  do {                                                   \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) {                            \
      fpm_set_error(cudaGetErrorString(e__));            \
      return (int)e__;                                   \
    }                                                    \
  } while (0)

static inline int fpm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

namespace fpm {

constexpr int kWarp = 32;
// ---- packed fp32 arithmetic of sm_100: one instruction, two lanes (FFMA2 / FADD2) --------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}

constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide reductions through a small shared scratch (>= 32 floats). All threads get the result.
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : kNegInf;
  return warp_max(r);
}
__device__ __forceinline__ float block_min(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : INFINITY;
  return warp_min(r);
}
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  return warp_sum(r);
}

// torch.logsumexp semantics for one (max, sum) pair: max + log(sum exp(x - max)), with the
// all--inf case returning -inf (torch substitutes 0 for an infinite max before subtracting).
__device__ __forceinline__ float lse_finish(float mx, float sum) {
  return (mx == kNegInf) ? kNegInf : mx + logf(sum);
}

__device__ __forceinline__ float softplus_torch(float x) {   // beta = 1, threshold = 20
  return x > 20.f ? x : log1pf(expf(x));
}

// softplus for values nobody reads back at full precision (the dead edge affinities Ke): fast exp / log, absolute
// error < 3e-6 (x > 8: x + log1p(e^-x) by its series; tiny e^x: series; otherwise log(1 + e^x) with __logf).
// log(1 + e^x) = max(x, 0) + log1p(e^-|x|), branch-free on ex2.approx / lg2.approx: absolute error < 3e-7
__device__ __forceinline__ float softplus_sfu(float x) {
  float t, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-fabsf(x) * 1.4426950408889634f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + t));
  return fmaf(l, 0.6931471805599453f, fmaxf(x, 0.f));
}
__device__ __forceinline__ float softplus_fast(float x) {
  if (x > 8.f) { const float u = __expf(-x); return x + u * (1.f - 0.5f * u); }
  const float t = __expf(x);
  return t < 1e-3f ? t * (1.f - 0.5f * t) : __logf(1.f + t);
}

}  // namespace fpm
