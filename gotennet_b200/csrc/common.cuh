// Shared device/host helpers for the gotennet_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gotennet_b200.h"

namespace goten {

extern thread_local char g_err[512];
int set_error(const char* fmt, ...);

#define GOTEN_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return goten::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

// every kernel launch is followed by this macro: it also feeds goten_launch_count()
extern unsigned long long g_launches;
#define GOTEN_CHECK_LAUNCH()                    \
  do {                                          \
    ++goten::g_launches;                        \
    GOTEN_CHECK_CUDA(cudaPeekAtLastError());    \
  } while (0)

#define GOTEN_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return goten::set_error(__VA_ARGS__);             \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device math
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// Fast forms for the BACKWARD kernels (gradient factors): ex2.approx + rcp.approx, relative error ~1e-6 for |x| < 10
// (the forward kernels keep the exact forms above).  __expf(-x) -> inf for x < -88 and the quotient -> 0, as exact.
__device__ __forceinline__ float sigmoid_fast_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_fast_(float x) { return x * sigmoid_fast_(x); }
// d/dx silu(x) = s + x s (1 - s)
__device__ __forceinline__ float dsiluf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of one float per thread; result valid in every thread.  `red` needs
// >= 33 floats of shared memory.  Contains two __syncthreads().
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// Block-wide sums of NV per-thread values with ONE barrier (geometry gradients: per edge, 1 + L channel sums).
// Every thread stores its values into `scratch` (row pitch blockDim + 4 floats: the quad reads below are bank-conflict
// free), one __syncthreads, then quads of threads add one row each and quad lane 0 hands the total to sink(v, sum).
// Fixed summation order: results are run-to-run identical.  Callers alternate between two scratch buffers on
// consecutive calls (a buffer is re-written only after the next call's barrier, i.e. after everybody has read it) and
// must call with all threads of the block.  Floats needed per buffer: NV * (blockDim + 4).
template <int NV, class Sink>
__device__ __forceinline__ void block_sums_one_barrier(const float (&vals)[NV], float* scratch, Sink sink) {
  const int pitch = blockDim.x + 4;
#pragma unroll
  for (int v = 0; v < NV; ++v) scratch[v * pitch + threadIdx.x] = vals[v];
  __syncthreads();
  const int r = threadIdx.x & 3;
  for (int v0 = 0; v0 < NV; v0 += (int)(blockDim.x >> 2)) {   // uniform trip count: the shuffles see all lanes
    const int v = v0 + (int)(threadIdx.x >> 2);
    float s = 0.f;
    if (v < NV) {
      const float* row = scratch + v * pitch;
      for (int i = r; i < (int)blockDim.x; i += 4) s += row[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (v < NV && r == 0) sink(v, s);
  }
}

// Running max |.| of the values a kernel writes into a GEMM operand: the split-fp16 GEMM (gemm_tc16.cu) scales its
// operands by a power of two derived from this bound.  `out` holds a non-negative float, compared as unsigned bits.
__device__ __forceinline__ float amax4(float m, float a, float b, float c, float d) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(a), fabsf(b))), fmaxf(fabsf(c), fabsf(d)));
}
__device__ __forceinline__ void amax_commit(float* out, float amx) {
  if (out != nullptr && amx > __ldcg(out)) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(amx));
}

// Block-wide form (one request per block): every thread of the block must call it.
__device__ __forceinline__ void block_amax_commit(float* out, float amx) {
  if (out == nullptr) return;  // block-uniform
  __shared__ float s_amax_red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  amx = warp_max(amx);
  if (lane == 0) s_amax_red[w] = amx;
  __syncthreads();
  if (w == 0) {
    float v = lane < nw ? s_amax_red[lane] : 0.f;
    v = warp_max(v);
    if (lane == 0) amax_commit(out, v);
  }
}

// degree-l block boundaries inside the L axis: l = 1..lmax occupies [l^2-1, (l+1)^2-1)
__host__ __device__ __forceinline__ constexpr int blk_lo(int l) { return l * l - 1; }            // l >= 1
__host__ __device__ __forceinline__ constexpr int blk_hi(int l) { return (l + 1) * (l + 1) - 1; }
__host__ __device__ __forceinline__ constexpr int deg_of(int m) { return m < 3 ? 1 : (m < 8 ? 2 : (m < 15 ? 3 : 4)); }

}  // namespace goten
