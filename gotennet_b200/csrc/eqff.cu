// EQFF element-wise stages (reference representation/gotennet.py:728-748).
// The three linear maps (W_vu, gamma_m.0, gamma_m.1) run through goten_gemm; these
// kernels are the norm/concat and gated-update stages and their backward passes.
// X-like tensors are degree-major: P[L][N][C].  One thread per (node, channel),
// channel fastest -> fully coalesced 128 B per warp; pure HBM-bound streaming.
#include "common.cuh"

namespace goten {

// All four kernels are templated on the vector width V (4 = 128-bit accesses when C % 4 == 0).
template <int V>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float* out) {
  if (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else {
    out[0] = p[0];
  }
}
template <int V>
__device__ __forceinline__ void stv(float* __restrict__ p, const float* in) {
  if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
  else p[0] = in[0];
}

template <int V>
__global__ void eqff_ctx_fwd_kernel(const float* __restrict__ h, const float* __restrict__ P, int N, int C, int L,
                                    float eps, float* __restrict__ ctx, float* __restrict__ ctx_amax) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  float amx = 0.f;   // running max |ctx| for the GEMM that consumes it (no separate absmax pass)
  if (idx < (int64_t)N * C) {
    const int64_t n = idx / C;
    const int c = (int)(idx % C);
    float s[V], hv[V];
#pragma unroll
    for (int q = 0; q < V; ++q) s[q] = 0.f;
    for (int m = 0; m < L; ++m) {
      float p[V];
      ldv<V>(P + ((int64_t)m * N + n) * C + c, p);
#pragma unroll
      for (int q = 0; q < V; ++q) s[q] = fmaf(p[q], p[q], s[q]);
    }
    ldv<V>(h + idx, hv);
    stv<V>(ctx + n * 2 * C + c, hv);
#pragma unroll
    for (int q = 0; q < V; ++q) { s[q] = sqrtf(s[q] + eps); amx = fmaxf(amx, fmaxf(fabsf(hv[q]), s[q])); }
    stv<V>(ctx + n * 2 * C + C + c, s);
  }
  block_amax_commit(ctx_amax, amx);
}

template <int V>
__global__ void eqff_update_fwd_kernel(const float* __restrict__ h, const float* __restrict__ Xd,
                                       const float* __restrict__ P, const float* __restrict__ mm, int N, int C, int L,
                                       float* __restrict__ h_out, float* __restrict__ Xd_out) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (idx >= (int64_t)N * C) return;
  const int64_t n = idx / C;
  const int c = (int)(idx % C);
  float hv[V], m1[V], m2[V];
  ldv<V>(h + idx, hv);
  ldv<V>(mm + n * 2 * C + c, m1);
  ldv<V>(mm + n * 2 * C + C + c, m2);
#pragma unroll
  for (int q = 0; q < V; ++q) hv[q] += m1[q];
  stv<V>(h_out + idx, hv);
  for (int m = 0; m < L; ++m) {
    const int64_t o = ((int64_t)m * N + n) * C + c;
    float x[V], p[V];
    ldv<V>(Xd + o, x);
    ldv<V>(P + o, p);
#pragma unroll
    for (int q = 0; q < V; ++q) x[q] = fmaf(m2[q], p[q], x[q]);
    stv<V>(Xd_out + o, x);
  }
}

template <int V>
__global__ void eqff_update_bwd_kernel(const float* __restrict__ g_h_out, const float* __restrict__ g_Xd_out,
                                       const float* __restrict__ P, int N, int C, int L, float* __restrict__ g_m,
                                       float* __restrict__ gm_amax) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  float amx = 0.f;
  if (idx < (int64_t)N * C) {
    const int64_t n = idx / C;
    const int c = (int)(idx % C);
    float s[V], g[V];
#pragma unroll
    for (int q = 0; q < V; ++q) s[q] = 0.f;
    for (int m = 0; m < L; ++m) {
      const int64_t o = ((int64_t)m * N + n) * C + c;
      float gx[V], p[V];
      ldv<V>(g_Xd_out + o, gx);
      ldv<V>(P + o, p);
#pragma unroll
      for (int q = 0; q < V; ++q) s[q] = fmaf(gx[q], p[q], s[q]);
    }
    ldv<V>(g_h_out + idx, g);
#pragma unroll
    for (int q = 0; q < V; ++q) amx = fmaxf(amx, fmaxf(fabsf(g[q]), fabsf(s[q])));
    stv<V>(g_m + n * 2 * C + c, g);
    stv<V>(g_m + n * 2 * C + C + c, s);
  }
  block_amax_commit(gm_amax, amx);
}

template <int V>
__global__ void eqff_ctx_bwd_kernel(const float* __restrict__ g_h_out, const float* __restrict__ g_Xd_out,
                                    const float* __restrict__ g_ctx, const float* __restrict__ P,
                                    const float* __restrict__ mm, const float* __restrict__ ctx, int N, int C, int L,
                                    float* __restrict__ g_P, float* __restrict__ g_h, float* __restrict__ gp_amax) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  float amx = 0.f;
  if (idx < (int64_t)N * C) {
  const int64_t n = idx / C;
  const int c = (int)(idx % C);
  float gh[V], gc[V], m2[V], gn[V], nn[V];
  ldv<V>(g_h_out + idx, gh);
  ldv<V>(g_ctx + n * 2 * C + c, gc);
#pragma unroll
  for (int q = 0; q < V; ++q) gh[q] += gc[q];
  stv<V>(g_h + idx, gh);
  ldv<V>(mm + n * 2 * C + C + c, m2);
  ldv<V>(g_ctx + n * 2 * C + C + c, gn);
  ldv<V>(ctx + n * 2 * C + C + c, nn);
#pragma unroll
  for (int q = 0; q < V; ++q) gn[q] = gn[q] / nn[q];  // d sqrt(s+eps)/dP = P / n
  for (int m = 0; m < L; ++m) {
    const int64_t o = ((int64_t)m * N + n) * C + c;
    float gx[V], p[V];
    ldv<V>(g_Xd_out + o, gx);
    ldv<V>(P + o, p);
#pragma unroll
    for (int q = 0; q < V; ++q) { gx[q] = fmaf(gx[q], m2[q], gn[q] * p[q]); amx = fmaxf(amx, fabsf(gx[q])); }
    stv<V>(g_P + o, gx);
  }
  }
  block_amax_commit(gp_amax, amx);
}

}  // namespace goten

using namespace goten;

extern "C" {

#define EQFF_LAUNCH(KERNEL, N, C, ...)                                                                  \
  do {                                                                                                  \
    if ((C) % 4 == 0)                                                                                   \
      KERNEL<4><<<(unsigned)cdiv64((int64_t)(N) * (C) / 4, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
    else                                                                                                \
      KERNEL<1><<<(unsigned)cdiv64((int64_t)(N) * (C), 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
  } while (0)

int goten_eqff_ctx_fwd(const float* h, const float* P, int N, int C, int L, float eps, float* ctx, float* ctx_amax,
                       void* stream) {
  if ((int64_t)N * C == 0) return 0;
  EQFF_LAUNCH(eqff_ctx_fwd_kernel, N, C, h, P, N, C, L, eps, ctx, ctx_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}
int goten_eqff_update_fwd(const float* h, const float* Xd, const float* P, const float* m, int N, int C, int L,
                          float* h_out, float* Xd_out, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  EQFF_LAUNCH(eqff_update_fwd_kernel, N, C, h, Xd, P, m, N, C, L, h_out, Xd_out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}
int goten_eqff_update_bwd(const float* g_h_out, const float* g_Xd_out, const float* P, int N, int C, int L,
                          float* g_m, float* gm_amax, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  EQFF_LAUNCH(eqff_update_bwd_kernel, N, C, g_h_out, g_Xd_out, P, N, C, L, g_m, gm_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}
int goten_eqff_ctx_bwd(const float* g_h_out, const float* g_Xd_out, const float* g_ctx, const float* P,
                       const float* m, const float* ctx, int N, int C, int L, float* g_P, float* g_h, float* gp_amax,
                       void* stream) {
  if ((int64_t)N * C == 0) return 0;
  EQFF_LAUNCH(eqff_ctx_bwd_kernel, N, C, g_h_out, g_Xd_out, g_ctx, P, m, ctx, N, C, L, g_P, g_h, gp_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
