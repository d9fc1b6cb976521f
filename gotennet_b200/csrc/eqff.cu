// EQFF element-wise stages (reference representation/gotennet.py:728-748).
// The three linear maps (W_vu, gamma_m.0, gamma_m.1) run through goten_gemm; these
// kernels are the norm/concat and gated-update stages and their backward passes.
// X-like tensors are degree-major: P[L][N][C].  One thread per (node, channel),
// channel fastest -> fully coalesced 128 B per warp; pure HBM-bound streaming.
#include "common.cuh"

namespace goten {

__global__ void eqff_ctx_fwd_kernel(const float* __restrict__ h, const float* __restrict__ P, int N, int C, int L,
                                    float eps, float* __restrict__ ctx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int64_t n = idx / C;
  const int c = (int)(idx % C);
  float s = 0.f;
  for (int m = 0; m < L; ++m) {
    const float p = P[((int64_t)m * N + n) * C + c];
    s = fmaf(p, p, s);
  }
  ctx[n * 2 * C + c] = h[idx];
  ctx[n * 2 * C + C + c] = sqrtf(s + eps);
}

__global__ void eqff_update_fwd_kernel(const float* __restrict__ h, const float* __restrict__ Xd,
                                       const float* __restrict__ P, const float* __restrict__ mm, int N, int C, int L,
                                       float* __restrict__ h_out, float* __restrict__ Xd_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int64_t n = idx / C;
  const int c = (int)(idx % C);
  h_out[idx] = h[idx] + mm[n * 2 * C + c];
  const float m2 = mm[n * 2 * C + C + c];
  for (int m = 0; m < L; ++m) {
    const int64_t o = ((int64_t)m * N + n) * C + c;
    Xd_out[o] = fmaf(m2, P[o], Xd[o]);
  }
}

__global__ void eqff_update_bwd_kernel(const float* __restrict__ g_h_out, const float* __restrict__ g_Xd_out,
                                       const float* __restrict__ P, int N, int C, int L, float* __restrict__ g_m) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int64_t n = idx / C;
  const int c = (int)(idx % C);
  float s = 0.f;
  for (int m = 0; m < L; ++m) {
    const int64_t o = ((int64_t)m * N + n) * C + c;
    s = fmaf(g_Xd_out[o], P[o], s);
  }
  g_m[n * 2 * C + c] = g_h_out[idx];
  g_m[n * 2 * C + C + c] = s;
}

__global__ void eqff_ctx_bwd_kernel(const float* __restrict__ g_h_out, const float* __restrict__ g_Xd_out,
                                    const float* __restrict__ g_ctx, const float* __restrict__ P,
                                    const float* __restrict__ mm, const float* __restrict__ ctx, int N, int C, int L,
                                    float* __restrict__ g_P, float* __restrict__ g_h) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int64_t n = idx / C;
  const int c = (int)(idx % C);
  g_h[idx] = g_h_out[idx] + g_ctx[n * 2 * C + c];
  const float m2 = mm[n * 2 * C + C + c];
  const float gn_over_n = g_ctx[n * 2 * C + C + c] / ctx[n * 2 * C + C + c];  // d sqrt(s+eps)/dP = P / n
  for (int m = 0; m < L; ++m) {
    const int64_t o = ((int64_t)m * N + n) * C + c;
    g_P[o] = fmaf(g_Xd_out[o], m2, gn_over_n * P[o]);
  }
}

}  // namespace goten

using namespace goten;

extern "C" {

#define EQFF_GRID(N, C) (unsigned)cdiv64((int64_t)(N) * (C), 256), 256, 0, as_stream(stream)

int goten_eqff_ctx_fwd(const float* h, const float* P, int N, int C, int L, float eps, float* ctx, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  eqff_ctx_fwd_kernel<<<EQFF_GRID(N, C)>>>(h, P, N, C, L, eps, ctx);
  GOTEN_CHECK_LAUNCH();
  return 0;
}
int goten_eqff_update_fwd(const float* h, const float* Xd, const float* P, const float* m, int N, int C, int L,
                          float* h_out, float* Xd_out, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  eqff_update_fwd_kernel<<<EQFF_GRID(N, C)>>>(h, Xd, P, m, N, C, L, h_out, Xd_out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}
int goten_eqff_update_bwd(const float* g_h_out, const float* g_Xd_out, const float* P, int N, int C, int L,
                          float* g_m, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  eqff_update_bwd_kernel<<<EQFF_GRID(N, C)>>>(g_h_out, g_Xd_out, P, N, C, L, g_m);
  GOTEN_CHECK_LAUNCH();
  return 0;
}
int goten_eqff_ctx_bwd(const float* g_h_out, const float* g_Xd_out, const float* g_ctx, const float* P,
                       const float* m, const float* ctx, int N, int C, int L, float* g_P, float* g_h, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  eqff_ctx_bwd_kernel<<<EQFF_GRID(N, C)>>>(g_h_out, g_Xd_out, g_ctx, P, m, ctx, N, C, L, g_P, g_h);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
