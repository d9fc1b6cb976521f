// Initialisation block: NodeInit message/aggregate, Dense+LayerNorm+SiLU, EdgeInit
// (reference components/layers.py:1658-1675, :523-528, :1704-1714) and backward.
// F[E][ldf] = phi [W_ndp ; W_erp]^T + b comes from goten_gemm; columns [0,C) feed
// NodeInit, columns [col0, col0+C) feed EdgeInit.
// Forward reductions run over the target CSR, backward scatters over the
// transposed (source) view: no atomics, bit-reproducible.
#include "common.cuh"

namespace goten {

// ------------------------------------------------------------- NodeInit ------
__global__ void node_init_agg_fwd_kernel(const float* __restrict__ F, int ldf, const float* __restrict__ hnbr,
                                         const float* __restrict__ fc, const int32_t* __restrict__ tgt_ptr,
                                         const int32_t* __restrict__ src, int N, int C, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int i = (int)(idx / C);
  const int c = (int)(idx % C);
  float acc = 0.f;
  for (int e = tgt_ptr[i]; e < tgt_ptr[i + 1]; ++e) {
    const int j = src[e];
    if (j == i) continue;  // self loops removed, layers.py:1660-1664
    acc = fmaf(hnbr[(int64_t)j * C + c] * F[(int64_t)e * ldf + c], fc[e], acc);
  }
  out[idx] = acc;
}

// one CTA per target node: writes gF rows of its incoming edges and (optionally) g_fc[e]
__global__ void node_init_agg_bwd_tgt_kernel(const float* __restrict__ g_m, const float* __restrict__ F, int ldf,
                                             const float* __restrict__ hnbr, const float* __restrict__ fc,
                                             const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src,
                                             int C, float* __restrict__ gF, int ldgf, float* __restrict__ g_fc) {
  __shared__ float red[33];
  const int i = blockIdx.x;
  for (int e = tgt_ptr[i]; e < tgt_ptr[i + 1]; ++e) {
    const int j = src[e];
    const float f = fc[e];
    float part = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float g = 0.f;
      if (j != i) {
        const float gm = g_m[(int64_t)i * C + c], hn = hnbr[(int64_t)j * C + c];
        g = gm * hn * f;
        part = fmaf(gm * hn, F[(int64_t)e * ldf + c], part);
      }
      gF[(int64_t)e * ldgf + c] = g;
    }
    if (g_fc != nullptr) {  // uniform branch
      const float s = block_sum(part, red);
      if (threadIdx.x == 0) g_fc[e] += s;
    }
  }
}

__global__ void node_init_agg_bwd_src_kernel(const float* __restrict__ g_m, const float* __restrict__ F, int ldf,
                                             const float* __restrict__ fc, const int32_t* __restrict__ src_ptr,
                                             const int32_t* __restrict__ src_perm, const int32_t* __restrict__ tgt,
                                             int N, int C, float* __restrict__ g_hnbr) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int j = (int)(idx / C);
  const int c = (int)(idx % C);
  float acc = 0.f;
  for (int p = src_ptr[j]; p < src_ptr[j + 1]; ++p) {
    const int e = src_perm[p];
    const int i = tgt[e];
    if (i == j) continue;
    acc = fmaf(g_m[(int64_t)i * C + c] * F[(int64_t)e * ldf + c], fc[e], acc);
  }
  g_hnbr[idx] = acc;
}

// ------------------------------------------------------ LayerNorm + SiLU -----
// one warp per row
__global__ void ln_silu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int64_t rows, int C, float eps, int act,
                                   float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mu = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mu;
    v = fmaf(d, d, v);
  }
  const float rs = rsqrtf(warp_sum(v) / (float)C + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  for (int c = lane; c < C; c += 32) {
    const float zz = (xr[c] - mu) * rs * gamma[c] + beta[c];
    y[row * C + c] = act ? siluf_(zz) : zz;
  }
}

// grid-stride over rows (one warp per row); each warp accumulates d gamma / d beta in its own
// shared-memory slice, slices are combined in warp order -> deterministic.  One partial row per block.
__global__ void ln_silu_bwd_kernel(const float* __restrict__ g_y, const float* __restrict__ x,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ rstd, int64_t rows, int C,
                                   int act, float* __restrict__ g_x, float* __restrict__ g_gamma_part,
                                   float* __restrict__ g_beta_part) {
  extern __shared__ float sm[];  // [nw][2][C]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* sg = sm + (size_t)w * 2 * C;
  float* sb = sg + C;
  for (int c = lane; c < C; c += 32) { sg[c] = 0.f; sb[c] = 0.f; }
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < rows; row += (int64_t)gridDim.x * nw) {
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[row * C + c] - mu) * rs;
      const float dz = act ? g_y[row * C + c] * dsiluf_(xh * gamma[c] + beta[c]) : g_y[row * C + c];
      const float dxh = dz * gamma[c];
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
      sg[c] += dz * xh;
      sb[c] += dz;
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[row * C + c] - mu) * rs;
      const float dz = act ? g_y[row * C + c] * dsiluf_(xh * gamma[c] + beta[c]) : g_y[row * C + c];
      g_x[row * C + c] = rs * (dz * gamma[c] - s1 - xh * s2);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, bsum = 0.f;
    for (int ww = 0; ww < nw; ++ww) {
      a += sm[(size_t)ww * 2 * C + c];
      bsum += sm[(size_t)ww * 2 * C + C + c];
    }
    g_gamma_part[(int64_t)blockIdx.x * C + c] = a;
    g_beta_part[(int64_t)blockIdx.x * C + c] = bsum;
  }
}

// ------------------------------------------------------------- EdgeInit ------
// V = 4: one 128-bit access per array and thread (C, ldf, col0 multiples of 4 and 16 B aligned bases); V = 1 fallback
template <int V>
__global__ void edge_init_fwd_kernel(const float* __restrict__ h, const float* __restrict__ F, int ldf, int col0,
                                     const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t E, int C,
                                     float* __restrict__ t, float* __restrict__ t_amax) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  float amx = 0.f;
  if (idx < E * C) {
    const int64_t e = idx / C;
    const int c = (int)(idx - e * C);
    if (V == 4) {
      const float4 hi = *reinterpret_cast<const float4*>(h + (int64_t)tgt[e] * C + c);
      const float4 hj = *reinterpret_cast<const float4*>(h + (int64_t)src[e] * C + c);
      const float4 f = *reinterpret_cast<const float4*>(F + e * ldf + col0 + c);
      const float4 v = make_float4((hi.x + hj.x) * f.x, (hi.y + hj.y) * f.y, (hi.z + hj.z) * f.z, (hi.w + hj.w) * f.w);
      *reinterpret_cast<float4*>(t + idx) = v;
      amx = amax4(0.f, v.x, v.y, v.z, v.w);
    } else {
      const float v = (h[(int64_t)tgt[e] * C + c] + h[(int64_t)src[e] * C + c]) * F[e * ldf + col0 + c];
      t[idx] = v;
      amx = fabsf(v);
    }
  }
  block_amax_commit(t_amax, amx);
}

template <int V>
__global__ void edge_init_bwd_edge_kernel(const float* __restrict__ g_t, const float* __restrict__ h,
                                          const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t E,
                                          int C, float* __restrict__ gF, int ldgf, int col0) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (idx >= E * C) return;
  const int64_t e = idx / C;
  const int c = (int)(idx - e * C);
  if (V == 4) {
    const float4 hi = *reinterpret_cast<const float4*>(h + (int64_t)tgt[e] * C + c);
    const float4 hj = *reinterpret_cast<const float4*>(h + (int64_t)src[e] * C + c);
    const float4 g = *reinterpret_cast<const float4*>(g_t + idx);
    *reinterpret_cast<float4*>(gF + e * ldgf + col0 + c) =
        make_float4(g.x * (hi.x + hj.x), g.y * (hi.y + hj.y), g.z * (hi.z + hj.z), g.w * (hi.w + hj.w));
  } else {
    gF[e * ldgf + col0 + c] = g_t[idx] * (h[(int64_t)tgt[e] * C + c] + h[(int64_t)src[e] * C + c]);
  }
}

__global__ void edge_init_bwd_node_kernel(const float* __restrict__ g_t, const float* __restrict__ F, int ldf, int col0,
                                          const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src_ptr,
                                          const int32_t* __restrict__ src_perm, int N, int C,
                                          float* __restrict__ g_h) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * C) return;
  const int n = (int)(idx / C);
  const int c = (int)(idx % C);
  float acc = 0.f;
  for (int e = tgt_ptr[n]; e < tgt_ptr[n + 1]; ++e)  // n is the target (h_i term)
    acc = fmaf(g_t[(int64_t)e * C + c], F[(int64_t)e * ldf + col0 + c], acc);
  for (int p = src_ptr[n]; p < src_ptr[n + 1]; ++p) {  // n is the source (h_j term)
    const int e = src_perm[p];
    acc = fmaf(g_t[(int64_t)e * C + c], F[(int64_t)e * ldf + col0 + c], acc);
  }
  g_h[idx] = acc;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace goten

using namespace goten;

extern "C" {

int goten_node_init_agg_fwd(const float* F, int ldf, const float* hnbr, const float* fc, const int32_t* tgt_ptr,
                            const int32_t* src, int N, int C, float* m, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  node_init_agg_fwd_kernel<<<(unsigned)cdiv64((int64_t)N * C, 256), 256, 0, as_stream(stream)>>>(F, ldf, hnbr, fc,
                                                                                                tgt_ptr, src, N, C, m);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_node_init_agg_bwd_tgt(const float* g_m, const float* F, int ldf, const float* hnbr, const float* fc,
                                const int32_t* tgt_ptr, const int32_t* src, int N, int C, float* gF, int ldgf,
                                float* g_fc, void* stream) {
  if ((int64_t)N * C == 0) return 0;
  int T = ((C + 31) / 32) * 32;
  if (T > 256) T = 256;
  node_init_agg_bwd_tgt_kernel<<<N, T, 0, as_stream(stream)>>>(g_m, F, ldf, hnbr, fc, tgt_ptr, src, C, gF, ldgf,
                                                               g_fc);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_node_init_agg_bwd_src(const float* g_m, const float* F, int ldf, const float* fc, const int32_t* src_ptr,
                                const int32_t* src_perm, const int32_t* tgt, int N, int C, float* g_hnbr,
                                void* stream) {
  if ((int64_t)N * C == 0) return 0;
  node_init_agg_bwd_src_kernel<<<(unsigned)cdiv64((int64_t)N * C, 256), 256, 0, as_stream(stream)>>>(
      g_m, F, ldf, fc, src_ptr, src_perm, tgt, N, C, g_hnbr);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_ln_silu_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int C, float eps, float* y,
                      float* mean, float* rstd, void* stream) {
  if (rows == 0) return 0;
  ln_silu_fwd_kernel<<<(unsigned)cdiv64(rows, 8), 256, 0, as_stream(stream)>>>(x, gamma, beta, rows, C, eps, 1, y,
                                                                               mean, rstd);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int C, float eps, float* y,
                        float* mean, float* rstd, void* stream) {
  if (rows == 0) return 0;
  ln_silu_fwd_kernel<<<(unsigned)cdiv64(rows, 8), 256, 0, as_stream(stream)>>>(x, gamma, beta, rows, C, eps, 0, y,
                                                                               mean, rstd);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_ln_silu_bwd(const float* g_y, const float* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, int64_t rows, int C, float* g_x, float* g_gamma_part, float* g_beta_part,
                      int n_part, void* stream) {
  GOTEN_REQUIRE(n_part >= 1, "n_part must be >= 1");
  ln_silu_bwd_kernel<<<n_part, 256, 8 * 2 * C * sizeof(float), as_stream(stream)>>>(g_y, x, gamma, beta, mean, rstd, rows,
                                                                                C, 1, g_x, g_gamma_part, g_beta_part);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_layernorm_bwd(const float* g_y, const float* x, const float* gamma, const float* beta, const float* mean,
                        const float* rstd, int64_t rows, int C, float* g_x, float* g_gamma_part, float* g_beta_part,
                        int n_part, void* stream) {
  GOTEN_REQUIRE(n_part >= 1, "n_part must be >= 1");
  ln_silu_bwd_kernel<<<n_part, 256, 8 * 2 * C * sizeof(float), as_stream(stream)>>>(g_y, x, gamma, beta, mean, rstd, rows,
                                                                                C, 0, g_x, g_gamma_part, g_beta_part);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_edge_init_fwd(const float* h, const float* F, int ldf, int col0, const int32_t* src, const int32_t* tgt,
                        int64_t E, int C, float* t, float* t_amax, void* stream) {
  if (E * C == 0) return 0;
  if (C % 4 == 0 && ldf % 4 == 0 && col0 % 4 == 0 && al16(h) && al16(F) && al16(t))
    edge_init_fwd_kernel<4><<<(unsigned)cdiv64(E * C / 4, 256), 256, 0, as_stream(stream)>>>(h, F, ldf, col0, src, tgt,
                                                                                            E, C, t, t_amax);
  else
    edge_init_fwd_kernel<1><<<(unsigned)cdiv64(E * C, 256), 256, 0, as_stream(stream)>>>(h, F, ldf, col0, src, tgt, E,
                                                                                        C, t, t_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_edge_init_bwd(const float* g_t, const float* h, const float* F, int ldf, int col0, const int32_t* tgt_ptr,
                        const int32_t* src, const int32_t* tgt, const int32_t* src_ptr, const int32_t* src_perm, int N,
                        int64_t E, int C, float* gF, int ldgf, float* g_h, void* stream) {
  cudaStream_t st = as_stream(stream);
  if ((int64_t)N * C == 0) return 0;
  if (E * C > 0) {
    if (C % 4 == 0 && ldgf % 4 == 0 && col0 % 4 == 0 && al16(h) && al16(g_t) && al16(gF))
      edge_init_bwd_edge_kernel<4><<<(unsigned)cdiv64(E * C / 4, 256), 256, 0, st>>>(g_t, h, src, tgt, E, C, gF, ldgf,
                                                                                    col0);
    else
      edge_init_bwd_edge_kernel<1><<<(unsigned)cdiv64(E * C, 256), 256, 0, st>>>(g_t, h, src, tgt, E, C, gF, ldgf, col0);
    GOTEN_CHECK_LAUNCH();
  }
  edge_init_bwd_node_kernel<<<(unsigned)cdiv64((int64_t)N * C, 256), 256, 0, st>>>(g_t, F, ldf, col0, tgt_ptr,
                                                                                  src_ptr, src_perm, N, C, g_h);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
