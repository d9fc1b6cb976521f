// Hierarchical tensor refinement of the edge scalars t_ij
// (reference representation/gotennet.py:351-364 vector_rejection, :561-611
// edge_update, :445 residual), forward and backward.
//
// EQ/EK are the [L][N][C] projections W_vq X / W_vk^l X^l (goten_gemm).  One CTA per
// node, one thread per channel: the node's own rows (EQ_i for the target pass, EK_j
// for the source pass) sit in registers, neighbour rows stream through L2 as
// coalesced 128 B warp loads, and the only HBM traffic is the [E][C] edge arrays
// (zt, t, t_out): the kernel is HBM-bound on  3*E*C*4 bytes.
// Forward/backward-target walk the target CSR; backward-source walks the
// transposed view.  No atomics.
#include "common.cuh"

namespace goten {

constexpr int HTR_SEP = 1, HTR_REJ = 2;

// accumulate  sum_m (q_m - a y_m)(k_m - b y_m)  over one group [lo,hi)
template <int LO, int HI>
__device__ __forceinline__ float group_weight(const float* q, const float* k, const float* y, bool rej) {
  float w = 0.f;
  if (rej) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int m = LO; m < HI; ++m) { a = fmaf(q[m], y[m], a); b = fmaf(k[m], y[m], b); }
#pragma unroll
    for (int m = LO; m < HI; ++m) w = fmaf(q[m] - a * y[m], k[m] - b * y[m], w);
  } else {
#pragma unroll
    for (int m = LO; m < HI; ++m) w = fmaf(q[m], k[m], w);
  }
  return w;
}

template <int LMAX>
__device__ __forceinline__ float htr_weight(const float* q, const float* k, const float* y, int flags) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const bool rej = flags & HTR_REJ;
  if (!(flags & HTR_SEP)) return group_weight<0, L>(q, k, y, rej);
  float w = group_weight<0, 3>(q, k, y, rej);
  if (LMAX >= 2) w += group_weight<3, 8>(q, k, y, rej);
  if (LMAX >= 3) w += group_weight<8, 15>(q, k, y, rej);
  return w;
}

// gradient of one group w.r.t. `q` given the other operand `k`:  out_m += dw * P (k - b y)_m with
// P = I - y y^T (rejection on) or out_m += dw * k_m (off).  Symmetric in q <-> k.
template <int LO, int HI>
__device__ __forceinline__ void group_grad(const float* k, const float* y, bool rej, float dw, float* out) {
  if (rej) {
    float b = 0.f;
#pragma unroll
    for (int m = LO; m < HI; ++m) b = fmaf(k[m], y[m], b);
    float kt[HI - LO];
    float s = 0.f;
#pragma unroll
    for (int m = LO; m < HI; ++m) { kt[m - LO] = dw * (k[m] - b * y[m]); s = fmaf(kt[m - LO], y[m], s); }
#pragma unroll
    for (int m = LO; m < HI; ++m) out[m] += kt[m - LO] - s * y[m];
  } else {
#pragma unroll
    for (int m = LO; m < HI; ++m) out[m] = fmaf(dw, k[m], out[m]);
  }
}

template <int LMAX>
__device__ __forceinline__ void htr_grad(const float* k, const float* y, int flags, float dw, float* out) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const bool rej = flags & HTR_REJ;
  if (!(flags & HTR_SEP)) { group_grad<0, L>(k, y, rej, dw, out); return; }
  group_grad<0, 3>(k, y, rej, dw, out);
  if (LMAX >= 2) group_grad<3, 8>(k, y, rej, dw, out);
  if (LMAX >= 3) group_grad<8, 15>(k, y, rej, dw, out);
}

// d w / d y_m for one group (rejection on): w = sum (q - a y)(k - b y), a = q.y, b = k.y
template <int LO, int HI>
__device__ __forceinline__ void group_grad_y(const float* q, const float* k, const float* y, float dw, float* gy) {
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int m = LO; m < HI; ++m) { a = fmaf(q[m], y[m], a); b = fmaf(k[m], y[m], b); }
  float sq = 0.f, sk = 0.f;  // sum y (q - a y), sum y (k - b y)
#pragma unroll
  for (int m = LO; m < HI; ++m) { sq = fmaf(y[m], q[m] - a * y[m], sq); sk = fmaf(y[m], k[m] - b * y[m], sk); }
#pragma unroll
  for (int m = LO; m < HI; ++m)
    gy[m] = dw * (-a * (k[m] - b * y[m]) - b * (q[m] - a * y[m]) - sk * q[m] - sq * k[m]);
}

// ---- V-wide channel vectors (V = 4: 128-bit loads/stores; V = 1 fallback)
template <int V>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float* out) {
  if (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else {
#pragma unroll
    for (int q = 0; q < V; ++q) out[q] = p[q];
  }
}
template <int V>
__device__ __forceinline__ void stv(float* __restrict__ p, const float* in) {
  if (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
  } else {
#pragma unroll
    for (int q = 0; q < V; ++q) p[q] = in[q];
  }
}
// gather column q of an [L][V] register tile into a contiguous [L] array (compile-time indices)
template <int L, int V>
__device__ __forceinline__ void col_of(const float (*a)[V], int q, float* out) {
#pragma unroll
  for (int m = 0; m < L; ++m) out[m] = a[m][q];
}

template <int LMAX, int V>
__global__ void htr_fwd_kernel(const float* __restrict__ EQ, const float* __restrict__ EK, const float* __restrict__ Y,
                               const float* __restrict__ Ze, int ldz, int zt_col0, const float* __restrict__ t,
                               const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C,
                               int flags, float* __restrict__ t_out) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const int i = blockIdx.x, c = threadIdx.x * V;
  if (c >= C) return;
  float q[L][V];
#pragma unroll
  for (int m = 0; m < L; ++m) ldv<V>(EQ + ((size_t)m * N + i) * C + c, q[m]);
  for (int e = tgt_ptr[i]; e < tgt_ptr[i + 1]; ++e) {
    const int j = src[e];
    float k[L][V], y[L], zt[V], tv[V];
#pragma unroll
    for (int m = 0; m < L; ++m) { ldv<V>(EK + ((size_t)m * N + j) * C + c, k[m]); y[m] = Y[(size_t)e * L + m]; }
    ldv<V>(Ze + (size_t)e * ldz + zt_col0 + c, zt);
    ldv<V>(t + (size_t)e * C + c, tv);
#pragma unroll
    for (int qq = 0; qq < V; ++qq) {
      float qc[L], kc[L];
      col_of<L, V>(q, qq, qc);
      col_of<L, V>(k, qq, kc);
      tv[qq] = fmaf(siluf_(zt[qq]), htr_weight<LMAX>(qc, kc, y, flags), tv[qq]);
    }
    stv<V>(t_out + (size_t)e * C + c, tv);
  }
}

template <int LMAX, int V>
__global__ void htr_bwd_tgt_kernel(const float* __restrict__ g_t_out, const float* __restrict__ EQ,
                                   const float* __restrict__ EK, const float* __restrict__ Y,
                                   const float* __restrict__ Ze, int ldz, int zt_col0,
                                   const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C,
                                   int flags, float* __restrict__ g_EQ, float* __restrict__ gZe, int ldgz,
                                   float* __restrict__ g_Y, float* __restrict__ gze_amax) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  __shared__ float red[33];
  const int i = blockIdx.x, c = threadIdx.x * V;
  const bool act = c < C;
  float amx = 0.f;
  float q[L][V], gq[L][V];
#pragma unroll
  for (int m = 0; m < L; ++m) {
#pragma unroll
    for (int qq = 0; qq < V; ++qq) { q[m][qq] = 0.f; gq[m][qq] = 0.f; }
    if (act) ldv<V>(EQ + ((size_t)m * N + i) * C + c, q[m]);
  }
  for (int e = tgt_ptr[i]; e < tgt_ptr[i + 1]; ++e) {
    const int j = src[e];
    float y[L], gy[L];
#pragma unroll
    for (int m = 0; m < L; ++m) { y[m] = Y[(size_t)e * L + m]; gy[m] = 0.f; }
    if (act) {
      float k[L][V], zt[V], dt[V], gz[V];
#pragma unroll
      for (int m = 0; m < L; ++m) ldv<V>(EK + ((size_t)m * N + j) * C + c, k[m]);
      ldv<V>(Ze + (size_t)e * ldz + zt_col0 + c, zt);
      ldv<V>(g_t_out + (size_t)e * C + c, dt);
#pragma unroll
      for (int qq = 0; qq < V; ++qq) {
        float qc[L], kc[L], gc[L];
        col_of<L, V>(q, qq, qc);
        col_of<L, V>(k, qq, kc);
        col_of<L, V>(gq, qq, gc);
        const float w = htr_weight<LMAX>(qc, kc, y, flags);
        gz[qq] = dt[qq] * w * dsiluf_(zt[qq]);
        amx = fmaxf(amx, fabsf(gz[qq]));
        const float dw = dt[qq] * siluf_(zt[qq]);
        htr_grad<LMAX>(kc, y, flags, dw, gc);
#pragma unroll
        for (int m = 0; m < L; ++m) gq[m][qq] = gc[m];
        if (g_Y != nullptr && (flags & HTR_REJ)) {
          float g1[L];
#pragma unroll
          for (int m = 0; m < L; ++m) g1[m] = 0.f;
          if (!(flags & HTR_SEP)) group_grad_y<0, L>(qc, kc, y, dw, g1);
          else {
            group_grad_y<0, 3>(qc, kc, y, dw, g1);
            if (LMAX >= 2) group_grad_y<3, 8>(qc, kc, y, dw, g1);
            if (LMAX >= 3) group_grad_y<8, 15>(qc, kc, y, dw, g1);
          }
#pragma unroll
          for (int m = 0; m < L; ++m) gy[m] += g1[m];
        }
      }
      stv<V>(gZe + (size_t)e * ldgz + zt_col0 + c, gz);
    }
    if (g_Y != nullptr) {  // block-uniform: geometry gradient for forces
#pragma unroll
      for (int m = 0; m < L; ++m) {
        const float s = block_sum(gy[m], red);
        if (threadIdx.x == 0) g_Y[(size_t)e * L + m] += s;
      }
    }
  }
  if (act) {
#pragma unroll
    for (int m = 0; m < L; ++m) stv<V>(g_EQ + ((size_t)m * N + i) * C + c, gq[m]);
  }
  amax_commit(gze_amax, amx);
}

template <int LMAX, int V>
__global__ void htr_bwd_src_kernel(const float* __restrict__ g_t_out, const float* __restrict__ EQ,
                                   const float* __restrict__ EK, const float* __restrict__ Y,
                                   const float* __restrict__ Ze, int ldz, int zt_col0,
                                   const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_perm,
                                   const int32_t* __restrict__ tgt, int N, int C, int flags,
                                   float* __restrict__ g_EK) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const int j = blockIdx.x, c = threadIdx.x * V;
  if (c >= C) return;
  float gk[L][V];
#pragma unroll
  for (int m = 0; m < L; ++m)
#pragma unroll
    for (int qq = 0; qq < V; ++qq) gk[m][qq] = 0.f;
  for (int p = src_ptr[j]; p < src_ptr[j + 1]; ++p) {
    const int e = src_perm[p];
    const int i = tgt[e];
    float q[L][V], y[L], zt[V], dt[V];
#pragma unroll
    for (int m = 0; m < L; ++m) { ldv<V>(EQ + ((size_t)m * N + i) * C + c, q[m]); y[m] = Y[(size_t)e * L + m]; }
    ldv<V>(Ze + (size_t)e * ldz + zt_col0 + c, zt);
    ldv<V>(g_t_out + (size_t)e * C + c, dt);
#pragma unroll
    for (int qq = 0; qq < V; ++qq) {
      float qc[L], gc[L];
      col_of<L, V>(q, qq, qc);
      col_of<L, V>(gk, qq, gc);
      htr_grad<LMAX>(qc, y, flags, dt[qq] * siluf_(zt[qq]), gc);  // d w / d k = P (q - a y) dw : q <-> k symmetric
#pragma unroll
      for (int m = 0; m < L; ++m) gk[m][qq] = gc[m];
    }
  }
#pragma unroll
  for (int m = 0; m < L; ++m) stv<V>(g_EK + ((size_t)m * N + j) * C + c, gk[m]);
}

static inline int block_for(int C, int V) { return (((C + V - 1) / V + 31) / 32) * 32; }

}  // namespace goten

using namespace goten;

#define HTR_DISPATCH(KERNEL, V4OK, ...)                                                        \
  do {                                                                                         \
    GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);                 \
    GOTEN_REQUIRE(C >= 1 && C <= 4096, "n_atom_basis=%d unsupported (<=4096)", C);             \
    if (N == 0) return 0;                                                                      \
    cudaStream_t st = as_stream(stream);                                                       \
    if (V4OK) {                                                                                \
      const int T = block_for(C, 4);                                                           \
      if (lmax == 1) KERNEL<1, 4><<<N, T, 0, st>>>(__VA_ARGS__);                               \
      else if (lmax == 2) KERNEL<2, 4><<<N, T, 0, st>>>(__VA_ARGS__);                          \
      else KERNEL<3, 4><<<N, T, 0, st>>>(__VA_ARGS__);                                         \
    } else {                                                                                   \
      const int T = block_for(C, 1);                                                           \
      GOTEN_REQUIRE(T <= 1024, "n_atom_basis=%d needs 16-byte aligned rows", C);               \
      if (lmax == 1) KERNEL<1, 1><<<N, T, 0, st>>>(__VA_ARGS__);                               \
      else if (lmax == 2) KERNEL<2, 1><<<N, T, 0, st>>>(__VA_ARGS__);                          \
      else KERNEL<3, 1><<<N, T, 0, st>>>(__VA_ARGS__);                                         \
    }                                                                                          \
    GOTEN_CHECK_LAUNCH();                                                                      \
    return 0;                                                                                  \
  } while (0)

extern "C" {

int goten_htr_fwd(const float* EQ, const float* EK, const float* Y, const float* Ze, int ldz, int zt_col0,
                  const float* t, const int32_t* tgt_ptr, const int32_t* src, int N, int C, int lmax, int flags,
                  float* t_out, void* stream) {
  HTR_DISPATCH(htr_fwd_kernel, (C % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0), EQ, EK, Y, Ze, ldz, zt_col0, t, tgt_ptr, src, N, C, flags, t_out);
}

int goten_htr_bwd_tgt(const float* g_t_out, const float* EQ, const float* EK, const float* Y, const float* Ze, int ldz,
                      int zt_col0, const int32_t* tgt_ptr, const int32_t* src, int N, int C, int lmax, int flags,
                      float* g_EQ, float* gZe, int ldgz, float* g_Y, float* gze_amax, void* stream) {
  HTR_DISPATCH(htr_bwd_tgt_kernel, (C % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0 && ldgz % 4 == 0), g_t_out, EQ, EK, Y, Ze, ldz, zt_col0, tgt_ptr, src, N, C, flags, g_EQ, gZe, ldgz,
               g_Y, gze_amax);
}

int goten_htr_bwd_src(const float* g_t_out, const float* EQ, const float* EK, const float* Y, const float* Ze, int ldz,
                      int zt_col0, const int32_t* src_ptr, const int32_t* src_perm, const int32_t* tgt, int N, int C,
                      int lmax, int flags, float* g_EK, void* stream) {
  HTR_DISPATCH(htr_bwd_src_kernel, (C % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0), g_t_out, EQ, EK, Y, Ze, ldz, zt_col0, src_ptr, src_perm, tgt, N, C, flags, g_EK);
}

}  // extern "C"
