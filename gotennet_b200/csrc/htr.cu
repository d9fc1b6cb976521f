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
// bits 2-3 of `flags`: gamma_w of the "gated" / "gatedt" / "act" edge-update variants (gotennet.py:283-289):
// 0 = identity (default), 1 = sigmoid, 2 = tanh, 3 = SiLU applied to the scalar weight w before it multiplies gamma_t(t)
__device__ __forceinline__ int htr_gate(int flags) { return (flags >> 2) & 3; }
// bit 4: gamma_t ends without an activation ("mlp" variant, gotennet.py:244-248): zt is used as is instead of SiLU(zt)
constexpr int HTR_T_LINEAR = 16;
// bit 5: emit / differentiate the raw weight w only (the "linw" family, gotennet.py:270-282: gamma_w is then a small
// network on w[E][evec_dim], run by the host as LayerNorm / activation / GEMM launches).  Forward: t_out[e][c] = w;
// backward: g_t_out IS dL/dw, no gZe is produced; Ze / t are not read.  C is then evec_dim.
constexpr int HTR_W_ONLY = 32;
__device__ __forceinline__ float gate_f(float w, int gate) {
  if (gate == 1) return sigmoidf_(w);
  if (gate == 2) return tanhf(w);
  if (gate == 3) return siluf_(w);
  return w;
}
__device__ __forceinline__ float gate_df(float w, int gate) {
  if (gate == 1) { const float s = sigmoidf_(w); return s * (1.0f - s); }
  if (gate == 2) { const float t = tanhf(w); return 1.0f - t * t; }
  if (gate == 3) return dsiluf_(w);
  return 1.0f;
}

// Rejection algebra.  With P = I - y y^T applied per group (degree l if sep_htr, else all of L):
//   (P q).(P k) = q.k - (q.y)(k.y) n,      n = 2 - |y|^2     (n = 0 switches the rejection off)
//   d/dq [(P q).(P k)] = k - n (k.y) y     (symmetric in q <-> k)
// n depends on the edge only, so the per-channel work is three dot products (forward) plus two FMAs per
// component (gradient) instead of forming both rejected vectors.
template <int LO, int HI>
__device__ __forceinline__ float group_n(const float* y, bool rej) {
  if (!rej) return 0.f;
  float s = 0.f;
#pragma unroll
  for (int m = LO; m < HI; ++m) s = fmaf(y[m], y[m], s);
  return 2.0f - s;
}
// n[g] for the (up to three) groups of one edge
template <int LMAX>
__device__ __forceinline__ void htr_coef(const float* y, int flags, float* n) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const bool rej = flags & HTR_REJ;
  n[0] = n[1] = n[2] = 0.f;
  if (!(flags & HTR_SEP)) { n[0] = group_n<0, L>(y, rej); return; }
  n[0] = group_n<0, 3>(y, rej);
  if (LMAX >= 2) n[1] = group_n<3, 8>(y, rej);
  if (LMAX >= 3) n[2] = group_n<8, 15>(y, rej);
}

template <int LO, int HI>
__device__ __forceinline__ float group_weight(const float* q, const float* k, const float* y, float n) {
  float a = 0.f, b = 0.f, s = 0.f;
#pragma unroll
  for (int m = LO; m < HI; ++m) { a = fmaf(q[m], y[m], a); b = fmaf(k[m], y[m], b); s = fmaf(q[m], k[m], s); }
  return fmaf(-n * a, b, s);
}

template <int LMAX>
__device__ __forceinline__ float htr_weight(const float* q, const float* k, const float* y, const float* n, int flags) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  if (!(flags & HTR_SEP)) return group_weight<0, L>(q, k, y, n[0]);
  float w = group_weight<0, 3>(q, k, y, n[0]);
  if (LMAX >= 2) w += group_weight<3, 8>(q, k, y, n[1]);
  if (LMAX >= 3) w += group_weight<8, 15>(q, k, y, n[2]);
  return w;
}

// out_m += dw * (k_m - n (k.y) y_m): gradient of one group w.r.t. `q` given the other operand `k`
template <int LO, int HI>
__device__ __forceinline__ void group_grad(const float* k, const float* y, float n, float dw, float* out) {
  float b = 0.f;
#pragma unroll
  for (int m = LO; m < HI; ++m) b = fmaf(k[m], y[m], b);
  const float cy = -dw * n * b;
#pragma unroll
  for (int m = LO; m < HI; ++m) out[m] = fmaf(dw, k[m], fmaf(cy, y[m], out[m]));
}

template <int LMAX>
__device__ __forceinline__ void htr_grad(const float* k, const float* y, const float* n, int flags, float dw, float* out) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  if (!(flags & HTR_SEP)) { group_grad<0, L>(k, y, n[0], dw, out); return; }
  group_grad<0, 3>(k, y, n[0], dw, out);
  if (LMAX >= 2) group_grad<3, 8>(k, y, n[1], dw, out);
  if (LMAX >= 3) group_grad<8, 15>(k, y, n[2], dw, out);
}

// weight and gradient w.r.t. q in one pass (target half of the backward): returns the group's weight.
// GY: also the gradient w.r.t. the harmonics (forces), from the same dot products:
//   d/dy_m [q.k - n (q.y)(k.y)] = -n (b q_m + a k_m) + 2 a b y_m        (a = q.y, b = k.y, dn/dy_m = -2 y_m)
template <int LO, int HI, bool GY>
__device__ __forceinline__ float group_weight_grad(const float* q, const float* k, const float* y, float n, bool rej,
                                                   float dw, float* out, float* gy) {
  float a = 0.f, b = 0.f, s = 0.f;
#pragma unroll
  for (int m = LO; m < HI; ++m) { a = fmaf(q[m], y[m], a); b = fmaf(k[m], y[m], b); s = fmaf(q[m], k[m], s); }
  const float nb = n * b, cy = -dw * nb;
#pragma unroll
  for (int m = LO; m < HI; ++m) out[m] = fmaf(dw, k[m], fmaf(cy, y[m], out[m]));
  if (GY) {
    const float ca = -dw * n * a, cab = rej ? 2.0f * dw * a * b : 0.f;  // rejection off: n is the constant 0
#pragma unroll
    for (int m = LO; m < HI; ++m) gy[m] += fmaf(cy, q[m], fmaf(ca, k[m], cab * y[m]));
  }
  return fmaf(-nb, a, s);
}

template <int LMAX, bool GY>
__device__ __forceinline__ float htr_weight_grad(const float* q, const float* k, const float* y, const float* n, int flags,
                                                 float dw, float* out, float* gy) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const bool rej = flags & HTR_REJ;
  if (!(flags & HTR_SEP)) return group_weight_grad<0, L, GY>(q, k, y, n[0], rej, dw, out, gy);
  float w = group_weight_grad<0, 3, GY>(q, k, y, n[0], rej, dw, out, gy);
  if (LMAX >= 2) w += group_weight_grad<3, 8, GY>(q, k, y, n[1], rej, dw, out, gy);
  if (LMAX >= 3) w += group_weight_grad<8, 15, GY>(q, k, y, n[2], rej, dw, out, gy);
  return w;
}

// ---- V-wide channel vectors (V = 4: 128-bit loads/stores; V = 1 fallback)
template <int V>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float* out) {
  if (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else if (V == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    out[0] = t.x; out[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < V; ++q) out[q] = p[q];
  }
}
template <int V>
__device__ __forceinline__ void stv(float* __restrict__ p, const float* in) {
  if (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
  } else if (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(in[0], in[1]);
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

// Edge metadata of one node, staged in shared memory in chunks of EC edges: neighbour index, harmonics and the
// rejection coefficients n[g].  Removes the dependent index -> row-address load from the per-edge loop, so the
// neighbour-row gathers of consecutive edges are independent and overlap.  Used by the target half of the backward
// (measured on B200, C = 256: 1.49 -> 1.35 ms per step); the forward and the source half are faster without it.
constexpr int EC = 32;
constexpr int GY_MAX_BLOCK = 256;  // largest block of the force-gradient instantiation (static scratch)
template <int L>
struct EdgeMeta {
  int nbr[EC];
  int eid[EC];
  float y[EC * L];
  float n[EC * 3];
};

// Cooperative fill for edges [p0, p0 + cnt) of a CSR segment.  perm == nullptr: edge id = position (target view,
// neighbour = src[e]); else edge id = perm[p] (transposed view, neighbour = nbr_of[e]).  Ends with __syncthreads().
template <int LMAX>
__device__ __forceinline__ void fill_meta(EdgeMeta<(LMAX + 1) * (LMAX + 1) - 1>& sm, int p0, int cnt,
                                          const int32_t* __restrict__ perm, const int32_t* __restrict__ nbr_of,
                                          const float* __restrict__ Y, int flags) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  __syncthreads();  // the previous chunk has been consumed
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
    const int e = perm ? perm[p0 + t] : p0 + t;
    sm.eid[t] = e;
    sm.nbr[t] = nbr_of[e];
    float y[L];
#pragma unroll
    for (int m = 0; m < L; ++m) { y[m] = Y[(size_t)e * L + m]; sm.y[t * L + m] = y[m]; }
    float nn[3];
    htr_coef<LMAX>(y, flags, nn);
    sm.n[t * 3 + 0] = nn[0]; sm.n[t * 3 + 1] = nn[1]; sm.n[t * 3 + 2] = nn[2];
  }
  __syncthreads();
}

// GATED: a gamma_w gate is configured (flags bits 2-3); the plain instantiation folds it away
template <int LMAX, int V, bool GATED>
__global__ void htr_fwd_kernel(const float* __restrict__ EQ, const float* __restrict__ EK, int ldp, const float* __restrict__ Y,
                               const float* __restrict__ Ze, int ldz, int zt_col0, const float* __restrict__ t,
                               const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C,
                               int flags, float* __restrict__ t_out, float* __restrict__ t_amax) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const int i = blockIdx.x, c = threadIdx.x * V;
  if (c >= C) return;
  const int gate = GATED ? htr_gate(flags) : 0;
  float amx = 0.f;
  float q[L][V];
#pragma unroll
  for (int m = 0; m < L; ++m) ldv<V>(EQ + ((size_t)m * N + i) * ldp + c, q[m]);
  for (int e = tgt_ptr[i]; e < tgt_ptr[i + 1]; ++e) {
    const int j = src[e];
    float k[L][V], y[L], zt[V], tv[V];
#pragma unroll
    for (int m = 0; m < L; ++m) { ldv<V>(EK + ((size_t)m * N + j) * ldp + c, k[m]); y[m] = Y[(size_t)e * L + m]; }
    const bool w_only = GATED && (flags & HTR_W_ONLY);
    if (!w_only) {
      ldv<V>(Ze + (size_t)e * ldz + zt_col0 + c, zt);
      ldv<V>(t + (size_t)e * C + c, tv);
    }
    float nn[3];
    htr_coef<LMAX>(y, flags, nn);
#pragma unroll
    for (int qq = 0; qq < V; ++qq) {
      float qc[L], kc[L];
      col_of<L, V>(q, qq, qc);
      col_of<L, V>(k, qq, kc);
      const float w = htr_weight<LMAX>(qc, kc, y, nn, flags);
      if (w_only) {
        tv[qq] = w;
      } else {
        const float gt_ = (GATED && (flags & HTR_T_LINEAR)) ? zt[qq] : siluf_(zt[qq]);
        tv[qq] = fmaf(gt_, gate_f(w, gate), tv[qq]);
      }
      amx = fmaxf(amx, fabsf(tv[qq]));
    }
    stv<V>(t_out + (size_t)e * C + c, tv);
  }
  amax_commit(t_amax, amx);
}

// GY: also produce the geometry gradient g_Y (forces); kept out of the common instantiation (registers)
template <int LMAX, int V, bool GY, bool GATED>
__device__ __forceinline__ void htr_bwd_tgt_body(const float* __restrict__ g_t_out, const float* __restrict__ EQ,
                                                 const float* __restrict__ EK, int ldp, const float* __restrict__ Y,
                                                 const float* __restrict__ Ze, int ldz, int zt_col0,
                                                 const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src,
                                                 int N, int C, int flags, float* __restrict__ g_EQ,
                                                 float* __restrict__ gZe, int ldgz, float* __restrict__ g_Y,
                                                 float* __restrict__ gze_amax, float* __restrict__ geq_amax) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  __shared__ EdgeMeta<L> sm;
  // geometry gradient: per-edge channel sums of gy[L] through block_sums_one_barrier (blocks of <= GY_MAX_BLOCK threads)
  __shared__ float gy_scratch[GY ? 2 * L * (GY_MAX_BLOCK + 4) : 1];
  const int i = blockIdx.x, c = threadIdx.x * V;
  const bool act = c < C;
  const int gate = GATED ? htr_gate(flags) : 0;
  float amx = 0.f;
  float q[L][V], gq[L][V];
#pragma unroll
  for (int m = 0; m < L; ++m) {
#pragma unroll
    for (int qq = 0; qq < V; ++qq) { q[m][qq] = 0.f; gq[m][qq] = 0.f; }
    if (act) ldv<V>(EQ + ((size_t)m * N + i) * ldp + c, q[m]);
  }
  const int e_end = tgt_ptr[i + 1];
  for (int e0 = tgt_ptr[i]; e0 < e_end; e0 += EC) {
    const int cnt = min(EC, e_end - e0);
    fill_meta<LMAX>(sm, e0, cnt, nullptr, src, Y, flags);
#pragma unroll 2
    for (int u = 0; u < cnt; ++u) {
      const int e = e0 + u, j = sm.nbr[u];
      float y[L], gy[L];
#pragma unroll
      for (int m = 0; m < L; ++m) { y[m] = sm.y[u * L + m]; gy[m] = 0.f; }
      if (act) {
        float k[L][V], zt[V], dt[V], gz[V];
#pragma unroll
        for (int m = 0; m < L; ++m) ldv<V>(EK + ((size_t)m * N + j) * ldp + c, k[m]);
        const bool w_only = GATED && (flags & HTR_W_ONLY);
        if (!w_only) ldv<V>(Ze + (size_t)e * ldz + zt_col0 + c, zt);
        ldv<V>(g_t_out + (size_t)e * C + c, dt);
        const float nn[3] = {sm.n[u * 3], sm.n[u * 3 + 1], sm.n[u * 3 + 2]};
#pragma unroll
        for (int qq = 0; qq < V; ++qq) {
          float qc[L], kc[L], gc[L];
          col_of<L, V>(q, qq, qc);
          col_of<L, V>(k, qq, kc);
          col_of<L, V>(gq, qq, gc);
          if (w_only) {  // dt is dL/dw
            htr_weight_grad<LMAX, GY>(qc, kc, y, nn, flags, dt[qq], gc, gy);
#pragma unroll
            for (int m = 0; m < L; ++m) gq[m][qq] = gc[m];
            continue;
          }
          const bool t_lin = GATED && (flags & HTR_T_LINEAR);
          const float sg = sigmoid_fast_(zt[qq]);
          float dw = t_lin ? dt[qq] * zt[qq] : dt[qq] * zt[qq] * sg;              // dt * gamma_t: zt or silu(zt)
          if (gate) dw *= gate_df(htr_weight<LMAX>(qc, kc, y, nn, flags), gate);   // ... * gamma_w'(w)
          const float w = htr_weight_grad<LMAX, GY>(qc, kc, y, nn, flags, dw, gc, gy);
          const float dact = t_lin ? 1.0f : sg * (1.0f + zt[qq] * (1.0f - sg));    // d gamma_t / d zt
          gz[qq] = dt[qq] * gate_f(w, gate) * dact;
          amx = fmaxf(amx, fabsf(gz[qq]));
#pragma unroll
          for (int m = 0; m < L; ++m) gq[m][qq] = gc[m];
        }
        if (!w_only) stv<V>(gZe + (size_t)e * ldgz + zt_col0 + c, gz);
      }
      if (GY) {  // geometry gradient for forces (block-uniform): L channel sums, one barrier
        block_sums_one_barrier<L>(gy, gy_scratch + (size_t)(u & 1) * L * (blockDim.x + 4),
                                  [&](int m, float sgy) { g_Y[(size_t)e * L + m] += sgy; });
      }
    }
  }
  if (act) {
#pragma unroll
    for (int m = 0; m < L; ++m) stv<V>(g_EQ + ((size_t)m * N + i) * ldp + c, gq[m]);
  }
  amax_commit(gze_amax, amx);
  if (geq_amax != nullptr) {
    float a2 = 0.f;
#pragma unroll
    for (int m = 0; m < L; ++m)
#pragma unroll
      for (int qq = 0; qq < V; ++qq) a2 = fmaxf(a2, fabsf(gq[m][qq]));
    amax_commit(geq_amax, a2);
  }
}

template <int LMAX, int V, bool GATED>
__global__ void htr_bwd_tgt_kernel(const float* __restrict__ g_t_out, const float* __restrict__ EQ,
                                   const float* __restrict__ EK, int ldp, const float* __restrict__ Y,
                                   const float* __restrict__ Ze, int ldz, int zt_col0,
                                   const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C,
                                   int flags, float* __restrict__ g_EQ, float* __restrict__ gZe, int ldgz,
                                   float* __restrict__ g_Y, float* __restrict__ gze_amax,
                                   float* __restrict__ geq_amax) {
  htr_bwd_tgt_body<LMAX, V, false, GATED>(g_t_out, EQ, EK, ldp, Y, Ze, ldz, zt_col0, tgt_ptr, src, N, C, flags, g_EQ, gZe, ldgz, g_Y,
                                   gze_amax, geq_amax);
}
template <int LMAX, int V, bool GATED>
__global__ void htr_bwd_tgt_gy_kernel(const float* __restrict__ g_t_out, const float* __restrict__ EQ,
                                      const float* __restrict__ EK, int ldp, const float* __restrict__ Y,
                                      const float* __restrict__ Ze, int ldz, int zt_col0,
                                      const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N,
                                      int C, int flags, float* __restrict__ g_EQ, float* __restrict__ gZe, int ldgz,
                                      float* __restrict__ g_Y, float* __restrict__ gze_amax,
                                      float* __restrict__ geq_amax) {
  htr_bwd_tgt_body<LMAX, V, true, GATED>(g_t_out, EQ, EK, ldp, Y, Ze, ldz, zt_col0, tgt_ptr, src, N, C, flags, g_EQ, gZe, ldgz, g_Y,
                                  gze_amax, geq_amax);
}

template <int LMAX, int V, bool GATED>
__global__ void htr_bwd_src_kernel(const float* __restrict__ g_t_out, const float* __restrict__ EQ,
                                   const float* __restrict__ EK, int ldp, const float* __restrict__ Y,
                                   const float* __restrict__ Ze, int ldz, int zt_col0,
                                   const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_perm,
                                   const int32_t* __restrict__ tgt, int N, int C, int flags,
                                   float* __restrict__ g_EK, float* __restrict__ gek_amax) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const int j = blockIdx.x, c = threadIdx.x * V;
  if (c >= C) return;
  const int gate = GATED ? htr_gate(flags) : 0;
  float gk[L][V], kown[L][V];   // kown: this node's EK rows, needed only to re-evaluate w for a gated gamma_w
#pragma unroll
  for (int m = 0; m < L; ++m) {
#pragma unroll
    for (int qq = 0; qq < V; ++qq) { gk[m][qq] = 0.f; kown[m][qq] = 0.f; }
    if (gate) ldv<V>(EK + ((size_t)m * N + j) * ldp + c, kown[m]);
  }
  for (int p = src_ptr[j]; p < src_ptr[j + 1]; ++p) {
    const int e = src_perm[p];
    const int i = tgt[e];
    float q[L][V], y[L], zt[V], dt[V];
#pragma unroll
    for (int m = 0; m < L; ++m) { ldv<V>(EQ + ((size_t)m * N + i) * ldp + c, q[m]); y[m] = Y[(size_t)e * L + m]; }
    const bool w_only = GATED && (flags & HTR_W_ONLY);
    if (!w_only) ldv<V>(Ze + (size_t)e * ldz + zt_col0 + c, zt);
    ldv<V>(g_t_out + (size_t)e * C + c, dt);
    float nn[3];
    htr_coef<LMAX>(y, flags, nn);
#pragma unroll
    for (int qq = 0; qq < V; ++qq) {
      float qc[L], gc[L];
      col_of<L, V>(q, qq, qc);
      col_of<L, V>(gk, qq, gc);
      float dw = w_only ? dt[qq]
                        : ((GATED && (flags & HTR_T_LINEAR)) ? dt[qq] * zt[qq] : dt[qq] * silu_fast_(zt[qq]));
      if (gate) {
        float kc[L];
        col_of<L, V>(kown, qq, kc);
        dw *= gate_df(htr_weight<LMAX>(qc, kc, y, nn, flags), gate);
      }
      htr_grad<LMAX>(qc, y, nn, flags, dw, gc);  // d w / d k: q <-> k symmetric
#pragma unroll
      for (int m = 0; m < L; ++m) gk[m][qq] = gc[m];
    }
  }
#pragma unroll
  for (int m = 0; m < L; ++m) stv<V>(g_EK + ((size_t)m * N + j) * ldp + c, gk[m]);
  if (gek_amax != nullptr) {
    float a2 = 0.f;
#pragma unroll
    for (int m = 0; m < L; ++m)
#pragma unroll
      for (int qq = 0; qq < V; ++qq) a2 = fmaxf(a2, fabsf(gk[m][qq]));
    amax_commit(gek_amax, a2);
  }
}

static inline int block_for(int C, int V) { return (((C + V - 1) / V + 31) / 32) * 32; }
// channels per thread of the vector kernels: the per-kernel default (measured best on B200 at C = 256: the
// backward kernels carry 2 L V accumulators and run faster with twice the threads) or GOTEN_HTR_V=2|4 (A/B timing)
static inline int htr_vec(int def) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GOTEN_HTR_V"); v = e ? atoi(e) : 0; }
  return (v == 2 || v == 4) ? v : def;
}

}  // namespace goten

using namespace goten;

#define HTR_LAUNCH(KERNEL, LM, VV, T, ...)                                                      \
  do {                                                                                         \
    if ((flags >> 2) & 15) KERNEL<LM, VV, true><<<N, T, 0, st>>>(__VA_ARGS__);                  \
    else KERNEL<LM, VV, false><<<N, T, 0, st>>>(__VA_ARGS__);                                  \
  } while (0)

#define HTR_DISPATCH(KERNEL, V4OK, VDEF, ...)                                                  \
  do {                                                                                         \
    GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);                 \
    GOTEN_REQUIRE(C >= 1 && C <= 4096, "n_atom_basis=%d unsupported (<=4096)", C);             \
    if (N == 0) return 0;                                                                      \
    cudaStream_t st = as_stream(stream);                                                       \
    if ((V4OK) && htr_vec(VDEF) == 2) {                                                        \
      const int T = block_for(C, 2);                                                           \
      GOTEN_REQUIRE(T <= 1024, "n_atom_basis=%d too wide for the 2-channel HTR kernels", C);   \
      if (lmax == 1) HTR_LAUNCH(KERNEL, 1, 2, T, __VA_ARGS__);                                 \
      else if (lmax == 2) HTR_LAUNCH(KERNEL, 2, 2, T, __VA_ARGS__);                            \
      else HTR_LAUNCH(KERNEL, 3, 2, T, __VA_ARGS__);                                           \
    } else if (V4OK) {                                                                         \
      const int T = block_for(C, 4);                                                           \
      if (lmax == 1) HTR_LAUNCH(KERNEL, 1, 4, T, __VA_ARGS__);                                 \
      else if (lmax == 2) HTR_LAUNCH(KERNEL, 2, 4, T, __VA_ARGS__);                            \
      else HTR_LAUNCH(KERNEL, 3, 4, T, __VA_ARGS__);                                           \
    } else {                                                                                   \
      const int T = block_for(C, 1);                                                           \
      GOTEN_REQUIRE(T <= 1024, "n_atom_basis=%d needs 16-byte aligned rows", C);               \
      if (lmax == 1) HTR_LAUNCH(KERNEL, 1, 1, T, __VA_ARGS__);                                 \
      else if (lmax == 2) HTR_LAUNCH(KERNEL, 2, 1, T, __VA_ARGS__);                            \
      else HTR_LAUNCH(KERNEL, 3, 1, T, __VA_ARGS__);                                           \
    }                                                                                          \
    GOTEN_CHECK_LAUNCH();                                                                      \
    return 0;                                                                                  \
  } while (0)

extern "C" {

int goten_htr_fwd(const float* EQ, const float* EK, int ldp, const float* Y, const float* Ze, int ldz, int zt_col0,
                  const float* t, const int32_t* tgt_ptr, const int32_t* src, int N, int C, int lmax, int flags,
                  float* t_out, float* t_amax, void* stream) {
  HTR_DISPATCH(htr_fwd_kernel, (C % 4 == 0 && ldp % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0), 4, EQ, EK, ldp, Y, Ze, ldz, zt_col0, t, tgt_ptr, src, N, C, flags, t_out, t_amax);
}

int goten_htr_bwd_tgt(const float* g_t_out, const float* EQ, const float* EK, int ldp, const float* Y, const float* Ze, int ldz,
                      int zt_col0, const int32_t* tgt_ptr, const int32_t* src, int N, int C, int lmax, int flags,
                      float* g_EQ, float* gZe, int ldgz, float* g_Y, float* gze_amax, float* geq_amax, void* stream) {
  if (g_Y != nullptr) {
    GOTEN_REQUIRE(block_for(C, htr_vec(lmax >= 3 ? 2 : 4)) <= GY_MAX_BLOCK,
                  "n_atom_basis=%d too wide for the force-gradient HTR kernel (block <= %d threads)", C, GY_MAX_BLOCK);
    GOTEN_REQUIRE(C % 4 == 0 && ldp % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0 && ldgz % 4 == 0,
                  "the force-gradient HTR kernel needs 16-byte aligned rows (C=%d)", C);
  }
  if (g_Y != nullptr)
    HTR_DISPATCH(htr_bwd_tgt_gy_kernel, (C % 4 == 0 && ldp % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0 && ldgz % 4 == 0), (lmax >= 3 ? 2 : 4), g_t_out, EQ, EK, ldp, Y, Ze, ldz, zt_col0, tgt_ptr, src, N, C, flags, g_EQ, gZe, ldgz,
                 g_Y, gze_amax, geq_amax);
  // lmax = 3: 2 x 15 x V accumulators per thread; two channels per thread keep them in registers
  HTR_DISPATCH(htr_bwd_tgt_kernel, (C % 4 == 0 && ldp % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0 && ldgz % 4 == 0), (lmax >= 3 ? 2 : 4), g_t_out, EQ, EK, ldp, Y, Ze, ldz, zt_col0, tgt_ptr, src, N, C, flags, g_EQ, gZe, ldgz,
               g_Y, gze_amax, geq_amax);
}

int goten_htr_bwd_src(const float* g_t_out, const float* EQ, const float* EK, int ldp, const float* Y, const float* Ze, int ldz,
                      int zt_col0, const int32_t* src_ptr, const int32_t* src_perm, const int32_t* tgt, int N, int C,
                      int lmax, int flags, float* g_EK, float* gek_amax, void* stream) {
  HTR_DISPATCH(htr_bwd_src_kernel, (C % 4 == 0 && ldp % 4 == 0 && ldz % 4 == 0 && zt_col0 % 4 == 0), 2, g_t_out, EQ, EK, ldp, Y, Ze, ldz, zt_col0, src_ptr, src_perm, tgt, N, C, flags, g_EK, gek_amax);
}

}  // extern "C"
