// GATA message passing: geometry-aware tensor attention + segment softmax +
// scatter-sum + residual, fused per target node; and its backward.
// Reference: representation/gotennet.py:452-559 (message), :503 (PyG softmax),
// :613-640 (aggregate), :426-427 (residual).
//
// Mapping: one CTA per node, one thread per V = 4 consecutive channels (128-bit
// loads/stores; V = 1 when C is not a multiple of 4).  For an edge e = (j -> i) every
// per-edge row (filter Ze[e], source rows x_j, v_j, k_j, X_j) is read as coalesced
// 512 B warp transactions; source rows of the same molecule are shared by
// neighbouring CTAs through L2, so HBM traffic is the node arrays once plus the
// [E][(S+1)C] edge array once.
//   forward      : target CSR; logits -> in-CTA softmax (smem) -> weighted sum in
//                  registers -> h_out, Xd_out.  alpha[E][H] is saved.
//   backward/tgt : target CSR; d alpha from group-wise shuffle reductions, softmax
//                  backward, dq (register reduction), per-edge d(filter), d(pre-act W_re).
//   backward/src : transposed view; dx, dv, dk, dX_in reduced in registers.
// No atomics anywhere: results are bit-reproducible run to run.
#include "common.cuh"

namespace goten {

template <int LMAX, bool SD, bool ST>
struct GataCfg {
  static constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  static constexpr int ND = SD ? LMAX : 1;   // direction chunks
  static constexpr int NT = ST ? LMAX : 1;   // tensor chunks
  static constexpr int S = 1 + ND + NT;      // gotennet.py:197-203
};

__device__ __forceinline__ constexpr int lo_of(int l) { return (l + 1) * (l + 1) - 1; }  // l = 0.. -> degree l+1
__device__ __forceinline__ constexpr int hi_of(int l) { return (l + 2) * (l + 2) - 1; }

// ---- V-wide vector helpers (V = 1 or 4)
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

__host__ __device__ inline int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// gradient reaching the S output chunks of one edge for V channels:
//   dout[0] = g_h ; dout[1+l] = sum_{m in blk l} Y_m gX_m ; dout[1+ND+l] = sum_{m in blk l} Xj_m gX_m
template <int LMAX, bool SD, bool ST, int V>
__device__ __forceinline__ void chunk_grads(const float* gh, const float (*gX)[V], const float* y,
                                            const float (*Xj)[V], float (*dout)[V]) {
  using Cf = GataCfg<LMAX, SD, ST>;
#pragma unroll
  for (int q = 0; q < V; ++q) dout[0][q] = gh[q];
#pragma unroll
  for (int k = 1; k < Cf::S; ++k)
#pragma unroll
    for (int q = 0; q < V; ++q) dout[k][q] = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
#pragma unroll
    for (int m = lo_of(l); m < hi_of(l); ++m) {
#pragma unroll
      for (int q = 0; q < V; ++q) {
        dout[1 + (SD ? l : 0)][q] = fmaf(y[m], gX[m][q], dout[1 + (SD ? l : 0)][q]);
        dout[1 + Cf::ND + (ST ? l : 0)][q] = fmaf(Xj[m][q], gX[m][q], dout[1 + Cf::ND + (ST ? l : 0)][q]);
      }
    }
  }
}

struct GataSmem {
  float* part;   // [max_deg][nparts]           logit partials / d alpha partials
  float* alpha;  // [max_deg][H]
  float* aux;    // [max_deg][H]   (backward: d alpha, then d logits)
  float* fc;     // [max_deg]
  float* kap;    // [max_deg]
  int* src;      // [max_deg]
  float* Y;      // [max_deg][L]
};

__host__ __device__ inline size_t gata_smem_floats(int max_deg, int nparts, int H, int L, bool bwd) {
  size_t n = (size_t)max_deg * (nparts + H + 3 + L);
  if (bwd) n += (size_t)max_deg * H;
  return n;
}
// extra floats of the target backward when geometry gradients are requested: two scratch buffers of
// (1 + L) x (block + 4) for block_sums_one_barrier
__host__ __device__ inline size_t gata_geo_floats(int block, int L) { return (size_t)2 * (1 + L) * (block + 4); }

__device__ __forceinline__ GataSmem carve(float* base, int max_deg, int nparts, int H, int L, bool bwd) {
  GataSmem s;
  s.part = base; base += (size_t)max_deg * nparts;
  s.alpha = base; base += (size_t)max_deg * H;
  s.fc = base; base += max_deg;
  s.kap = base; base += max_deg;
  s.src = reinterpret_cast<int*>(base); base += max_deg;
  s.Y = base; base += (size_t)max_deg * L;
  s.aux = bwd ? base : nullptr;
  return s;
}

// ------------------------------------------------------------------ forward ---
template <int LMAX, bool SD, bool ST, int V>
__global__ void gata_fwd_kernel(const float* __restrict__ h, const float* __restrict__ Xd, const float* __restrict__ qk,
                                int ldqk, const float* __restrict__ x, const float* __restrict__ v,
                                const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                const float* __restrict__ fc, const float* __restrict__ kappa, const float* __restrict__ drop,
                                const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C, int H,
                                int max_deg, float* __restrict__ h_out, float* __restrict__ Xd_out, float* __restrict__ xd_amax,
                                float* __restrict__ alpha_out) {
  using Cf = GataCfg<LMAX, SD, ST>;
  constexpr int L = Cf::L, S = Cf::S;
  extern __shared__ float smem_f[];
  const int i = blockIdx.x, c = threadIdx.x * V;
  const bool act = c < C;
  const int D = C / H;
  const int Dt = D / V;                       // threads per head in q/k space
  const int W = Dt < 32 ? Dt : 32;            // shuffle segment width (threads)
  const int nparts = (C / V) / W, segs = Dt / W;
  const int SC = S * C, SD_ = S * D;          // SD_ = value columns per head
  GataSmem sm = carve(smem_f, max_deg, nparts, H, L, false);
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (deg > max_deg) __trap();  // host passed a too small max in-degree

  for (int t = threadIdx.x; t < deg; t += blockDim.x) {
    sm.src[t] = src[e0 + t]; sm.fc[t] = fc[e0 + t]; sm.kap[t] = kappa[e0 + t];
  }
  for (int t = threadIdx.x; t < deg * L; t += blockDim.x) sm.Y[t] = Y[(size_t)e0 * L + t];
  __syncthreads();

  // ---- attention logits: a[e][hd] = sum_d q_i k_j silu(W_re t)   (gotennet.py:502)
  float qi[V];
#pragma unroll
  for (int q = 0; q < V; ++q) qi[q] = 0.f;
  if (act) ldv<V>(qk + (size_t)i * ldqk + c, qi);
  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    float p = 0.f;
    if (act) {
      float kj[V], z[V];
      ldv<V>(qk + (size_t)j * ldqk + C + c, kj);
      ldv<V>(Ze + (size_t)(e0 + t) * ldz + c, z);
#pragma unroll
      for (int q = 0; q < V; ++q) p = fmaf(qi[q] * kj[q], siluf_(z[q]), p);
    }
    for (int o = W >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (act && (threadIdx.x % W) == 0) sm.part[t * nparts + threadIdx.x / W] = p;
  }
  __syncthreads();

  // ---- segment softmax over the incoming edges, one warp per head (gotennet.py:503; +1e-16)
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int hd = w; hd < H; hd += nw) {
      float mx = -INFINITY;
      for (int t = lane; t < deg; t += 32) {
        float a = 0.f;
        for (int s = 0; s < segs; ++s) a += sm.part[t * nparts + hd * segs + s];
        sm.alpha[t * H + hd] = a;
        mx = fmaxf(mx, a);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int t = lane; t < deg; t += 32) {
        const float ex = expf(sm.alpha[t * H + hd] - mx);
        sm.alpha[t * H + hd] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      const float den = sum + 1e-16f;
      for (int t = lane; t < deg; t += 32) {
        const float al = sm.alpha[t * H + hd] / den;
        alpha_out[(size_t)(e0 + t) * H + hd] = al;
        // attention dropout (gotennet.py:513): `drop` = mask / (1 - p) per (edge, head); the messages use the dropped weights
        sm.alpha[t * H + hd] = drop ? al * drop[(size_t)(e0 + t) * H + hd] : al;
      }
    }
  }
  __syncthreads();
  if (!act) return;

  // ---- messages + aggregation in registers (gotennet.py:516-558, :638-639)
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = (k * C + c) / SD_;
  float acc_h[V], accX[L][V];
#pragma unroll
  for (int q = 0; q < V; ++q) acc_h[q] = 0.f;
#pragma unroll
  for (int m = 0; m < L; ++m)
#pragma unroll
    for (int q = 0; q < V; ++q) accX[m][q] = 0.f;

  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    const float* ze = Ze + (size_t)(e0 + t) * ldz + C + c;
    const float* xj = x + (size_t)j * SC + c;
    const float* vj = v + (size_t)j * SC + c;
    const float f = sm.fc[t], kap = sm.kap[t];
    float o[S][V];
#pragma unroll
    for (int k = 0; k < S; ++k) {
      float tf[V], xv[V], vv[V];
      ldv<V>(ze + k * C, tf);
      ldv<V>(xj + k * C, xv);
      ldv<V>(vj + k * C, vv);
      const float al = sm.alpha[t * H + hd_of[k]] * kap;
#pragma unroll
      for (int q = 0; q < V; ++q) o[k][q] = tf[q] * xv[q] * f + al * vv[q];
    }
#pragma unroll
    for (int q = 0; q < V; ++q) acc_h[q] += o[0][q];
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
#pragma unroll
      for (int m = lo_of(l); m < hi_of(l); ++m) {
        float Xj[V];
        ldv<V>(Xd + ((size_t)m * N + j) * C + c, Xj);
        const float y = sm.Y[t * L + m];
#pragma unroll
        for (int q = 0; q < V; ++q)
          accX[m][q] += y * o[1 + (SD ? l : 0)][q] + Xj[q] * o[1 + Cf::ND + (ST ? l : 0)][q];
      }
    }
  }
  {
    float hv[V];
    ldv<V>(h + (size_t)i * C + c, hv);
#pragma unroll
    for (int q = 0; q < V; ++q) hv[q] += acc_h[q];
    stv<V>(h_out + (size_t)i * C + c, hv);
  }
  float xamx = 0.f;
#pragma unroll
  for (int m = 0; m < L; ++m) {
    const size_t o_ = ((size_t)m * N + i) * C + c;
    float xv[V];
    ldv<V>(Xd + o_, xv);
#pragma unroll
    for (int q = 0; q < V; ++q) { xv[q] += accX[m][q]; xamx = fmaxf(xamx, fabsf(xv[q])); }
    stv<V>(Xd_out + o_, xv);
  }
  amax_commit(xd_amax, xamx);
}

// --------------------------------------------------------- backward, target ---
// `part` here holds d alpha~ partials: [max_deg][S][n_grp], one per (edge, chunk, column group); a column group is
// g_cols = gcd(S*D, 32*V) consecutive value columns, which never straddles a head boundary.
template <int LMAX, bool SD, bool ST, int V>
__global__ void gata_bwd_tgt_kernel(const float* __restrict__ g_h, const float* __restrict__ g_Xd,
                                    const float* __restrict__ Xd, const float* __restrict__ qk, int ldqk,
                                    const float* __restrict__ x, const float* __restrict__ v,
                                    const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                    const float* __restrict__ fc, const float* __restrict__ kappa, const float* __restrict__ drop,
                                    const float* __restrict__ alpha, const int32_t* __restrict__ tgt_ptr,
                                    const int32_t* __restrict__ src, int N, int C, int H, int max_deg, int g_cols,
                                    float* __restrict__ g_qk, int ldgqk, float* __restrict__ gZe, int ldgz,
                                    float* __restrict__ da_out, float* __restrict__ g_fc, float* __restrict__ g_Y,
                                    float* __restrict__ gze_amax) {
  using Cf = GataCfg<LMAX, SD, ST>;
  float amx = 0.f;
  constexpr int L = Cf::L, S = Cf::S;
  extern __shared__ float smem_f[];
  const int i = blockIdx.x, c = threadIdx.x * V;
  const bool act = c < C;
  const int D = C / H;
  const int SC = S * C, SD_ = S * D;
  const int gt = g_cols / V;             // threads per column group (power of two <= 32)
  const int n_grp = C / g_cols;          // groups per chunk
  const int nparts = S * n_grp;
  GataSmem sm = carve(smem_f, max_deg, nparts, H, L, true);
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (deg > max_deg) __trap();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;

  for (int t = threadIdx.x; t < deg; t += blockDim.x) {
    sm.src[t] = src[e0 + t]; sm.fc[t] = fc[e0 + t]; sm.kap[t] = kappa[e0 + t];
  }
  for (int t = threadIdx.x; t < deg * L; t += blockDim.x) sm.Y[t] = Y[(size_t)e0 * L + t];
  for (int t = threadIdx.x; t < deg * H; t += blockDim.x) sm.alpha[t] = alpha[(size_t)e0 * H + t];
  float gh[V], gX[L][V];
#pragma unroll
  for (int q = 0; q < V; ++q) gh[q] = 0.f;
#pragma unroll
  for (int m = 0; m < L; ++m)
#pragma unroll
    for (int q = 0; q < V; ++q) gX[m][q] = 0.f;
  if (act) {
    ldv<V>(g_h + (size_t)i * C + c, gh);
#pragma unroll
    for (int m = 0; m < L; ++m) ldv<V>(g_Xd + ((size_t)m * N + i) * C + c, gX[m]);
  }
  __syncthreads();

  // ---- pass A: d alpha~ partials,  sum over the group's columns of dout[e][col] * v_j[col]
  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    float pk[S];
#pragma unroll
    for (int k = 0; k < S; ++k) pk[k] = 0.f;
    if (act) {
      float Xj[L][V], y[L], dout[S][V];
#pragma unroll
      for (int m = 0; m < L; ++m) { ldv<V>(Xd + ((size_t)m * N + j) * C + c, Xj[m]); y[m] = sm.Y[t * L + m]; }
      chunk_grads<LMAX, SD, ST, V>(gh, gX, y, Xj, dout);
      const float* vj = v + (size_t)j * SC + c;
#pragma unroll
      for (int k = 0; k < S; ++k) {
        float vv[V];
        ldv<V>(vj + k * C, vv);
#pragma unroll
        for (int q = 0; q < V; ++q) pk[k] = fmaf(dout[k][q], vv[q], pk[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < S; ++k) {
      float p = pk[k];
      for (int o = gt >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (act && (threadIdx.x % gt) == 0) sm.part[(t * S + k) * n_grp + threadIdx.x / gt] = p;
    }
  }
  __syncthreads();
  // d alpha[e][hd] = kappa * sum of the partials whose column group lies in head hd
  for (int idx = threadIdx.x; idx < deg * H; idx += blockDim.x) {
    const int t = idx / H, hd = idx - t * H;
    float s = 0.f;
    for (int k = 0; k < S; ++k)
      for (int g = 0; g < n_grp; ++g)
        if ((k * C + g * g_cols) / SD_ == hd) s += sm.part[(t * S + k) * n_grp + g];
    sm.aux[idx] = s * sm.kap[t] * (drop ? drop[(size_t)(e0 + t) * H + hd] : 1.0f);  // alpha~ = alpha * kappa * dropout
  }
  __syncthreads();
  // ---- softmax backward: da = alpha * (dalpha - sum_e alpha dalpha)
  for (int hd = w; hd < H; hd += nw) {
    float dot = 0.f;
    for (int t = lane; t < deg; t += 32) dot = fmaf(sm.alpha[t * H + hd], sm.aux[t * H + hd], dot);
    dot = warp_sum(dot);
    for (int t = lane; t < deg; t += 32) {
      const float da = sm.alpha[t * H + hd] * (sm.aux[t * H + hd] - dot);
      sm.aux[t * H + hd] = da;
      da_out[(size_t)(e0 + t) * H + hd] = da;
    }
  }
  __syncthreads();

  // ---- pass B: dq, d(pre-act W_re), d(filter), optional geometry gradients
  float qi[V], gq[V];
#pragma unroll
  for (int q = 0; q < V; ++q) { qi[q] = 0.f; gq[q] = 0.f; }
  if (act) ldv<V>(qk + (size_t)i * ldqk + c, qi);
  const int hq = act ? c / D : 0;
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = act ? (k * C + c) / SD_ : 0;
  const bool geom = (g_fc != nullptr) || (g_Y != nullptr);
  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    const size_t e = (size_t)(e0 + t);
    const float f = sm.fc[t];
    float gfc_part = 0.f, gy_part[L];
#pragma unroll
    for (int m = 0; m < L; ++m) gy_part[m] = 0.f;
    if (act) {
      float kj[V], zre[V], gz[V];
      ldv<V>(qk + (size_t)j * ldqk + C + c, kj);
      ldv<V>(Ze + e * ldz + c, zre);
      const float dal = sm.aux[t * H + hq];
#pragma unroll
      for (int q = 0; q < V; ++q) {
        gq[q] = fmaf(dal * kj[q], siluf_(zre[q]), gq[q]);
        gz[q] = dal * qi[q] * kj[q] * dsiluf_(zre[q]);
        amx = fmaxf(amx, fabsf(gz[q]));
      }
      stv<V>(gZe + e * ldgz + c, gz);
      float Xj[L][V], y[L], dout[S][V];
#pragma unroll
      for (int m = 0; m < L; ++m) { ldv<V>(Xd + ((size_t)m * N + j) * C + c, Xj[m]); y[m] = sm.Y[t * L + m]; }
      chunk_grads<LMAX, SD, ST, V>(gh, gX, y, Xj, dout);
      float o[S][V];
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const int col = k * C + c;
        float xv[V], gtf[V];
        ldv<V>(x + (size_t)j * SC + col, xv);
#pragma unroll
        for (int q = 0; q < V; ++q) { gtf[q] = dout[k][q] * xv[q] * f; amx = fmaxf(amx, fabsf(gtf[q])); }
        stv<V>(gZe + e * ldgz + C + col, gtf);
        if (geom) {
          float tf[V], vv[V];
          ldv<V>(Ze + e * ldz + C + col, tf);
          ldv<V>(v + (size_t)j * SC + col, vv);
          const float al = sm.alpha[t * H + hd_of[k]] * sm.kap[t] * (drop ? drop[e * H + hd_of[k]] : 1.0f);
#pragma unroll
          for (int q = 0; q < V; ++q) {
            gfc_part = fmaf(dout[k][q], tf[q] * xv[q], gfc_part);
            o[k][q] = tf[q] * xv[q] * f + al * vv[q];
          }
        }
      }
      if (geom) {
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) {
#pragma unroll
            for (int q = 0; q < V; ++q) gy_part[m] = fmaf(o[1 + (SD ? l : 0)][q], gX[m][q], gy_part[m]);
          }
        }
      }
    }
    if (geom) {  // block-uniform: the 1 + L channel sums of this edge with one barrier
      float vals[1 + L];
      vals[0] = gfc_part;
#pragma unroll
      for (int m = 0; m < L; ++m) vals[1 + m] = gy_part[m];
      float* scratch = smem_f + gata_smem_floats(max_deg, nparts, H, L, true) + (size_t)(t & 1) * (1 + L) * (blockDim.x + 4);
      block_sums_one_barrier<1 + L>(vals, scratch, [&](int v, float s) {
        if (v == 0) { if (g_fc != nullptr) g_fc[e] += s; }
        else if (g_Y != nullptr) g_Y[e * L + (v - 1)] += s;
      });
    }
  }
  if (act) stv<V>(g_qk + (size_t)i * ldgqk + c, gq);
  amax_commit(gze_amax, amx);
}

// --------------------------------------------------------- backward, source ---
constexpr int SRC_CHUNK = 32;

template <int LMAX, bool SD, bool ST, int V>
__global__ void gata_bwd_src_kernel(const float* __restrict__ g_h, const float* __restrict__ g_Xd,
                                    const float* __restrict__ Xd, const float* __restrict__ qk, int ldqk,
                                    const float* __restrict__ x, const float* __restrict__ v,
                                    const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                    const float* __restrict__ fc, const float* __restrict__ kappa, const float* __restrict__ drop,
                                    const float* __restrict__ alpha, const float* __restrict__ da,
                                    const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_perm,
                                    const int32_t* __restrict__ tgt, int N, int C, int H, float* __restrict__ g_qk,
                                    int ldgqk, float* __restrict__ g_x, float* __restrict__ g_v,
                                    float* __restrict__ g_Xd_in, float* __restrict__ gx_amax,
                                    float* __restrict__ gv_amax) {
  using Cf = GataCfg<LMAX, SD, ST>;
  constexpr int L = Cf::L, S = Cf::S, ND = Cf::ND;
  extern __shared__ float smem_f[];
  // [CHUNK] e, tgt, fc, kap ; [CHUNK][L] Y ; [CHUNK][H] alpha ; [CHUNK][H] da
  int* s_e = reinterpret_cast<int*>(smem_f);
  int* s_i = s_e + SRC_CHUNK;
  float* s_fc = smem_f + 2 * SRC_CHUNK;
  float* s_kap = s_fc + SRC_CHUNK;
  float* s_Y = s_kap + SRC_CHUNK;
  float* s_al = s_Y + SRC_CHUNK * L;
  float* s_da = s_al + SRC_CHUNK * H;

  const int j = blockIdx.x, c = threadIdx.x * V;
  const bool act = c < C;
  const int D = C / H, SC = S * C, SD_ = S * D;
  float xo[S][V], vo[S][V], Xo[L][V], gx[S][V], gv[S][V], gXin[L][V], gk[V];
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
#pragma unroll
    for (int q = 0; q < V; ++q) { xo[k][q] = 0.f; vo[k][q] = 0.f; gx[k][q] = 0.f; gv[k][q] = 0.f; }
    if (act) { ldv<V>(x + (size_t)j * SC + k * C + c, xo[k]); ldv<V>(v + (size_t)j * SC + k * C + c, vo[k]); }
    hd_of[k] = act ? (k * C + c) / SD_ : 0;
  }
#pragma unroll
  for (int m = 0; m < L; ++m) {
#pragma unroll
    for (int q = 0; q < V; ++q) { Xo[m][q] = 0.f; gXin[m][q] = 0.f; }
    if (act) ldv<V>(Xd + ((size_t)m * N + j) * C + c, Xo[m]);
  }
#pragma unroll
  for (int q = 0; q < V; ++q) gk[q] = 0.f;
  const int hq = act ? c / D : 0;

  const int p_begin = src_ptr[j], p_end = src_ptr[j + 1];
  for (int p0 = p_begin; p0 < p_end; p0 += SRC_CHUNK) {
    const int n = min(SRC_CHUNK, p_end - p0);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const int e = src_perm[p0 + t];
      s_e[t] = e; s_i[t] = tgt[e]; s_fc[t] = fc[e]; s_kap[t] = kappa[e];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n * L; q += blockDim.x) s_Y[q] = Y[(size_t)s_e[q / L] * L + (q % L)];
    for (int q = threadIdx.x; q < n * H; q += blockDim.x) {
      const size_t o_ = (size_t)s_e[q / H] * H + (q % H);
      s_al[q] = drop ? alpha[o_] * drop[o_] : alpha[o_];
      s_da[q] = da[o_];
    }
    __syncthreads();
    if (act) {
      for (int t = 0; t < n; ++t) {
        const size_t e = (size_t)s_e[t];
        const int i = s_i[t];
        const float f = s_fc[t], kap = s_kap[t];
        float gh[V], gX[L][V], y[L], dout[S][V];
        ldv<V>(g_h + (size_t)i * C + c, gh);
#pragma unroll
        for (int m = 0; m < L; ++m) { ldv<V>(g_Xd + ((size_t)m * N + i) * C + c, gX[m]); y[m] = s_Y[t * L + m]; }
        chunk_grads<LMAX, SD, ST, V>(gh, gX, y, Xo, dout);
        float ot[Cf::NT][V];
#pragma unroll
        for (int k = 0; k < S; ++k) {
          float tf[V];
          ldv<V>(Ze + e * ldz + C + k * C + c, tf);
          const float al = s_al[t * H + hd_of[k]] * kap;
#pragma unroll
          for (int q = 0; q < V; ++q) {
            gx[k][q] = fmaf(dout[k][q], tf[q] * f, gx[k][q]);
            gv[k][q] = fmaf(dout[k][q], al, gv[k][q]);
            if (k >= 1 + ND) ot[k - 1 - ND][q] = tf[q] * xo[k][q] * f + al * vo[k][q];
          }
        }
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) {
#pragma unroll
            for (int q = 0; q < V; ++q) gXin[m][q] = fmaf(ot[ST ? l : 0][q], gX[m][q], gXin[m][q]);
          }
        }
        float qi[V], zre[V];
        ldv<V>(qk + (size_t)i * ldqk + c, qi);
        ldv<V>(Ze + e * ldz + c, zre);
        const float dal = s_da[t * H + hq];
#pragma unroll
        for (int q = 0; q < V; ++q) gk[q] = fmaf(dal * qi[q], siluf_(zre[q]), gk[q]);
      }
    }
  }
  if (act) {
#pragma unroll
    float ax = 0.f, av = 0.f;
    for (int k = 0; k < S; ++k) {
      stv<V>(g_x + (size_t)j * SC + k * C + c, gx[k]);
      stv<V>(g_v + (size_t)j * SC + k * C + c, gv[k]);
#pragma unroll
      for (int q = 0; q < V; ++q) { ax = fmaxf(ax, fabsf(gx[k][q])); av = fmaxf(av, fabsf(gv[k][q])); }
    }
    amax_commit(gx_amax, ax);
    amax_commit(gv_amax, av);
    stv<V>(g_qk + (size_t)j * ldgqk + C + c, gk);
#pragma unroll
    for (int m = 0; m < L; ++m) {
      const size_t o_ = ((size_t)m * N + j) * C + c;
      float gxd[V];
      ldv<V>(g_Xd + o_, gxd);
#pragma unroll
      for (int q = 0; q < V; ++q) gxd[q] += gXin[m][q];
      stv<V>(g_Xd_in + o_, gxd);
    }
  }
}

// vector width: 4 when every row pointer stays 16 B aligned and heads are 4-aligned
static inline int gata_vec(int C, int H, int ldqk, int ldz) {
  const int D = C / H;
  return (C % 4 == 0 && D % 4 == 0 && ldqk % 4 == 0 && ldz % 4 == 0) ? 4 : 1;
}
static inline int gata_block(int C, int V) { return (((C + V - 1) / V + 31) / 32) * 32; }

static int gata_check(int C, int H, int lmax, int V) {
  GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);
  GOTEN_REQUIRE(C >= 1 && C <= 4096 && H >= 1 && C % H == 0, "n_atom_basis=%d / num_heads=%d unsupported", C, H);
  const int Dt = (C / H) / V;  // threads per head
  GOTEN_REQUIRE(Dt >= 1 && ((Dt <= 32 && (Dt & (Dt - 1)) == 0) || (Dt % 32 == 0)),
                "head width %d unsupported (power of two, or a multiple of %d)", C / H, 32 * V);
  return 0;
}

// TMA-staged production variants (gata_staged.cu); *handled = false -> shape outside their contract
int gata_fwd_staged(const float* h, const float* Xd, const float* qk, int ldqk, const float* x, const float* v,
                    const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                    const int32_t* tgt_ptr, const int32_t* src, int N, int C, int H, int lmax, int flags,
                    int max_deg_in, float* h_out, float* Xd_out, float* alpha, float* xd_amax, cudaStream_t st,
                    bool* handled);

int gata_bwd_tgt_staged(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk, const float* x,
                        const float* v, const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                        const float* alpha, const int32_t* tgt_ptr, const int32_t* src, int N, int C, int H, int lmax,
                        int flags, int max_deg_in, float* g_qk, int ldgqk, float* gZe, int ldgz, float* da,
                        float* gze_amax, float* g_fc, float* g_Y, cudaStream_t st, bool* handled);
int gata_bwd_src_staged(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk, const float* x,
                        const float* v, const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                        const float* alpha, const float* da, const int32_t* src_ptr, const int32_t* src_perm,
                        const int32_t* tgt, int N, int C, int H, int lmax, int flags, float* g_qk, int ldgqk, float* g_x,
                        float* g_v, float* g_Xd_in, float* gx_amax, float* gv_amax, cudaStream_t st, bool* handled);

// GOTEN_GATA=legacy forces the register-gather kernels of this file (A/B timing, tests)
static bool use_staged() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GOTEN_GATA");
    v = (e && strcmp(e, "legacy") == 0) ? 0 : 1;
  }
  return v == 1;
}

}  // namespace goten

using namespace goten;

#define GATA_LAUNCH(KERNEL, LM, SD, ST, GRID, SMEM, ...)                                                \
  do {                                                                                                  \
    if (V == 4) {                                                                                       \
      auto kfn = KERNEL<LM, SD, ST, 4>;                                                                 \
      if ((SMEM) > 48 * 1024)                                                                           \
        GOTEN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
      kfn<<<GRID, gata_block(C, 4), SMEM, st>>>(__VA_ARGS__);                                           \
    } else {                                                                                            \
      auto kfn = KERNEL<LM, SD, ST, 1>;                                                                 \
      if ((SMEM) > 48 * 1024)                                                                           \
        GOTEN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
      kfn<<<GRID, gata_block(C, 1), SMEM, st>>>(__VA_ARGS__);                                           \
    }                                                                                                   \
  } while (0)

// dispatch over (lmax, sep_dir, sep_tensor); for lmax == 1 the sep flags do not change the layout
#define GATA_DISPATCH(KERNEL, GRID, SMEM, ...)                                                          \
  do {                                                                                                  \
    const bool sd = (flags & 1) && lmax > 1, stn = (flags & 2) && lmax > 1;                             \
    if (lmax == 1) { GATA_LAUNCH(KERNEL, 1, false, false, GRID, SMEM, __VA_ARGS__); }                   \
    else if (lmax == 2 && !sd && !stn) { GATA_LAUNCH(KERNEL, 2, false, false, GRID, SMEM, __VA_ARGS__); } \
    else if (lmax == 2 && sd && !stn) { GATA_LAUNCH(KERNEL, 2, true, false, GRID, SMEM, __VA_ARGS__); } \
    else if (lmax == 2 && !sd && stn) { GATA_LAUNCH(KERNEL, 2, false, true, GRID, SMEM, __VA_ARGS__); } \
    else if (lmax == 2) { GATA_LAUNCH(KERNEL, 2, true, true, GRID, SMEM, __VA_ARGS__); }                \
    else if (!sd && !stn) { GATA_LAUNCH(KERNEL, 3, false, false, GRID, SMEM, __VA_ARGS__); }            \
    else if (sd && !stn) { GATA_LAUNCH(KERNEL, 3, true, false, GRID, SMEM, __VA_ARGS__); }              \
    else if (!sd && stn) { GATA_LAUNCH(KERNEL, 3, false, true, GRID, SMEM, __VA_ARGS__); }              \
    else { GATA_LAUNCH(KERNEL, 3, true, true, GRID, SMEM, __VA_ARGS__); }                               \
  } while (0)

static inline int multiplier_of(int lmax, int flags) {
  return 3 + ((flags & 1) ? lmax - 1 : 0) + ((flags & 2) ? lmax - 1 : 0);
}

extern "C" {

int goten_gata_fwd(const float* h, const float* Xd, const float* qk, int ldqk, const float* x, const float* v,
                   const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                   const int32_t* tgt_ptr, const int32_t* src, int N, int C, int H, int lmax, int flags,
                   int max_deg_in, float* h_out, float* Xd_out, float* alpha, float* xd_amax, void* stream) {
  const int V = gata_vec(C, H, ldqk, ldz);
  if (gata_check(C, H, lmax, V)) return 1;
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (use_staged()) {
    bool handled = false;
    if (gata_fwd_staged(h, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, tgt_ptr, src, N, C, H, lmax, flags, max_deg_in,
                        h_out, Xd_out, alpha, xd_amax, st, &handled))
      return 1;
    if (handled) return 0;
  }
  const int Dt = (C / H) / V, W = Dt < 32 ? Dt : 32, nparts = (C / V) / W, L = (lmax + 1) * (lmax + 1) - 1;
  if (max_deg_in < 1) max_deg_in = 1;
  const size_t smem = gata_smem_floats(max_deg_in, nparts, H, L, false) * sizeof(float);
  GOTEN_REQUIRE(smem <= 200 * 1024, "max in-degree %d needs %zu B of shared memory", max_deg_in, smem);
  GATA_DISPATCH(gata_fwd_kernel, N, smem, h, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, tgt_ptr, src, N, C, H,
                max_deg_in, h_out, Xd_out, xd_amax, alpha);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_gata_bwd_tgt(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk,
                       const float* x, const float* v, const float* Ze, int ldz, const float* Y, const float* fc,
                       const float* kappa, const float* drop, const float* alpha, const int32_t* tgt_ptr, const int32_t* src, int N,
                       int C, int H, int lmax, int flags, int max_deg_in, float* g_qk, int ldgqk, float* gZe,
                       int ldgz, float* da, float* g_fc, float* g_Y, float* gze_amax, void* stream) {
  const int V = (gata_vec(C, H, ldqk, ldz) == 4 && ldgqk % 4 == 0 && ldgz % 4 == 0) ? 4 : 1;
  if (gata_check(C, H, lmax, V)) return 1;
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (use_staged()) {
    bool handled = false;
    if (gata_bwd_tgt_staged(g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha, tgt_ptr, src, N, C, H, lmax,
                            flags, max_deg_in, g_qk, ldgqk, gZe, ldgz, da, gze_amax, g_fc, g_Y, st, &handled))
      return 1;
    if (handled) return 0;
  }
  const int L = (lmax + 1) * (lmax + 1) - 1, S = multiplier_of(lmax, lmax > 1 ? flags : 0);
  const int g_cols = gcd_i(S * (C / H), 32 * V);  // column group that never straddles a head; multiple of V
  GOTEN_REQUIRE(g_cols % V == 0 && C % g_cols == 0, "unsupported head / channel combination (C=%d H=%d S=%d)", C, H, S);
  const int nparts = S * (C / g_cols);
  if (max_deg_in < 1) max_deg_in = 1;
  const bool geo = g_fc != nullptr || g_Y != nullptr;
  const int block_t = gata_block(C, V);
  const size_t smem = (gata_smem_floats(max_deg_in, nparts, H, L, true) + (geo ? gata_geo_floats(block_t, L) : 0)) * sizeof(float);
  GOTEN_REQUIRE(smem <= 200 * 1024, "max in-degree %d needs %zu B of shared memory", max_deg_in, smem);
  GATA_DISPATCH(gata_bwd_tgt_kernel, N, smem, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha, tgt_ptr,
                src, N, C, H, max_deg_in, g_cols, g_qk, ldgqk, gZe, ldgz, da, g_fc, g_Y, gze_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_gata_bwd_src(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk,
                       const float* x, const float* v, const float* Ze, int ldz, const float* Y, const float* fc,
                       const float* kappa, const float* drop, const float* alpha, const float* da, const int32_t* src_ptr,
                       const int32_t* src_perm, const int32_t* tgt, int N, int C, int H, int lmax, int flags,
                       float* g_qk, int ldgqk, float* g_x, float* g_v, float* g_Xd_in, float* gx_amax, float* gv_amax,
                       void* stream) {
  const int V = (gata_vec(C, H, ldqk, ldz) == 4 && ldgqk % 4 == 0) ? 4 : 1;
  if (gata_check(C, H, lmax, V)) return 1;
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (use_staged()) {
    bool handled = false;
    if (gata_bwd_src_staged(g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha, da, src_ptr, src_perm, tgt, N, C,
                            H, lmax, flags, g_qk, ldgqk, g_x, g_v, g_Xd_in, gx_amax, gv_amax, st, &handled))
      return 1;
    if (handled) return 0;
  }
  const int L = (lmax + 1) * (lmax + 1) - 1;
  const size_t smem = (size_t)SRC_CHUNK * (4 + L + 2 * H) * sizeof(float);
  GATA_DISPATCH(gata_bwd_src_kernel, N, smem, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha, da,
                src_ptr, src_perm, tgt, N, C, H, g_qk, ldgqk, g_x, g_v, g_Xd_in, gx_amax, gv_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
