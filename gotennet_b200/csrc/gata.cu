// GATA message passing: geometry-aware tensor attention + segment softmax +
// scatter-sum + residual, fused per target node; and its backward.
// Reference: representation/gotennet.py:452-559 (message), :503 (PyG softmax),
// :613-640 (aggregate), :426-427 (residual).
//
// Mapping: one CTA per node, one thread per channel c (blockDim = C rounded to a
// warp).  For an edge e = (j -> i) every per-edge row (filter Ze[e], source rows
// x_j, v_j, k_j, X_j) is read as coalesced 128 B warp transactions; source rows of
// the same molecule are shared by neighbouring CTAs through L2, so HBM traffic is
// the node arrays once plus the [E][(S+1)C] edge array once.
//   forward      : target CSR; logits -> in-CTA softmax (smem) -> weighted sum in
//                  registers -> h_out, Xd_out.  alpha[E][H] is saved.
//   backward/tgt : target CSR; d alpha via a warp-per-head pass, softmax backward,
//                  dq (register reduction), per-edge d(filter), d(pre-act W_re).
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

// gradient reaching the S output chunks of one edge for channel c:
//   dout[0] = g_h ; dout[1+l] = sum_{m in blk l} Y_m gX_m ; dout[1+ND+l] = sum_{m in blk l} Xj_m gX_m
template <int LMAX, bool SD, bool ST>
__device__ __forceinline__ void chunk_grads(float gh, const float* gX, const float* y, const float* Xj, float* dout) {
  using Cf = GataCfg<LMAX, SD, ST>;
  dout[0] = gh;
#pragma unroll
  for (int k = 1; k < Cf::S; ++k) dout[k] = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    float sd = 0.f, st = 0.f;
#pragma unroll
    for (int m = lo_of(l); m < hi_of(l); ++m) { sd = fmaf(y[m], gX[m], sd); st = fmaf(Xj[m], gX[m], st); }
    dout[1 + (SD ? l : 0)] += sd;
    dout[1 + Cf::ND + (ST ? l : 0)] += st;
  }
}

struct GataSmem {
  float* part;   // [max_deg][nparts]
  float* alpha;  // [max_deg][H]
  float* aux;    // [max_deg][H]   (backward: d alpha / d logits)
  float* fc;     // [max_deg]
  float* kap;    // [max_deg]
  int* src;      // [max_deg]
  float* Y;      // [max_deg][L]
  float* gh;     // [C]            (backward)
  float* gX;     // [L][C]         (backward)
};

__host__ __device__ inline size_t gata_smem_floats(int max_deg, int nparts, int H, int L, int C, bool bwd) {
  size_t n = (size_t)max_deg * (nparts + H + 3 + L);
  if (bwd) n += (size_t)max_deg * H + (size_t)(1 + L) * C;
  return n;
}

__device__ __forceinline__ GataSmem carve(float* base, int max_deg, int nparts, int H, int L, int C, bool bwd) {
  GataSmem s;
  s.part = base; base += (size_t)max_deg * nparts;
  s.alpha = base; base += (size_t)max_deg * H;
  s.fc = base; base += max_deg;
  s.kap = base; base += max_deg;
  s.src = reinterpret_cast<int*>(base); base += max_deg;
  s.Y = base; base += (size_t)max_deg * L;
  s.aux = nullptr; s.gh = nullptr; s.gX = nullptr;
  if (bwd) {
    s.aux = base; base += (size_t)max_deg * H;
    s.gh = base; base += C;
    s.gX = base;
  }
  return s;
}

// ------------------------------------------------------------------ forward ---
template <int LMAX, bool SD, bool ST>
__global__ void gata_fwd_kernel(const float* __restrict__ h, const float* __restrict__ Xd, const float* __restrict__ qk,
                                int ldqk, const float* __restrict__ x, const float* __restrict__ v,
                                const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                const float* __restrict__ fc, const float* __restrict__ kappa,
                                const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C, int H,
                                int max_deg, float* __restrict__ h_out, float* __restrict__ Xd_out,
                                float* __restrict__ alpha_out) {
  using Cf = GataCfg<LMAX, SD, ST>;
  constexpr int L = Cf::L, S = Cf::S;
  extern __shared__ float smem_f[];
  const int i = blockIdx.x, c = threadIdx.x;
  const bool act = c < C;
  const int D = C / H, W = D < 32 ? D : 32, nparts = C / W, segs = D / W;
  const int SC = S * C, SD_ = S * D;  // SD_ = value columns per head
  GataSmem sm = carve(smem_f, max_deg, nparts, H, L, C, false);
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (deg > max_deg) __trap();  // host passed a too small max in-degree

  for (int t = threadIdx.x; t < deg; t += blockDim.x) {
    sm.src[t] = src[e0 + t]; sm.fc[t] = fc[e0 + t]; sm.kap[t] = kappa[e0 + t];
  }
  for (int t = threadIdx.x; t < deg * L; t += blockDim.x) sm.Y[t] = Y[(size_t)e0 * L + t];
  __syncthreads();

  // ---- attention logits: a[e][hd] = sum_d q_i k_j silu(W_re t)   (gotennet.py:502)
  const float qi = act ? qk[(size_t)i * ldqk + c] : 0.f;
  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    float p = 0.f;
    if (act) p = qi * qk[(size_t)j * ldqk + C + c] * siluf_(Ze[(size_t)(e0 + t) * ldz + c]);
    for (int o = W >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (act && (c % W) == 0) sm.part[t * nparts + c / W] = p;
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
        sm.alpha[t * H + hd] = al;
        alpha_out[(size_t)(e0 + t) * H + hd] = al;
      }
    }
  }
  __syncthreads();
  if (!act) return;

  // ---- messages + aggregation in registers (gotennet.py:516-558, :638-639)
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = (k * C + c) / SD_;
  float acc_h = 0.f, accX[L];
#pragma unroll
  for (int m = 0; m < L; ++m) accX[m] = 0.f;

  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    const size_t e = (size_t)(e0 + t);
    const float f = sm.fc[t], kap = sm.kap[t];
    float o[S];
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const int col = k * C + c;
      const float spatial = Ze[e * ldz + C + col] * x[(size_t)j * SC + col] * f;
      const float sea = (sm.alpha[t * H + hd_of[k]] * kap) * v[(size_t)j * SC + col];
      o[k] = spatial + sea;
    }
    acc_h += o[0];
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      const float od = o[1 + (SD ? l : 0)], ot = o[1 + Cf::ND + (ST ? l : 0)];
#pragma unroll
      for (int m = lo_of(l); m < hi_of(l); ++m)
        accX[m] += sm.Y[t * L + m] * od + Xd[((size_t)m * N + j) * C + c] * ot;
    }
  }
  h_out[(size_t)i * C + c] = h[(size_t)i * C + c] + acc_h;
#pragma unroll
  for (int m = 0; m < L; ++m) {
    const size_t o_ = ((size_t)m * N + i) * C + c;
    Xd_out[o_] = Xd[o_] + accX[m];
  }
}

// --------------------------------------------------------- backward, target ---
template <int LMAX, bool SD, bool ST>
__global__ void gata_bwd_tgt_kernel(const float* __restrict__ g_h, const float* __restrict__ g_Xd,
                                    const float* __restrict__ Xd, const float* __restrict__ qk, int ldqk,
                                    const float* __restrict__ x, const float* __restrict__ v,
                                    const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                    const float* __restrict__ fc, const float* __restrict__ kappa,
                                    const float* __restrict__ alpha, const int32_t* __restrict__ tgt_ptr,
                                    const int32_t* __restrict__ src, int N, int C, int H, int max_deg,
                                    float* __restrict__ g_qk, int ldgqk, float* __restrict__ gZe, int ldgz,
                                    float* __restrict__ da_out, float* __restrict__ g_fc, float* __restrict__ g_Y) {
  using Cf = GataCfg<LMAX, SD, ST>;
  constexpr int L = Cf::L, S = Cf::S, ND = Cf::ND;
  extern __shared__ float smem_f[];
  __shared__ float red[33];
  const int i = blockIdx.x, c = threadIdx.x;
  const bool act = c < C;
  const int D = C / H, W = D < 32 ? D : 32, nparts = C / W;
  const int SC = S * C, SD_ = S * D;
  GataSmem sm = carve(smem_f, max_deg, nparts, H, L, C, true);
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (deg > max_deg) __trap();  // host passed a too small max in-degree
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;

  for (int t = threadIdx.x; t < deg; t += blockDim.x) {
    sm.src[t] = src[e0 + t]; sm.fc[t] = fc[e0 + t]; sm.kap[t] = kappa[e0 + t];
  }
  for (int t = threadIdx.x; t < deg * L; t += blockDim.x) sm.Y[t] = Y[(size_t)e0 * L + t];
  for (int t = threadIdx.x; t < deg * H; t += blockDim.x) sm.alpha[t] = alpha[(size_t)e0 * H + t];
  float gh = 0.f, gX[L];
#pragma unroll
  for (int m = 0; m < L; ++m) gX[m] = 0.f;
  if (act) {
    gh = g_h[(size_t)i * C + c];
    sm.gh[c] = gh;
#pragma unroll
    for (int m = 0; m < L; ++m) { gX[m] = g_Xd[((size_t)m * N + i) * C + c]; sm.gX[m * C + c] = gX[m]; }
  }
  __syncthreads();

  // ---- d alpha~[e][hd] = sum_{col in head hd} dout[e][col] * v_j[col]; one warp per head
  for (int hd = w; hd < H; hd += nw) {
    for (int t = 0; t < deg; ++t) {
      const int j = sm.src[t];
      float acc = 0.f;
      for (int col = hd * SD_ + lane; col < (hd + 1) * SD_; col += 32) {
        const int k = col / C, cc = col - k * C;
        float d;
        if (k == 0) {
          d = sm.gh[cc];
        } else if (k < 1 + ND) {
          const int l = SD ? k - 1 : -1;
          const int m0 = l < 0 ? 0 : lo_of(l), m1 = l < 0 ? L : hi_of(l);
          d = 0.f;
          for (int m = m0; m < m1; ++m) d = fmaf(sm.Y[t * L + m], sm.gX[m * C + cc], d);
        } else {
          const int l = ST ? k - 1 - ND : -1;
          const int m0 = l < 0 ? 0 : lo_of(l), m1 = l < 0 ? L : hi_of(l);
          d = 0.f;
          for (int m = m0; m < m1; ++m) d = fmaf(Xd[((size_t)m * N + j) * C + cc], sm.gX[m * C + cc], d);
        }
        acc = fmaf(d, v[(size_t)j * SC + col], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) sm.aux[t * H + hd] = acc * sm.kap[t];  // alpha~ = alpha * kappa
    }
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

  // ---- per-channel pass: dq, d(pre-act W_re), d(filter), optional geometry gradients
  const float qi = act ? qk[(size_t)i * ldqk + c] : 0.f;
  const int hq = act ? c / D : 0;
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = act ? (k * C + c) / SD_ : 0;
  float gq = 0.f;
  const bool geom = (g_fc != nullptr) || (g_Y != nullptr);
  for (int t = 0; t < deg; ++t) {
    const int j = sm.src[t];
    const size_t e = (size_t)(e0 + t);
    const float f = sm.fc[t];
    float gfc_part = 0.f, gy_part[L];
#pragma unroll
    for (int m = 0; m < L; ++m) gy_part[m] = 0.f;
    if (act) {
      const float kj = qk[(size_t)j * ldqk + C + c];
      const float zre = Ze[e * ldz + c];
      const float dal = sm.aux[t * H + hq];
      gq = fmaf(dal * kj, siluf_(zre), gq);
      gZe[e * ldgz + c] = dal * qi * kj * dsiluf_(zre);
      float Xj[L], y[L], dout[S];
#pragma unroll
      for (int m = 0; m < L; ++m) { Xj[m] = Xd[((size_t)m * N + j) * C + c]; y[m] = sm.Y[t * L + m]; }
      chunk_grads<LMAX, SD, ST>(gh, gX, y, Xj, dout);
      float o[S];
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const int col = k * C + c;
        const float tf = Ze[e * ldz + C + col], xj = x[(size_t)j * SC + col];
        gZe[e * ldgz + C + col] = dout[k] * xj * f;
        if (geom) {
          gfc_part = fmaf(dout[k], tf * xj, gfc_part);
          o[k] = tf * xj * f + (sm.alpha[t * H + hd_of[k]] * sm.kap[t]) * v[(size_t)j * SC + col];
        }
      }
      if (geom) {
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
          const float od = o[1 + (SD ? l : 0)];
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) gy_part[m] = od * gX[m];
        }
      }
    }
    if (g_fc != nullptr) {  // block-uniform branches
      const float s = block_sum(gfc_part, red);
      if (threadIdx.x == 0) g_fc[e] += s;
    }
    if (g_Y != nullptr) {
#pragma unroll
      for (int m = 0; m < L; ++m) {
        const float s = block_sum(gy_part[m], red);
        if (threadIdx.x == 0) g_Y[e * L + m] += s;
      }
    }
  }
  if (act) g_qk[(size_t)i * ldgqk + c] = gq;
}

// --------------------------------------------------------- backward, source ---
constexpr int SRC_CHUNK = 32;

template <int LMAX, bool SD, bool ST>
__global__ void gata_bwd_src_kernel(const float* __restrict__ g_h, const float* __restrict__ g_Xd,
                                    const float* __restrict__ Xd, const float* __restrict__ qk, int ldqk,
                                    const float* __restrict__ x, const float* __restrict__ v,
                                    const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                    const float* __restrict__ fc, const float* __restrict__ kappa,
                                    const float* __restrict__ alpha, const float* __restrict__ da,
                                    const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_perm,
                                    const int32_t* __restrict__ tgt, int N, int C, int H, float* __restrict__ g_qk,
                                    int ldgqk, float* __restrict__ g_x, float* __restrict__ g_v,
                                    float* __restrict__ g_Xd_in) {
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

  const int j = blockIdx.x, c = threadIdx.x;
  const bool act = c < C;
  const int D = C / H, SC = S * C, SD_ = S * D;
  float xo[S], vo[S], Xo[L], gx[S], gv[S], gXin[L];
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
    xo[k] = act ? x[(size_t)j * SC + k * C + c] : 0.f;
    vo[k] = act ? v[(size_t)j * SC + k * C + c] : 0.f;
    gx[k] = 0.f; gv[k] = 0.f;
    hd_of[k] = act ? (k * C + c) / SD_ : 0;
  }
#pragma unroll
  for (int m = 0; m < L; ++m) { Xo[m] = act ? Xd[((size_t)m * N + j) * C + c] : 0.f; gXin[m] = 0.f; }
  const float kj = act ? qk[(size_t)j * ldqk + C + c] : 0.f;
  (void)kj;
  const int hq = act ? c / D : 0;
  float gk = 0.f;

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
      s_al[q] = alpha[o_]; s_da[q] = da[o_];
    }
    __syncthreads();
    if (act) {
      for (int t = 0; t < n; ++t) {
        const size_t e = (size_t)s_e[t];
        const int i = s_i[t];
        const float f = s_fc[t], kap = s_kap[t];
        float gX[L], y[L], dout[S];
        const float gh = g_h[(size_t)i * C + c];
#pragma unroll
        for (int m = 0; m < L; ++m) { gX[m] = g_Xd[((size_t)m * N + i) * C + c]; y[m] = s_Y[t * L + m]; }
        chunk_grads<LMAX, SD, ST>(gh, gX, y, Xo, dout);
        float o[S];
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const float tf = Ze[e * ldz + C + k * C + c];
          const float al = s_al[t * H + hd_of[k]] * kap;
          gx[k] = fmaf(dout[k], tf * f, gx[k]);
          gv[k] = fmaf(dout[k], al, gv[k]);
          o[k] = tf * xo[k] * f + al * vo[k];
        }
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
          const float ot = o[1 + ND + (ST ? l : 0)];
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) gXin[m] = fmaf(ot, gX[m], gXin[m]);
        }
        gk = fmaf(s_da[t * H + hq] * qk[(size_t)i * ldqk + c], siluf_(Ze[e * ldz + c]), gk);
      }
    }
  }
  if (act) {
#pragma unroll
    for (int k = 0; k < S; ++k) {
      g_x[(size_t)j * SC + k * C + c] = gx[k];
      g_v[(size_t)j * SC + k * C + c] = gv[k];
    }
    g_qk[(size_t)j * ldgqk + C + c] = gk;
#pragma unroll
    for (int m = 0; m < L; ++m) {
      const size_t o_ = ((size_t)m * N + j) * C + c;
      g_Xd_in[o_] = g_Xd[o_] + gXin[m];
    }
  }
}

static inline int gata_block(int C) { return ((C + 31) / 32) * 32; }

static int gata_check(int C, int H, int lmax) {
  GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);
  GOTEN_REQUIRE(C >= 1 && C <= 1024 && H >= 1 && C % H == 0, "n_atom_basis=%d / num_heads=%d unsupported", C, H);
  const int D = C / H;
  GOTEN_REQUIRE((D <= 32 && (D & (D - 1)) == 0) || (D % 32 == 0),
                "head width %d unsupported (power of two <= 32, or a multiple of 32)", D);
  return 0;
}

}  // namespace goten

using namespace goten;

// dispatch over (lmax, sep_dir, sep_tensor); for lmax == 1 the sep flags do not change the layout
#define GATA_DISPATCH(KERNEL, GRID, BLOCK, SMEM, ...)                                                   \
  do {                                                                                                  \
    const bool sd = (flags & 1) && lmax > 1, stn = (flags & 2) && lmax > 1;                             \
    if (lmax == 1) { GATA_LAUNCH(KERNEL, 1, false, false, GRID, BLOCK, SMEM, __VA_ARGS__); }            \
    else if (lmax == 2 && !sd && !stn) { GATA_LAUNCH(KERNEL, 2, false, false, GRID, BLOCK, SMEM, __VA_ARGS__); } \
    else if (lmax == 2 && sd && !stn) { GATA_LAUNCH(KERNEL, 2, true, false, GRID, BLOCK, SMEM, __VA_ARGS__); }   \
    else if (lmax == 2 && !sd && stn) { GATA_LAUNCH(KERNEL, 2, false, true, GRID, BLOCK, SMEM, __VA_ARGS__); }   \
    else if (lmax == 2) { GATA_LAUNCH(KERNEL, 2, true, true, GRID, BLOCK, SMEM, __VA_ARGS__); }         \
    else if (!sd && !stn) { GATA_LAUNCH(KERNEL, 3, false, false, GRID, BLOCK, SMEM, __VA_ARGS__); }     \
    else if (sd && !stn) { GATA_LAUNCH(KERNEL, 3, true, false, GRID, BLOCK, SMEM, __VA_ARGS__); }       \
    else if (!sd && stn) { GATA_LAUNCH(KERNEL, 3, false, true, GRID, BLOCK, SMEM, __VA_ARGS__); }       \
    else { GATA_LAUNCH(KERNEL, 3, true, true, GRID, BLOCK, SMEM, __VA_ARGS__); }                        \
  } while (0)

#define GATA_LAUNCH(KERNEL, LM, SD, ST, GRID, BLOCK, SMEM, ...)                                         \
  do {                                                                                                  \
    auto kfn = KERNEL<LM, SD, ST>;                                                                      \
    if ((SMEM) > 48 * 1024)                                                                             \
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
    kfn<<<GRID, BLOCK, SMEM, st>>>(__VA_ARGS__);                                                        \
  } while (0)

extern "C" {

int goten_gata_fwd(const float* h, const float* Xd, const float* qk, int ldqk, const float* x, const float* v,
                   const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa,
                   const int32_t* tgt_ptr, const int32_t* src, int N, int C, int H, int lmax, int flags,
                   int max_deg_in, float* h_out, float* Xd_out, float* alpha, void* stream) {
  if (gata_check(C, H, lmax)) return 1;
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int D = C / H, W = D < 32 ? D : 32, nparts = C / W, L = (lmax + 1) * (lmax + 1) - 1;
  if (max_deg_in < 1) max_deg_in = 1;
  const size_t smem = gata_smem_floats(max_deg_in, nparts, H, L, C, false) * sizeof(float);
  GOTEN_REQUIRE(smem <= 200 * 1024, "max in-degree %d needs %zu B of shared memory", max_deg_in, smem);
  GATA_DISPATCH(gata_fwd_kernel, N, gata_block(C), smem, h, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, tgt_ptr, src,
                N, C, H, max_deg_in, h_out, Xd_out, alpha);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_gata_bwd_tgt(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk,
                       const float* x, const float* v, const float* Ze, int ldz, const float* Y, const float* fc,
                       const float* kappa, const float* alpha, const int32_t* tgt_ptr, const int32_t* src, int N,
                       int C, int H, int lmax, int flags, int max_deg_in, float* g_qk, int ldgqk, float* gZe,
                       int ldgz, float* da, float* g_fc, float* g_Y, void* stream) {
  if (gata_check(C, H, lmax)) return 1;
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int D = C / H, W = D < 32 ? D : 32, nparts = C / W, L = (lmax + 1) * (lmax + 1) - 1;
  if (max_deg_in < 1) max_deg_in = 1;
  const size_t smem = gata_smem_floats(max_deg_in, nparts, H, L, C, true) * sizeof(float);
  GOTEN_REQUIRE(smem <= 200 * 1024, "max in-degree %d needs %zu B of shared memory", max_deg_in, smem);
  GATA_DISPATCH(gata_bwd_tgt_kernel, N, gata_block(C), smem, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa,
                alpha, tgt_ptr, src, N, C, H, max_deg_in, g_qk, ldgqk, gZe, ldgz, da, g_fc, g_Y);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_gata_bwd_src(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk,
                       const float* x, const float* v, const float* Ze, int ldz, const float* Y, const float* fc,
                       const float* kappa, const float* alpha, const float* da, const int32_t* src_ptr,
                       const int32_t* src_perm, const int32_t* tgt, int N, int C, int H, int lmax, int flags,
                       float* g_qk, int ldgqk, float* g_x, float* g_v, float* g_Xd_in, void* stream) {
  if (gata_check(C, H, lmax)) return 1;
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int L = (lmax + 1) * (lmax + 1) - 1;
  const size_t smem = (size_t)SRC_CHUNK * (4 + L + 2 * H) * sizeof(float);
  GATA_DISPATCH(gata_bwd_src_kernel, N, gata_block(C), smem, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa,
                alpha, da, src_ptr, src_perm, tgt, N, C, H, g_qk, ldgqk, g_x, g_v, g_Xd_in);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
