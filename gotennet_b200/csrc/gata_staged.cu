// GATA message passing, TMA-staged variant (the production path for 16 B aligned rows).
// Reference: representation/gotennet.py:452-559 (message), :503 (PyG softmax), :613-640 (aggregate),
// :426-427 (residual); backward = autograd of the same.
//
// Why a second variant: the register-gather kernels in gata.cu are latency bound (ncu: 25-35 % of the
// HBM roofline, 12-30 % occupancy) because every byte in flight occupies a register of a resident warp.
// Here one elected thread per CTA issues cp.async.bulk (TMA, 1-D) copies of whole rows -- the per-edge
// filter row Ze[e, C:(S+1)C], the neighbour rows x_j, v_j and the L rows X_j^m -- into a ring of shared
// memory stages, completion counted on mbarriers; the CTA's threads consume one staged edge while the
// next ones are in flight.  Bytes in flight per SM = (resident CTAs) x (ring depth) x (stage bytes),
// i.e. ~100-180 KB at C=256, independent of the register file.
//
//   forward   : attn kernel (logits + segment softmax -> alpha[E][H], 2 KB/edge, plain loads)
//               msg kernel  (ring-staged; weighted sum in registers -> h_out, Xd_out)
// No atomics: results are bit-reproducible run to run.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"   // tensor-map TMA (cp.async.bulk.tensor) + cuTensorMapEncodeTiled entry point

namespace goten {

namespace staged {

constexpr int EC = 32;  // edges per scalar-staging chunk (src, fc, kappa, Y, alpha rows of <= EC edges in smem)

template <int LMAX, bool SD, bool ST>
struct Cfg {
  static constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  static constexpr int ND = SD ? LMAX : 1;
  static constexpr int NT = ST ? LMAX : 1;
  static constexpr int S = 1 + ND + NT;  // gotennet.py:197-203
};

__device__ __forceinline__ constexpr int lo_of(int l) { return (l + 1) * (l + 1) - 1; }  // l = 0.. -> degree l+1
__device__ __forceinline__ constexpr int hi_of(int l) { return (l + 2) * (l + 2) - 1; }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 fma4(float a, float4 b, float4 c) {
  return make_float4(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y), fmaf(a, b.z, c.z), fmaf(a, b.w, c.w));
}
__device__ __forceinline__ float4 fma44(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ------------------------------------------------------------- attention ------
// alpha[e][hd] = softmax over the edges of one target of  sum_{d in head} q_i k_j silu(W_re t)   (gotennet.py:502-503)
// CTA per target, thread per 4 channels.  smem: part[deg][nparts] + logit[deg][H].
__global__ void gata_attn_fwd_kernel(const float* __restrict__ qk, int ldqk, const float* __restrict__ Ze, int ldz,
                                     const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int C, int H,
                                     int max_deg, float* __restrict__ alpha_out) {
  extern __shared__ float smem_f[];
  const int i = blockIdx.x, c = threadIdx.x * 4;
  const bool act = c < C;
  const int D = C / H, Dt = D / 4;
  const int W = Dt < 32 ? Dt : 32;  // shuffle segment width (threads)
  const int nparts = (C / 4) / W, segs = Dt / W;
  float* part = smem_f;                           // [max_deg][nparts]
  float* logit = smem_f + (size_t)max_deg * nparts;  // [max_deg][H]
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (deg > max_deg) __trap();
  float4 qi = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) qi = ld4(qk + (size_t)i * ldqk + c);
  for (int t0 = 0; t0 < deg; t0 += 4) {
    float p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      p[u] = 0.f;
      if (act && t0 + u < deg) {
        const int j = src[e0 + t0 + u];
        const float4 kj = ld4(qk + (size_t)j * ldqk + C + c);
        const float4 z = ld4(Ze + (size_t)(e0 + t0 + u) * ldz + c);
        p[u] = qi.x * kj.x * siluf_(z.x) + qi.y * kj.y * siluf_(z.y) + qi.z * kj.z * siluf_(z.z) + qi.w * kj.w * siluf_(z.w);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      for (int o = W >> 1; o > 0; o >>= 1) p[u] += __shfl_xor_sync(0xffffffffu, p[u], o);
      if (act && t0 + u < deg && (threadIdx.x % W) == 0) part[(t0 + u) * nparts + threadIdx.x / W] = p[u];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int hd = w; hd < H; hd += nw) {
    float mx = -INFINITY;
    for (int t = lane; t < deg; t += 32) {
      float a = 0.f;
      for (int s = 0; s < segs; ++s) a += part[t * nparts + hd * segs + s];
      logit[t * H + hd] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = lane; t < deg; t += 32) {
      const float ex = expf(logit[t * H + hd] - mx);
      logit[t * H + hd] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float den = sum + 1e-16f;  // PyG softmax epsilon
    for (int t = lane; t < deg; t += 32) alpha_out[(size_t)(e0 + t) * H + hd] = logit[t * H + hd] / den;
  }
}

// ------------------------------------------------------------- messages -------
// CTA = (target i, channel slice sl of CH = C / n_slices channels).  Stage layout (floats, this slice only):
//   filter Ze[e, C + k*C + sl*CH ..] [S][CH] | x_j [S][CH] | v_j [S][CH] | X_j [L][CH]
// Each of the four pieces arrives by ONE tensor-map TMA copy (3-D boxes over Ze viewed as [E][ldz/C][C], x / v as
// [N][S][C], Xd as [L][N][C]), so a stage costs four issued instructions whatever S and L are, and a slice of half the
// channels halves the stage (n_slices(), measured neutral: the kernels are bound by bytes in flight per SM - ring depth 2
// instead of 1 costs 40 % - and a half-width CTA keeps half the bytes in flight).
template <int LMAX, bool SD, bool ST>
__global__ void gata_msg_fwd_kernel(const __grid_constant__ CUtensorMap tmZe, const __grid_constant__ CUtensorMap tmX,
                                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmXd,
                                    const float* __restrict__ h, const float* __restrict__ Xd,
                                    const float* __restrict__ Y, const float* __restrict__ fc,
                                    const float* __restrict__ kappa, const float* __restrict__ drop, const float* __restrict__ alpha,
                                    const int32_t* __restrict__ tgt_ptr, const int32_t* __restrict__ src, int N, int C,
                                    int CH, int H, int R, float* __restrict__ h_out, float* __restrict__ Xd_out,
                                    float* __restrict__ xd_amax) {
  using Cf = Cfg<LMAX, SD, ST>;
  constexpr int L = Cf::L, S = Cf::S, ND = Cf::ND;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, CH4 = CH >> 2;
  const int c0s = blockIdx.y * CH;            // first channel of this slice
  const int c = c0s + tid * 4;
  const bool act = tid * 4 < CH;
  const int stage_floats = (3 * S + L) * CH;
  const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)R * stage_bytes);
  int* s_src = reinterpret_cast<int*>(bars + 8);  // 8 barrier slots (R <= 8)
  float* s_fc = reinterpret_cast<float*>(s_src + EC);
  float* s_kap = s_fc + EC;
  float* s_Y = s_kap + EC;        // [EC][L]
  float* s_al = s_Y + EC * L;     // [EC][H]
  const uint32_t bar0 = tma::smem_u32(bars);
  const uint32_t stage0 = tma::smem_u32(stages);

  const int i = blockIdx.x;
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (tid == 0) {
    for (int s = 0; s < R; ++s) tma::mbar_init(bar0 + 8 * s, 1);
    tma::fence_barrier_init();
  }
  const int SD_ = S * (C / H);  // value columns per head (gotennet.py:516-519)
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = act ? (k * C + c) / SD_ : 0;
  float4 acc_h = make_float4(0.f, 0.f, 0.f, 0.f), accX[L];
#pragma unroll
  for (int m = 0; m < L; ++m) accX[m] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto issue = [&](int gt, int t_local) {  // elected thread: the four boxes of edge gt -> stage gt % R
    const int s = gt % R;
    const uint32_t bar = bar0 + 8 * s, dst = stage0 + (uint32_t)s * stage_bytes;
    const int j = s_src[t_local];
    const uint32_t piece = (uint32_t)(S * CH) * 4u;
    tma::mbar_expect_tx(bar, stage_bytes);
    tc::tma_load_3d(dst, &tmZe, bar, c0s, 1, e0 + gt);            // chunks 1 .. S of the edge's projection row
    tc::tma_load_3d(dst + piece, &tmX, bar, c0s, 0, j);
    tc::tma_load_3d(dst + 2u * piece, &tmV, bar, c0s, 0, j);
    tc::tma_load_3d(dst + 3u * piece, &tmXd, bar, c0s, j, 0);     // the L degree rows of node j
  };

  for (int c0 = 0; c0 < deg; c0 += EC) {
    const int n = min(EC, deg - c0);
    __syncthreads();  // previous chunk fully consumed (also orders the mbarrier init before first use)
    for (int t = tid; t < n; t += blockDim.x) {
      s_src[t] = src[e0 + c0 + t]; s_fc[t] = fc[e0 + c0 + t]; s_kap[t] = kappa[e0 + c0 + t];
    }
    for (int t = tid; t < n * L; t += blockDim.x) s_Y[t] = Y[(size_t)(e0 + c0) * L + t];
    // attention dropout (gotennet.py:513): `drop` holds mask / (1 - p) per (edge, head); NULL = eval / p = 0
    for (int t = tid; t < n * H; t += blockDim.x) {
      const size_t o_ = (size_t)(e0 + c0) * H + t;
      s_al[t] = drop ? alpha[o_] * drop[o_] : alpha[o_];
    }
    __syncthreads();
    if (tid == 0) {
      const int pre = n < R ? n : R;
      for (int u = 0; u < pre; ++u) issue(c0 + u, u);
    }
    for (int t = 0; t < n; ++t) {
      const int gt = c0 + t, s = gt % R;
      if (act) {
        tma::mbar_wait(bar0 + 8 * s, (uint32_t)((gt / R) & 1));
        const float4* st = reinterpret_cast<const float4*>(stages + (size_t)s * stage_floats);
        const float f = s_fc[t], kap = s_kap[t];
        float4 o[S];
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const float4 tf = st[k * CH4 + tid], xv = st[(S + k) * CH4 + tid], vv = st[(2 * S + k) * CH4 + tid];
          const float al = s_al[t * H + hd_of[k]] * kap;
          o[k] = make_float4(tf.x * xv.x * f + al * vv.x, tf.y * xv.y * f + al * vv.y, tf.z * xv.z * f + al * vv.z,
                             tf.w * xv.w * f + al * vv.w);
        }
        acc_h = add4(acc_h, o[0]);
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) {
            const float4 Xj = st[(3 * S) * CH4 + m * CH4 + tid];
            const float y = s_Y[t * L + m];
            const float4 od = o[1 + (SD ? l : 0)], ot = o[1 + ND + (ST ? l : 0)];
            accX[m].x += y * od.x + Xj.x * ot.x;
            accX[m].y += y * od.y + Xj.y * ot.y;
            accX[m].z += y * od.z + Xj.z * ot.z;
            accX[m].w += y * od.w + Xj.w * ot.w;
          }
        }
      }
      __syncthreads();  // every thread is done with stage s
      if (tid == 0 && t + R < n) issue(gt + R, t + R);
    }
  }
  if (!act) return;
  st4(h_out + (size_t)i * C + c, add4(ld4(h + (size_t)i * C + c), acc_h));
  float xamx = 0.f;
#pragma unroll
  for (int m = 0; m < L; ++m) {
    const size_t o_ = ((size_t)m * N + i) * C + c;
    const float4 xo = add4(ld4(Xd + o_), accX[m]);
    xamx = amax4(xamx, xo.x, xo.y, xo.z, xo.w);
    st4(Xd_out + o_, xo);
  }
  amax_commit(xd_amax, xamx);
}

// --------------------------------------------------------- backward, target ---
// Stage layout (floats): X_j [L*C] | v_j [S*C] | x_j [S*C].   Per target i:
//   pass A (ring-staged): dout[k] for every edge (gotennet.py:516-558 differentiated), d(filter) written straight to
//           gZe[e, C:(S+1)C], d alpha~ partials per (edge, chunk, column group) into smem;
//   softmax backward (gotennet.py:503-511)  -> da[e][hd] (also consumed by the source pass);
//   pass B (plain loads, 2 KB/edge): dq_i and d(pre-activation of W_re) -> gZe[e, 0:C].
// smem tail: part[EC][S][n_grp] (one chunk of edges) | al[max_deg][H] | aux[max_deg][H]
// GEO: also the geometry gradients (forces): g_fc[e] += sum_c dout . (filter x_j) and g_Y[e][m] += sum_c o_d . gX_i[m]
// (o_d = the forward message of the direction chunk), reduced over the block with one extra barrier per edge; the
// filter row of the edge is read straight from Ze (it is used once, by one thread).
template <int LMAX, bool SD, bool ST, bool GEO>
__device__ __forceinline__ void gata_bwd_tgt_staged_body(
    const CUtensorMap& tmXd, const float* __restrict__ g_h, const float* __restrict__ g_Xd, const float* __restrict__ Xd,
    const float* __restrict__ qk, int ldqk, const float* __restrict__ x, const float* __restrict__ v,
    const float* __restrict__ Ze, int ldz, const float* __restrict__ Y, const float* __restrict__ fc,
    const float* __restrict__ kappa, const float* __restrict__ drop, const float* __restrict__ alpha, const int32_t* __restrict__ tgt_ptr,
    const int32_t* __restrict__ src, int N, int C, int H, int R, int max_deg, int g_cols, float* __restrict__ g_qk,
    int ldgqk, float* __restrict__ gZe, int ldgz, float* __restrict__ da_out, float* __restrict__ gze_amax,
    float* __restrict__ g_fc, float* __restrict__ g_Y) {
  using Cf = Cfg<LMAX, SD, ST>;
  float amx = 0.f;
  constexpr int L = Cf::L, S = Cf::S, ND = Cf::ND;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, c = tid * 4, C4 = C >> 2;
  const bool act = c < C;
  const int SC = S * C, D = C / H, SD_ = S * D;
  const int gt_ = g_cols / 4;      // threads per column group (power of two <= 32)
  const int n_grp = C / g_cols;    // groups per chunk
  const int stage_floats = (2 * S + L) * C;
  const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)R * stage_bytes);
  int* s_src = reinterpret_cast<int*>(bars + 8);  // 8 barrier slots (R <= 8)
  float* s_fc = reinterpret_cast<float*>(s_src + EC);
  float* s_Y = s_fc + EC;                               // [EC][L]
  float* part = s_Y + EC * L;                           // [EC][S][n_grp]: group sums of the current chunk of edges
  float* s_al = part + (size_t)EC * S * n_grp;          // [max_deg][H]
  float* s_aux = s_al + (size_t)max_deg * H;            // [max_deg][H]
  // [2][S][blockDim]: per-thread d alpha~ partials of one edge (16 B aligned for the vector group sums)
  // (offset arithmetic on the shared array, not an integer round trip of the pointer: the latter makes every access to
  //  s_pk / s_geo a generic LD.E / ST.E instead of LDS / STS)
  float* s_pk = reinterpret_cast<float*>(smem_raw) +
                (((size_t)((s_aux + (size_t)max_deg * H) - reinterpret_cast<float*>(smem_raw)) + 3) & ~size_t(3));
  float* s_geo = s_pk + (size_t)2 * S * blockDim.x;      // GEO: [2][(1 + L)][blockDim + 4]
  const uint32_t bar0 = tma::smem_u32(bars);
  const uint32_t stage0 = tma::smem_u32(stages);

  const int i = blockIdx.x;
  const int e0 = tgt_ptr[i];
  const int deg = tgt_ptr[i + 1] - e0;
  if (deg > max_deg) __trap();
  if (tid == 0) {
    tma::mbar_init(bar0, 1);
    tma::mbar_init(bar0 + 8, 1);
    tma::fence_barrier_init();
  }
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = act ? (k * C + c) / SD_ : 0;
  if (GEO) {  // attention weights of every edge of this target (pass A needs alpha~ = alpha * kappa)
    for (int idx = tid; idx < deg * H; idx += blockDim.x) s_al[idx] = alpha[(size_t)e0 * H + idx];
  }
  float4 gh = make_float4(0.f, 0.f, 0.f, 0.f), gX[L];
#pragma unroll
  for (int m = 0; m < L; ++m) gX[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) {
    gh = ld4(g_h + (size_t)i * C + c);
#pragma unroll
    for (int m = 0; m < L; ++m) gX[m] = ld4(g_Xd + ((size_t)m * N + i) * C + c);
  }

  // ONE stage, refilled in two pieces with their own barriers (slots 0 / 1): piece X = the L rows of X_j, consumed by the
  // first half of an edge's work (dout), piece VX = v_j | x_j, consumed by the second half.  As soon as every thread is
  // past the first half, the next edge's X piece is requested, so its round trip overlaps the second half - the overlap
  // of a deeper ring without its shared memory (ring depth 2 halves the resident CTAs: +40 %).  (R is unused here.)
  const uint32_t barX = bar0, barV = bar0 + 8;
  auto issueX = [&](int t_local) {
    tma::mbar_expect_tx(barX, (uint32_t)(L * C) * 4u);
    tc::tma_load_3d(stage0, &tmXd, barX, 0, s_src[t_local], 0);   // the L degree rows of node j in one tensor-map copy
  };
  auto issueV = [&](int t_local) {
    const int j = s_src[t_local];
    tma::mbar_expect_tx(barV, 2u * (uint32_t)SC * 4u);
    tma::bulk_g2s(stage0 + (uint32_t)(L * C) * 4u, v + (size_t)j * SC, (uint32_t)SC * 4u, barV);
    tma::bulk_g2s(stage0 + (uint32_t)(L * C + SC) * 4u, x + (size_t)j * SC, (uint32_t)SC * 4u, barV);
  };

  // ---- pass A
  for (int c0 = 0; c0 < deg; c0 += EC) {
    const int n = min(EC, deg - c0);
    __syncthreads();
    for (int t = tid; t < n; t += blockDim.x) { s_src[t] = src[e0 + c0 + t]; s_fc[t] = fc[e0 + c0 + t]; }
    for (int t = tid; t < n * L; t += blockDim.x) s_Y[t] = Y[(size_t)(e0 + c0) * L + t];
    __syncthreads();
    if (tid == 0) { issueX(0); issueV(0); }
    for (int t = 0; t < n; ++t) {
      const int gt = c0 + t;
      constexpr int s = 0;
      float pk[S];
#pragma unroll
      for (int k = 0; k < S; ++k) pk[k] = 0.f;
      float geo[1 + L];  // GEO: [d fc | d Y_m] partials of this thread
#pragma unroll
      for (int m = 0; m <= L; ++m) geo[m] = 0.f;
      const float4* st = reinterpret_cast<const float4*>(stages + (size_t)s * stage_floats);
      float4 dout[S];
      dout[0] = gh;
#pragma unroll
      for (int k = 1; k < S; ++k) dout[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) {
        tma::mbar_wait(barX, (uint32_t)(gt & 1));
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) {
            const float y = s_Y[t * L + m];
            const float4 Xj = st[m * C4 + tid];
            dout[1 + (SD ? l : 0)] = fma4(y, gX[m], dout[1 + (SD ? l : 0)]);
            dout[1 + ND + (ST ? l : 0)] = fma44(Xj, gX[m], dout[1 + ND + (ST ? l : 0)]);
          }
        }
      }
      __syncthreads();                                   // every thread is past the X piece
      if (tid == 0 && t + 1 < n) issueX(t + 1);
      if (act) {
        tma::mbar_wait(barV, (uint32_t)(gt & 1));
        const float f = s_fc[t];
        float* gz = gZe + (size_t)(e0 + gt) * ldgz + C + c;
        float4 od[ND];  // GEO: forward messages of the direction chunks (gotennet.py:522-526, 532)
        const float kap = GEO ? kappa[e0 + gt] : 0.f;
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const float4 vv = st[(L + k) * C4 + tid], xv = st[(L + S + k) * C4 + tid];
          pk[k] = dout[k].x * vv.x + dout[k].y * vv.y + dout[k].z * vv.z + dout[k].w * vv.w;
          const float4 gv = make_float4(dout[k].x * xv.x * f, dout[k].y * xv.y * f, dout[k].z * xv.z * f, dout[k].w * xv.w * f);
          amx = amax4(amx, gv.x, gv.y, gv.z, gv.w);
          st4(gz + k * C, gv);
          if (GEO) {
            const float4 tf = ld4(Ze + (size_t)(e0 + gt) * ldz + C + k * C + c);
            const float4 sx = make_float4(tf.x * xv.x, tf.y * xv.y, tf.z * xv.z, tf.w * xv.w);
            geo[0] += dout[k].x * sx.x + dout[k].y * sx.y + dout[k].z * sx.z + dout[k].w * sx.w;
            if (k >= 1 && k <= ND) {
              float al = s_al[gt * H + hd_of[k]] * kap;
              if (drop) al *= drop[(size_t)(e0 + gt) * H + hd_of[k]];
              od[k - 1] = make_float4(fmaf(sx.x, f, al * vv.x), fmaf(sx.y, f, al * vv.y), fmaf(sx.z, f, al * vv.z),
                                      fmaf(sx.w, f, al * vv.w));
            }
          }
        }
        if (GEO) {
#pragma unroll
          for (int l = 0; l < LMAX; ++l) {
#pragma unroll
            for (int m = lo_of(l); m < hi_of(l); ++m) {
              const float4 o4 = od[SD ? l : 0];
              geo[1 + m] = o4.x * gX[m].x + o4.y * gX[m].y + o4.z * gX[m].z + o4.w * gX[m].w;
            }
          }
        }
      }
      if (GEO) {  // block-uniform: 1 + L channel sums of this edge
        const size_t e = (size_t)(e0 + gt);
        block_sums_one_barrier<1 + L>(geo, s_geo + (size_t)(gt & 1) * (1 + L) * (blockDim.x + 4), [&](int vi, float sum) {
          if (vi == 0) { if (g_fc != nullptr) g_fc[e] += sum; }
          else if (g_Y != nullptr) g_Y[e * L + (vi - 1)] += sum;
        });
      }
      // group sums of the partials (gt_ consecutive threads, never across a warp): every thread stores its S values,
      // then S * n_grp threads each add one group's gt_ values (conflict-free smem reads instead of S log2(gt_)
      // dependent shuffles per thread)
      // (double buffered on the edge parity: the buffer written for edge t is re-written for edge t + 2, which a
      //  thread reaches only after the barrier of edge t + 1, i.e. after every thread has finished reading it)
      float* pkb = s_pk + (size_t)(gt & 1) * S * blockDim.x;
#pragma unroll
      for (int k = 0; k < S; ++k) pkb[k * blockDim.x + tid] = pk[k];
      __syncthreads();
      for (int o = tid; o < S * n_grp; o += blockDim.x) {
        const int k = o / n_grp, g = o - k * n_grp;
        const float* pp = pkb + k * blockDim.x + g * gt_;
        float p = 0.f;
        if ((gt_ & 3) == 0) {
          for (int u = 0; u < gt_; u += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(pp + u);
            p += (q4.x + q4.y) + (q4.z + q4.w);
          }
        } else {
          for (int u = 0; u < gt_; ++u) p += pp[u];
        }
        part[(size_t)t * S * n_grp + o] = p;
      }
      if (tid == 0 && t + 1 < n) issueV(t + 1);           // (after the barrier above: everybody is past v_j | x_j)
    }
    // ---- chunk epilogue: d alpha~[e][hd] = kappa_e * sum of the group sums whose columns lie in head hd
    // (per chunk, so `part` holds EC edges whatever the in-degree: 160-neighbour molecules keep 4 CTAs per SM)
    __syncthreads();
    for (int idx = tid; idx < n * H; idx += blockDim.x) {
      const int t = idx / H, hd = idx - t * H;
      float sacc = 0.f;
      for (int k = 0; k < S; ++k)
        for (int g = 0; g < n_grp; ++g)
          if ((k * C + g * g_cols) / SD_ == hd) sacc += part[((size_t)t * S + k) * n_grp + g];
      // d alpha = d alpha~ * kappa * (dropout mask / (1 - p)); the softmax backward below uses the un-dropped alpha
      float dal_ = sacc * kappa[e0 + c0 + t];
      if (drop) dal_ *= drop[(size_t)(e0 + c0 + t) * H + hd];
      s_aux[(size_t)(c0 + t) * H + hd] = dal_;
    }
  }
  __syncthreads();
  if (!GEO) {
    for (int idx = tid; idx < deg * H; idx += blockDim.x) s_al[idx] = alpha[(size_t)e0 * H + idx];
  }
  __syncthreads();
  // ---- softmax backward: da = alpha * (dalpha - sum_e alpha dalpha)
  {
    const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    for (int hd = w; hd < H; hd += nw) {
      float dot = 0.f;
      for (int t = lane; t < deg; t += 32) dot = fmaf(s_al[t * H + hd], s_aux[t * H + hd], dot);
      dot = warp_sum(dot);
      for (int t = lane; t < deg; t += 32) {
        const float da = s_al[t * H + hd] * (s_aux[t * H + hd] - dot);
        s_aux[t * H + hd] = da;
        da_out[(size_t)(e0 + t) * H + hd] = da;
      }
    }
  }
  __syncthreads();
  if (!act) return;
  // ---- pass B: dq_i, d(pre-act W_re)
  const float4 qi = ld4(qk + (size_t)i * ldqk + c);
  float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
  const int hq = c / D;
  for (int t0 = 0; t0 < deg; t0 += 4) {
    float4 kj[4], zr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u < deg) {
        const int j = src[e0 + t0 + u];
        kj[u] = ld4(qk + (size_t)j * ldqk + C + c);
        zr[u] = ld4(Ze + (size_t)(e0 + t0 + u) * ldz + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u < deg) {
        const float dal = s_aux[(t0 + u) * H + hq];
        // one sigmoid per element serves silu (for dq) and silu' (for d pre-act W_re)
        const float4 sg = make_float4(sigmoid_fast_(zr[u].x), sigmoid_fast_(zr[u].y), sigmoid_fast_(zr[u].z), sigmoid_fast_(zr[u].w));
        gq.x = fmaf(dal * kj[u].x, zr[u].x * sg.x, gq.x);
        gq.y = fmaf(dal * kj[u].y, zr[u].y * sg.y, gq.y);
        gq.z = fmaf(dal * kj[u].z, zr[u].z * sg.z, gq.z);
        gq.w = fmaf(dal * kj[u].w, zr[u].w * sg.w, gq.w);
        const float4 gv = make_float4(dal * qi.x * kj[u].x * sg.x * (1.0f + zr[u].x * (1.0f - sg.x)),
                                      dal * qi.y * kj[u].y * sg.y * (1.0f + zr[u].y * (1.0f - sg.y)),
                                      dal * qi.z * kj[u].z * sg.z * (1.0f + zr[u].z * (1.0f - sg.z)),
                                      dal * qi.w * kj[u].w * sg.w * (1.0f + zr[u].w * (1.0f - sg.w)));
        amx = amax4(amx, gv.x, gv.y, gv.z, gv.w);
        st4(gZe + (size_t)(e0 + t0 + u) * ldgz + c, gv);
      }
    }
  }
  st4(g_qk + (size_t)i * ldgqk + c, gq);
  amax_commit(gze_amax, amx);
}

#define GOTEN_BWD_TGT_ARGS                                                                                            \
  const __grid_constant__ CUtensorMap tmXd, const float *__restrict__ g_h, const float *__restrict__ g_Xd, const float *__restrict__ Xd,                       \
      const float *__restrict__ qk, int ldqk, const float *__restrict__ x, const float *__restrict__ v,              \
      const float *__restrict__ Ze, int ldz, const float *__restrict__ Y, const float *__restrict__ fc,              \
      const float *__restrict__ kappa, const float *__restrict__ drop, const float *__restrict__ alpha, const int32_t *__restrict__ tgt_ptr,         \
      const int32_t *__restrict__ src, int N, int C, int H, int R, int max_deg, int g_cols, float *__restrict__ g_qk, \
      int ldgqk, float *__restrict__ gZe, int ldgz, float *__restrict__ da_out, float *__restrict__ gze_amax,        \
      float *__restrict__ g_fc, float *__restrict__ g_Y
#define GOTEN_BWD_TGT_PASS                                                                                          \
  tmXd, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha, tgt_ptr, src, N, C, H, R, max_deg, g_cols, g_qk, ldgqk, \
      gZe, ldgz, da_out, gze_amax, g_fc, g_Y
template <int LMAX, bool SD, bool ST>
__global__ void gata_bwd_tgt_staged_kernel(GOTEN_BWD_TGT_ARGS) {
  gata_bwd_tgt_staged_body<LMAX, SD, ST, false>(GOTEN_BWD_TGT_PASS);
}
template <int LMAX, bool SD, bool ST>
__global__ void gata_bwd_tgt_staged_geo_kernel(GOTEN_BWD_TGT_ARGS) {
  gata_bwd_tgt_staged_body<LMAX, SD, ST, true>(GOTEN_BWD_TGT_PASS);
}
#undef GOTEN_BWD_TGT_ARGS
#undef GOTEN_BWD_TGT_PASS

// --------------------------------------------------------- backward, source ---
// Per source j over the transposed view.  Stage layout (floats):
//   Ze[e, 0:(S+1)C] [(S+1)*C] | g_h_i [C] | g_Xd_i [L*C] | q_i [C]
// Own rows (X_j, and the tensor chunks of x_j / v_j) sit in shared memory; dx_j, dv_j, dk_j, dX_j accumulate in registers.
template <int LMAX, bool SD, bool ST>
__global__ void gata_bwd_src_staged_kernel(const __grid_constant__ CUtensorMap tmGX, const float* __restrict__ g_h,
                                           const float* __restrict__ g_Xd,
                                           const float* __restrict__ Xd, const float* __restrict__ qk, int ldqk,
                                           const float* __restrict__ x, const float* __restrict__ v,
                                           const float* __restrict__ Ze, int ldz, const float* __restrict__ Y,
                                           const float* __restrict__ fc, const float* __restrict__ kappa, const float* __restrict__ drop,
                                           const float* __restrict__ alpha, const float* __restrict__ da,
                                           const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_perm,
                                           const int32_t* __restrict__ tgt, int N, int C, int H, int R,
                                           float* __restrict__ g_qk, int ldgqk, float* __restrict__ g_x,
                                           float* __restrict__ g_v, float* __restrict__ g_Xd_in,
                                           float* __restrict__ gx_amax, float* __restrict__ gv_amax) {
  using Cf = Cfg<LMAX, SD, ST>;
  constexpr int L = Cf::L, S = Cf::S, ND = Cf::ND, NT = Cf::NT;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, c = tid * 4, C4 = C >> 2;
  const bool act = c < C;
  const int SC = S * C, D = C / H, SD_ = S * D;
  const int stage_floats = (S + 3 + L) * C;
  const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)R * stage_bytes);
  float* own = reinterpret_cast<float*>(bars + 8);  // X_j [L*C] | x_j tensor chunks [NT*C] | v_j tensor chunks [NT*C]
  int* s_e = reinterpret_cast<int*>(own + (size_t)(L + 2 * NT) * C);
  int* s_i = s_e + EC;
  float* s_fc = reinterpret_cast<float*>(s_i + EC);
  float* s_kap = s_fc + EC;
  float* s_Y = s_kap + EC;       // [EC][L]
  float* s_al = s_Y + EC * L;    // [EC][H]
  float* s_da = s_al + EC * H;   // [EC][H]
  const uint32_t bar0 = tma::smem_u32(bars);
  const uint32_t stage0 = tma::smem_u32(stages);

  const int j = blockIdx.x;
  if (tid == 0) {
    for (int s = 0; s < R; ++s) tma::mbar_init(bar0 + 8 * s, 1);
    tma::fence_barrier_init();
  }
  float4* own4 = reinterpret_cast<float4*>(own);
  if (act) {  // each thread only ever reads back its own columns: no barrier needed for `own`
#pragma unroll
    for (int m = 0; m < L; ++m) own4[m * C4 + tid] = ld4(Xd + ((size_t)m * N + j) * C + c);
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      own4[(L + q) * C4 + tid] = ld4(x + (size_t)j * SC + (1 + ND + q) * C + c);
      own4[(L + NT + q) * C4 + tid] = ld4(v + (size_t)j * SC + (1 + ND + q) * C + c);
    }
  }
  int hd_of[S];
#pragma unroll
  for (int k = 0; k < S; ++k) hd_of[k] = act ? (k * C + c) / SD_ : 0;
  const int hq = act ? c / D : 0;
  float4 gx[S], gv[S], gXin[L], gk = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < S; ++k) { gx[k] = make_float4(0.f, 0.f, 0.f, 0.f); gv[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
  for (int m = 0; m < L; ++m) gXin[m] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto issue = [&](int gp, int t_local) {
    const int s = gp % R;
    const uint32_t bar = bar0 + 8 * s, dst = stage0 + (uint32_t)s * stage_bytes;
    const int e = s_e[t_local], i = s_i[t_local];
    tma::mbar_expect_tx(bar, stage_bytes);
    tma::bulk_g2s(dst, Ze + (size_t)e * ldz, (uint32_t)(S + 1) * (uint32_t)C * 4u, bar);
    tma::bulk_g2s(dst + (uint32_t)((S + 1) * C) * 4u, g_h + (size_t)i * C, (uint32_t)C * 4u, bar);
    tc::tma_load_3d(dst + (uint32_t)((S + 2) * C) * 4u, &tmGX, bar, 0, i, 0);   // the L rows of g_Xd_i in one copy
    tma::bulk_g2s(dst + (uint32_t)((S + 2 + L) * C) * 4u, qk + (size_t)i * ldqk, (uint32_t)C * 4u, bar);
  };

  const int p_begin = src_ptr[j], cnt = src_ptr[j + 1] - p_begin;
  for (int c0 = 0; c0 < cnt; c0 += EC) {
    const int n = min(EC, cnt - c0);
    __syncthreads();
    for (int t = tid; t < n; t += blockDim.x) {
      const int e = src_perm[p_begin + c0 + t];
      s_e[t] = e; s_i[t] = tgt[e]; s_fc[t] = fc[e]; s_kap[t] = kappa[e];
    }
    __syncthreads();
    for (int q = tid; q < n * L; q += blockDim.x) s_Y[q] = Y[(size_t)s_e[q / L] * L + (q % L)];
    for (int q = tid; q < n * H; q += blockDim.x) {
      const size_t o_ = (size_t)s_e[q / H] * H + (q % H);
      s_al[q] = drop ? alpha[o_] * drop[o_] : alpha[o_];  // the message used the dropped weights
      s_da[q] = da[o_];
    }
    if (tid == 0) {
      const int pre = n < R ? n : R;
      for (int u = 0; u < pre; ++u) issue(c0 + u, u);
    }
    __syncthreads();
    for (int t = 0; t < n; ++t) {
      const int gp = c0 + t, s = gp % R;
      if (act) {
        tma::mbar_wait(bar0 + 8 * s, (uint32_t)((gp / R) & 1));
        const float4* st = reinterpret_cast<const float4*>(stages + (size_t)s * stage_floats);
        const float f = s_fc[t], kap = s_kap[t];
        const float4 gh = st[(S + 1) * C4 + tid];
        float4 dout[S];
        dout[0] = gh;
#pragma unroll
        for (int k = 1; k < S; ++k) dout[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) {
            const float y = s_Y[t * L + m];
            const float4 gXm = st[(S + 2 + m) * C4 + tid], Xo = own4[m * C4 + tid];
            dout[1 + (SD ? l : 0)] = fma4(y, gXm, dout[1 + (SD ? l : 0)]);
            dout[1 + ND + (ST ? l : 0)] = fma44(Xo, gXm, dout[1 + ND + (ST ? l : 0)]);
          }
        }
        float4 ot[NT];
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const float4 tf = st[(1 + k) * C4 + tid];
          const float al = s_al[t * H + hd_of[k]] * kap;
          gx[k].x = fmaf(dout[k].x, tf.x * f, gx[k].x); gx[k].y = fmaf(dout[k].y, tf.y * f, gx[k].y);
          gx[k].z = fmaf(dout[k].z, tf.z * f, gx[k].z); gx[k].w = fmaf(dout[k].w, tf.w * f, gx[k].w);
          gv[k] = fma4(al, dout[k], gv[k]);
          if (k >= 1 + ND) {
            const float4 xo = own4[(L + k - 1 - ND) * C4 + tid], vo = own4[(L + NT + k - 1 - ND) * C4 + tid];
            ot[k - 1 - ND] = make_float4(tf.x * xo.x * f + al * vo.x, tf.y * xo.y * f + al * vo.y,
                                         tf.z * xo.z * f + al * vo.z, tf.w * xo.w * f + al * vo.w);
          }
        }
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
#pragma unroll
          for (int m = lo_of(l); m < hi_of(l); ++m) gXin[m] = fma44(ot[ST ? l : 0], st[(S + 2 + m) * C4 + tid], gXin[m]);
        }
        const float4 qi = st[(S + 2 + L) * C4 + tid], zre = st[tid];
        const float dal = s_da[t * H + hq];
        gk.x = fmaf(dal * qi.x, silu_fast_(zre.x), gk.x); gk.y = fmaf(dal * qi.y, silu_fast_(zre.y), gk.y);
        gk.z = fmaf(dal * qi.z, silu_fast_(zre.z), gk.z); gk.w = fmaf(dal * qi.w, silu_fast_(zre.w), gk.w);
      }
      __syncthreads();
      if (tid == 0 && t + R < n) issue(gp + R, t + R);
    }
  }
  if (!act) return;
  float ax = 0.f, av = 0.f;  // running max |g_x|, |g_v|: operand scales of the gamma_s.1 / gamma_v.1 gradient GEMMs
#pragma unroll
  for (int k = 0; k < S; ++k) {
    st4(g_x + (size_t)j * SC + k * C + c, gx[k]);
    st4(g_v + (size_t)j * SC + k * C + c, gv[k]);
    ax = amax4(ax, gx[k].x, gx[k].y, gx[k].z, gx[k].w);
    av = amax4(av, gv[k].x, gv[k].y, gv[k].z, gv[k].w);
  }
  amax_commit(gx_amax, ax);
  amax_commit(gv_amax, av);
  st4(g_qk + (size_t)j * ldgqk + C + c, gk);
#pragma unroll
  for (int m = 0; m < L; ++m) {
    const size_t o_ = ((size_t)m * N + j) * C + c;
    st4(g_Xd_in + o_, add4(ld4(g_Xd + o_), gXin[m]));
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// channel slices per node CTA.  Default 1: halving the slice (twice the CTAs, half the stage each) was measured
// neutral-to-slower on B200 (cfg2 forward 2.65 -> 2.79 ms): bytes in flight per SM stay the same and the per-edge
// barrier / issue cost doubles.  GOTEN_GATA_SLICES=2 selects halves of >= 128 channels for experiments.
static int n_slices(int C) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("GOTEN_GATA_SLICES"); env = e ? atoi(e) : 0; }
  int n = env > 0 ? env : 1;
  while (n > 1 && (C % n != 0 || (C / n) % 128 != 0)) --n;
  return n < 1 ? 1 : n;
}

// 3-D fp32 tensor map over P with extents (d0, d1, d2) (d0 contiguous), byte strides (s1, s2), box (b0, b1, b2), no swizzle
static bool map3d(CUtensorMap* m, const float* P, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                  uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(P), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int ring_depth() {
  static int r = 0;
  if (r == 0) {
    const char* e = getenv("GOTEN_GATA_STAGES");
    r = e ? atoi(e) : 1;  // measured on B200 (cfg2): depth 1 wins, more resident CTAs beat a deeper ring
    if (r < 1) r = 1;
    if (r > 8) r = 8;
  }
  return r;
}

}  // namespace staged

#define STAGED_DISPATCH(KERNEL, GRID, BLOCK, SMEM, ...)                                                     \
  do {                                                                                                      \
    const bool sd = (flags & 1) && lmax > 1, stn = (flags & 2) && lmax > 1;                                 \
    auto launch = [&](auto kfn) -> int {                                                                    \
      if ((SMEM) > 48 * 1024)                                                                               \
        GOTEN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
      kfn<<<GRID, BLOCK, SMEM, st>>>(__VA_ARGS__);                                                          \
      return 0;                                                                                             \
    };                                                                                                      \
    int rc_;                                                                                                \
    if (lmax == 1) rc_ = launch(staged::KERNEL<1, false, false>);                                           \
    else if (lmax == 2 && !sd && !stn) rc_ = launch(staged::KERNEL<2, false, false>);                       \
    else if (lmax == 2 && sd && !stn) rc_ = launch(staged::KERNEL<2, true, false>);                         \
    else if (lmax == 2 && !sd && stn) rc_ = launch(staged::KERNEL<2, false, true>);                         \
    else if (lmax == 2) rc_ = launch(staged::KERNEL<2, true, true>);                                        \
    else if (!sd && !stn) rc_ = launch(staged::KERNEL<3, false, false>);                                    \
    else if (sd && !stn) rc_ = launch(staged::KERNEL<3, true, false>);                                      \
    else if (!sd && stn) rc_ = launch(staged::KERNEL<3, false, true>);                                      \
    else rc_ = launch(staged::KERNEL<3, true, true>);                                                       \
    if (rc_) return rc_;                                                                                    \
  } while (0)

static inline int multiplier_of_(int lmax, int flags) {
  return 3 + ((flags & 1) ? lmax - 1 : 0) + ((flags & 2) ? lmax - 1 : 0);
}

// Returns 0 on success with *handled = true when the staged path ran; *handled = false means the
// shape is outside its contract (rows not 16 B aligned, head layout) and the caller uses gata.cu.
int gata_fwd_staged(const float* h, const float* Xd, const float* qk, int ldqk, const float* x, const float* v,
                    const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                    const int32_t* tgt_ptr, const int32_t* src, int N, int C, int H, int lmax, int flags,
                    int max_deg_in, float* h_out, float* Xd_out, float* alpha, float* xd_amax, cudaStream_t st,
                    bool* handled) {
  *handled = false;
  const int D = C / H;
  if (C % 4 != 0 || D % 4 != 0 || ldqk % 4 != 0 || ldz % 4 != 0) return 0;
  const int Dt = D / 4;
  if (!((Dt <= 32 && (Dt & (Dt - 1)) == 0) || (Dt % 32 == 0))) return 0;
  if (!staged::aligned16(h) || !staged::aligned16(Xd) || !staged::aligned16(qk) || !staged::aligned16(x) ||
      !staged::aligned16(v) || !staged::aligned16(Ze) || !staged::aligned16(h_out) || !staged::aligned16(Xd_out))
    return 0;
  const int L = (lmax + 1) * (lmax + 1) - 1, S = multiplier_of_(lmax, lmax > 1 ? flags : 0);
  const int block = ((C / 4 + 31) / 32) * 32;
  if (block > 1024) return 0;
  // attention kernel
  const int W = Dt < 32 ? Dt : 32, nparts = (C / 4) / W;
  if (max_deg_in < 1) max_deg_in = 1;
  const size_t smem_a = (size_t)max_deg_in * (nparts + H) * sizeof(float);
  if (smem_a > 200 * 1024) return 0;
  // message kernel: CTAs = (target, channel slice); ring of R stages + mbarriers + per-chunk scalars
  const int nsl = staged::n_slices(C);
  const int CH = C / nsl;
  if (ldz % C != 0 || CH > 256 || get_encode() == nullptr) return 0;   // (a TMA box holds <= 256 elements per dim)
  int R = staged::ring_depth();
  const size_t stage_bytes = (size_t)(3 * S + L) * CH * 4;
  const size_t tail = (size_t)8 * 8 + (size_t)staged::EC * (3 + L + H) * 4;
  while (R > 1 && R * stage_bytes + tail > 220 * 1024) --R;
  const size_t smem_m = R * stage_bytes + tail;
  if (smem_m > 220 * 1024) return 0;
  CUtensorMap mZe, mX, mV, mXd;
  {
    const int SC = S * C;
    // the host does not know E here: N * max in-degree bounds it (the map's outer extent only guards out-of-range boxes)
    const uint64_t e_bound = (uint64_t)N * (uint64_t)max_deg_in;
    const bool ok = staged::map3d(&mZe, Ze, C, ldz / C, e_bound, (uint64_t)C * 4, (uint64_t)ldz * 4, CH, S, 1) &&
                    staged::map3d(&mX, x, C, S, N, (uint64_t)C * 4, (uint64_t)SC * 4, CH, S, 1) &&
                    staged::map3d(&mV, v, C, S, N, (uint64_t)C * 4, (uint64_t)SC * 4, CH, S, 1) &&
                    staged::map3d(&mXd, Xd, C, N, L, (uint64_t)C * 4, (uint64_t)N * C * 4, CH, 1, L);
    GOTEN_REQUIRE(ok, "cuTensorMapEncodeTiled failed (GATA message kernel, N=%d C=%d ldz=%d)", N, C, ldz);
  }
  {
    auto kfn = staged::gata_attn_fwd_kernel;
    if (smem_a > 48 * 1024)
      GOTEN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    kfn<<<N, block, smem_a, st>>>(qk, ldqk, Ze, ldz, tgt_ptr, src, C, H, max_deg_in, alpha);
    GOTEN_CHECK_LAUNCH();
  }
  const int block_m = ((CH / 4 + 31) / 32) * 32;
  STAGED_DISPATCH(gata_msg_fwd_kernel, dim3((unsigned)N, (unsigned)nsl), block_m, smem_m, mZe, mX, mV, mXd, h, Xd, Y, fc, kappa,
                  drop, alpha, tgt_ptr, src, N, C, CH, H, R, h_out, Xd_out, xd_amax);
  GOTEN_CHECK_LAUNCH();
  *handled = true;
  return 0;
}

static inline int gcd_i_(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

int gata_bwd_tgt_staged(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk, const float* x,
                        const float* v, const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                        const float* alpha, const int32_t* tgt_ptr, const int32_t* src, int N, int C, int H, int lmax,
                        int flags, int max_deg_in, float* g_qk, int ldgqk, float* gZe, int ldgz, float* da,
                        float* gze_amax, float* g_fc, float* g_Y, cudaStream_t st, bool* handled) {
  *handled = false;
  const int D = C / H;
  if (C % 4 != 0 || D % 4 != 0 || ldqk % 4 != 0 || ldz % 4 != 0 || ldgqk % 4 != 0 || ldgz % 4 != 0) return 0;
  if (!staged::aligned16(g_h) || !staged::aligned16(g_Xd) || !staged::aligned16(Xd) || !staged::aligned16(qk) ||
      !staged::aligned16(x) || !staged::aligned16(v) || !staged::aligned16(Ze) || !staged::aligned16(g_qk) ||
      !staged::aligned16(gZe))
    return 0;
  const int L = (lmax + 1) * (lmax + 1) - 1, S = multiplier_of_(lmax, lmax > 1 ? flags : 0);
  const int g_cols = gcd_i_(S * D, 128);
  if (g_cols % 4 != 0 || C % g_cols != 0) return 0;
  const int n_grp = C / g_cols;
  const int block = ((C / 4 + 31) / 32) * 32;
  if (block > 1024) return 0;
  if (max_deg_in < 1) max_deg_in = 1;
  int R = 1;   // one stage, refilled in two pieces (see the kernel)
  const size_t stage_bytes = (size_t)(2 * S + L) * C * 4;
  const bool geo = g_fc != nullptr || g_Y != nullptr;
  const size_t tail = (size_t)8 * 8 + (size_t)staged::EC * (2 + L + S * n_grp) * 4 + (size_t)max_deg_in * 2 * H * 4 +
                      (size_t)2 * S * block * 4 + 16 + (geo ? (size_t)2 * (1 + L) * (block + 4) * 4 : 0);
  while (R > 1 && R * stage_bytes + tail > 220 * 1024) --R;
  const size_t smem = R * stage_bytes + tail;
  if (R * stage_bytes + tail > 220 * 1024) return 0;
  if (C > 256 || get_encode() == nullptr) return 0;   // (a TMA box holds at most 256 elements per dimension)
  CUtensorMap mXd;
  GOTEN_REQUIRE(staged::map3d(&mXd, Xd, C, N, L, (uint64_t)C * 4, (uint64_t)N * C * 4, C, 1, L),
                "cuTensorMapEncodeTiled failed (GATA backward, N=%d C=%d)", N, C);
  if (geo)
    STAGED_DISPATCH(gata_bwd_tgt_staged_geo_kernel, N, block, smem, mXd, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop,
                    alpha, tgt_ptr, src, N, C, H, R, max_deg_in, g_cols, g_qk, ldgqk, gZe, ldgz, da, gze_amax, g_fc, g_Y);
  else
    STAGED_DISPATCH(gata_bwd_tgt_staged_kernel, N, block, smem, mXd, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha,
                    tgt_ptr, src, N, C, H, R, max_deg_in, g_cols, g_qk, ldgqk, gZe, ldgz, da, gze_amax, g_fc, g_Y);
  GOTEN_CHECK_LAUNCH();
  *handled = true;
  return 0;
}

int gata_bwd_src_staged(const float* g_h, const float* g_Xd, const float* Xd, const float* qk, int ldqk, const float* x,
                        const float* v, const float* Ze, int ldz, const float* Y, const float* fc, const float* kappa, const float* drop,
                        const float* alpha, const float* da, const int32_t* src_ptr, const int32_t* src_perm,
                        const int32_t* tgt, int N, int C, int H, int lmax, int flags, float* g_qk, int ldgqk, float* g_x,
                        float* g_v, float* g_Xd_in, float* gx_amax, float* gv_amax, cudaStream_t st, bool* handled) {
  *handled = false;
  const int D = C / H;
  if (C % 4 != 0 || D % 4 != 0 || ldqk % 4 != 0 || ldz % 4 != 0 || ldgqk % 4 != 0) return 0;
  if (!staged::aligned16(g_h) || !staged::aligned16(g_Xd) || !staged::aligned16(Xd) || !staged::aligned16(qk) ||
      !staged::aligned16(x) || !staged::aligned16(v) || !staged::aligned16(Ze) || !staged::aligned16(g_qk) ||
      !staged::aligned16(g_x) || !staged::aligned16(g_v) || !staged::aligned16(g_Xd_in))
    return 0;
  const int L = (lmax + 1) * (lmax + 1) - 1, S = multiplier_of_(lmax, lmax > 1 ? flags : 0);
  const int NT = ((flags & 2) && lmax > 1) ? lmax : 1;
  const int block = ((C / 4 + 31) / 32) * 32;
  if (block > 1024) return 0;
  int R = staged::ring_depth();
  const size_t stage_bytes = (size_t)(S + 3 + L) * C * 4;
  const size_t tail = (size_t)8 * 8 + (size_t)(L + 2 * NT) * C * 4 + (size_t)staged::EC * (4 + L + 2 * H) * 4;
  while (R > 1 && R * stage_bytes + tail > 220 * 1024) --R;
  if (R * stage_bytes + tail > 220 * 1024) return 0;
  const size_t smem = R * stage_bytes + tail;
  if (C > 256 || get_encode() == nullptr) return 0;   // (a TMA box holds at most 256 elements per dimension)
  CUtensorMap mGX;
  GOTEN_REQUIRE(staged::map3d(&mGX, g_Xd, C, N, L, (uint64_t)C * 4, (uint64_t)N * C * 4, C, 1, L),
                "cuTensorMapEncodeTiled failed (GATA backward, N=%d C=%d)", N, C);
  STAGED_DISPATCH(gata_bwd_src_staged_kernel, N, block, smem, mGX, g_h, g_Xd, Xd, qk, ldqk, x, v, Ze, ldz, Y, fc, kappa, drop, alpha,
                  da, src_ptr, src_perm, tgt, N, C, H, R, g_qk, ldgqk, g_x, g_v, g_Xd_in, gx_amax, gv_amax);
  GOTEN_CHECK_LAUNCH();
  *handled = true;
  return 0;
}

}  // namespace goten
