// Optional pre-norms of the GATA block (reference representation/gotennet.py:306-315, :397-398):
//   layernorm       nn.LayerNorm(C) on the scalars h              -> goten_layernorm_fwd/bwd (init.cu kernels, act = 0)
//   steerable_norm  TensorLayerNorm on X (components/layers.py:1497-1563): per degree l and node, the channel norms
//                   d_c = |X^l[:, c]| are max-min normalised over the channels,
//                        out^l[m][c] = w_c * relu((d_c - min_c d) / (max_c d - min_c d)) * X^l[m][c] / d_c
//                   (d clamped at 1e-12; a zero range is replaced by 1).  The reference's global "all zero" early-out
//                   is a host synchronisation that returns what the formula gives anyway (0), so it is not replicated.
// One CTA per node, one thread per channel, X in degree-major layout [L][N][C]; the backward recomputes the forward
// statistics (same arithmetic, so the arg-max / arg-min channels agree) and routes the max / min gradients to the first
// maximal / minimal channel like torch.max / torch.min.
#include "common.cuh"

namespace goten {

constexpr float TLN_EPS = 1e-12f;

struct ValIdx {
  float v;
  int i;
};
__device__ __forceinline__ ValIdx pick_max(ValIdx a, ValIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }
__device__ __forceinline__ ValIdx pick_min(ValIdx a, ValIdx b) { return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

// block-wide (max, argmax) and (min, argmin) of one value per thread; inactive threads pass i = INT_MAX sentinels
__device__ __forceinline__ void block_maxmin(float v, int idx, bool valid, ValIdx& mx, ValIdx& mn, ValIdx* smem /*[2][32]*/) {
  ValIdx a{valid ? v : -INFINITY, valid ? idx : 0x7fffffff}, b{valid ? v : INFINITY, valid ? idx : 0x7fffffff};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx oa{__shfl_xor_sync(0xffffffffu, a.v, o), __shfl_xor_sync(0xffffffffu, a.i, o)};
    ValIdx ob{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
    a = pick_max(a, oa);
    b = pick_min(b, ob);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { smem[w] = a; smem[32 + w] = b; }
  __syncthreads();
  a = smem[0];
  b = smem[32];
  for (int k = 1; k < nw; ++k) { a = pick_max(a, smem[k]); b = pick_min(b, smem[32 + k]); }
  mx = a;
  mn = b;
}

template <int LMAX>
__global__ void tln_fwd_kernel(const float* __restrict__ Xd, const float* __restrict__ weight, int N, int C,
                               float* __restrict__ out) {
  __shared__ ValIdx red[64];
  const int n = blockIdx.x, c = threadIdx.x;
  const bool act = c < C;
  const float w = act ? weight[c] : 0.f;
#pragma unroll
  for (int l = 1; l <= LMAX; ++l) {
    const int lo = blk_lo(l), hi = blk_hi(l);
    float x[7];
    float ss = 0.f;
#pragma unroll
    for (int m = lo; m < hi; ++m) {
      x[m - lo] = act ? Xd[((size_t)m * N + n) * C + c] : 0.f;
      ss = fmaf(x[m - lo], x[m - lo], ss);
    }
    const float d = fmaxf(sqrtf(ss), TLN_EPS);
    ValIdx mx, mn;
    block_maxmin(d, c, act, mx, mn, red);
    float delta = mx.v - mn.v;
    if (delta == 0.f) delta = 1.f;
    const float s = fmaxf((d - mn.v) / delta, 0.f);
    if (act) {
      const float f = w * s / d;
#pragma unroll
      for (int m = lo; m < hi; ++m) out[((size_t)m * N + n) * C + c] = f * x[m - lo];
    }
  }
}

template <int LMAX>
__global__ void tln_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ Xd,
                               const float* __restrict__ weight, int N, int C, float* __restrict__ g_X) {
  __shared__ ValIdx red[64];
  __shared__ float sred[33];
  const int n = blockIdx.x, c = threadIdx.x;
  const bool act = c < C;
  const float w = act ? weight[c] : 0.f;
#pragma unroll
  for (int l = 1; l <= LMAX; ++l) {
    const int lo = blk_lo(l), hi = blk_hi(l);
    float x[7], g[7];
    float ss = 0.f, a = 0.f;
#pragma unroll
    for (int m = lo; m < hi; ++m) {
      const size_t o = ((size_t)m * N + n) * C + c;
      x[m - lo] = act ? Xd[o] : 0.f;
      g[m - lo] = act ? g_out[o] : 0.f;
      ss = fmaf(x[m - lo], x[m - lo], ss);
      a = fmaf(g[m - lo], x[m - lo], a);   // sum_m G X
    }
    const float dist = sqrtf(ss);
    const float d = fmaxf(dist, TLN_EPS);
    ValIdx mx, mn;
    block_maxmin(d, c, act, mx, mn, red);
    const float range = mx.v - mn.v;
    const float delta = range == 0.f ? 1.f : range;
    const float u = (d - mn.v) / delta;
    const float s = fmaxf(u, 0.f);
    const float r = u > 0.f ? 1.f : 0.f;
    const float b = w * a / d;                     // dL/ds
    const float S1 = block_sum(act ? b * r : 0.f, sred);
    const float S2 = block_sum(act ? b * r * u : 0.f, sred);
    const float g_delta = range == 0.f ? 0.f : -S2 / delta;
    const float g_mx = g_delta, g_mn = -S1 / delta - g_delta;
    float gd = -w * a * s / (d * d) + b * r / delta;
    if (c == mx.i) gd += g_mx;
    if (c == mn.i) gd += g_mn;
    const float g_dist = dist >= TLN_EPS ? gd : 0.f;     // clamp(min = eps) backward
    const float k2 = dist > 0.f ? g_dist / dist : 0.f;   // d|x|/dx = x / |x| (0 at the origin)
    if (act) {
      const float f = w * s / d;
#pragma unroll
      for (int m = lo; m < hi; ++m) g_X[((size_t)m * N + n) * C + c] = fmaf(f, g[m - lo], k2 * x[m - lo]);
    }
  }
}

}  // namespace goten

using namespace goten;

extern "C" {

int goten_tensor_layernorm_fwd(const float* Xd, const float* weight, int N, int C, int lmax, float* out, void* stream) {
  GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);
  GOTEN_REQUIRE(C >= 1 && C <= 1024, "n_atom_basis=%d unsupported by the tensor layer norm (<= 1024)", C);
  if (N == 0) return 0;
  const int T = ((C + 31) / 32) * 32;
  cudaStream_t st = as_stream(stream);
  if (lmax == 1) tln_fwd_kernel<1><<<N, T, 0, st>>>(Xd, weight, N, C, out);
  else if (lmax == 2) tln_fwd_kernel<2><<<N, T, 0, st>>>(Xd, weight, N, C, out);
  else tln_fwd_kernel<3><<<N, T, 0, st>>>(Xd, weight, N, C, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_tensor_layernorm_bwd(const float* g_out, const float* Xd, const float* weight, int N, int C, int lmax,
                               float* g_X, void* stream) {
  GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);
  GOTEN_REQUIRE(C >= 1 && C <= 1024, "n_atom_basis=%d unsupported by the tensor layer norm (<= 1024)", C);
  if (N == 0) return 0;
  const int T = ((C + 31) / 32) * 32;
  cudaStream_t st = as_stream(stream);
  if (lmax == 1) tln_bwd_kernel<1><<<N, T, 0, st>>>(g_out, Xd, weight, N, C, g_X);
  else if (lmax == 2) tln_bwd_kernel<2><<<N, T, 0, st>>>(g_out, Xd, weight, N, C, g_X);
  else tln_bwd_kernel<3><<<N, T, 0, st>>>(g_out, Xd, weight, N, C, g_X);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
