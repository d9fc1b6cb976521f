// FP32 SIMT GEMM (128x128x16 tiles, 8x8 register micro-tiles, double-buffered
// shared memory) with the epilogues the GotenNet path needs.  This is the
// exact-fp32 arm of goten_gemm (the reference runs every nn.Linear in strict
// fp32, scripts/train.py:16); the tcgen05 3xTF32 arm lives in gemm_tc.cu.
//
//   C[M][N] = opA(A) * opB(B) (+ bias[N]);   optional silu side output; optional
//   fused column sums of A^T (bias gradient inside the weight-gradient GEMM);
//   deterministic split-K (partials + ordered reduction, no atomics).
#include "common.cuh"

namespace goten {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;

struct Epilogue {
  const float* bias;
  const float* add_src;
  int ld_add;
  float* act_out;
  int ld_act, act_lo, act_hi;
};

__device__ __forceinline__ void store_out(float* __restrict__ C, int ldc, int m, int n, float v, const Epilogue& ep) {
  if (ep.bias) v += ep.bias[n];
  if (ep.add_src) v += ep.add_src[(size_t)m * ep.ld_add + n];
  C[(size_t)m * ldc + n] = v;
  if (ep.act_out && n >= ep.act_lo && n < ep.act_hi) ep.act_out[(size_t)m * ep.ld_act + (n - ep.act_lo)] = siluf_(v);
}

// guarded 4-wide load along the contiguous direction of a row-major matrix
__device__ __forceinline__ float4 load4(const float* __restrict__ P, int ld, int row, int col, int nrows, int ncols,
                                        bool vec_ok) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows) {
    const float* p = P + (size_t)row * ld + col;
    if (vec_ok && col + 3 < ncols) {
      r = *reinterpret_cast<const float4*>(p);
    } else {
      if (col + 0 < ncols) r.x = p[0];
      if (col + 1 < ncols) r.y = p[1];
      if (col + 2 < ncols) r.z = p[2];
      if (col + 3 < ncols) r.w = p[3];
    }
  }
  return r;
}

// TA=false: A[M][K]; TA=true: A[K][M].  TB=true: B[N][K]; TB=false: B[K][N].
template <bool TA, bool TB>
__global__ void __launch_bounds__(GT) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                   int ldb, float* __restrict__ C, int ldc, int M, int N, int K,
                                                   Epilogue ep, float* __restrict__ colsum, int k_len,
                                                   float* __restrict__ partial, float* __restrict__ partial_colsum,
                                                   bool a_vec, bool b_vec) {
  __shared__ __align__(16) float As[2][BK][BM + 4];  // +4: halves the transposed-store bank conflicts
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kz0 = blockIdx.z * k_len;
  const int kz1 = min(K, kz0 + k_len);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float csum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) csum[i] = 0.f;
  const bool do_colsum = TA && colsum != nullptr && blockIdx.x == 0;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * GT;
      if (!TA) {
        const int row = idx >> 2, kq = (idx & 3) * 4;
        ra[it] = load4(A, lda, m0 + row, k0 + kq, M, kz1, a_vec);
      } else {
        const int k = idx >> 5, mq = (idx & 31) * 4;
        ra[it] = load4(A, lda, k0 + k, m0 + mq, kz1, M, a_vec);
      }
      if (TB) {
        const int row = idx >> 2, kq = (idx & 3) * 4;
        rb[it] = load4(B, ldb, n0 + row, k0 + kq, N, kz1, b_vec);
      } else {
        const int k = idx >> 5, nq = (idx & 31) * 4;
        rb[it] = load4(B, ldb, k0 + k, n0 + nq, kz1, N, b_vec);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * GT;
      if (!TA) {
        const int row = idx >> 2, kq = (idx & 3) * 4;
        As[buf][kq + 0][row] = ra[it].x; As[buf][kq + 1][row] = ra[it].y;
        As[buf][kq + 2][row] = ra[it].z; As[buf][kq + 3][row] = ra[it].w;
      } else {
        const int k = idx >> 5, mq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][mq]) = ra[it];
      }
      if (TB) {
        const int row = idx >> 2, kq = (idx & 3) * 4;
        Bs[buf][kq + 0][row] = rb[it].x; Bs[buf][kq + 1][row] = rb[it].y;
        Bs[buf][kq + 2][row] = rb[it].z; Bs[buf][kq + 3][row] = rb[it].w;
      } else {
        const int k = idx >> 5, nq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][nq]) = rb[it];
      }
    }
  };

  const int nkt = (kz1 > kz0) ? (kz1 - kz0 + BK - 1) / BK : 0;
  if (nkt > 0) {
    gload(kz0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) gload(kz0 + (kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_colsum) {
#pragma unroll
        for (int i = 0; i < 8; ++i) csum[i] += a[i];
      }
    }
    if (kt + 1 < nkt) sstore(buf ^ 1);
    __syncthreads();
  }

  const bool split = partial != nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      if (split) partial[((size_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      else store_out(C, ldc, m, n, acc[i][j], ep);
    }
    if (do_colsum && tx == 0) {
      if (split) partial_colsum[(size_t)blockIdx.z * M + m] = csum[i];
      else colsum[m] = csum[i];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_colsum,
                                     int splits, float* __restrict__ C, int ldc, int M, int N, Epilogue ep,
                                     float* __restrict__ colsum) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t MN = (size_t)M * N;
  if (idx < MN) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * MN + idx];
    store_out(C, ldc, (int)(idx / N), (int)(idx % N), s, ep);
  }
  if (colsum && idx < (size_t)M) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial_colsum[(size_t)z * M + idx];
    colsum[idx] = s;
  }
}

static int pick_splits(int M, int N, int K, int trans_a) {
  // only the weight-gradient shape (huge reduction, small output) is split
  if (!trans_a) return 1;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int want = (148 * 3 + tiles - 1) / tiles;
  int max_by_k = (K + 4 * BK - 1) / (4 * BK);
  int s = want < max_by_k ? want : max_by_k;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

int gemm_simt(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc, int M,
              int N, int K, const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
              int act_hi, float* colsum,
              void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  GOTEN_REQUIRE(!(colsum && !trans_a), "colsum is only fused into the trans_a=1 GEMM");
  Epilogue ep{bias, add_src, ld_add, act_out, ld_act, act_lo, act_hi};
  int splits = pick_splits(M, N, K, trans_a);
  float* partial = nullptr;
  float* partial_cs = nullptr;
  if (splits > 1) {
    const int64_t need = (int64_t)splits * ((int64_t)M * N + M) * 4;
    if (workspace == nullptr || workspace_bytes < need) splits = 1;
    else {
      partial = reinterpret_cast<float*>(workspace);
      partial_cs = partial + (size_t)splits * M * N;
    }
  }
  int k_len = (K + splits - 1) / splits;
  k_len = ((k_len + BK - 1) / BK) * BK;
  if (k_len <= 0) k_len = BK;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  auto aligned = [](const void* p, int ld) { return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0); };
  const bool a_vec = aligned(A, lda), b_vec = aligned(B, ldb);
#define LAUNCH(TA_, TB_)                                                                                      \
  sgemm_kernel<TA_, TB_><<<grid, GT, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, ep, colsum, k_len, partial,    \
                                              partial_cs, a_vec, b_vec)
  if (!trans_a && trans_b) LAUNCH(false, true);
  else if (!trans_a && !trans_b) LAUNCH(false, false);
  else if (trans_a && !trans_b) LAUNCH(true, false);
  else LAUNCH(true, true);
#undef LAUNCH
  GOTEN_CHECK_LAUNCH();
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, partial_cs, splits, C, ldc, M, N,
                                                                          ep, colsum);
    GOTEN_CHECK_LAUNCH();
  }
  return 0;
}

int splitk_finish(const float* partial, const float* partial_cs, int splits, float* C, int ldc, int M, int N,
                  const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo, int act_hi,
                  float* colsum, cudaStream_t st) {
  Epilogue ep{bias, add_src, ld_add, act_out, ld_act, act_lo, act_hi};
  const size_t total = (size_t)M * N;
  splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, partial_cs, splits, C, ldc, M, N, ep,
                                                                        colsum);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int64_t gemm_simt_workspace_bytes(int M, int N, int K, int trans_a) {
  const int s = pick_splits(M, N, K, trans_a);
  return s > 1 ? (int64_t)s * ((int64_t)M * N + M) * 4 : 0;
}

// ------------------------------------------------------------- small helpers
__global__ void dsilu_mul_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ pre, int ldp,
                                 float* __restrict__ out, int ldo, int64_t M, int N, float* __restrict__ out_amax) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float amx = 0.f;
  if (idx < M * N) {
    const int64_t m = idx / N;
    const int n = (int)(idx % N);
    const float v = g[m * ldg + n] * dsiluf_(pre[m * ldp + n]);
    out[m * ldo + n] = v;
    amx = fabsf(v);
  }
  block_amax_commit(out_amax, amx);
}

constexpr int CS_ROWS = 256;  // rows per partial block
__global__ void colsum_partial_kernel(const float* __restrict__ A, int lda, int64_t M, int N, float* __restrict__ part) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = r0 + CS_ROWS < M ? r0 + CS_ROWS : M;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += A[r * lda + n];
  part[(size_t)blockIdx.y * N + n] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int nparts, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * N + n];
  out[n] = s;
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

__global__ void permute_nlc_kernel(const float* __restrict__ in, float* __restrict__ out, int n_nodes, int L, int C,
                                   int to_dm) {
  // to_dm: in[N][L][C] -> out[L][N][C]; else the inverse
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n_nodes * L * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int64_t r = idx / C;
  if (to_dm) {
    const int l = (int)(r % L);
    const int64_t n = r / L;
    out[((int64_t)l * n_nodes + n) * C + c] = in[idx];
  } else {
    const int64_t n = r % n_nodes;
    const int l = (int)(r / n_nodes);
    out[(n * L + l) * C + c] = in[idx];
  }
}

__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t n, int C,
                                     float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int64_t row = i / C;
  const int c = (int)(i % C);
  out[i] = table[idx[row] * C + c];
}

constexpr int EMB_ROWS = 64;  // nodes per partial block
// part[b][z][c] = sum over nodes of block b with idx == z; one thread per channel -> no conflicts
__global__ void embedding_bwd_partial_kernel(const float* __restrict__ g, const int64_t* __restrict__ idx, int64_t n,
                                             int C, int n_rows, float* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float* mine = part + (size_t)blockIdx.y * n_rows * C;
  const int64_t r0 = (int64_t)blockIdx.y * EMB_ROWS;
  const int64_t r1 = r0 + EMB_ROWS < n ? r0 + EMB_ROWS : n;
  for (int64_t r = r0; r < r1; ++r) mine[idx[r] * C + c] += g[r * C + c];
}
__global__ void embedding_bwd_final_kernel(const float* __restrict__ part, int nparts, int64_t total,
                                           float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * total + i];
  out[i] = s;
}

}  // namespace goten

using namespace goten;

namespace goten {
int gemm_tc(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc, int M,
            int N, int K, const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
              int act_hi, float* colsum,
            void* workspace, int64_t workspace_bytes, cudaStream_t st, bool* handled);
int64_t gemm_tc16_workspace_bytes(int M, int N, int K, int trans_a, int trans_b);
int gemm_tc16(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc, int M, int N,
              int K, const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
              int act_hi, float* colsum, const float* a_amax, const float* b_amax, void* workspace,
              int64_t workspace_bytes, cudaStream_t st, bool* handled);
int64_t gemm_tc_workspace_bytes(int M, int N, int K, int trans_a, int trans_b);
}  // namespace goten

extern "C" {

int64_t goten_gemm_workspace_bytes(int M, int N, int K, int trans_a, int trans_b) {
  int64_t a = gemm_simt_workspace_bytes(M, N, K, trans_a);
  int64_t b = gemm_tc_workspace_bytes(M, N, K, trans_a, trans_b);
  int64_t c = gemm_tc16_workspace_bytes(M, N, K, trans_a, trans_b);
  a = a > b ? a : b;
  return a > c ? a : c;
}

// impl: 0 = auto (split-fp16 tcgen05 -> 3xTF32 tcgen05 -> fp32 SIMT, first arm that accepts the shape),
//       1 = fp32 SIMT, 2 = 3xTF32 tcgen05, 3 = split-fp16 tcgen05.  GOTEN_TC16=0 removes the fp16 arm from auto.
int goten_gemm_scaled(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc,
                      int M, int N, int K, const float* bias, const float* add_src, int ld_add, float* act_out,
                      int ld_act, int act_lo, int act_hi, float* colsum, const float* a_amax, const float* b_amax,
                      void* workspace, int64_t workspace_bytes, int impl, void* stream) {
  cudaStream_t st = as_stream(stream);
  static int auto16 = -1;
  if (auto16 < 0) { const char* e = getenv("GOTEN_TC16"); auto16 = e ? atoi(e) : 1; }
  // negative impl = "prefer arm |impl|, fall back to the fp32 SIMT arm for shapes it declines" (GOTEN_GEMM=tc|tc16)
  const bool strict = impl > 0;
  if (impl < 0) impl = -impl;
  if (impl == 3 || (impl == 0 && auto16)) {
    bool handled = false;
    int rc = gemm_tc16(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, add_src, ld_add, act_out, ld_act,
                       act_lo, act_hi, colsum, a_amax, b_amax, workspace, workspace_bytes, st, &handled);
    if (rc) return rc;
    if (handled) return 0;
    GOTEN_REQUIRE(!strict, "tcgen05 fp16 GEMM does not support this shape/layout (M=%d N=%d K=%d ta=%d tb=%d)", M, N,
                  K, trans_a, trans_b);
  }
  if (impl == 0 || impl == 2) {
    bool handled = false;
    int rc = gemm_tc(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, add_src, ld_add, act_out, ld_act, act_lo,
                     act_hi, colsum, workspace, workspace_bytes, st, &handled);
    if (rc) return rc;
    if (handled) return 0;
    GOTEN_REQUIRE(!strict, "tcgen05 GEMM does not support this shape/layout (M=%d N=%d K=%d ta=%d tb=%d)", M, N, K,
                  trans_a, trans_b);
  }
  return gemm_simt(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, add_src, ld_add, act_out, ld_act, act_lo,
                   act_hi, colsum, workspace, workspace_bytes, st);
}

int goten_gemm(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc, int M,
               int N, int K, const float* bias, const float* add_src, int ld_add, float* act_out, int ld_act, int act_lo,
               int act_hi, float* colsum, void* workspace, int64_t workspace_bytes, int impl, void* stream) {
  return goten_gemm_scaled(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, add_src, ld_add, act_out, ld_act,
                           act_lo, act_hi, colsum, nullptr, nullptr, workspace, workspace_bytes, impl, stream);
}

int goten_dsilu_mul(const float* g, int ldg, const float* pre, int ldp, float* out, int ldo, int64_t M, int N,
                    float* out_amax, void* stream) {
  if (M * N == 0) return 0;
  dsilu_mul_kernel<<<(unsigned)cdiv64(M * N, 256), 256, 0, as_stream(stream)>>>(g, ldg, pre, ldp, out, ldo, M, N, out_amax);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_colsum(const float* A, int lda, int64_t M, int N, float* out, float* workspace, int64_t workspace_bytes,
                 void* stream) {
  cudaStream_t st = as_stream(stream);
  if (N == 0) return 0;
  const int nparts = (int)cdiv64(M > 0 ? M : 1, CS_ROWS);
  GOTEN_REQUIRE(workspace_bytes >= (int64_t)nparts * N * 4, "colsum workspace too small");
  dim3 grid((N + 127) / 128, nparts);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(A, lda, M, N, workspace);
  GOTEN_CHECK_LAUNCH();
  colsum_final_kernel<<<(N + 127) / 128, 128, 0, st>>>(workspace, nparts, N, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  add_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, as_stream(stream)>>>(a, b, out, n);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_permute_nlc(const float* in, float* out, int n_nodes, int L, int C, int to_degree_major, void* stream) {
  const int64_t total = (int64_t)n_nodes * L * C;
  if (total == 0) return 0;
  permute_nlc_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, as_stream(stream)>>>(in, out, n_nodes, L, C,
                                                                                 to_degree_major);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_embedding_fwd(const float* table, const int64_t* idx, int64_t n, int C, float* out, void* stream) {
  if (n * C == 0) return 0;
  embedding_fwd_kernel<<<(unsigned)cdiv64(n * C, 256), 256, 0, as_stream(stream)>>>(table, idx, n, C, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_embedding_bwd(const float* g_out, const int64_t* idx, int64_t n, int C, int n_rows, float* g_table,
                        float* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int nparts = (int)cdiv64(n > 0 ? n : 1, EMB_ROWS);
  const int64_t total = (int64_t)n_rows * C;
  GOTEN_REQUIRE(workspace_bytes >= (int64_t)nparts * total * 4, "embedding_bwd workspace too small");
  GOTEN_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)nparts * total * 4, st));
  if (n > 0) {
    dim3 grid((C + 127) / 128, nparts);
    embedding_bwd_partial_kernel<<<grid, 128, 0, st>>>(g_out, idx, n, C, n_rows, workspace);
    GOTEN_CHECK_LAUNCH();
  }
  embedding_bwd_final_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(workspace, nparts, total, g_table);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
