// Radius graph (bit-exact integers), CSR/CSC construction and per-edge geometry.
//
// Replaces torch_cluster.radius_graph as called from Distance.forward
// (reference components/layers.py:1588-1590) plus the geometry prologue of
// GotenNet.forward (representation/gotennet.py:974-989).
//
// Layout: edges are emitted sorted by (target, source): one thread per target
// scans the atoms of its own molecule in ascending index, so `tgt_ptr` is the
// CSR over targets for free.  The transposed view (edge ids grouped by source,
// ascending target) is built without any sort: one thread per source scans the
// same molecule and binary-searches itself in each target's (sorted) source
// list.  Everything is integer-deterministic; no float atomics anywhere.
#include <stdarg.h>

#include "common.cuh"

namespace goten {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

// ------------------------------------------------------------------ scan ----
constexpr int SCAN_T = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = SCAN_T * SCAN_ITEMS;

// exclusive scan of one int per thread across the block; returns exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sm /* >= 33 ints */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = (lane < nw) ? sm[lane] : 0;
    int sinc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, sinc, o);
      if (lane >= o) sinc += t;
    }
    if (lane < nw) sm[lane] = sinc - s;  // exclusive warp offsets
    if (lane == 31) sm[32] = sinc;       // block total
  }
  __syncthreads();
  *total = sm[32];
  return sm[w] + inc - v;
}

__global__ void scan_partial_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ block_sums) {
  __shared__ int sm[33];
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_excl_scan(s, &total, sm);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void scan_sums_kernel(int32_t* __restrict__ block_sums, int nb) {
  __shared__ int sm[33];
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += SCAN_T) {
    int i = b0 + threadIdx.x;
    int v = (i < nb) ? block_sums[i] : 0;
    int total;
    int ex = block_excl_scan(v, &total, sm);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[nb] = carry;
}

// out[i] = exclusive prefix of in[0..i), i in [0, n]; out[n] = total
__global__ void scan_final_kernel(const int32_t* __restrict__ in, int n, const int32_t* __restrict__ block_offs,
                                  int32_t* __restrict__ out) {
  __shared__ int sm[33];
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int ex = block_excl_scan(s, &total, sm) + block_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_offs[gridDim.x];
}

// scratch: >= cdiv(n, SCAN_CHUNK) + 1 ints.  in and out may NOT alias.
static int exclusive_scan(const int32_t* in, int n, int32_t* out, int32_t* scratch, cudaStream_t st) {
  if (n <= 0) {
    GOTEN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
    return 0;
  }
  const int nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  scan_partial_kernel<<<nb, SCAN_T, 0, st>>>(in, n, scratch);
  GOTEN_CHECK_LAUNCH();
  scan_sums_kernel<<<1, SCAN_T, 0, st>>>(scratch, nb);
  GOTEN_CHECK_LAUNCH();
  scan_final_kernel<<<nb, SCAN_T, 0, st>>>(in, n, scratch, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------ molecules -----
__global__ void mol_flags_kernel(const int64_t* __restrict__ batch, int n, int32_t* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || batch[i] != batch[i - 1]) ? 1 : 0;
}
// rank[i] = exclusive scan of flags -> molecule id of node i is rank[i] + flag[i] - 1
__global__ void mol_ptr_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ rank, int n,
                               int32_t* __restrict__ mol_of, int32_t* __restrict__ mol_ptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int m = rank[i] + flag[i] - 1;
    mol_of[i] = m;
    if (flag[i]) mol_ptr[m] = i;
  }
  if (i == n) mol_ptr[rank[n]] = n;  // rank[n] = number of molecules
}

// squared distance with the evaluation order of a plain fp32 loop (no FMA contraction):
// ((dx*dx + dy*dy) + dz*dz); (a-b)^2 == (b-a)^2 exactly, so target/source order is irrelevant.
__device__ __forceinline__ float dist2(const float* __restrict__ pos, int a, int b) {
  float dx = __fsub_rn(pos[3 * a + 0], pos[3 * b + 0]);
  float dy = __fsub_rn(pos[3 * a + 1], pos[3 * b + 1]);
  float dz = __fsub_rn(pos[3 * a + 2], pos[3 * b + 2]);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__global__ void rg_count_kernel(const float* __restrict__ pos, const int32_t* __restrict__ mol_ptr,
                                const int32_t* __restrict__ mol_of, int n, float r2, int K, int loop,
                                int32_t* __restrict__ deg_in) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = mol_of[i];
  const int a = mol_ptr[m], b = mol_ptr[m + 1];
  int cnt = 0;
  for (int j = a; j < b && cnt < K; ++j) {
    if (!loop && j == i) continue;
    if (dist2(pos, i, j) < r2) ++cnt;
  }
  deg_in[i] = cnt;
}

__global__ void rg_fill_kernel(const float* __restrict__ pos, const int32_t* __restrict__ mol_ptr,
                               const int32_t* __restrict__ mol_of, const int32_t* __restrict__ tgt_ptr, int n,
                               int64_t E, float r2, int K, int loop, int32_t* __restrict__ src,
                               int32_t* __restrict__ tgt, int64_t* __restrict__ edge_index,
                               int32_t* __restrict__ deg_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = mol_of[i];
  const int a = mol_ptr[m], b = mol_ptr[m + 1];
  const int base = tgt_ptr[i];
  int cnt = 0;
  for (int j = a; j < b && cnt < K; ++j) {
    if (!loop && j == i) continue;
    if (dist2(pos, i, j) < r2) {
      const int e = base + cnt;
      src[e] = j;
      tgt[e] = i;
      edge_index[e] = j;
      edge_index[E + e] = i;
      atomicAdd(&deg_out[j], 1);  // integer: result is order independent
      ++cnt;
    }
  }
}

// lower_bound of key in sorted src[lo, hi)
__device__ __forceinline__ int find_in(const int32_t* __restrict__ src, int lo, int hi, int key) {
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (src[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void rg_src_perm_kernel(const float* __restrict__ pos, const int32_t* __restrict__ mol_ptr,
                                   const int32_t* __restrict__ mol_of, const int32_t* __restrict__ tgt_ptr,
                                   const int32_t* __restrict__ src, const int32_t* __restrict__ src_ptr, int n,
                                   float r2, int loop, int32_t* __restrict__ src_perm) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int m = mol_of[j];
  const int a = mol_ptr[m], b = mol_ptr[m + 1];
  int out = src_ptr[j];
  for (int i = a; i < b; ++i) {
    if (!loop && i == j) continue;
    if (dist2(pos, i, j) < r2) {
      const int lo = tgt_ptr[i], hi = tgt_ptr[i + 1];
      const int e = find_in(src, lo, hi, j);
      if (e < hi && src[e] == j) src_perm[out++] = e;  // survives the first-K truncation of target i
    }
  }
}

// ---------------------------------------------------- external edge lists ---
__global__ void count_by_key_kernel(const int32_t* __restrict__ key, int64_t E, int32_t* __restrict__ cnt) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) atomicAdd(&cnt[key[e]], 1);
}
__global__ void copy_i32_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = in[e];
}

// --------------------------------------------------------------- geometry ---
template <int LMAX>
__device__ __forceinline__ void sph_harm(float x, float y, float z, float* Y) {
  // reference components/layers.py:822-869 (degrees 1..3, no l=0 term)
  Y[0] = x; Y[1] = y; Y[2] = z;
  if (LMAX >= 2) {
    const float s3 = 1.7320508075688772f;
    const float y2 = y * y, x2z2 = x * x + z * z;
    const float s20 = s3 * x * z, s21 = s3 * x * y, s22 = y2 - 0.5f * x2z2, s23 = s3 * y * z,
                s24 = (s3 / 2.0f) * (z * z - x * x);
    Y[3] = s20; Y[4] = s21; Y[5] = s22; Y[6] = s23; Y[7] = s24;
    if (LMAX >= 3) {
      const float c42 = 1.0801234497346435f /* sqrt(42)/6 */, c7 = 2.6457513110645907f,
                  c168 = 1.6201851746019651f /* sqrt(168)/8 */;
      Y[8] = c42 * (s20 * z + s24 * x);
      Y[9] = c7 * s20 * y;
      Y[10] = c168 * (4.0f * y2 - x2z2) * x;
      Y[11] = 0.5f * c7 * y * (2.0f * y2 - 3.0f * x2z2);
      Y[12] = c168 * z * (4.0f * y2 - x2z2);
      Y[13] = c7 * s24 * y;
      Y[14] = c42 * (s24 * z - s20 * x);
    }
  }
}

template <int LMAX>
__global__ void edge_geometry_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ vec_in,
                                         const float* __restrict__ r_in, const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                         const int32_t* __restrict__ deg_out, int64_t E, float rc, int R, int basis,
                                         const float* __restrict__ means, const float* __restrict__ betas,
                                         int scale_edge, float inv_sqrt_c, float* __restrict__ r_out,
                                         float* __restrict__ u_out, float* __restrict__ Y_out,
                                         float* __restrict__ fc_out, float* __restrict__ kappa_out,
                                         float* __restrict__ phi_out) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int j = src[e], i = tgt[e];
  float vx, vy, vz;
  if (vec_in) {
    vx = vec_in[3 * e + 0]; vy = vec_in[3 * e + 1]; vz = vec_in[3 * e + 2];
  } else {
    vx = pos[3 * j + 0] - pos[3 * i + 0];
    vy = pos[3 * j + 1] - pos[3 * i + 1];
    vz = pos[3 * j + 2] - pos[3 * i + 2];
  }
  float r = 0.f, ux = vx, uy = vy, uz = vz;
  if (i != j) {  // self loops keep r = 0 and the raw (zero) vector, gotennet.py:978-980
    r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    ux = vx / r; uy = vy / r; uz = vz / r;
  }
  if (r_in) r = r_in[e];
  r_out[e] = r;
  u_out[3 * e + 0] = ux; u_out[3 * e + 1] = uy; u_out[3 * e + 2] = uz;
  float Y[L];
  sph_harm<LMAX>(ux, uy, uz, Y);
#pragma unroll
  for (int m = 0; m < L; ++m) Y_out[e * L + m] = Y[m];
  const float fc = (r < rc) ? 0.5f * (cosf(r * 3.14159265358979323846f / rc) + 1.0f) : 0.f;
  fc_out[e] = fc;
  kappa_out[e] = scale_edge ? sqrtf((float)deg_out[j]) * inv_sqrt_c : inv_sqrt_c;
  if (basis == 0) {         // ExpNormalSmearing (layers.py:744-746): cutoff included; means / betas
    const float ex = expf((5.0f / rc) * (-r));
    for (int k = 0; k < R; ++k) {
      const float d = ex - means[k];
      phi_out[e * R + k] = fc * expf(-betas[k] * d * d);
    }
  } else if (basis == 1) {  // BesselBasis (layers.py:349-358): sin(freq r) / r, r = 0 -> divisor 1; means = freqs
    const float inv = 1.0f / (r == 0.f ? 1.0f : r);
    for (int k = 0; k < R; ++k) phi_out[e * R + k] = sinf(means[k] * r) * inv;
  } else {                  // GaussianRBF (layers.py:276-291): means = offsets, betas = widths
    for (int k = 0; k < R; ++k) {
      const float d = r - means[k];
      phi_out[e * R + k] = expf((-0.5f / (betas[k] * betas[k])) * d * d);
    }
  }
}

// d(loss)/d(edge vector).  Self loops (r = 0) receive no gradient (the reference masks them out
// before the norm, layers.py:1598-1600 / gotennet.py:978-980).
template <int LMAX>
__global__ void edge_geometry_bwd_kernel(const float* __restrict__ r_in, const float* __restrict__ u_in,
                                         const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                         int64_t E, float rc, int R, int basis, const float* __restrict__ means,
                                         const float* __restrict__ betas, const float* __restrict__ g_phi,
                                         const float* __restrict__ g_fc, const float* __restrict__ g_Y,
                                         const float* __restrict__ g_r_in, float* __restrict__ g_vec) {
  constexpr int L = (LMAX + 1) * (LMAX + 1) - 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  if (src[e] == tgt[e]) {
    g_vec[3 * e + 0] = 0.f; g_vec[3 * e + 1] = 0.f; g_vec[3 * e + 2] = 0.f;
    return;
  }
  const float r = r_in[e];
  const float x = u_in[3 * e + 0], y = u_in[3 * e + 1], z = u_in[3 * e + 2];
  // ---- radial part: dL/dr
  const float pi_rc = 3.14159265358979323846f / rc;
  const bool inside = r < rc;
  const float fc = inside ? 0.5f * (cosf(r * pi_rc) + 1.0f) : 0.f;
  const float dfc = inside ? -0.5f * sinf(r * pi_rc) * pi_rc : 0.f;
  float g_r = (g_fc ? g_fc[e] * dfc : 0.f) + (g_r_in ? g_r_in[e] : 0.f);   // g_r_in: gradient of the returned distance itself
  if (g_phi && basis == 0) {
    const float alpha = 5.0f / rc;
    const float ex = expf(-alpha * r);
    for (int k = 0; k < R; ++k) {
      const float d = ex - means[k];
      const float gk = expf(-betas[k] * d * d);
      // phi = fc * gk ; dgk/dr = gk * (-2 beta d) * (-alpha ex)
      g_r += g_phi[e * R + k] * (dfc * gk + fc * gk * (2.0f * betas[k] * d * alpha * ex));
    }
  } else if (g_phi && basis == 1) {  // d/dr [sin(a r) / r] = (a r cos(a r) - sin(a r)) / r^2   (r > 0 on non-loop edges)
    const float inv = 1.0f / r;
    for (int k = 0; k < R; ++k) {
      const float ar = means[k] * r;
      g_r += g_phi[e * R + k] * (ar * cosf(ar) - sinf(ar)) * inv * inv;
    }
  } else if (g_phi) {                // d/dr exp(c (r - o)^2) = 2 c (r - o) phi
    for (int k = 0; k < R; ++k) {
      const float c = -0.5f / (betas[k] * betas[k]), d = r - means[k];
      g_r += g_phi[e * R + k] * 2.0f * c * d * expf(c * d * d);
    }
  }
  // ---- angular part: dL/du through the harmonics
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (g_Y) {
    const float* g = g_Y + e * L;
    gx += g[0]; gy += g[1]; gz += g[2];
    if (LMAX >= 2) {
      const float s3 = 1.7320508075688772f;
      const float y2 = y * y, x2z2 = x * x + z * z;
      const float s20 = s3 * x * z, s24 = (s3 / 2.0f) * (z * z - x * x);
      float g20 = g[3], g21 = g[4], g22 = g[5], g23 = g[6], g24 = g[7];
      float gy2 = 0.f, gx2z2 = 0.f;  // gradients flowing into y2 and x2z2
      if (LMAX >= 3) {
        const float c42 = 1.0801234497346435f, c7 = 2.6457513110645907f, c168 = 1.6201851746019651f;
        const float q = 4.0f * y2 - x2z2;
        // sh_3_0 = c42 (s20 z + s24 x)
        g20 += g[8] * c42 * z; g24 += g[8] * c42 * x; gz += g[8] * c42 * s20; gx += g[8] * c42 * s24;
        // sh_3_1 = c7 s20 y
        g20 += g[9] * c7 * y; gy += g[9] * c7 * s20;
        // sh_3_2 = c168 q x
        gx += g[10] * c168 * q; gy2 += g[10] * c168 * x * 4.0f; gx2z2 -= g[10] * c168 * x;
        // sh_3_3 = 0.5 c7 y (2 y2 - 3 x2z2)
        gy += g[11] * 0.5f * c7 * (2.0f * y2 - 3.0f * x2z2);
        gy2 += g[11] * 0.5f * c7 * y * 2.0f; gx2z2 -= g[11] * 0.5f * c7 * y * 3.0f;
        // sh_3_4 = c168 z q
        gz += g[12] * c168 * q; gy2 += g[12] * c168 * z * 4.0f; gx2z2 -= g[12] * c168 * z;
        // sh_3_5 = c7 s24 y
        g24 += g[13] * c7 * y; gy += g[13] * c7 * s24;
        // sh_3_6 = c42 (s24 z - s20 x)
        g24 += g[14] * c42 * z; gz += g[14] * c42 * s24; g20 -= g[14] * c42 * x; gx -= g[14] * c42 * s20;
      }
      // s20 = s3 x z ; s21 = s3 x y ; s22 = y2 - 0.5 x2z2 ; s23 = s3 y z ; s24 = s3/2 (z^2 - x^2)
      gx += g20 * s3 * z; gz += g20 * s3 * x;
      gx += g21 * s3 * y; gy += g21 * s3 * x;
      gy2 += g22; gx2z2 -= 0.5f * g22;
      gy += g23 * s3 * z; gz += g23 * s3 * y;
      gz += g24 * s3 * z; gx -= g24 * s3 * x;
      gy += gy2 * 2.0f * y;
      gx += gx2z2 * 2.0f * x; gz += gx2z2 * 2.0f * z;
    }
  }
  // u = v / r  ->  dL/dv = (g_u - (g_u . u) u) / r + g_r u
  const float dot = gx * x + gy * y + gz * z;
  const float inv_r = 1.0f / r;
  g_vec[3 * e + 0] = (gx - dot * x) * inv_r + g_r * x;
  g_vec[3 * e + 1] = (gy - dot * y) * inv_r + g_r * y;
  g_vec[3 * e + 2] = (gz - dot * z) * inv_r + g_r * z;
}

// edge_vec = pos[src] - pos[tgt]  ->  g_pos[n] = sum_{e: src=n} g_vec[e] - sum_{e: tgt=n} g_vec[e]
__global__ void edge_vec_to_pos_kernel(const float* __restrict__ g_vec, const int32_t* __restrict__ tgt_ptr,
                                       const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_perm,
                                       int n, float* __restrict__ g_pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int p = src_ptr[i]; p < src_ptr[i + 1]; ++p) {
    const int e = src_perm[p];
    ax += g_vec[3 * (int64_t)e + 0]; ay += g_vec[3 * (int64_t)e + 1]; az += g_vec[3 * (int64_t)e + 2];
  }
  for (int e = tgt_ptr[i]; e < tgt_ptr[i + 1]; ++e) {
    ax -= g_vec[3 * (int64_t)e + 0]; ay -= g_vec[3 * (int64_t)e + 1]; az -= g_vec[3 * (int64_t)e + 2];
  }
  g_pos[3 * i + 0] = ax; g_pos[3 * i + 1] = ay; g_pos[3 * i + 2] = az;
}

}  // namespace goten

using namespace goten;

extern "C" {

int goten_abi_version(void) { return GOTEN_ABI_VERSION; }
const char* goten_last_error(void) { return goten::g_err; }
int64_t goten_launch_count(void) { return (int64_t)goten::g_launches; }

int goten_device_info(int* out3) {
  int dev = 0;
  GOTEN_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GOTEN_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  out3[0] = p.multiProcessorCount;
  out3[1] = (int)p.sharedMemPerBlockOptin;
  out3[2] = p.major * 10 + p.minor;
  return 0;
}

int goten_radius_graph_count(const float* pos, const int64_t* batch, int n, float cutoff, int K, int loop,
                             int32_t* mol_ptr, int32_t* mol_of, int32_t* tgt_ptr, int32_t* scratch,
                             int64_t* n_edges_out, int32_t* n_mol_out, void* stream) {
  cudaStream_t st = as_stream(stream);
  *n_edges_out = 0;
  *n_mol_out = 0;
  if (n == 0) {
    GOTEN_CHECK_CUDA(cudaMemsetAsync(tgt_ptr, 0, sizeof(int32_t), st));
    GOTEN_CHECK_CUDA(cudaMemsetAsync(mol_ptr, 0, sizeof(int32_t), st));
    return 0;
  }
  GOTEN_REQUIRE(K >= 1, "max_num_neighbors must be >= 1");
  int32_t* flag = scratch;            // [n]
  int32_t* rank = scratch + n;        // [n+1]
  int32_t* sc = scratch + 2 * n + 1;  // scan scratch
  const int T = 256, nb = (n + T) / T;  // n+1 threads
  mol_flags_kernel<<<nb, T, 0, st>>>(batch, n, flag);
  GOTEN_CHECK_LAUNCH();
  if (exclusive_scan(flag, n, rank, sc, st)) return 1;
  mol_ptr_kernel<<<nb, T, 0, st>>>(flag, rank, n, mol_of, mol_ptr);
  GOTEN_CHECK_LAUNCH();
  int32_t* deg_in = flag;  // reuse
  const float r2 = cutoff * cutoff;
  rg_count_kernel<<<(n + 127) / 128, 128, 0, st>>>(pos, mol_ptr, mol_of, n, r2, K, loop, deg_in);
  GOTEN_CHECK_LAUNCH();
  if (exclusive_scan(deg_in, n, tgt_ptr, sc, st)) return 1;
  int32_t hE = 0, hM = 0;
  GOTEN_CHECK_CUDA(cudaMemcpyAsync(&hE, tgt_ptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GOTEN_CHECK_CUDA(cudaMemcpyAsync(&hM, rank + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GOTEN_CHECK_CUDA(cudaStreamSynchronize(st));
  *n_edges_out = hE;
  *n_mol_out = hM;
  return 0;
}

int goten_radius_graph_fill(const float* pos, const int32_t* mol_ptr, const int32_t* mol_of,
                            const int32_t* tgt_ptr, int n, int64_t E, float cutoff, int K, int loop,
                            int32_t* src, int32_t* tgt, int64_t* edge_index, int32_t* deg_out,
                            int32_t* src_ptr, int32_t* src_perm, int32_t* scratch, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    GOTEN_CHECK_CUDA(cudaMemsetAsync(src_ptr, 0, sizeof(int32_t), st));
    return 0;
  }
  const float r2 = cutoff * cutoff;
  GOTEN_CHECK_CUDA(cudaMemsetAsync(deg_out, 0, sizeof(int32_t) * n, st));
  rg_fill_kernel<<<(n + 127) / 128, 128, 0, st>>>(pos, mol_ptr, mol_of, tgt_ptr, n, E, r2, K, loop, src, tgt,
                                                  edge_index, deg_out);
  GOTEN_CHECK_LAUNCH();
  if (exclusive_scan(deg_out, n, src_ptr, scratch, st)) return 1;
  rg_src_perm_kernel<<<(n + 127) / 128, 128, 0, st>>>(pos, mol_ptr, mol_of, tgt_ptr, src, src_ptr, n, r2, loop,
                                                      src_perm);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_csr_from_sorted(const int32_t* src, const int32_t* tgt, const int32_t* order_by_src, int n, int64_t E,
                          int32_t* tgt_ptr, int32_t* deg_out, int32_t* src_ptr, int32_t* src_perm,
                          int32_t* scratch, void* stream) {
  cudaStream_t st = as_stream(stream);
  int32_t* cnt = scratch;         // [n]
  int32_t* sc = scratch + n + 1;  // scan scratch
  const int T = 256;
  const unsigned nbE = (unsigned)cdiv64(E > 0 ? E : 1, T);
  GOTEN_CHECK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (n > 0 ? n : 1), st));
  if (E > 0) { count_by_key_kernel<<<nbE, T, 0, st>>>(tgt, E, cnt); GOTEN_CHECK_LAUNCH(); }
  if (exclusive_scan(cnt, n, tgt_ptr, sc, st)) return 1;
  GOTEN_CHECK_CUDA(cudaMemsetAsync(deg_out, 0, sizeof(int32_t) * (n > 0 ? n : 1), st));
  if (E > 0) { count_by_key_kernel<<<nbE, T, 0, st>>>(src, E, deg_out); GOTEN_CHECK_LAUNCH(); }
  if (exclusive_scan(deg_out, n, src_ptr, sc, st)) return 1;
  if (E > 0) { copy_i32_kernel<<<nbE, T, 0, st>>>(order_by_src, E, src_perm); GOTEN_CHECK_LAUNCH(); }
  return 0;
}

int goten_edge_geometry_fwd(const float* pos, const float* edge_vec_in, const float* r_in, const int32_t* src,
                            const int32_t* tgt,
                            const int32_t* deg_out, int64_t E, int lmax, float cutoff, int n_rbf, int basis,
                            const float* means, const float* betas, int scale_edge, int C, float* r, float* u,
                            float* Y, float* fc, float* kappa, float* phi, void* stream) {
  if (E == 0) return 0;
  GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);
  cudaStream_t st = as_stream(stream);
  const int T = 128;
  const unsigned nb = (unsigned)cdiv64(E, T);
  const float isc = 1.0f / sqrtf((float)C);
#define LAUNCH(LM)                                                                                              \
  edge_geometry_fwd_kernel<LM><<<nb, T, 0, st>>>(pos, edge_vec_in, r_in, src, tgt, deg_out, E, cutoff, n_rbf, basis, means,   \
                                                betas, scale_edge, isc, r, u, Y, fc, kappa, phi)
  if (lmax == 1) LAUNCH(1); else if (lmax == 2) LAUNCH(2); else LAUNCH(3);
#undef LAUNCH
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_edge_geometry_bwd(const float* r, const float* u, const int32_t* src, const int32_t* tgt, int64_t E,
                            int lmax, float cutoff, int n_rbf, int basis, const float* means, const float* betas,
                            const float* g_phi, const float* g_fc, const float* g_Y, const float* g_r, float* g_vec,
                            void* stream) {
  if (E == 0) return 0;
  GOTEN_REQUIRE(lmax >= 1 && lmax <= 3, "lmax=%d unsupported (1..3)", lmax);
  cudaStream_t st = as_stream(stream);
  const int T = 128;
  const unsigned nb = (unsigned)cdiv64(E, T);
#define LAUNCH(LM)                                                                                        \
  edge_geometry_bwd_kernel<LM><<<nb, T, 0, st>>>(r, u, src, tgt, E, cutoff, n_rbf, basis, means, betas, g_phi,    \
                                                g_fc, g_Y, g_r, g_vec)
  if (lmax == 1) LAUNCH(1); else if (lmax == 2) LAUNCH(2); else LAUNCH(3);
#undef LAUNCH
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_edge_vec_to_pos_bwd(const float* g_vec, const int32_t* tgt_ptr, const int32_t* src_ptr,
                              const int32_t* src_perm, int n, float* g_pos, void* stream) {
  if (n == 0) return 0;
  edge_vec_to_pos_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(g_vec, tgt_ptr, src_ptr, src_perm, n,
                                                                         g_pos);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
