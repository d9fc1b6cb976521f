// Atom-wise read-out head pieces (reference models/components/outputs.py:323-376 Atomwise.forward,
// components/layers.py:172-202 ScaleShift, :69-81 shifted_softplus): standardisation + single-atom reference +
// per-molecule segment reduction, and the element-wise activations of the head's MLP.
// Molecules are contiguous runs of atoms (PyG batches are sorted by graph id), so the scatter of the reference
// (torch_scatter, atomics) becomes one deterministic segmented sum per molecule.
#include "common.cuh"

namespace goten {

// mol_ptr[m] = first atom with batch id >= m  (binary search over the sorted batch vector), m = 0..n_mol
__global__ void mol_ptr_kernel(const int64_t* __restrict__ batch, int n_nodes, int n_mol, int32_t* __restrict__ mol_ptr) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > n_mol) return;
  int lo = 0, hi = n_nodes;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (batch[mid] < (int64_t)m) lo = mid + 1; else hi = mid;
  }
  mol_ptr[m] = lo;
}

// sortedness check of the batch vector: flag[0] = 1 if any batch[i] < batch[i-1]
__global__ void batch_sorted_kernel(const int64_t* __restrict__ batch, int n_nodes, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1 && i < n_nodes && batch[i] < batch[i - 1]) flag[0] = 1;
}

// yi[n][o] = raw[n][o] * stddev + mean (+ atomref[z[n]][o]);   y[m][o] = sum / mean over the atoms of molecule m
// one warp per (molecule, output column)
__global__ void atomwise_reduce_fwd_kernel(const float* __restrict__ raw, const int64_t* __restrict__ z,
                                           const float* __restrict__ atomref, int atomref_rows,
                                           const float* __restrict__ mean, const float* __restrict__ stddev, int n_stat,
                                           const int32_t* __restrict__ mol_ptr, int n_mol, int n_out, int mode,
                                           float* __restrict__ yi, float* __restrict__ y) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_mol * n_out) return;
  const int m = w / n_out, o = w - m * n_out;
  const float sd = stddev ? stddev[n_stat > 1 ? o : 0] : 1.f, mu = mean ? mean[n_stat > 1 ? o : 0] : 0.f;
  const int a0 = mol_ptr[m], a1 = mol_ptr[m + 1];
  float acc = 0.f;
  for (int n = a0 + lane; n < a1; n += 32) {
    float v = raw[(size_t)n * n_out + o] * sd + mu;
    if (atomref) {
      int64_t zz = z[n];
      zz = zz < 0 ? 0 : (zz >= atomref_rows ? atomref_rows - 1 : zz);
      v += atomref[(size_t)zz * n_out + o];
    }
    yi[(size_t)n * n_out + o] = v;
    acc += v;
  }
  acc = warp_sum(acc);
  if (lane == 0 && y) y[(size_t)m * n_out + o] = (mode == 2 && a1 > a0) ? acc / (float)(a1 - a0) : acc;
}

// no aggregation: element-wise only
__global__ void atomwise_scale_kernel(const float* __restrict__ raw, const int64_t* __restrict__ z,
                                      const float* __restrict__ atomref, int atomref_rows, const float* __restrict__ mean,
                                      const float* __restrict__ stddev, int n_stat, int64_t n_nodes, int n_out,
                                      float* __restrict__ yi) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_nodes * n_out) return;
  const int64_t n = idx / n_out;
  const int o = (int)(idx - n * n_out);
  float v = raw[idx] * (stddev ? stddev[n_stat > 1 ? o : 0] : 1.f) + (mean ? mean[n_stat > 1 ? o : 0] : 0.f);
  if (atomref) {
    int64_t zz = z[n];
    zz = zz < 0 ? 0 : (zz >= atomref_rows ? atomref_rows - 1 : zz);
    v += atomref[(size_t)zz * n_out + o];
  }
  yi[idx] = v;
}

// g_raw[n][o] = (g_y[mol(n)][o] (/count) + g_yi[n][o]) * stddev
__global__ void atomwise_reduce_bwd_kernel(const float* __restrict__ g_y, const float* __restrict__ g_yi,
                                           const float* __restrict__ stddev, int n_stat,
                                           const int32_t* __restrict__ mol_ptr, int n_mol, int n_out, int mode,
                                           float* __restrict__ g_raw) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_mol * n_out) return;
  const int m = w / n_out, o = w - m * n_out;
  const float sd = stddev ? stddev[n_stat > 1 ? o : 0] : 1.f;
  const int a0 = mol_ptr[m], a1 = mol_ptr[m + 1];
  float gm = g_y ? g_y[(size_t)m * n_out + o] : 0.f;
  if (mode == 2 && a1 > a0) gm /= (float)(a1 - a0);
  for (int n = a0 + lane; n < a1; n += 32) {
    const float gi = g_yi ? g_yi[(size_t)n * n_out + o] : 0.f;
    g_raw[(size_t)n * n_out + o] = (gm + gi) * sd;
  }
}

// mode 0 (no aggregation): g_raw = (g_y + g_yi) * stddev, element-wise
__global__ void atomwise_scale_bwd_kernel(const float* __restrict__ g_y, const float* __restrict__ g_yi,
                                          const float* __restrict__ stddev, int n_stat, int64_t total, int n_out,
                                          float* __restrict__ g_raw) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o = (int)(idx % n_out);
  const float g = (g_y ? g_y[idx] : 0.f) + (g_yi ? g_yi[idx] : 0.f);
  g_raw[idx] = g * (stddev ? stddev[n_stat > 1 ? o : 0] : 1.f);
}

// element-wise activations (head MLP; gamma_w of the "linw" edge updates): kind 1 = SiLU, 2 = shifted softplus
// ln(1 + e^x) - ln 2, 3 = sigmoid, 4 = tanh
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch threshold = 20
__device__ __forceinline__ float act_value(int kind, float v) {
  if (kind == 1) return siluf_(v);
  if (kind == 2) return softplusf_(v) - 0.69314718055994530942f;
  if (kind == 3) return sigmoidf_(v);
  return tanhf(v);
}
__device__ __forceinline__ float act_slope(int kind, float v) {
  if (kind == 1) return dsiluf_(v);
  if (kind == 2) return v > 20.f ? 1.f : sigmoidf_(v);
  if (kind == 3) { const float s = sigmoidf_(v); return s * (1.0f - s); }
  const float t = tanhf(v);
  return 1.0f - t * t;
}
__global__ void act_fwd_kernel(int kind, const float* __restrict__ x, int64_t n, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = act_value(kind, x[i]);
}
__global__ void act_bwd_kernel(int kind, const float* __restrict__ g, const float* __restrict__ x, int64_t n,
                               float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = g[i] * act_slope(kind, x[i]);
}

// t' = a * b + c and its gradients g_a = g * b, g_b = g * a (g_c = g): the residual edge update
// t_ij + gamma_t(t_ij) * gamma_w(w_ij) (gotennet.py:611, :445) when gamma_w is a host-composed network
__global__ void mul_add_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                   int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = fmaf(a[i], b[i], c[i]);
}
__global__ void mul_add_bwd_kernel(const float* __restrict__ g, const float* __restrict__ a, const float* __restrict__ b,
                                   int64_t n, float* __restrict__ g_a, float* __restrict__ g_b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  g_a[i] = gi * b[i];
  g_b[i] = gi * a[i];
}

}  // namespace goten

using namespace goten;

extern "C" {

int goten_mol_ptr(const int64_t* batch, int n_nodes, int n_mol, int32_t* mol_ptr, int32_t* unsorted_flag, void* stream) {
  GOTEN_REQUIRE(n_nodes >= 0 && n_mol >= 0, "bad sizes n_nodes=%d n_mol=%d", n_nodes, n_mol);
  cudaStream_t st = as_stream(stream);
  if (unsorted_flag) {
    GOTEN_CHECK_CUDA(cudaMemsetAsync(unsorted_flag, 0, sizeof(int32_t), st));
    if (n_nodes > 1) {
      batch_sorted_kernel<<<(n_nodes + 255) / 256, 256, 0, st>>>(batch, n_nodes, unsorted_flag);
      GOTEN_CHECK_LAUNCH();
    }
  }
  mol_ptr_kernel<<<(n_mol + 1 + 255) / 256, 256, 0, st>>>(batch, n_nodes, n_mol, mol_ptr);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_atomwise_reduce_fwd(const float* raw, const int64_t* z, const float* atomref, int atomref_rows,
                              const float* mean, const float* stddev, int n_stat, const int32_t* mol_ptr, int n_nodes,
                              int n_mol, int n_out, int mode, float* yi, float* y, void* stream) {
  GOTEN_REQUIRE(n_out >= 1 && (n_stat == 1 || n_stat == n_out), "mean/stddev must have 1 or n_out=%d entries", n_out);
  GOTEN_REQUIRE(mode >= 0 && mode <= 2, "aggregation mode %d unsupported (0 none, 1 sum, 2 mean)", mode);
  GOTEN_REQUIRE(atomref == nullptr || z != nullptr, "atomref needs atomic numbers");
  cudaStream_t st = as_stream(stream);
  if (n_nodes == 0) return 0;
  if (mode == 0) {
    const int64_t tot = (int64_t)n_nodes * n_out;
    atomwise_scale_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(raw, z, atomref, atomref_rows, mean, stddev, n_stat,
                                                                     n_nodes, n_out, yi);
  } else {
    const int64_t warps = (int64_t)n_mol * n_out;
    if (warps == 0) return 0;
    atomwise_reduce_fwd_kernel<<<(unsigned)cdiv64(warps * 32, 256), 256, 0, st>>>(raw, z, atomref, atomref_rows, mean,
                                                                                 stddev, n_stat, mol_ptr, n_mol, n_out,
                                                                                 mode, yi, y);
  }
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_atomwise_reduce_bwd(const float* g_y, const float* g_yi, const float* stddev, int n_stat,
                              const int32_t* mol_ptr, int n_nodes, int n_mol, int n_out, int mode, float* g_raw,
                              void* stream) {
  GOTEN_REQUIRE(mode >= 0 && mode <= 2, "aggregation mode %d unsupported in the backward", mode);
  cudaStream_t st = as_stream(stream);
  if (mode == 0) {   // no aggregation: y is yi, both gradients add element-wise
    const int64_t tot = (int64_t)n_nodes * n_out;
    if (tot == 0) return 0;
    atomwise_scale_bwd_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(g_y, g_yi, stddev, n_stat, tot, n_out, g_raw);
    GOTEN_CHECK_LAUNCH();
    return 0;
  }
  const int64_t warps = (int64_t)n_mol * n_out;
  if (n_nodes == 0 || warps == 0) return 0;
  atomwise_reduce_bwd_kernel<<<(unsigned)cdiv64(warps * 32, 256), 256, 0, st>>>(g_y, g_yi, stddev, n_stat, mol_ptr, n_mol,
                                                                               n_out, mode, g_raw);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_act_fwd(int kind, const float* x, int64_t n, float* y, void* stream) {
  GOTEN_REQUIRE(kind >= 1 && kind <= 4, "activation kind %d unsupported (1 silu, 2 shifted softplus, 3 sigmoid, 4 tanh)", kind);
  if (n == 0) return 0;
  act_fwd_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, as_stream(stream)>>>(kind, x, n, y);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_act_bwd(int kind, const float* g, const float* x, int64_t n, float* out, void* stream) {
  GOTEN_REQUIRE(kind >= 1 && kind <= 4, "activation kind %d unsupported (1 silu, 2 shifted softplus, 3 sigmoid, 4 tanh)", kind);
  if (n == 0) return 0;
  act_bwd_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, as_stream(stream)>>>(kind, g, x, n, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_mul_add_fwd(const float* a, const float* b, const float* c, int64_t n, float* out, void* stream) {
  if (n == 0) return 0;
  mul_add_fwd_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, as_stream(stream)>>>(a, b, c, n, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_mul_add_bwd(const float* g, const float* a, const float* b, int64_t n, float* g_a, float* g_b, void* stream) {
  if (n == 0) return 0;
  mul_add_bwd_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, as_stream(stream)>>>(g, a, b, n, g_a, g_b);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
