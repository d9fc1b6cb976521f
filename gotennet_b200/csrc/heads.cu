// Equivariant read-out heads of the QM9 task: the element-wise / per-molecule stages of
// GatedEquivariantBlock (reference models/components/outputs.py:24-104), Dipole (:379-468) and
// ElectronicSpatialExtentV2 (:471-542).  The dense layers of the blocks run on the GEMM kernels; these kernels are the
// glue between them (vector norms, gating, dipole assembly, mass-weighted centroids), forward and backward.
#include "common.cuh"

namespace goten {

__device__ __forceinline__ float softplus_h_(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float act_h_(int kind, float v) {
  return kind == 1 ? siluf_(v) : (kind == 2 ? softplus_h_(v) - 0.69314718055994530942f : v);
}
__device__ __forceinline__ float dact_h_(int kind, float v) {
  return kind == 1 ? dsiluf_(v) : (kind == 2 ? (v > 20.f ? 1.f : sigmoidf_(v)) : 1.f);
}

// ctx[n] = [ scalars[n] | ||V[n,:,j]||_2 ],  vmix [N,3,2*nv] = [V | W]            (outputs.py:89-92)
__global__ void geb_ctx_fwd_kernel(const float* __restrict__ scalars, const float* __restrict__ vmix, int64_t N, int ns,
                                   int nv, float* __restrict__ ctx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = ns + nv;
  if (idx >= N * w) return;
  const int64_t n = idx / w;
  const int c = (int)(idx - n * w);
  if (c < ns) {
    ctx[idx] = scalars[n * ns + c];
  } else {
    const int j = c - ns;
    const float* v = vmix + n * 6 * nv + j;
    const float a = v[0], b = v[2 * nv], d = v[4 * nv];
    ctx[idx] = sqrtf(a * a + b * b + d * d);
  }
}

// g_scalars = g_ctx[:, :ns];  g_V[n,k,j] = g_ctx[n,ns+j] V[n,k,j] / ||V[n,:,j]||  (0 where the norm is 0, as torch)
__global__ void geb_ctx_bwd_kernel(const float* __restrict__ g_ctx, const float* __restrict__ vmix, int64_t N, int ns,
                                   int nv, float* __restrict__ g_scalars, float* __restrict__ g_vmix) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = ns + nv;
  if (idx >= N * w) return;
  const int64_t n = idx / w;
  const int c = (int)(idx - n * w);
  if (c < ns) {
    g_scalars[n * ns + c] = g_ctx[idx];
  } else {
    const int j = c - ns;
    const float* v = vmix + n * 6 * nv + j;
    float* g = g_vmix + n * 6 * nv + j;
    const float a = v[0], b = v[2 * nv], d = v[4 * nv];
    const float nrm = sqrtf(a * a + b * b + d * d);
    const float f = nrm > 0.f ? g_ctx[idx] / nrm : 0.f;
    g[0] = f * a;
    g[2 * nv] = f * b;
    g[4 * nv] = f * d;
  }
}

// s_out = sact(x[:, :nso]);  v_out[n,k,j] = x[n,nso+j] * W[n,k,j]                  (outputs.py:94-99)
__global__ void geb_gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ vmix, int64_t N, int nso,
                                    int nv, int sact, float* __restrict__ s_out, float* __restrict__ v_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = nso + nv;
  if (idx >= N * w) return;
  const int64_t n = idx / w;
  const int c = (int)(idx - n * w);
  const float xv = x[idx];
  if (c < nso) {
    s_out[n * nso + c] = act_h_(sact, xv);
  } else {
    const int j = c - nso;
    const float* wv = vmix + n * 6 * nv + nv + j;
    float* o = v_out + n * 3 * nv + j;
    o[0] = xv * wv[0];
    o[nv] = xv * wv[2 * nv];
    o[2 * nv] = xv * wv[4 * nv];
  }
}

__global__ void geb_gate_bwd_kernel(const float* __restrict__ g_s, const float* __restrict__ g_v,
                                    const float* __restrict__ x, const float* __restrict__ vmix, int64_t N, int nso, int nv,
                                    int sact, float* __restrict__ g_x, float* __restrict__ g_vmix) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = nso + nv;
  if (idx >= N * w) return;
  const int64_t n = idx / w;
  const int c = (int)(idx - n * w);
  const float xv = x[idx];
  if (c < nso) {
    g_x[idx] = (g_s ? g_s[n * nso + c] : 0.f) * dact_h_(sact, xv);
  } else {
    const int j = c - nso;
    const float* wv = vmix + n * 6 * nv + nv + j;
    float* gw = g_vmix + n * 6 * nv + nv + j;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float g = g_v ? g_v[n * 3 * nv + k * nv + j] : 0.f;
      acc += g * wv[2 * k * nv];
      gw[2 * k * nv] = g * xv;
    }
    g_x[idx] = acc;
  }
}

// Dipole: yi[n,k] = mu_atom[n,k] + pos[n,k] * q[n],  q = stddev * l0 + mean (when standardised)   (outputs.py:446-453)
__global__ void dipole_atom_fwd_kernel(const float* __restrict__ l1, const float* __restrict__ l0,
                                       const float* __restrict__ pos, float sd, float mu, int64_t N, float* __restrict__ yi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 3) return;
  const int64_t n = i / 3;
  yi[i] = l1[i] + pos[i] * (sd * l0[n] + mu);
}
__global__ void dipole_atom_bwd_kernel(const float* __restrict__ g_yi, const float* __restrict__ l0,
                                       const float* __restrict__ pos, float sd, float mu, int64_t N, float* __restrict__ g_l1,
                                       float* __restrict__ g_l0, float* __restrict__ g_pos) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float q = sd * l0[n] + mu;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float g = g_yi[n * 3 + k];
    g_l1[n * 3 + k] = g;
    acc += g * pos[n * 3 + k];
    if (g_pos) g_pos[n * 3 + k] = g * q;
  }
  g_l0[n] = sd * acc;
}

// y[m] = || v[m,:] ||_2  (predict_magnitude, outputs.py:460-461)
__global__ void rownorm_fwd_kernel(const float* __restrict__ v, int64_t M, int D, float* __restrict__ y) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += v[m * D + d] * v[m * D + d];
  y[m] = sqrtf(s);
}
__global__ void rownorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v, int64_t M, int D,
                                   float* __restrict__ g_v) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += v[m * D + d] * v[m * D + d];
  const float nrm = sqrtf(s), f = nrm > 0.f ? g[m] / nrm : 0.f;
  for (int d = 0; d < D; ++d) g_v[m * D + d] = f * v[m * D + d];
}

// Electronic spatial extent (outputs.py:522-541): c_m = sum_i m_i p_i / sum_i m_i over the molecule,
// yi = |p_i - c_m|^2 x_i,  y_m = sum_i yi.  One warp per molecule, fixed summation order.
__device__ __forceinline__ float mass_of(const float* __restrict__ table, int rows, int64_t z) {
  z = z < 0 ? 0 : (z >= rows ? rows - 1 : z);
  return table[z];
}
__global__ void ese_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pos, const int64_t* __restrict__ z,
                               const float* __restrict__ mass, int mass_rows, const int32_t* __restrict__ mol_ptr,
                               int n_mol, float* __restrict__ yi, float* __restrict__ y, float* __restrict__ cen) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= n_mol) return;
  const int a0 = mol_ptr[m], a1 = mol_ptr[m + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f, sm = 0.f;
  for (int n = a0 + lane; n < a1; n += 32) {
    const float w = mass_of(mass, mass_rows, z[n]);
    sx += w * pos[n * 3];
    sy += w * pos[n * 3 + 1];
    sz += w * pos[n * 3 + 2];
    sm += w;
  }
  sx = warp_sum(sx), sy = warp_sum(sy), sz = warp_sum(sz), sm = warp_sum(sm);
  const float cx = sx / sm, cy = sy / sm, cz = sz / sm;
  float acc = 0.f;
  for (int n = a0 + lane; n < a1; n += 32) {
    const float dx = pos[n * 3] - cx, dy = pos[n * 3 + 1] - cy, dz = pos[n * 3 + 2] - cz;
    const float r = sqrtf(dx * dx + dy * dy + dz * dz);   // the reference squares the norm (outputs.py:528-529)
    const float v = r * r * x[n];
    yi[n] = v;
    acc += v;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    y[m] = acc;
    cen[m * 4] = cx, cen[m * 4 + 1] = cy, cen[m * 4 + 2] = cz, cen[m * 4 + 3] = sm;
  }
}
// g_x[n] = g_y[m] r_n^2;  g_pos[k] = g_y[m] (2 x_k (p_k - c) - 2 (m_k / M) sum_i x_i (p_i - c))
__global__ void ese_bwd_kernel(const float* __restrict__ g_y, const float* __restrict__ x, const float* __restrict__ pos,
                               const int64_t* __restrict__ z, const float* __restrict__ mass, int mass_rows,
                               const int32_t* __restrict__ mol_ptr, int n_mol, const float* __restrict__ cen,
                               float* __restrict__ g_x, float* __restrict__ g_pos) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= n_mol) return;
  const int a0 = mol_ptr[m], a1 = mol_ptr[m + 1];
  const float cx = cen[m * 4], cy = cen[m * 4 + 1], cz = cen[m * 4 + 2], M = cen[m * 4 + 3], g = g_y[m];
  float tx = 0.f, ty = 0.f, tz = 0.f;
  for (int n = a0 + lane; n < a1; n += 32) {
    const float dx = pos[n * 3] - cx, dy = pos[n * 3 + 1] - cy, dz = pos[n * 3 + 2] - cz;
    g_x[n] = g * (dx * dx + dy * dy + dz * dz);
    tx += x[n] * dx, ty += x[n] * dy, tz += x[n] * dz;
  }
  if (g_pos == nullptr) return;
  tx = warp_sum(tx), ty = warp_sum(ty), tz = warp_sum(tz);
  for (int n = a0 + lane; n < a1; n += 32) {
    const float w = mass_of(mass, mass_rows, z[n]) / M;
    g_pos[n * 3] = 2.f * g * (x[n] * (pos[n * 3] - cx) - w * tx);
    g_pos[n * 3 + 1] = 2.f * g * (x[n] * (pos[n * 3 + 1] - cy) - w * ty);
    g_pos[n * 3 + 2] = 2.f * g * (x[n] * (pos[n * 3 + 2] - cz) - w * tz);
  }
}

}  // namespace goten

using namespace goten;

extern "C" {

int goten_geb_ctx_fwd(const float* scalars, const float* vmix, int64_t n_nodes, int ns, int nv, float* ctx, void* stream) {
  GOTEN_REQUIRE(ns >= 0 && nv >= 1, "bad widths ns=%d nv=%d", ns, nv);
  const int64_t tot = n_nodes * (ns + nv);
  if (tot == 0) return 0;
  geb_ctx_fwd_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, as_stream(stream)>>>(scalars, vmix, n_nodes, ns, nv, ctx);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_geb_ctx_bwd(const float* g_ctx, const float* vmix, int64_t n_nodes, int ns, int nv, float* g_scalars,
                      float* g_vmix, void* stream) {
  GOTEN_REQUIRE(ns >= 0 && nv >= 1, "bad widths ns=%d nv=%d", ns, nv);
  const int64_t tot = n_nodes * (ns + nv);
  if (tot == 0) return 0;
  geb_ctx_bwd_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, as_stream(stream)>>>(g_ctx, vmix, n_nodes, ns, nv, g_scalars,
                                                                               g_vmix);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_geb_gate_fwd(const float* x, const float* vmix, int64_t n_nodes, int nso, int nv, int sact, float* s_out,
                       float* v_out, void* stream) {
  GOTEN_REQUIRE(nso >= 0 && nv >= 1 && sact >= 0 && sact <= 2, "bad arguments nso=%d nv=%d sact=%d", nso, nv, sact);
  const int64_t tot = n_nodes * (nso + nv);
  if (tot == 0) return 0;
  geb_gate_fwd_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, as_stream(stream)>>>(x, vmix, n_nodes, nso, nv, sact, s_out,
                                                                                v_out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_geb_gate_bwd(const float* g_s, const float* g_v, const float* x, const float* vmix, int64_t n_nodes, int nso,
                       int nv, int sact, float* g_x, float* g_vmix, void* stream) {
  GOTEN_REQUIRE(nso >= 0 && nv >= 1 && sact >= 0 && sact <= 2, "bad arguments nso=%d nv=%d sact=%d", nso, nv, sact);
  const int64_t tot = n_nodes * (nso + nv);
  if (tot == 0) return 0;
  geb_gate_bwd_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, as_stream(stream)>>>(g_s, g_v, x, vmix, n_nodes, nso, nv, sact,
                                                                                g_x, g_vmix);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_dipole_atom_fwd(const float* l1, const float* l0, const float* pos, float stddev, float mean, int64_t n_nodes,
                          float* yi, void* stream) {
  if (n_nodes == 0) return 0;
  dipole_atom_fwd_kernel<<<(unsigned)cdiv64(n_nodes * 3, 256), 256, 0, as_stream(stream)>>>(l1, l0, pos, stddev, mean,
                                                                                           n_nodes, yi);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_dipole_atom_bwd(const float* g_yi, const float* l0, const float* pos, float stddev, float mean, int64_t n_nodes,
                          float* g_l1, float* g_l0, float* g_pos, void* stream) {
  if (n_nodes == 0) return 0;
  dipole_atom_bwd_kernel<<<(unsigned)cdiv64(n_nodes, 256), 256, 0, as_stream(stream)>>>(g_yi, l0, pos, stddev, mean, n_nodes,
                                                                                       g_l1, g_l0, g_pos);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_rownorm_fwd(const float* v, int64_t rows, int dim, float* y, void* stream) {
  GOTEN_REQUIRE(dim >= 1, "bad row width %d", dim);
  if (rows == 0) return 0;
  rownorm_fwd_kernel<<<(unsigned)cdiv64(rows, 256), 256, 0, as_stream(stream)>>>(v, rows, dim, y);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_rownorm_bwd(const float* g, const float* v, int64_t rows, int dim, float* g_v, void* stream) {
  GOTEN_REQUIRE(dim >= 1, "bad row width %d", dim);
  if (rows == 0) return 0;
  rownorm_bwd_kernel<<<(unsigned)cdiv64(rows, 256), 256, 0, as_stream(stream)>>>(g, v, rows, dim, g_v);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_ese_fwd(const float* x, const float* pos, const int64_t* z, const float* mass, int mass_rows,
                  const int32_t* mol_ptr, int n_mol, float* yi, float* y, float* centroid, void* stream) {
  GOTEN_REQUIRE(mass_rows >= 1, "empty atomic mass table");
  if (n_mol == 0) return 0;
  ese_fwd_kernel<<<(unsigned)cdiv64((int64_t)n_mol * 32, 256), 256, 0, as_stream(stream)>>>(x, pos, z, mass, mass_rows,
                                                                                           mol_ptr, n_mol, yi, y, centroid);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_ese_bwd(const float* g_y, const float* x, const float* pos, const int64_t* z, const float* mass, int mass_rows,
                  const int32_t* mol_ptr, int n_mol, const float* centroid, float* g_x, float* g_pos, void* stream) {
  GOTEN_REQUIRE(mass_rows >= 1, "empty atomic mass table");
  if (n_mol == 0) return 0;
  ese_bwd_kernel<<<(unsigned)cdiv64((int64_t)n_mol * 32, 256), 256, 0, as_stream(stream)>>>(g_y, x, pos, z, mass, mass_rows,
                                                                                           mol_ptr, n_mol, centroid, g_x,
                                                                                           g_pos);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
