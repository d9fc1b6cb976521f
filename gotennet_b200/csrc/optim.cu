// Optimiser step of the reference training loop on ONE flat fp32 parameter buffer:
// torch.optim.AdamW(eps=1e-7) + global-norm gradient clipping (reference models/goten_model.py:521-578,
// configs/trainer/default.yaml:10 gradient_clip_val 5.0).  The flat gradient buffer is the one the data-parallel
// all-reduce produces (parallel.FlatGradBuffer), so the step is two launches right behind the collective:
//   goten_sumsq      deterministic two-stage sum of squares (no atomics)  -> device scalar
//   goten_adamw_step clip coefficient from that scalar (no host read) + AdamW update of p, m, v in place
// Both are pure HBM streams: 4 B/param read for the norm, 16 B read + 12 B written per parameter for the update.
#include "common.cuh"

namespace goten {

constexpr int SUMSQ_BLOCKS = 592;  // 4 per SM on a 148-SM B200; also the length of the partial buffer

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[33];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail
    const float v = g[(n4 << 2) + threadIdx.x];
    s = fmaf(v, v, s);
  }
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(1024) sumsq_final_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ out) {
  __shared__ float red[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = t;
}

struct AdamW {
  float lr, beta1, beta2, omb1, omb2, eps, weight_decay, bias_c1, bias_c2_sqrt, max_norm, grad_scale;
};

// torch.optim.AdamW single-tensor update order: decay, lerp of the first moment, second moment, bias-corrected step
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamW& a) {
  p *= 1.0f - a.lr * a.weight_decay;
  m = fmaf(g - m, a.omb1, m);
  v = fmaf(a.omb2, g * g, v * a.beta2);
  const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
  p -= (a.lr / a.bias_c1) * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, AdamW a, const float* __restrict__ sumsq) {
  float coef = a.grad_scale;
  if (a.max_norm > 0.f && sumsq != nullptr) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6) clamped to 1 (norm of the scaled gradient)
    const float norm = sqrtf(*sumsq) * a.grad_scale;
    coef *= fminf(1.0f, a.max_norm / (norm + 1e-6f));
  }
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adamw_one(pp.x, gg.x * coef, mm.x, vv.x, a);
    adamw_one(pp.y, gg.y * coef, mm.y, vv.y, a);
    adamw_one(pp.z, gg.z * coef, mm.z, vv.z, a);
    adamw_one(pp.w, gg.w * coef, mm.w, vv.w, a);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    adamw_one(p[i], g[i] * coef, m[i], v[i], a);
  }
}

}  // namespace goten

using namespace goten;

extern "C" {

int goten_sumsq(const float* g, int64_t n, float* partial, float* out, void* stream) {
  cudaStream_t st = as_stream(stream);
  GOTEN_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "goten_sumsq needs a 16-byte aligned buffer");
  sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, st>>>(g, n, partial);
  GOTEN_CHECK_LAUNCH();
  sumsq_final_kernel<<<1, 1024, 0, st>>>(partial, SUMSQ_BLOCKS, out);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

int goten_sumsq_workspace_floats(void) { return SUMSQ_BLOCKS; }

int goten_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                     double weight_decay, float bias_c1, float bias_c2, float max_norm, const float* sumsq,
                     float grad_scale, void* stream) {
  if (n == 0) return 0;
  GOTEN_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0, "goten_adamw_step needs 16-byte aligned buffers");
  GOTEN_REQUIRE(bias_c1 > 0.f && bias_c2 > 0.f, "bias corrections must be positive (step >= 1)");
  // 1 - beta is formed in double like torch does on the host (float(1 - 0.999f) is 1.3e-5 off 1e-3)
  AdamW a{(float)lr, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay,
          bias_c1, sqrtf(bias_c2), max_norm, grad_scale};
  int64_t grid = cdiv64(n >> 2, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  adamw_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(p, g, m, v, n, a, sumsq);
  GOTEN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
