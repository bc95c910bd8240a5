// Fused multi-tensor Adam over a flat parameter bucket (Keras 2.1.6 semantics) and the
// `Spectral` kernel regulariser of the discriminators (layers/spectralnorm.py:199-246).
#include "common.cuh"

namespace dafk {

constexpr int OT = 256;

__global__ void __launch_bounds__(OT) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v,
                                                  __nv_bfloat16* __restrict__ shadow, int64_t n, float lr_host,
                                                  const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                                  float gscale) {
  const float lr_t = lr_dev != nullptr ? __ldg(lr_dev) : lr_host;
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = *reinterpret_cast<float4*>(p + 4 * i);
    float4 gv = ldg_stream4(g + 4 * i);
    float4 mv = *reinterpret_cast<float4*>(m + 4 * i);
    float4 vv = *reinterpret_cast<float4*>(v + 4 * i);
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = ga[k] * gscale;
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      pa[k] = pa[k] - lr_t * ma[k] / (sqrtf(va[k]) + eps);
    }
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    *reinterpret_cast<float4*>(m + 4 * i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + 4 * i) = make_float4(va[0], va[1], va[2], va[3]);
    if (shadow) Vec4<__nv_bfloat16>::store(shadow + 4 * i, pa);
  }
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float gk = g[t] * gscale;
    float mk = b1 * m[t] + (1.f - b1) * gk;
    float vk = b2 * v[t] + (1.f - b2) * gk * gk;
    float pk = p[t] - lr_t * mk / (sqrtf(vk) + eps);
    p[t] = pk; m[t] = mk; v[t] = vk;
    if (shadow) shadow[t] = __float2bfloat16_rn(pk);
  }
}

// state[0] = t (step count), state[1] = lr_t = lr*sqrt(1-b2^t)/(1-b1^t): the Keras schedule advanced on the device, so
// that a whole optimizer step is CUDA-graph capturable (no host scalar changes between replays)
__global__ void adam_tick_kernel(float* state, float lr, float b1, float b2) {
  const double t = (double)state[0] + 1.0;
  state[0] = (float)t;
  state[1] = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
}

// ---------------------------------------------------------------- Spectral
// v[c] += sum_{d in chunk} W[d,c]*u[d]
__global__ void __launch_bounds__(OT) spec_wtu_kernel(const float* __restrict__ W, const float* __restrict__ u,
                                                      float* __restrict__ v, int dim, int cout, int rows_per_block) {
  int d0 = blockIdx.x * rows_per_block;
  int d1 = min(dim, d0 + rows_per_block);
  for (int c = threadIdx.x; c < cout; c += OT) {
    float acc = 0.f;
    for (int d = d0; d < d1; ++d) acc = fmaf(W[(int64_t)d * cout + c], u[d], acc);
    atomicAdd(v + c, acc);
  }
}
// t[d] = sum_c W[d,c]*v[c]; one warp per row
__global__ void __launch_bounds__(OT) spec_wv_kernel(const float* __restrict__ W, const float* __restrict__ v,
                                                     float* __restrict__ t, int dim, int cout) {
  int warp = (blockIdx.x * OT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= dim) return;
  float acc = 0.f;
  for (int c = lane; c < cout; c += 32) acc = fmaf(W[(int64_t)warp * cout + c], v[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) t[warp] = acc;
}
// x <- x/||x|| (single CTA); if dot_with != NULL writes sum(x*dot_with) to *dot_out instead
__global__ void __launch_bounds__(OT) spec_norm_kernel(float* __restrict__ x, int n, const float* __restrict__ dot_with,
                                                       float* __restrict__ dot_out) {
  __shared__ float red[OT / 32];
  __shared__ float total;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += OT) acc += dot_with ? x[i] * dot_with[i] : x[i] * x[i];
  float t = block_sum<OT>(acc, red);
  if (threadIdx.x == 0) total = t;
  __syncthreads();
  if (dot_with) {
    if (threadIdx.x == 0) *dot_out = total;
    return;
  }
  float inv = 1.f / sqrtf(total);
  for (int i = threadIdx.x; i < n; i += OT) x[i] *= inv;
}
__global__ void __launch_bounds__(OT) spec_apply_kernel(const float* __restrict__ W, const float* __restrict__ sigma,
                                                        float alpha, float* __restrict__ loss, float* __restrict__ dW,
                                                        int64_t n) {
  __shared__ float red[OT / 32];
  const float inv_sigma = 1.f / sigma[0];
  const float gs = alpha / (float)n;
  float acc = 0.f;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = W[i];
    float d = w * inv_sigma - w;   // target - x
    acc += fabsf(d);
    // d|t - x|/dx = -sign(t - x)   (t is stop_gradient)
    if (dW) dW[i] += gs * (d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f));
  }
  float t = block_sum<OT>(acc, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, gs * t);
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, int64_t n, float lr_t,
                   float beta1, float beta2, float eps, float grad_scale, void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_adam_step: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(p && g && m && v, DAFK_ERR_BAD_ARG, "dafk_adam_step: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(p) && DAFK_ALIGNED16(g) && DAFK_ALIGNED16(m) && DAFK_ALIGNED16(v) &&
                   DAFK_ALIGNED16(bf16_shadow),
               DAFK_ERR_ALIGN, "dafk_adam_step: pointers must be 16-byte aligned");
  adam_kernel<<<bw_grid((n + 3) / 4, OT), OT, 0, as_stream(stream)>>>(p, g, m, v, (__nv_bfloat16*)bf16_shadow, n, lr_t,
                                                                     nullptr, beta1, beta2, eps, grad_scale);
  return check_launch("dafk_adam_step");
}

int dafk_adam_tick(float* state, float lr, float beta1, float beta2, void* stream) {
  DAFK_REQUIRE(state != nullptr, DAFK_ERR_BAD_ARG, "dafk_adam_tick: null pointer");
  adam_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(state, lr, beta1, beta2);
  return check_launch("dafk_adam_tick");
}

int dafk_adam_step_dev(float* p, const float* g, float* m, float* v, void* bf16_shadow, int64_t n,
                       const float* state, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_adam_step_dev: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(p && g && m && v && state, DAFK_ERR_BAD_ARG, "dafk_adam_step_dev: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(p) && DAFK_ALIGNED16(g) && DAFK_ALIGNED16(m) && DAFK_ALIGNED16(v) &&
                   DAFK_ALIGNED16(bf16_shadow),
               DAFK_ERR_ALIGN, "dafk_adam_step_dev: pointers must be 16-byte aligned");
  adam_kernel<<<bw_grid((n + 3) / 4, OT), OT, 0, as_stream(stream)>>>(p, g, m, v, (__nv_bfloat16*)bf16_shadow, n, 0.f,
                                                                     state + 1, beta1, beta2, eps, grad_scale);
  return check_launch("dafk_adam_step_dev");
}

int dafk_spectral_reg(const float* W, const float* u0, float alpha, float* loss, float* dW, float* ws, int dim,
                      int cout, void* stream) {
  DAFK_REQUIRE(dim > 0 && cout > 0 && W && u0 && ws, DAFK_ERR_BAD_ARG, "dafk_spectral_reg: bad argument");
  cudaStream_t s = as_stream(stream);
  float* u = ws;               // [dim]
  float* t = ws + dim;         // [dim]
  float* v = ws + 2 * dim;     // [cout]
  float* sigma = v + cout;     // [1]
  cudaMemcpyAsync(u, u0, sizeof(float) * dim, cudaMemcpyDeviceToDevice, s);
  int rows_per_block = (dim + kNumSMs - 1) / kNumSMs;
  if (rows_per_block < 8) rows_per_block = 8;
  int wtu_blocks = (dim + rows_per_block - 1) / rows_per_block;
  int wv_blocks = (dim * 32 + OT - 1) / OT;
  int rc;
  for (int it = 0; it < 3; ++it) {
    cudaMemsetAsync(v, 0, sizeof(float) * cout, s);
    spec_wtu_kernel<<<wtu_blocks, OT, 0, s>>>(W, u, v, dim, cout, rows_per_block);
    if ((rc = check_launch("dafk_spectral_reg(wtu)"))) return rc;
    spec_norm_kernel<<<1, OT, 0, s>>>(v, cout, nullptr, nullptr);
    if ((rc = check_launch("dafk_spectral_reg(norm v)"))) return rc;
    spec_wv_kernel<<<wv_blocks, OT, 0, s>>>(W, v, u, dim, cout);
    if ((rc = check_launch("dafk_spectral_reg(wv)"))) return rc;
    spec_norm_kernel<<<1, OT, 0, s>>>(u, dim, nullptr, nullptr);
    if ((rc = check_launch("dafk_spectral_reg(norm u)"))) return rc;
  }
  // sigma = u^T W v
  spec_wv_kernel<<<wv_blocks, OT, 0, s>>>(W, v, t, dim, cout);
  if ((rc = check_launch("dafk_spectral_reg(wv2)"))) return rc;
  spec_norm_kernel<<<1, OT, 0, s>>>(t, dim, u, sigma);
  if ((rc = check_launch("dafk_spectral_reg(sigma)"))) return rc;
  int64_t n = (int64_t)dim * cout;
  spec_apply_kernel<<<bw_grid(n, OT, 4), OT, 0, s>>>(W, sigma, alpha, loss, dW, n);
  return check_launch("dafk_spectral_reg(apply)");
}

}  // extern "C"
