// BatchNormalization (training + inference, forward + backward) and the per-sample
// instance normalisation fused with SPADE conditioning.  HBM-bound: every pass reads each
// activation exactly once with 128-bit loads; statistics are reduced warp -> block -> global
// (double atomics, one per channel per block).
#include "common.cuh"
#include "reduce.cuh"
#include "norm_wide.cuh"

namespace dafk {

constexpr int TPB = 256;

static inline bool bn_channels_ok(int C) { return C >= 4 && C <= 1024 && (1024 % C) == 0; }

// grid for [M,C] float4 streams: multiple of the SM count
static inline int bn_grid(int64_t n4) { return bw_grid(n4, TPB, 8); }

// ---------------------------------------------------------------- BN statistics
template <typename TX>
__global__ void __launch_bounds__(TPB) bn_stats_kernel(const TX* __restrict__ x, double* __restrict__ acc,
                                                       int64_t n4, int C) {
  extern __shared__ float sm[];
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    Vec4<TX>::load(x + 4 * i, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[k] += v[k]; q[k] += v[k] * v[k]; }
  }
  channel_reduce2<TPB>(s, q, C, sm, acc, acc + C);
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int64_t M, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ mm,
                                   float* __restrict__ mv) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = acc[c] / (double)M;
  double var = acc[C + c] / (double)M - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (mm && mv) {
    // moving variance receives the Bessel-corrected batch variance (fused TF path)
    double unb = (M > 1) ? var * ((double)M / (double)(M - 1)) : var;
    mm[c] = mm[c] * momentum + (float)m * (1.f - momentum);
    mv[c] = mv[c] * momentum + (float)unb * (1.f - momentum);
  }
}

__global__ void bn_rstd_kernel(const float* __restrict__ var, float* __restrict__ rstd, int C, float eps) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) rstd[c] = 1.f / sqrtf(var[c] + eps);
}

// ---------------------------------------------------------------- BN apply
template <typename TX, typename TO>
__global__ void __launch_bounds__(TPB) bn_apply_kernel(const TX* __restrict__ x, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, TO* __restrict__ out,
                                                       int64_t n4, int C, int act) {
  const int c0 = (threadIdx.x * 4) % C;
  float mu[4], rs[4], g[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { mu[k] = mean[c0 + k]; rs[k] = rstd[c0 + k]; g[k] = gamma[c0 + k]; b[k] = beta[c0 + k]; }
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    Vec4<TX>::load(x + 4 * i, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float z = (v[k] - mu[k]) * rs[k] * g[k] + b[k];
      v[k] = (act == DAFK_ACT_RELU) ? fmaxf(z, 0.f) : z;
    }
    Vec4<TO>::store(out + 4 * i, v);
  }
}

// ---------------------------------------------------------------- BN backward
template <typename TD, typename TX>
__global__ void __launch_bounds__(TPB) bn_bwd_reduce_kernel(const TD* __restrict__ dout, const TX* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            double* __restrict__ acc, int64_t n4, int C, int act) {
  extern __shared__ float sm[];
  const int c0 = (threadIdx.x * 4) % C;
  float mu[4], rs[4], g[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { mu[k] = mean[c0 + k]; rs[k] = rstd[c0 + k]; g[k] = gamma[c0 + k]; b[k] = beta[c0 + k]; }
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    Vec4<TX>::load(x + 4 * i, v);
    float d[4];
    Vec4<TD>::load(dout + 4 * i, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xh = (v[k] - mu[k]) * rs[k];
      float z = xh * g[k] + b[k];
      float dz = (act == DAFK_ACT_RELU && !(z > 0.f)) ? 0.f : d[k];
      s0[k] += dz;
      s1[k] += dz * xh;
    }
  }
  channel_reduce2<TPB>(s0, s1, C, sm, acc, acc + C);
}

template <typename TD, typename TX, typename TO>
__global__ void __launch_bounds__(TPB) bn_bwd_apply_kernel(const TD* __restrict__ dout, const TX* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const double* __restrict__ acc, TO* __restrict__ dx,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           float* __restrict__ dbias, int64_t n4, int64_t M, int C,
                                                           int act) {
  extern __shared__ float sm[];
  float sdx[4] = {0, 0, 0, 0};
  const int c0 = (threadIdx.x * 4) % C;
  float mu[4], rs[4], g[4], b[4], m0[4], m1[4];
  const double invM = 1.0 / (double)M;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mu[k] = mean[c0 + k]; rs[k] = rstd[c0 + k]; g[k] = gamma[c0 + k]; b[k] = beta[c0 + k];
    m0[k] = (float)(acc[c0 + k] * invM);
    m1[k] = (float)(acc[C + c0 + k] * invM);
  }
  if (blockIdx.x == 0 && dgamma && dbeta) {
    for (int c = threadIdx.x; c < C; c += TPB) {
      dbeta[c] += (float)acc[c];
      dgamma[c] += (float)acc[C + c];
    }
  }
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    Vec4<TX>::load(x + 4 * i, v);
    float d[4];
    Vec4<TD>::load(dout + 4 * i, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xh = (v[k] - mu[k]) * rs[k];
      float z = xh * g[k] + b[k];
      float dz = (act == DAFK_ACT_RELU && !(z > 0.f)) ? 0.f : d[k];
      v[k] = g[k] * rs[k] * (dz - m0[k] - xh * m1[k]);
      sdx[k] += v[k];
    }
    Vec4<TO>::store(dx + 4 * i, v);
  }
  if (dbias) {
    // block-level per-channel sum of dx, then one float atomic per channel per CTA
    for (int i = threadIdx.x; i < C; i += TPB) sm[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&sm[c0 + k], sdx[k]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += TPB) atomicAdd(dbias + i, sm[i]);
  }
}

__global__ void __launch_bounds__(TPB) bn_bwd_frozen_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ dx, int64_t n4, int C, int act) {
  const int c0 = (threadIdx.x * 4) % C;
  float mu[4], rs[4], g[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { mu[k] = mean[c0 + k]; rs[k] = rstd[c0 + k]; g[k] = gamma[c0 + k]; b[k] = beta[c0 + k]; }
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 t = ldg_stream4(x + 4 * i);
    float4 dd = ldg_stream4(dout + 4 * i);
    float v[4] = {t.x, t.y, t.z, t.w};
    float d[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float z = (v[k] - mu[k]) * rs[k] * g[k] + b[k];
      float dz = (act == DAFK_ACT_RELU && !(z > 0.f)) ? 0.f : d[k];
      v[k] = dz * g[k] * rs[k];
    }
    Vec4<float>::store(dx + 4 * i, v);
  }
}

// ---------------------------------------------------------------- instance norm (axis=None) + SPADE
// grid = (chunks, B)
__global__ void __launch_bounds__(TPB) in_stats_kernel(const float* __restrict__ x, double* __restrict__ acc,
                                                       int64_t HWC) {
  __shared__ float red[TPB / 32];
  int b = blockIdx.y;
  const float* xb = x + (int64_t)b * HWC;
  float s = 0.f, q = 0.f;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldg_stream4(xb + 4 * i);
    s += (v.x + v.y) + (v.z + v.w);
    q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  float ts = block_sum<TPB>(s, red);
  float tq = block_sum<TPB>(q, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 2 * b, (double)ts);
    atomicAdd(acc + 2 * b + 1, (double)tq);
  }
}

__device__ __forceinline__ void in_moments(const double* acc, int b, int64_t HWC, float eps, float& mean,
                                           float& sd, float& inv) {
  double m = acc[2 * b] / (double)HWC;
  double var = acc[2 * b + 1] / (double)HWC - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  sd = (float)sqrt(var);
  inv = 1.f / (sd + eps);
}

__device__ __forceinline__ float act_apply(float z, int act, float alpha) {
  if (act == DAFK_ACT_RELU) return z > 0.f ? z : 0.f;
  if (act == DAFK_ACT_LRELU) return z > 0.f ? z : alpha * z;
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act, float alpha) {
  if (act == DAFK_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == DAFK_ACT_LRELU) return z > 0.f ? 1.f : (z < 0.f ? alpha : 0.f);
  return 1.f;
}

__global__ void __launch_bounds__(TPB) spade_fwd_kernel(const float* __restrict__ x, const double* __restrict__ acc,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ y, int64_t HWC, float eps, int act,
                                                        float alpha) {
  int b = blockIdx.y;
  float mean, sd, inv;
  in_moments(acc, b, HWC, eps, mean, sd, inv);
  int64_t off = (int64_t)b * HWC;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldg_stream4(x + off + 4 * i);
    float4 g = ldg_stream4(gamma + off + 4 * i);
    float4 t = ldg_stream4(beta + off + 4 * i);
    float4 r;
    r.x = act_apply((v.x - mean) * inv * (1.f + g.x) + t.x, act, alpha);
    r.y = act_apply((v.y - mean) * inv * (1.f + g.y) + t.y, act, alpha);
    r.z = act_apply((v.z - mean) * inv * (1.f + g.z) + t.z, act, alpha);
    r.w = act_apply((v.w - mean) * inv * (1.f + g.w) + t.w, act, alpha);
    stg_stream4(y + off + 4 * i, r);
  }
}

// pass 1: dgamma, dbeta (full-res), S1 = sum dxn, S2 = sum dxn*(x-mean) per sample
__global__ void __launch_bounds__(TPB) spade_bwd1_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const double* __restrict__ acc, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, double* __restrict__ ws, int64_t HWC,
                                                         float eps, int act, float alpha) {
  __shared__ float red[TPB / 32];
  int b = blockIdx.y;
  float mean, sd, inv;
  in_moments(acc, b, HWC, eps, mean, sd, inv);
  int64_t off = (int64_t)b * HWC;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v4 = ldg_stream4(x + off + 4 * i);
    float4 g4 = ldg_stream4(gamma + off + 4 * i);
    float4 t4 = ldg_stream4(beta + off + 4 * i);
    float4 d4 = ldg_stream4(dy + off + 4 * i);
    float v[4] = {v4.x, v4.y, v4.z, v4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
    float t[4] = {t4.x, t4.y, t4.z, t4.w}, d[4] = {d4.x, d4.y, d4.z, d4.w};
    float og[4], ob[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xc = v[k] - mean;
      float xn = xc * inv;
      float z = xn * (1.f + g[k]) + t[k];
      float dz = d[k] * act_grad(z, act, alpha);
      og[k] = dz * xn;
      ob[k] = dz;
      float dxn = dz * (1.f + g[k]);
      s1 += dxn;
      s2 += dxn * xc;
    }
    Vec4<float>::store(dgamma + off + 4 * i, og);
    Vec4<float>::store(dbeta + off + 4 * i, ob);
  }
  float t1 = block_sum<TPB>(s1, red);
  float t2 = block_sum<TPB>(s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * b, (double)t1);
    atomicAdd(ws + 2 * b + 1, (double)t2);
  }
}

// pass 2: dx_i = (dxn_i - S1/n)/s' - (x_i-mean)*S2/(n*sd*s'^2),  s' = sd+eps
__global__ void __launch_bounds__(TPB) spade_bwd2_kernel(const float* __restrict__ dbeta_dz, const float* __restrict__ x,
                                                         const double* __restrict__ acc, const float* __restrict__ gamma,
                                                         const double* __restrict__ ws, float* __restrict__ dx,
                                                         int64_t HWC, float eps) {
  int b = blockIdx.y;
  float mean, sd, inv;
  in_moments(acc, b, HWC, eps, mean, sd, inv);
  const float m1 = (float)(ws[2 * b] / (double)HWC);
  const float k2 = (float)(ws[2 * b + 1] / (double)HWC) * inv * inv / sd;
  int64_t off = (int64_t)b * HWC;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v4 = ldg_stream4(x + off + 4 * i);
    float4 g4 = ldg_stream4(gamma + off + 4 * i);
    float4 z4 = *reinterpret_cast<const float4*>(dbeta_dz + off + 4 * i);  // dz was stored as dbeta
    float4 r;
    r.x = (z4.x * (1.f + g4.x) - m1) * inv - (v4.x - mean) * k2;
    r.y = (z4.y * (1.f + g4.y) - m1) * inv - (v4.y - mean) * k2;
    r.z = (z4.z * (1.f + g4.z) - m1) * inv - (v4.z - mean) * k2;
    r.w = (z4.w * (1.f + g4.w) - m1) * inv - (v4.w - mean) * k2;
    stg_stream4(dx + off + 4 * i, r);
  }
}


// ---------------------------------------------------------------- SPADE_COND alone (layers/spade.py:41-58)
// y = x*(1+gamma) + beta on an already-normalised input; backward dx = dy*(1+gamma), dgamma = dy*x (dbeta = dy)
__global__ void __launch_bounds__(TPB) spade_cond_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y,
                                                             int64_t n) {
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldg_stream4(x + 4 * i), g = ldg_stream4(gamma + 4 * i), t = ldg_stream4(beta + 4 * i), r;
    r.x = v.x * (1.f + g.x) + t.x; r.y = v.y * (1.f + g.y) + t.y;
    r.z = v.z * (1.f + g.z) + t.z; r.w = v.w * (1.f + g.w) + t.w;
    stg_stream4(y + 4 * i, r);
  }
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) y[t] = x[t] * (1.f + gamma[t]) + beta[t];
}
__global__ void __launch_bounds__(TPB) spade_cond_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ gamma, float* __restrict__ dx,
                                                             float* __restrict__ dgamma, int64_t n) {
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 d = ldg_stream4(dy + 4 * i), v = ldg_stream4(x + 4 * i), g = ldg_stream4(gamma + 4 * i), a, b;
    a.x = d.x * (1.f + g.x); a.y = d.y * (1.f + g.y); a.z = d.z * (1.f + g.z); a.w = d.w * (1.f + g.w);
    b.x = d.x * v.x; b.y = d.y * v.y; b.z = d.z * v.z; b.w = d.w * v.w;
    stg_stream4(dx + 4 * i, a);
    stg_stream4(dgamma + 4 * i, b);
  }
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { dx[t] = dy[t] * (1.f + gamma[t]); dgamma[t] = dy[t] * x[t]; }
}

// ---------------------------------------------------------------- InstanceNormalization(axis=None) with its scalar affine
// utils/model_utils.py:6-12 normalise('instance') = keras_contrib InstanceNormalization(): per-sample statistics over
// H,W,C jointly, y = act(gamma * (x-mean)/(std+eps) + beta), gamma and beta of shape (1,).  grid = (chunks, B)
__global__ void __launch_bounds__(TPB) in_affine_fwd_kernel(const float* __restrict__ x, const double* __restrict__ acc,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ y, int64_t HWC, float eps, int act,
                                                            float alpha) {
  int b = blockIdx.y;
  float mean, sd, inv;
  in_moments(acc, b, HWC, eps, mean, sd, inv);
  const float g = gamma[0], t = beta[0];
  int64_t off = (int64_t)b * HWC;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldg_stream4(x + off + 4 * i), r;
    r.x = act_apply((v.x - mean) * inv * g + t, act, alpha);
    r.y = act_apply((v.y - mean) * inv * g + t, act, alpha);
    r.z = act_apply((v.z - mean) * inv * g + t, act, alpha);
    r.w = act_apply((v.w - mean) * inv * g + t, act, alpha);
    stg_stream4(y + off + 4 * i, r);
  }
}
// pass 1: dz = dy*act'(z) (stored in dx), per-sample S1 = sum dz*gamma, S2 = sum dz*gamma*(x-mean); dgamma += sum dz*xn,
// dbeta += sum dz over the whole batch (double accumulators wsg[0..1])
__global__ void __launch_bounds__(TPB) in_affine_bwd1_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const double* __restrict__ acc, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ dz_out,
                                                             double* __restrict__ ws, double* __restrict__ wsg,
                                                             int64_t HWC, float eps, int act, float alpha) {
  __shared__ float red[TPB / 32];
  int b = blockIdx.y;
  float mean, sd, inv;
  in_moments(acc, b, HWC, eps, mean, sd, inv);
  const float g = gamma[0], t = beta[0];
  int64_t off = (int64_t)b * HWC;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float s1 = 0.f, s2 = 0.f, sg = 0.f, sb = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v4 = ldg_stream4(x + off + 4 * i), d4 = ldg_stream4(dy + off + 4 * i);
    float v[4] = {v4.x, v4.y, v4.z, v4.w}, d[4] = {d4.x, d4.y, d4.z, d4.w}, o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xc = v[k] - mean, xn = xc * inv;
      float dz = d[k] * act_grad(xn * g + t, act, alpha);
      o[k] = dz;
      sg += dz * xn;
      sb += dz;
      s1 += dz * g;
      s2 += dz * g * xc;
    }
    Vec4<float>::store(dz_out + off + 4 * i, o);
  }
  float t1 = block_sum<TPB>(s1, red), t2 = block_sum<TPB>(s2, red);
  float tg = block_sum<TPB>(sg, red), tb = block_sum<TPB>(sb, red);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * b, (double)t1);
    atomicAdd(ws + 2 * b + 1, (double)t2);
    atomicAdd(wsg, (double)tg);
    atomicAdd(wsg + 1, (double)tb);
  }
}
// pass 2 (in place on dz): dx = (dz*gamma - S1/n)/s' - (x-mean)*S2/(n*sd*s'^2); block (0,0) adds the parameter gradients
__global__ void __launch_bounds__(TPB) in_affine_bwd2_kernel(float* __restrict__ dx, const float* __restrict__ x,
                                                             const double* __restrict__ acc, const float* __restrict__ gamma,
                                                             const double* __restrict__ ws, const double* __restrict__ wsg,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             int64_t HWC, float eps) {
  int b = blockIdx.y;
  float mean, sd, inv;
  in_moments(acc, b, HWC, eps, mean, sd, inv);
  const float g = gamma[0];
  const float m1 = (float)(ws[2 * b] / (double)HWC);
  const float k2 = sd > 0.f ? (float)(ws[2 * b + 1] / (double)HWC) * inv * inv / sd : 0.f;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    if (dgamma) dgamma[0] += (float)wsg[0];
    if (dbeta) dbeta[0] += (float)wsg[1];
  }
  int64_t off = (int64_t)b * HWC;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v4 = ldg_stream4(x + off + 4 * i);
    float4 z4 = *reinterpret_cast<const float4*>(dx + off + 4 * i);
    float4 r;
    r.x = (z4.x * g - m1) * inv - (v4.x - mean) * k2;
    r.y = (z4.y * g - m1) * inv - (v4.y - mean) * k2;
    r.z = (z4.z * g - m1) * inv - (v4.z - mean) * k2;
    r.w = (z4.w * g - m1) * inv - (v4.w - mean) * k2;
    *reinterpret_cast<float4*>(dx + off + 4 * i) = r;
  }
}

// model_components/balancer.py:33-38
__global__ void __launch_bounds__(TPB) pair_dice_kernel(const float* __restrict__ a, const float* __restrict__ bb,
                                                        double* __restrict__ ws, int64_t HWC) {
  __shared__ float red[TPB / 32];
  int b = blockIdx.y;
  int64_t off = (int64_t)b * HWC;
  float si = 0.f, sa = 0.f, sb = 0.f;
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 u = ldg_stream4(a + off + 4 * i);
    float4 v = ldg_stream4(bb + off + 4 * i);
    si += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
    sa += (u.x + u.y) + (u.z + u.w);
    sb += (v.x + v.y) + (v.z + v.w);
  }
  float t0 = block_sum<TPB>(si, red), t1 = block_sum<TPB>(sa, red), t2 = block_sum<TPB>(sb, red);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 3 * b, (double)t0);
    atomicAdd(ws + 3 * b + 1, (double)t1);
    atomicAdd(ws + 3 * b + 2, (double)t2);
  }
}
__global__ void pair_dice_finish_kernel(const double* __restrict__ ws, float* __restrict__ out, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = (float)((2.0 * ws[3 * b] + 1e-12) / (ws[3 * b + 1] + ws[3 * b + 2] + 1e-12));
}

static inline int per_sample_chunks(int64_t n4, int B) {
  int64_t chunks = (n4 + TPB - 1) / TPB;
  int64_t cap = ((int64_t)kNumSMs * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

// any channel count (f32 in / out): one thread per element.  Only reached when C is not a power of two in [4,1024]
// (e.g. a 2-filter UNet level); forward only.
__global__ void __launch_bounds__(TPB) bn_apply_any_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ out,
                                                           int64_t n, int C, int act) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c = (int)(i % C);
    float z = (x[i] - mean[c]) * rstd[c] * gamma[c] + beta[c];
    out[i] = (act == DAFK_ACT_RELU && !(z > 0.f)) ? 0.f : z;
  }
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_bn_stats(const void* x, int x_dt, double* acc, int64_t M, int C, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_stats: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(x && acc, DAFK_ERR_BAD_ARG, "dafk_bn_stats: null pointer");
  DAFK_REQUIRE(bn_channels_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_stats: C must be a power of two in [4,1024] (got %d)", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(x), DAFK_ERR_ALIGN, "dafk_bn_stats: x must be 16-byte aligned");
  int64_t n4 = M * C / 4;
  if (x_dt == DAFK_F32)
    bn_stats_kernel<float><<<bn_grid(n4), TPB, 2 * C * sizeof(float), as_stream(stream)>>>((const float*)x, acc, n4, C);
  else if (x_dt == DAFK_BF16)
    bn_stats_kernel<__nv_bfloat16><<<bn_grid(n4), TPB, 2 * C * sizeof(float), as_stream(stream)>>>((const __nv_bfloat16*)x, acc, n4, C);
  else { set_error("dafk_bn_stats: bad dtype"); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_bn_stats");
}

int dafk_bn_finalize(const double* acc, int64_t M, int C, float eps, float momentum, float* mean, float* rstd,
                     float* moving_mean, float* moving_var, void* stream) {
  DAFK_REQUIRE(M > 0 && C > 0 && acc && mean && rstd, DAFK_ERR_BAD_ARG, "dafk_bn_finalize: bad argument");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(acc, M, C, eps, momentum, mean, rstd,
                                                                   moving_mean, moving_var);
  return check_launch("dafk_bn_finalize");
}

int dafk_bn_rstd_from_var(const float* var, float* rstd, int C, float eps, void* stream) {
  DAFK_REQUIRE(C > 0 && var && rstd, DAFK_ERR_BAD_ARG, "dafk_bn_rstd_from_var: bad argument");
  bn_rstd_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(var, rstd, C, eps);
  return check_launch("dafk_bn_rstd_from_var");
}

int dafk_bn_apply(const void* x, int x_dt, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  void* out, int out_dt, int64_t M, int C, int act, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_apply: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(x && mean && rstd && gamma && beta && out, DAFK_ERR_BAD_ARG, "dafk_bn_apply: null pointer");
  DAFK_REQUIRE(act == DAFK_ACT_NONE || act == DAFK_ACT_RELU, DAFK_ERR_UNSUPPORTED, "dafk_bn_apply: act must be NONE or RELU");
  if (!bn_channels_ok(C) && x_dt == DAFK_F32 && out_dt == DAFK_F32) {
    bn_apply_any_kernel<<<bw_grid(M * C, TPB), TPB, 0, as_stream(stream)>>>((const float*)x, mean, rstd, gamma, beta,
                                                                           (float*)out, M * C, C, act);
    return check_launch("dafk_bn_apply(any C)");
  }
  DAFK_REQUIRE(bn_channels_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_apply: C must be a power of two in [4,1024] (got %d)", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(out), DAFK_ERR_ALIGN, "dafk_bn_apply: alignment");
  int64_t n4 = M * C / 4;
  cudaStream_t s = as_stream(stream);
  if (bn_wide_ok(C) && (x_dt == DAFK_F32 || x_dt == DAFK_BF16) && (out_dt == DAFK_F32 || out_dt == DAFK_BF16)) {
    const int64_t n8 = n4 / 2;
    const int grid = bn_wide_grid(n8, 1, 4);
    if (out_dt == DAFK_F32)
      { if (x_dt == DAFK_BF16) bn_apply_wide_kernel<__nv_bfloat16, float><<<grid, BW_T, 0, s>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (float*)out, n8, C, act);
        else bn_apply_wide_kernel<float, float><<<grid, BW_T, 0, s>>>((const float*)x, mean, rstd, gamma, beta, (float*)out, n8, C, act); }
    else
      { if (x_dt == DAFK_BF16) bn_apply_wide_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, BW_T, 0, s>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (__nv_bfloat16*)out, n8, C, act);
        else bn_apply_wide_kernel<float, __nv_bfloat16><<<grid, BW_T, 0, s>>>((const float*)x, mean, rstd, gamma, beta, (__nv_bfloat16*)out, n8, C, act); }
    return check_launch("dafk_bn_apply");
  }
  if (out_dt == DAFK_F32)
    { if (x_dt == DAFK_BF16) bn_apply_kernel<__nv_bfloat16, float><<<bn_grid(n4), TPB, 0, s>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (float*)out, n4, C, act);
      else bn_apply_kernel<float, float><<<bn_grid(n4), TPB, 0, s>>>((const float*)x, mean, rstd, gamma, beta, (float*)out, n4, C, act); }
  else if (out_dt == DAFK_BF16)
    { if (x_dt == DAFK_BF16) bn_apply_kernel<__nv_bfloat16, __nv_bfloat16><<<bn_grid(n4), TPB, 0, s>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (__nv_bfloat16*)out, n4, C, act);
      else bn_apply_kernel<float, __nv_bfloat16><<<bn_grid(n4), TPB, 0, s>>>((const float*)x, mean, rstd, gamma, beta, (__nv_bfloat16*)out, n4, C, act); }
  else { set_error("dafk_bn_apply: bad out dtype %d", out_dt); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_bn_apply");
}

int dafk_bn_bwd_reduce(const void* dout, int dout_dt, const void* x, int x_dt, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, double* acc, int64_t M, int C, int act, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_reduce: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(dout && x && mean && rstd && gamma && beta && acc, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_reduce: null pointer");
  DAFK_REQUIRE(bn_channels_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_bwd_reduce: unsupported C %d", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dout), DAFK_ERR_ALIGN, "dafk_bn_bwd_reduce: alignment");
  int64_t n4 = M * C / 4;
  cudaStream_t s = as_stream(stream);
  size_t smem = 2 * C * sizeof(float);
  if (dout_dt == DAFK_F32)
    { if (x_dt == DAFK_BF16) bn_bwd_reduce_kernel<float, __nv_bfloat16><<<bn_grid(n4), TPB, smem, s>>>((const float*)dout, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, acc, n4, C, act);
      else bn_bwd_reduce_kernel<float, float><<<bn_grid(n4), TPB, smem, s>>>((const float*)dout, (const float*)x, mean, rstd, gamma, beta, acc, n4, C, act); }
  else if (dout_dt == DAFK_BF16)
    { if (x_dt == DAFK_BF16) bn_bwd_reduce_kernel<__nv_bfloat16, __nv_bfloat16><<<bn_grid(n4), TPB, smem, s>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, acc, n4, C, act);
      else bn_bwd_reduce_kernel<__nv_bfloat16, float><<<bn_grid(n4), TPB, smem, s>>>((const __nv_bfloat16*)dout, (const float*)x, mean, rstd, gamma, beta, acc, n4, C, act); }
  else { set_error("dafk_bn_bwd_reduce: bad dtype"); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_bn_bwd_reduce");
}

int dafk_bn_bwd_apply(const void* dout, int dout_dt, const void* x, int x_dt, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, const double* acc, void* dx, int dx_dt, float* dgamma,
                      float* dbeta, float* dbias_prev, int64_t M, int C, int act, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_apply: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(dout && x && mean && rstd && gamma && beta && acc && dx, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_apply: null pointer");
  DAFK_REQUIRE(bn_channels_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_bwd_apply: unsupported C %d", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dout) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN, "dafk_bn_bwd_apply: alignment");
  int64_t n4 = M * C / 4;
  cudaStream_t s = as_stream(stream);
  int grid = bn_grid(n4);
  if (bn_wide_ok(C)) {
    const int64_t n8 = n4 / 2;
    const int wgrid = bn_wide_grid(n8, 1, 2);
#define LAUNCHW(TD, TO)                                                                                                 \
  do {                                                                                                                  \
    if (x_dt == DAFK_BF16)                                                                                              \
      bn_bwd_apply_wide_kernel<TD, __nv_bfloat16, TO><<<wgrid, BW_T, 0, s>>>(                                           \
          (const TD*)dout, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, acc, (TO*)dx, dgamma, dbeta, dbias_prev, n8, M, C, act); \
    else                                                                                                                \
      bn_bwd_apply_wide_kernel<TD, float, TO><<<wgrid, BW_T, 0, s>>>(                                                   \
          (const TD*)dout, (const float*)x, mean, rstd, gamma, beta, acc, (TO*)dx, dgamma, dbeta, dbias_prev, n8, M, C, act); \
  } while (0)
    if (dout_dt == DAFK_F32 && dx_dt == DAFK_F32) LAUNCHW(float, float);
    else if (dout_dt == DAFK_F32 && dx_dt == DAFK_BF16) LAUNCHW(float, __nv_bfloat16);
    else if (dout_dt == DAFK_BF16 && dx_dt == DAFK_F32) LAUNCHW(__nv_bfloat16, float);
    else if (dout_dt == DAFK_BF16 && dx_dt == DAFK_BF16) LAUNCHW(__nv_bfloat16, __nv_bfloat16);
    else { set_error("dafk_bn_bwd_apply: bad dtype"); return DAFK_ERR_BAD_ARG; }
#undef LAUNCHW
    return check_launch("dafk_bn_bwd_apply");
  }
#define LAUNCH(TD, TO)                                                                                                  \
  do {                                                                                                                  \
    if (x_dt == DAFK_BF16)                                                                                              \
      bn_bwd_apply_kernel<TD, __nv_bfloat16, TO><<<grid, TPB, C * sizeof(float), s>>>(                                  \
          (const TD*)dout, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, acc, (TO*)dx, dgamma, dbeta, dbias_prev, n4, M, C, act); \
    else                                                                                                                \
      bn_bwd_apply_kernel<TD, float, TO><<<grid, TPB, C * sizeof(float), s>>>(                                          \
          (const TD*)dout, (const float*)x, mean, rstd, gamma, beta, acc, (TO*)dx, dgamma, dbeta, dbias_prev, n4, M, C, act); \
  } while (0)
  if (dout_dt == DAFK_F32 && dx_dt == DAFK_F32) LAUNCH(float, float);
  else if (dout_dt == DAFK_F32 && dx_dt == DAFK_BF16) LAUNCH(float, __nv_bfloat16);
  else if (dout_dt == DAFK_BF16 && dx_dt == DAFK_F32) LAUNCH(__nv_bfloat16, float);
  else if (dout_dt == DAFK_BF16 && dx_dt == DAFK_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else { set_error("dafk_bn_bwd_apply: bad dtype"); return DAFK_ERR_BAD_ARG; }
#undef LAUNCH
  return check_launch("dafk_bn_bwd_apply");
}

int dafk_bn_wide_supported(int C) { return bn_wide_ok(C) ? 1 : 0; }
int64_t dafk_bn_wide_ws_bytes(int C) { return (int64_t)bn_wide_ws_bytes(C); }

int dafk_bn_stats_fused(const void* x, int x_dt, void* ws, int64_t ws_bytes, int64_t M, int C, float eps, float momentum,
                        float* mean, float* rstd, float* moving_mean, float* moving_var, void* stream) {
  DAFK_REQUIRE(M > 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_stats_fused: bad shape");
  DAFK_REQUIRE(bn_wide_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_stats_fused: C must be a power of two in [8,1024] (got %d)", C);
  DAFK_REQUIRE(x && ws && mean && rstd, DAFK_ERR_BAD_ARG, "dafk_bn_stats_fused: null pointer");
  DAFK_REQUIRE(ws_bytes >= (int64_t)bn_wide_ws_bytes(C), DAFK_ERR_BAD_ARG, "dafk_bn_stats_fused: workspace too small");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(ws), DAFK_ERR_ALIGN, "dafk_bn_stats_fused: alignment");
  const int64_t n8 = M * C / 8;
  unsigned* ticket = (unsigned*)ws;
  double* wacc = (double*)((char*)ws + 16);
  cudaStream_t s = as_stream(stream);
  const int grid = bn_wide_grid(n8, 2, 4);
  if (x_dt == DAFK_F32)
    bn_stats_wide_kernel<float><<<grid, BW_T, 0, s>>>((const float*)x, ticket, wacc, n8, C, M, eps, momentum, mean, rstd, moving_mean, moving_var);
  else if (x_dt == DAFK_BF16)
    bn_stats_wide_kernel<__nv_bfloat16><<<grid, BW_T, 0, s>>>((const __nv_bfloat16*)x, ticket, wacc, n8, C, M, eps, momentum, mean, rstd, moving_mean, moving_var);
  else { set_error("dafk_bn_stats_fused: bad dtype"); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_bn_stats_fused");
}

int dafk_bn_bwd_reduce_fused(const void* dout, int dout_dt, const void* x, int x_dt, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, double* acc, void* ws, int64_t ws_bytes, int64_t M,
                             int C, int act, void* stream) {
  DAFK_REQUIRE(M > 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_reduce_fused: bad shape");
  DAFK_REQUIRE(bn_wide_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_bwd_reduce_fused: C must be a power of two in [8,1024] (got %d)", C);
  DAFK_REQUIRE(dout && x && mean && rstd && gamma && beta && acc && ws, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_reduce_fused: null pointer");
  DAFK_REQUIRE(ws_bytes >= (int64_t)bn_wide_ws_bytes(C), DAFK_ERR_BAD_ARG, "dafk_bn_bwd_reduce_fused: workspace too small");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dout) && DAFK_ALIGNED16(ws), DAFK_ERR_ALIGN, "dafk_bn_bwd_reduce_fused: alignment");
  const int64_t n8 = M * C / 8;
  unsigned* ticket = (unsigned*)ws;
  double* wacc = (double*)((char*)ws + 16);
  cudaStream_t s = as_stream(stream);
  const int grid = bn_wide_grid(n8, 1, 2);
  if (dout_dt == DAFK_F32)
    { if (x_dt == DAFK_BF16) bn_bwd_reduce_wide_kernel<float, __nv_bfloat16><<<grid, BW_T, 0, s>>>((const float*)dout, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, ticket, wacc, acc, n8, C, act);
      else bn_bwd_reduce_wide_kernel<float, float><<<grid, BW_T, 0, s>>>((const float*)dout, (const float*)x, mean, rstd, gamma, beta, ticket, wacc, acc, n8, C, act); }
  else if (dout_dt == DAFK_BF16)
    { if (x_dt == DAFK_BF16) bn_bwd_reduce_wide_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, BW_T, 0, s>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, ticket, wacc, acc, n8, C, act);
      else bn_bwd_reduce_wide_kernel<__nv_bfloat16, float><<<grid, BW_T, 0, s>>>((const __nv_bfloat16*)dout, (const float*)x, mean, rstd, gamma, beta, ticket, wacc, acc, n8, C, act); }
  else { set_error("dafk_bn_bwd_reduce_fused: bad dtype"); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_bn_bwd_reduce_fused");
}

int dafk_bn_bwd_frozen(const float* dout, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, float* dx, int64_t M, int C, int act, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_frozen: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(dout && x && mean && rstd && gamma && beta && dx, DAFK_ERR_BAD_ARG, "dafk_bn_bwd_frozen: null pointer");
  DAFK_REQUIRE(bn_channels_ok(C), DAFK_ERR_UNSUPPORTED, "dafk_bn_bwd_frozen: unsupported C %d", C);
  int64_t n4 = M * C / 4;
  bn_bwd_frozen_kernel<<<bn_grid(n4), TPB, 0, as_stream(stream)>>>(dout, x, mean, rstd, gamma, beta, dx, n4, C, act);
  return check_launch("dafk_bn_bwd_frozen");
}

int dafk_in_stats(const float* x, double* acc, int B, int64_t HWC, void* stream) {
  DAFK_REQUIRE(B >= 0 && HWC >= 0, DAFK_ERR_BAD_ARG, "dafk_in_stats: bad shape");
  if (B == 0 || HWC == 0) return DAFK_OK;
  DAFK_REQUIRE(x && acc, DAFK_ERR_BAD_ARG, "dafk_in_stats: null pointer");
  DAFK_REQUIRE(HWC % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_in_stats: H*W*C must be a multiple of 4");
  DAFK_REQUIRE(DAFK_ALIGNED16(x), DAFK_ERR_ALIGN, "dafk_in_stats: alignment");
  in_stats_kernel<<<dim3(per_sample_chunks(HWC / 4, B), B), TPB, 0, as_stream(stream)>>>(x, acc, HWC);
  return check_launch("dafk_in_stats");
}

int dafk_spade_fwd(const float* x, const double* acc, const float* gamma, const float* beta, float* y, int B,
                   int64_t HWC, float eps, int act, float alpha, void* stream) {
  DAFK_REQUIRE(B >= 0 && HWC >= 0, DAFK_ERR_BAD_ARG, "dafk_spade_fwd: bad shape");
  if (B == 0 || HWC == 0) return DAFK_OK;
  DAFK_REQUIRE(x && acc && gamma && beta && y, DAFK_ERR_BAD_ARG, "dafk_spade_fwd: null pointer");
  DAFK_REQUIRE(HWC % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_spade_fwd: H*W*C must be a multiple of 4");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(beta) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN,
               "dafk_spade_fwd: alignment");
  spade_fwd_kernel<<<dim3(per_sample_chunks(HWC / 4, B), B), TPB, 0, as_stream(stream)>>>(x, acc, gamma, beta, y, HWC,
                                                                                         eps, act, alpha);
  return check_launch("dafk_spade_fwd");
}

int dafk_spade_bwd(const float* dy, const float* x, const double* acc, const float* gamma, const float* beta,
                   float* dx, float* dgamma, float* dbeta, double* ws, int B, int64_t HWC, float eps, int act,
                   float alpha, void* stream) {
  DAFK_REQUIRE(B >= 0 && HWC >= 0, DAFK_ERR_BAD_ARG, "dafk_spade_bwd: bad shape");
  if (B == 0 || HWC == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && x && acc && gamma && beta && dx && dgamma && dbeta && ws, DAFK_ERR_BAD_ARG, "dafk_spade_bwd: null pointer");
  DAFK_REQUIRE(HWC % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_spade_bwd: H*W*C must be a multiple of 4");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B, s);
  dim3 grid(per_sample_chunks(HWC / 4, B), B);
  spade_bwd1_kernel<<<grid, TPB, 0, s>>>(dy, x, acc, gamma, beta, dgamma, dbeta, ws, HWC, eps, act, alpha);
  int rc = check_launch("dafk_spade_bwd(1)");
  if (rc) return rc;
  spade_bwd2_kernel<<<grid, TPB, 0, s>>>(dbeta, x, acc, gamma, ws, dx, HWC, eps);
  return check_launch("dafk_spade_bwd(2)");
}

int dafk_spade_cond_fwd(const float* x, const float* gamma, const float* beta, float* y, int64_t n, void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_spade_cond_fwd: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(x && gamma && beta && y, DAFK_ERR_BAD_ARG, "dafk_spade_cond_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(beta) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN,
               "dafk_spade_cond_fwd: alignment");
  spade_cond_fwd_kernel<<<bw_grid((n + 3) / 4, TPB), TPB, 0, as_stream(stream)>>>(x, gamma, beta, y, n);
  return check_launch("dafk_spade_cond_fwd");
}

int dafk_spade_cond_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, int64_t n,
                        void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_spade_cond_bwd: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && x && gamma && dx && dgamma, DAFK_ERR_BAD_ARG, "dafk_spade_cond_bwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(x) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(dx) &&
               DAFK_ALIGNED16(dgamma), DAFK_ERR_ALIGN, "dafk_spade_cond_bwd: alignment");
  spade_cond_bwd_kernel<<<bw_grid((n + 3) / 4, TPB), TPB, 0, as_stream(stream)>>>(dy, x, gamma, dx, dgamma, n);
  return check_launch("dafk_spade_cond_bwd");
}

int dafk_in_affine_fwd(const float* x, const double* acc, const float* gamma, const float* beta, float* y, int B,
                       int64_t HWC, float eps, int act, float alpha, void* stream) {
  DAFK_REQUIRE(B >= 0 && HWC >= 0, DAFK_ERR_BAD_ARG, "dafk_in_affine_fwd: bad shape");
  if (B == 0 || HWC == 0) return DAFK_OK;
  DAFK_REQUIRE(x && acc && gamma && beta && y, DAFK_ERR_BAD_ARG, "dafk_in_affine_fwd: null pointer");
  DAFK_REQUIRE(HWC % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_in_affine_fwd: H*W*C must be a multiple of 4");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "dafk_in_affine_fwd: alignment");
  in_affine_fwd_kernel<<<dim3(per_sample_chunks(HWC / 4, B), B), TPB, 0, as_stream(stream)>>>(x, acc, gamma, beta, y,
                                                                                             HWC, eps, act, alpha);
  return check_launch("dafk_in_affine_fwd");
}

int dafk_in_affine_bwd(const float* dy, const float* x, const double* acc, const float* gamma, const float* beta,
                       float* dx, float* dgamma, float* dbeta, double* ws, int B, int64_t HWC, float eps, int act,
                       float alpha, void* stream) {
  DAFK_REQUIRE(B >= 0 && HWC >= 0, DAFK_ERR_BAD_ARG, "dafk_in_affine_bwd: bad shape");
  if (B == 0 || HWC == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && x && acc && gamma && beta && dx && ws, DAFK_ERR_BAD_ARG, "dafk_in_affine_bwd: null pointer");
  DAFK_REQUIRE(HWC % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_in_affine_bwd: H*W*C must be a multiple of 4");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN, "dafk_in_affine_bwd: alignment");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * (2 * B + 2), s);
  dim3 grid(per_sample_chunks(HWC / 4, B), B);
  in_affine_bwd1_kernel<<<grid, TPB, 0, s>>>(dy, x, acc, gamma, beta, dx, ws, ws + 2 * B, HWC, eps, act, alpha);
  int rc = check_launch("dafk_in_affine_bwd(1)");
  if (rc) return rc;
  in_affine_bwd2_kernel<<<grid, TPB, 0, s>>>(dx, x, acc, gamma, ws, ws + 2 * B, dgamma, dbeta, HWC, eps);
  return check_launch("dafk_in_affine_bwd(2)");
}

int dafk_pair_dice(const float* a, const float* b, float* out, double* ws, int B, int64_t HWC, void* stream) {
  DAFK_REQUIRE(B >= 0 && HWC >= 0, DAFK_ERR_BAD_ARG, "dafk_pair_dice: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(a && b && out && ws, DAFK_ERR_BAD_ARG, "dafk_pair_dice: null pointer");
  DAFK_REQUIRE(HWC % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_pair_dice: H*W*C must be a multiple of 4");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * 3 * B, s);
  pair_dice_kernel<<<dim3(per_sample_chunks(HWC / 4, B), B), TPB, 0, s>>>(a, b, ws, HWC);
  int rc = check_launch("dafk_pair_dice");
  if (rc) return rc;
  pair_dice_finish_kernel<<<(B + 127) / 128, 128, 0, s>>>(ws, out, B);
  return check_launch("dafk_pair_dice(finish)");
}

}  // extern "C"
