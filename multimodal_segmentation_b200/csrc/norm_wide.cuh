// Wide BatchNorm kernels: 8 channels (one 128-bit load of bf16, two of f32) per thread, 512-thread CTAs, four
// independent loads in flight per thread (64 KB per SM), one CTA per SM (two for the statistics pass).
// Used whenever C is a power of two in [8, 1024] (every BatchNorm of the UNet / segmentor: 64 ... 1024 channels);
// the 4-wide kernels of norm.cu remain for the other channel counts.
//
// Each thread owns the fixed channel group c0 = (8 * threadIdx.x) % C (4096 % C == 0).  Cross-thread reduction:
// registers -> shared-memory transpose [16][512] -> per-channel sums -> one double atomicAdd per channel per CTA into a
// persistent, self-resetting workspace; the last CTA to arrive (ticket counter) finishes the job in the same launch:
// statistics -> mean / rstd / moving averages (no separate finalize launch), backward sums -> `acc`.  The workspace is
// zero before and after every launch, so no memset launches are needed either.
#pragma once
#include "common.cuh"

namespace dafk {

constexpr int BW_T = 512;

static inline bool bn_wide_ok(int C) { return C >= 8 && C <= 1024 && (C & (C - 1)) == 0; }
// workspace: [0] ticket counter (uint32, padded to 16 B) | 2*C doubles
static inline size_t bn_wide_ws_bytes(int C) { return 16 + sizeof(double) * 2 * (size_t)C; }

template <typename T> struct W8;     // raw 8-channel vector
template <> struct W8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
  }
};
template <> struct W8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = ldg_stream4(p); b = ldg_stream4(p + 4); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    stg_stream4(p, make_float4(v[0], v[1], v[2], v[3]));
    stg_stream4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
};

// per-channel sums of a[8], b[8] over the CTA -> double atomics into ws_acc[0:C], ws_acc[C:2C]; returns true in
// the LAST CTA of the grid (all of its threads), after every CTA's contribution is visible
__device__ __forceinline__ bool bn_wide_reduce(const float (&a)[8], const float (&b)[8], int C, float* red /*[16*BW_T]*/,
                                               unsigned* ticket, double* ws_acc) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[k * BW_T + threadIdx.x] = a[k];
    red[(8 + k) * BW_T + threadIdx.x] = b[k];
  }
  __syncthreads();
  const int G = C >> 3;                 // channel groups; threads tid, tid+G, tid+2G ... share a group
  for (int o = threadIdx.x; o < 2 * C; o += BW_T) {
    const int which = o >= C ? 1 : 0;
    const int c = o - which * C;
    const float* row = red + (which * 8 + (c & 7)) * BW_T + (c >> 3);
    float s = 0.f;
    for (int j = 0; j < BW_T; j += G) s += row[j];
    atomicAdd(ws_acc + o, (double)s);
  }
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

template <typename TX>
__global__ void __launch_bounds__(BW_T, 2) bn_stats_wide_kernel(const TX* __restrict__ x, unsigned* ticket,
                                                                double* ws_acc, int64_t n8, int C, int64_t M, float eps,
                                                                float momentum, float* __restrict__ mean,
                                                                float* __restrict__ rstd, float* __restrict__ mm,
                                                                float* __restrict__ mv) {
  __shared__ float red[16 * BW_T];
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
  const int64_t stride = (int64_t)gridDim.x * BW_T;
  int64_t i = (int64_t)blockIdx.x * BW_T + threadIdx.x;
  for (; i + 3 * stride < n8; i += 4 * stride) {
    W8<TX> r0, r1, r2, r3;
    r0.load(x + 8 * i); r1.load(x + 8 * (i + stride)); r2.load(x + 8 * (i + 2 * stride)); r3.load(x + 8 * (i + 3 * stride));
    float v[8];
    r0.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    r1.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    r2.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    r3.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
  }
  for (; i < n8; i += stride) {
    W8<TX> r0;
    r0.load(x + 8 * i);
    float v[8];
    r0.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
  }
  if (!bn_wide_reduce(s, q, C, red, ticket, ws_acc)) return;
  // last CTA: finalize (keras BatchNormalization: biased batch variance; the moving variance receives the
  // Bessel-corrected one, as on TF's fused path) and reset the workspace
  for (int c = threadIdx.x; c < C; c += BW_T) {
    const double sa = __ldcg(ws_acc + c), sq = __ldcg(ws_acc + C + c);
    ws_acc[c] = 0.0;
    ws_acc[C + c] = 0.0;
    const double m = sa / (double)M;
    double var = sq / (double)M - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (mm && mv) {
      const double unb = (M > 1) ? var * ((double)M / (double)(M - 1)) : var;
      mm[c] = mm[c] * momentum + (float)m * (1.f - momentum);
      mv[c] = mv[c] * momentum + (float)unb * (1.f - momentum);
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

struct BnP8 {
  float mu[8], rs[8], g[8], b[8];
};
__device__ __forceinline__ void bn_load_p8(BnP8& P, const float* mean, const float* rstd, const float* gamma,
                                           const float* beta, int c0) {
#pragma unroll
  for (int k = 0; k < 8; ++k) { P.mu[k] = mean[c0 + k]; P.rs[k] = rstd[c0 + k]; P.g[k] = gamma[c0 + k]; P.b[k] = beta[c0 + k]; }
}

template <typename TX, typename TO>
__global__ void __launch_bounds__(BW_T, 1) bn_apply_wide_kernel(const TX* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, TO* __restrict__ out,
                                                                int64_t n8, int C, int act) {
  BnP8 P;
  bn_load_p8(P, mean, rstd, gamma, beta, (threadIdx.x * 8) % C);
  const int64_t stride = (int64_t)gridDim.x * BW_T;
  int64_t i = (int64_t)blockIdx.x * BW_T + threadIdx.x;
  auto one = [&](const W8<TX>& r, int64_t at) {
    float v[8];
    r.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float z = (v[k] - P.mu[k]) * P.rs[k] * P.g[k] + P.b[k];
      v[k] = (act == DAFK_ACT_RELU) ? fmaxf(z, 0.f) : z;
    }
    W8<TO>::store(out + 8 * at, v);
  };
  for (; i + 3 * stride < n8; i += 4 * stride) {
    W8<TX> r0, r1, r2, r3;
    r0.load(x + 8 * i); r1.load(x + 8 * (i + stride)); r2.load(x + 8 * (i + 2 * stride)); r3.load(x + 8 * (i + 3 * stride));
    one(r0, i); one(r1, i + stride); one(r2, i + 2 * stride); one(r3, i + 3 * stride);
  }
  for (; i < n8; i += stride) {
    W8<TX> r0;
    r0.load(x + 8 * i);
    one(r0, i);
  }
}

template <typename TD, typename TX>
__global__ void __launch_bounds__(BW_T, 1) bn_bwd_reduce_wide_kernel(const TD* __restrict__ dout, const TX* __restrict__ x,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, unsigned* ticket,
                                                                     double* ws_acc, double* __restrict__ acc, int64_t n8,
                                                                     int C, int act) {
  __shared__ float red[16 * BW_T];
  BnP8 P;
  bn_load_p8(P, mean, rstd, gamma, beta, (threadIdx.x * 8) % C);
  float s0[8], s1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s0[k] = 0.f; s1[k] = 0.f; }
  const int64_t stride = (int64_t)gridDim.x * BW_T;
  int64_t i = (int64_t)blockIdx.x * BW_T + threadIdx.x;
  auto one = [&](const W8<TX>& rx, const W8<TD>& rd) {
    float v[8], d[8];
    rx.get(v);
    rd.get(d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (v[k] - P.mu[k]) * P.rs[k];
      const float z = xh * P.g[k] + P.b[k];
      const float dz = (act == DAFK_ACT_RELU && !(z > 0.f)) ? 0.f : d[k];
      s0[k] += dz;
      s1[k] = fmaf(dz, xh, s1[k]);
    }
  };
  for (; i + stride < n8; i += 2 * stride) {
    W8<TX> x0, x1;
    W8<TD> d0, d1;
    x0.load(x + 8 * i); d0.load(dout + 8 * i); x1.load(x + 8 * (i + stride)); d1.load(dout + 8 * (i + stride));
    one(x0, d0);
    one(x1, d1);
  }
  for (; i < n8; i += stride) {
    W8<TX> x0;
    W8<TD> d0;
    x0.load(x + 8 * i); d0.load(dout + 8 * i);
    one(x0, d0);
  }
  if (!bn_wide_reduce(s0, s1, C, red, ticket, ws_acc)) return;
  for (int o = threadIdx.x; o < 2 * C; o += BW_T) {
    acc[o] = __ldcg(ws_acc + o);
    ws_acc[o] = 0.0;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

template <typename TD, typename TX, typename TO>
__global__ void __launch_bounds__(BW_T, 1) bn_bwd_apply_wide_kernel(const TD* __restrict__ dout, const TX* __restrict__ x,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta,
                                                                    const double* __restrict__ acc, TO* __restrict__ dx,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                    float* __restrict__ dbias, int64_t n8, int64_t M, int C,
                                                                    int act) {
  __shared__ float red[8 * BW_T];
  const int c0 = (threadIdx.x * 8) % C;
  BnP8 P;
  bn_load_p8(P, mean, rstd, gamma, beta, c0);
  float m0[8], m1[8], sdx[8];
  const double invM = 1.0 / (double)M;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m0[k] = (float)(acc[c0 + k] * invM);
    m1[k] = (float)(acc[C + c0 + k] * invM);
    sdx[k] = 0.f;
  }
  if (blockIdx.x == 0 && dgamma && dbeta) {
    for (int c = threadIdx.x; c < C; c += BW_T) {
      dbeta[c] += (float)acc[c];
      dgamma[c] += (float)acc[C + c];
    }
  }
  const int64_t stride = (int64_t)gridDim.x * BW_T;
  int64_t i = (int64_t)blockIdx.x * BW_T + threadIdx.x;
  auto one = [&](const W8<TX>& rx, const W8<TD>& rd, int64_t at) {
    float v[8], d[8];
    rx.get(v);
    rd.get(d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (v[k] - P.mu[k]) * P.rs[k];
      const float z = xh * P.g[k] + P.b[k];
      const float dz = (act == DAFK_ACT_RELU && !(z > 0.f)) ? 0.f : d[k];
      v[k] = P.g[k] * P.rs[k] * (dz - m0[k] - xh * m1[k]);
      sdx[k] += v[k];
    }
    W8<TO>::store(dx + 8 * at, v);
  };
  for (; i + stride < n8; i += 2 * stride) {
    W8<TX> x0, x1;
    W8<TD> d0, d1;
    x0.load(x + 8 * i); d0.load(dout + 8 * i); x1.load(x + 8 * (i + stride)); d1.load(dout + 8 * (i + stride));
    one(x0, d0, i);
    one(x1, d1, i + stride);
  }
  for (; i < n8; i += stride) {
    W8<TX> x0;
    W8<TD> d0;
    x0.load(x + 8 * i); d0.load(dout + 8 * i);
    one(x0, d0, i);
  }
  if (dbias) {
    // gradient of the producing convolution's bias = per-channel sum of dx
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k * BW_T + threadIdx.x] = sdx[k];
    __syncthreads();
    const int G = C >> 3;
    for (int c = threadIdx.x; c < C; c += BW_T) {
      const float* row = red + (c & 7) * BW_T + (c >> 3);
      float s = 0.f;
      for (int j = 0; j < BW_T; j += G) s += row[j];
      atomicAdd(dbias + c, s);
    }
  }
}

// out[c] += sum over rows of x[row][c] (bias gradients of the convolutions, models/discriminator.py:24,39): same
// 8-channels-per-thread stream as the statistics pass, one float atomic per channel per CTA
template <typename TX>
__global__ void __launch_bounds__(BW_T, 2) colsum_wide_kernel(const TX* __restrict__ x, float* __restrict__ out,
                                                              int64_t n8, int C) {
  __shared__ float red[8 * BW_T];
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  const int64_t stride = (int64_t)gridDim.x * BW_T;
  int64_t i = (int64_t)blockIdx.x * BW_T + threadIdx.x;
  for (; i + 3 * stride < n8; i += 4 * stride) {
    W8<TX> r0, r1, r2, r3;
    r0.load(x + 8 * i); r1.load(x + 8 * (i + stride)); r2.load(x + 8 * (i + 2 * stride)); r3.load(x + 8 * (i + 3 * stride));
    float v[8];
    r0.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
    r1.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
    r2.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
    r3.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
  }
  for (; i < n8; i += stride) {
    W8<TX> r0;
    r0.load(x + 8 * i);
    float v[8];
    r0.get(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[k * BW_T + threadIdx.x] = s[k];
  __syncthreads();
  const int G = C >> 3;
  for (int c = threadIdx.x; c < C; c += BW_T) {
    const float* row = red + (c & 7) * BW_T + (c >> 3);
    float t = 0.f;
    for (int j = 0; j < BW_T; j += G) t += row[j];
    atomicAdd(out + c, t);
  }
}

static inline int bn_wide_grid(int64_t n8, int per_sm, int unroll) {
  int64_t need = (n8 + (int64_t)BW_T * unroll - 1) / ((int64_t)BW_T * unroll);
  int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace dafk
