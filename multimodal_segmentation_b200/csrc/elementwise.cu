// HBM-bandwidth-bound pointwise kernels: rounding, softmax, activations, FiLM, Maximum,
// channel copies.  All use 128-bit accesses on the main body, a scalar tail, and
// grid-stride loops over a grid sized to a multiple of the SM count.
#include "common.cuh"
#include "reduce.cuh"

namespace dafk {

constexpr int TPB = 256;

// ---------------------------------------------------------------- generic maps
template <typename F>
__global__ void __launch_bounds__(TPB) map1_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                   int64_t n, F f) {
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldg_stream4(x + 4 * i);
    v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
    stg_stream4(y + 4 * i, v);
  }
  // tail
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) y[t] = f(x[t]);
}

template <typename F>
__global__ void __launch_bounds__(TPB) map2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ y, int64_t n, F f) {
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 u = *reinterpret_cast<const float4*>(a + 4 * i);
    float4 v = *reinterpret_cast<const float4*>(b + 4 * i);
    float4 r;
    r.x = f(u.x, v.x); r.y = f(u.y, v.y); r.z = f(u.z, v.z); r.w = f(u.w, v.w);
    *reinterpret_cast<float4*>(y + 4 * i) = r;
  }
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) y[t] = f(a[t], b[t]);
}

struct RoundOp { __device__ float operator()(float v) const { return rintf(v); } };
struct ActFwd {
  int act; float alpha;
  __device__ float operator()(float v) const {
    if (act == DAFK_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == DAFK_ACT_LRELU) return v > 0.f ? v : alpha * v;   // relu(x) - a*relu(-x)
    if (act == DAFK_ACT_TANH) return tanhf(v);
    return v;
  }
};
struct ActBwd {  // (dy, y) -> dx, derivative at exactly 0 is 0 for relu/lrelu (Keras 2.1.6)
  int act; float alpha;
  __device__ float operator()(float dy, float y) const {
    if (act == DAFK_ACT_RELU) return y > 0.f ? dy : 0.f;
    if (act == DAFK_ACT_LRELU) return y > 0.f ? dy : (y < 0.f ? alpha * dy : 0.f);
    if (act == DAFK_ACT_TANH) return dy * (1.f - y * y);
    return dy;
  }
};
// (dy, y) -> bf16(dx): the activation backward of a tensor-core layer, written directly in the operand dtype of the
// weight- and data-gradient kernels that consume it (no fp32 intermediate)
__global__ void __launch_bounds__(TPB) act_bwd_bf16_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           __nv_bfloat16* __restrict__ dx, int64_t n4, ActBwd f) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 u = ldg_stream4(dy + 4 * i);
    float4 v = ldg_stream4(y + 4 * i);
    float r[4] = {f(u.x, v.x), f(u.y, v.y), f(u.z, v.z), f(u.w, v.w)};
    Vec4<__nv_bfloat16>::store(dx + 4 * i, r);
  }
}
// (dy1 + dy2) * act'(y): a feature map with two consumers (the residual blocks of the FiLM decoder,
// model_components/decoder.py:44-54) receives two gradients; summing them while the activation backward runs saves
// one pass over the map (read 3, write 1 instead of read 4, write 2)
__global__ void __launch_bounds__(TPB) add_act_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          const float* __restrict__ y, float* __restrict__ dx, int64_t n4,
                                                          ActBwd f) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 u = ldg_stream4(a + 4 * i);
    const float4 w = ldg_stream4(b + 4 * i);
    const float4 v = ldg_stream4(y + 4 * i);
    stg_stream4(dx + 4 * i, make_float4(f(u.x + w.x, v.x), f(u.y + w.y, v.y), f(u.z + w.z, v.z), f(u.w + w.w, v.w)));
  }
}
struct AddOp { __device__ float operator()(float a, float b) const { return a + b; } };
struct MaxOp { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct AxpbyOp { float a, b; __device__ float operator()(float x, float y) const { return a * x + b * y; } };
struct FillOp { float v; __device__ float operator()(float) const { return v; } };

template <typename F>
static int launch_map1(const float* x, float* y, int64_t n, F f, void* stream, const char* name) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "%s: negative size", name);
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "%s: null pointer", name);
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "%s: pointers must be 16-byte aligned", name);
  int grid = bw_grid((n + 3) / 4, TPB);
  map1_kernel<<<grid, TPB, 0, as_stream(stream)>>>(x, y, n, f);
  return check_launch(name);
}
template <typename F>
static int launch_map2(const float* a, const float* b, float* y, int64_t n, F f, void* stream, const char* name) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "%s: negative size", name);
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(a && b && y, DAFK_ERR_BAD_ARG, "%s: null pointer", name);
  DAFK_REQUIRE(DAFK_ALIGNED16(a) && DAFK_ALIGNED16(b) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN,
               "%s: pointers must be 16-byte aligned", name);
  int grid = bw_grid((n + 3) / 4, TPB);
  map2_kernel<<<grid, TPB, 0, as_stream(stream)>>>(a, b, y, n, f);
  return check_launch(name);
}

// ---------------------------------------------------------------- Maximum backward
__global__ void __launch_bounds__(TPB) max_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      const float* __restrict__ g, float* __restrict__ da,
                                                      float* __restrict__ db, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float av = a[i], bv = b[i], gv = g[i];
    bool first = av >= bv;   // tf.maximum gradient: ties go to the first input
    da[i] = first ? gv : 0.f;
    db[i] = first ? 0.f : gv;
  }
}

// ---------------------------------------------------------------- softmax over the last axis
template <int C>
__global__ void __launch_bounds__(TPB) softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ p,
                                                          float* __restrict__ r, int64_t M) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    float v[C];
    if constexpr (C % 4 == 0) {
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        float4 t = *reinterpret_cast<const float4*>(x + m * C + c);
        v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = x[m * C + c];
    }
    float mx = v[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, v[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = v[c] / s;
    if constexpr (C % 4 == 0) {
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        *reinterpret_cast<float4*>(p + m * C + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        if (r) *reinterpret_cast<float4*>(r + m * C + c) =
            make_float4(rintf(v[c]), rintf(v[c + 1]), rintf(v[c + 2]), rintf(v[c + 3]));
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) { p[m * C + c] = v[c]; if (r) r[m * C + c] = rintf(v[c]); }
    }
  }
}

template <int C>
__global__ void __launch_bounds__(TPB) softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                          float* __restrict__ dx, int64_t M) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    float pv[C], gv[C];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { pv[c] = p[m * C + c]; gv[c] = dp[m * C + c]; dot += pv[c] * gv[c]; }
#pragma unroll
    for (int c = 0; c < C; ++c) dx[m * C + c] = pv[c] * (gv[c] - dot);
  }
}

// ---------------------------------------------------------------- cast
template <typename TI, typename TO>
__global__ void __launch_bounds__(TPB) cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    Vec4<TI>::load(x + 4 * i, v);
    Vec4<TO>::store(y + 4 * i, v);
  }
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) y[t] = from_f<TO>(to_f<TI>(x[t]));
}

template <typename T>
__global__ void __launch_bounds__(TPB) add_dt_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n) {
  int64_t n4 = n >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float u[4], v[4];
    Vec4<T>::load(a + 4 * i, u);
    Vec4<T>::load(b + 4 * i, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] += v[k];
    Vec4<T>::store(y + 4 * i, u);
  }
  int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) y[t] = from_f<T>(to_f<T>(a[t]) + to_f<T>(b[t]));
}

// ---------------------------------------------------------------- channel copies
__global__ void __launch_bounds__(TPB) copy_channels_kernel(const float* __restrict__ src, int src_c, int src_off,
                                                            float* __restrict__ dst, int dst_c, int dst_off, int c,
                                                            int64_t M, int accumulate) {
  int64_t total = M * c;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t m = i / c;
    int ch = (int)(i - m * c);
    float v = src[m * src_c + src_off + ch];
    float* d = dst + m * dst_c + dst_off + ch;
    *d = accumulate ? (*d + v) : v;
  }
}

__global__ void __launch_bounds__(TPB) gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          float* __restrict__ dst, int64_t rows, int64_t row_elems) {
  // row_elems % 4 == 0 (checked by the host wrapper)
  int64_t per_row = row_elems >> 2;
  int64_t total = rows * per_row;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t r = i / per_row, k = i - r * per_row;
    int64_t s = idx[r];
    *reinterpret_cast<float4*>(dst + r * row_elems + 4 * k) =
        *reinterpret_cast<const float4*>(src + s * row_elems + 4 * k);
  }
}

// ---------------------------------------------------------------- FiLM
// x:[B,HW,C]; one float4 per thread-iteration; C % 4 == 0.
// optional fused tail of the decoder's FiLM layer (model_components/decoder.py:50-54): y = res + lrelu(x*gamma + beta)
__device__ __forceinline__ float film_act(float z, int act, float alpha) {
  // __fmul_rn: the product is rounded on its own (as in the stand-alone activation kernel), never contracted into the
  // residual add that follows
  return (act == DAFK_ACT_LRELU) ? (z > 0.f ? z : __fmul_rn(alpha, z)) : ((act == DAFK_ACT_RELU) ? fmaxf(z, 0.f) : z);
}
__device__ __forceinline__ float film_act_grad(float z, int act, float alpha) {
  if (act == DAFK_ACT_LRELU) return z > 0.f ? 1.f : (z < 0.f ? alpha : 0.f);
  if (act == DAFK_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}
__global__ void __launch_bounds__(TPB) film_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ res,
                                                       float* __restrict__ y, int64_t HWC, int C, int64_t n4, int act,
                                                       float alpha) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int64_t e = i << 2;
    int64_t b = e / HWC;
    int c = (int)(e % C);
    float4 v = ldg_stream4(x + e);
    float4 g = *reinterpret_cast<const float4*>(gamma + b * C + c);
    float4 t = *reinterpret_cast<const float4*>(beta + b * C + c);
    v.x = film_act(v.x * g.x + t.x, act, alpha); v.y = film_act(v.y * g.y + t.y, act, alpha);
    v.z = film_act(v.z * g.z + t.z, act, alpha); v.w = film_act(v.w * g.w + t.w, act, alpha);
    if (res) {
      float4 r = ldg_stream4(res + e);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    stg_stream4(y + e, v);
  }
}

// grid = (chunks, B).  Each thread owns a fixed group of 4 channels.
__global__ void __launch_bounds__(TPB) film_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ dx, double* __restrict__ ws, int64_t HWC, int C,
                                                       int act, float alpha) {
  extern __shared__ float sm[];
  int b = blockIdx.y;
  const float* dyb = dy + (int64_t)b * HWC;
  const float* xb = x + (int64_t)b * HWC;
  float* dxb = dx + (int64_t)b * HWC;
  int c = (threadIdx.x * 4) % C;
  float4 g = *reinterpret_cast<const float4*>(gamma + (int64_t)b * C + c);
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act != DAFK_ACT_NONE) t = *reinterpret_cast<const float4*>(beta + (int64_t)b * C + c);
  float ag[4] = {0, 0, 0, 0}, ab[4] = {0, 0, 0, 0};
  int64_t n4 = HWC >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 d = ldg_stream4(dyb + 4 * i);
    float4 v = ldg_stream4(xb + 4 * i);
    if (act != DAFK_ACT_NONE) {      // gradient of the fused activation at z = x*gamma + beta
      d.x *= film_act_grad(v.x * g.x + t.x, act, alpha); d.y *= film_act_grad(v.y * g.y + t.y, act, alpha);
      d.z *= film_act_grad(v.z * g.z + t.z, act, alpha); d.w *= film_act_grad(v.w * g.w + t.w, act, alpha);
    }
    ag[0] += d.x * v.x; ag[1] += d.y * v.y; ag[2] += d.z * v.z; ag[3] += d.w * v.w;
    ab[0] += d.x; ab[1] += d.y; ab[2] += d.z; ab[3] += d.w;
    d.x *= g.x; d.y *= g.y; d.z *= g.z; d.w *= g.w;
    stg_stream4(dxb + 4 * i, d);
  }
  channel_reduce2<TPB>(ag, ab, C, sm, ws + (int64_t)b * 2 * C, ws + (int64_t)b * 2 * C + C);
}

__global__ void film_bwd_finish_kernel(const double* __restrict__ ws, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int B, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * C) {
    int b = i / C, c = i % C;
    dgamma[i] = (float)ws[(int64_t)b * 2 * C + c];
    dbeta[i] = (float)ws[(int64_t)b * 2 * C + C + c];
  }
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_round_fwd(const float* x, float* y, int64_t n, void* stream) {
  return launch_map1(x, y, n, RoundOp{}, stream, "dafk_round_fwd");
}

int dafk_act_fwd(const float* x, float* y, int64_t n, int act, float alpha, void* stream) {
  DAFK_REQUIRE(act >= 0 && act <= 3, DAFK_ERR_BAD_ARG, "dafk_act_fwd: bad activation %d", act);
  return launch_map1(x, y, n, ActFwd{act, alpha}, stream, "dafk_act_fwd");
}

int dafk_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, float alpha, void* stream) {
  DAFK_REQUIRE(act >= 0 && act <= 3, DAFK_ERR_BAD_ARG, "dafk_act_bwd: bad activation %d", act);
  return launch_map2(dy, y, dx, n, ActBwd{act, alpha}, stream, "dafk_act_bwd");
}

int dafk_act_bwd_bf16(const float* dy, const float* y, void* dx, int64_t n, int act, float alpha, void* stream) {
  DAFK_REQUIRE(act >= 0 && act <= 3, DAFK_ERR_BAD_ARG, "dafk_act_bwd_bf16: bad activation %d", act);
  DAFK_REQUIRE(n >= 0 && n % 4 == 0, DAFK_ERR_BAD_ARG, "dafk_act_bwd_bf16: size must be a multiple of 4");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && y && dx, DAFK_ERR_BAD_ARG, "dafk_act_bwd_bf16: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(y) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN, "dafk_act_bwd_bf16: alignment");
  act_bwd_bf16_kernel<<<bw_grid(n / 4, TPB), TPB, 0, as_stream(stream)>>>(dy, y, (__nv_bfloat16*)dx, n / 4, ActBwd{act, alpha});
  return check_launch("dafk_act_bwd_bf16");
}

int dafk_add_act_bwd(const float* dy1, const float* dy2, const float* y, float* dx, int64_t n, int act, float alpha,
                     void* stream) {
  DAFK_REQUIRE(act >= 0 && act <= 3, DAFK_ERR_BAD_ARG, "dafk_add_act_bwd: bad activation %d", act);
  DAFK_REQUIRE(n >= 0 && n % 4 == 0, DAFK_ERR_BAD_ARG, "dafk_add_act_bwd: size must be a multiple of 4");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(dy1 && dy2 && y && dx, DAFK_ERR_BAD_ARG, "dafk_add_act_bwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy1) && DAFK_ALIGNED16(dy2) && DAFK_ALIGNED16(y) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN,
               "dafk_add_act_bwd: alignment");
  add_act_bwd_kernel<<<bw_grid(n / 4, TPB), TPB, 0, as_stream(stream)>>>(dy1, dy2, y, dx, n / 4, ActBwd{act, alpha});
  return check_launch("dafk_add_act_bwd");
}

int dafk_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  return launch_map2(a, b, out, n, AddOp{}, stream, "dafk_add");
}

int dafk_add_dt(const void* a, const void* b, void* out, int dt, int64_t n, void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_add_dt: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(a && b && out, DAFK_ERR_BAD_ARG, "dafk_add_dt: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(a) && DAFK_ALIGNED16(b) && DAFK_ALIGNED16(out), DAFK_ERR_ALIGN, "dafk_add_dt: alignment");
  int grid = bw_grid((n + 3) / 4, TPB);
  if (dt == DAFK_F32) add_dt_kernel<float><<<grid, TPB, 0, as_stream(stream)>>>((const float*)a, (const float*)b, (float*)out, n);
  else if (dt == DAFK_BF16) add_dt_kernel<__nv_bfloat16><<<grid, TPB, 0, as_stream(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, n);
  else { set_error("dafk_add_dt: bad dtype %d", dt); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_add_dt");
}

int dafk_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream) {
  return launch_map2(x, y, y, n, AxpbyOp{a, b}, stream, "dafk_axpby");
}

int dafk_fill(float* x, float v, int64_t n, void* stream) {
  return launch_map1(x, x, n, FillOp{v}, stream, "dafk_fill");
}

int dafk_max_fwd(const float* a, const float* b, float* out, int64_t n, void* stream) {
  return launch_map2(a, b, out, n, MaxOp{}, stream, "dafk_max_fwd");
}

int dafk_max_bwd(const float* a, const float* b, const float* dout, float* da, float* db, int64_t n,
                 void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_max_bwd: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(a && b && dout && da && db, DAFK_ERR_BAD_ARG, "dafk_max_bwd: null pointer");
  max_bwd_kernel<<<bw_grid(n, TPB), TPB, 0, as_stream(stream)>>>(a, b, dout, da, db, n);
  return check_launch("dafk_max_bwd");
}

int dafk_softmax_fwd(const float* x, float* p, float* r, int64_t M, int C, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_softmax_fwd: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(x && p, DAFK_ERR_BAD_ARG, "dafk_softmax_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(p) && DAFK_ALIGNED16(r), DAFK_ERR_ALIGN,
               "dafk_softmax_fwd: pointers must be 16-byte aligned");
  int grid = bw_grid(M, TPB);
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 2: softmax_fwd_kernel<2><<<grid, TPB, 0, s>>>(x, p, r, M); break;
    case 3: softmax_fwd_kernel<3><<<grid, TPB, 0, s>>>(x, p, r, M); break;
    case 4: softmax_fwd_kernel<4><<<grid, TPB, 0, s>>>(x, p, r, M); break;
    case 5: softmax_fwd_kernel<5><<<grid, TPB, 0, s>>>(x, p, r, M); break;
    case 8: softmax_fwd_kernel<8><<<grid, TPB, 0, s>>>(x, p, r, M); break;
    case 16: softmax_fwd_kernel<16><<<grid, TPB, 0, s>>>(x, p, r, M); break;
    default:
      set_error("dafk_softmax_fwd: unsupported channel count %d (2,3,4,5,8,16)", C);
      return DAFK_ERR_UNSUPPORTED;
  }
  return check_launch("dafk_softmax_fwd");
}

int dafk_softmax_bwd(const float* p, const float* dp, float* dx, int64_t M, int C, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_softmax_bwd: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(p && dp && dx, DAFK_ERR_BAD_ARG, "dafk_softmax_bwd: null pointer");
  int grid = bw_grid(M, TPB);
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 2: softmax_bwd_kernel<2><<<grid, TPB, 0, s>>>(p, dp, dx, M); break;
    case 3: softmax_bwd_kernel<3><<<grid, TPB, 0, s>>>(p, dp, dx, M); break;
    case 4: softmax_bwd_kernel<4><<<grid, TPB, 0, s>>>(p, dp, dx, M); break;
    case 5: softmax_bwd_kernel<5><<<grid, TPB, 0, s>>>(p, dp, dx, M); break;
    case 8: softmax_bwd_kernel<8><<<grid, TPB, 0, s>>>(p, dp, dx, M); break;
    case 16: softmax_bwd_kernel<16><<<grid, TPB, 0, s>>>(p, dp, dx, M); break;
    default:
      set_error("dafk_softmax_bwd: unsupported channel count %d (2,3,4,5,8,16)", C);
      return DAFK_ERR_UNSUPPORTED;
  }
  return check_launch("dafk_softmax_bwd");
}

int dafk_cast(const void* x, int x_dt, void* y, int y_dt, int64_t n, void* stream) {
  DAFK_REQUIRE(n >= 0, DAFK_ERR_BAD_ARG, "dafk_cast: negative size");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "dafk_cast: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "dafk_cast: pointers must be 16-byte aligned");
  int grid = bw_grid((n + 3) / 4, TPB);
  cudaStream_t s = as_stream(stream);
  if (x_dt == DAFK_F32 && y_dt == DAFK_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, TPB, 0, s>>>((const float*)x, (__nv_bfloat16*)y, n);
  else if (x_dt == DAFK_BF16 && y_dt == DAFK_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, TPB, 0, s>>>((const __nv_bfloat16*)x, (float*)y, n);
  else if (x_dt == DAFK_F32 && y_dt == DAFK_F32)
    cast_kernel<float, float><<<grid, TPB, 0, s>>>((const float*)x, (float*)y, n);
  else if (x_dt == DAFK_BF16 && y_dt == DAFK_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, TPB, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  else {
    set_error("dafk_cast: unknown dtype codes %d -> %d", x_dt, y_dt);
    return DAFK_ERR_BAD_ARG;
  }
  return check_launch("dafk_cast");
}

int dafk_copy_channels(const float* src, int src_c, int src_off, float* dst, int dst_c, int dst_off,
                       int c, int64_t M, int accumulate, void* stream) {
  DAFK_REQUIRE(M >= 0 && c >= 0 && src_off >= 0 && dst_off >= 0 && src_off + c <= src_c &&
                   dst_off + c <= dst_c,
               DAFK_ERR_BAD_ARG, "dafk_copy_channels: channel window out of range");
  if (M == 0 || c == 0) return DAFK_OK;
  DAFK_REQUIRE(src && dst, DAFK_ERR_BAD_ARG, "dafk_copy_channels: null pointer");
  copy_channels_kernel<<<bw_grid(M * c, TPB), TPB, 0, as_stream(stream)>>>(src, src_c, src_off, dst, dst_c,
                                                                          dst_off, c, M, accumulate);
  return check_launch("dafk_copy_channels");
}

int dafk_gather_rows(const float* src, const int32_t* idx, float* dst, int64_t rows, int64_t row_elems,
                     void* stream) {
  DAFK_REQUIRE(rows >= 0 && row_elems >= 0, DAFK_ERR_BAD_ARG, "dafk_gather_rows: negative size");
  if (rows == 0 || row_elems == 0) return DAFK_OK;
  DAFK_REQUIRE(src && idx && dst, DAFK_ERR_BAD_ARG, "dafk_gather_rows: null pointer");
  DAFK_REQUIRE(row_elems % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_gather_rows: row_elems must be a multiple of 4");
  DAFK_REQUIRE(DAFK_ALIGNED16(src) && DAFK_ALIGNED16(dst), DAFK_ERR_ALIGN, "dafk_gather_rows: alignment");
  gather_rows_kernel<<<bw_grid(rows * (row_elems / 4), TPB), TPB, 0, as_stream(stream)>>>(src, idx, dst, rows,
                                                                                        row_elems);
  return check_launch("dafk_gather_rows");
}

int dafk_film_fwd(const float* x, const float* gamma, const float* beta, float* y, int B, int64_t HW, int C,
                  void* stream) {
  DAFK_REQUIRE(B >= 0 && HW >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_film_fwd: bad shape");
  if (B == 0 || HW == 0) return DAFK_OK;
  DAFK_REQUIRE(x && gamma && beta && y, DAFK_ERR_BAD_ARG, "dafk_film_fwd: null pointer");
  DAFK_REQUIRE(C % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_film_fwd: C must be a multiple of 4 (got %d)", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(beta),
               DAFK_ERR_ALIGN, "dafk_film_fwd: pointers must be 16-byte aligned");
  int64_t n4 = (int64_t)B * HW * C / 4;
  film_fwd_kernel<<<bw_grid(n4, TPB), TPB, 0, as_stream(stream)>>>(x, gamma, beta, nullptr, y, HW * C, C, n4, DAFK_ACT_NONE, 0.f);
  return check_launch("dafk_film_fwd");
}

int dafk_film_act_add_fwd(const float* x, const float* gamma, const float* beta, const float* res, float* y, int B,
                          int64_t HW, int C, int act, float alpha, void* stream) {
  DAFK_REQUIRE(B >= 0 && HW >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_film_act_add_fwd: bad shape");
  if (B == 0 || HW == 0) return DAFK_OK;
  DAFK_REQUIRE(x && gamma && beta && y, DAFK_ERR_BAD_ARG, "dafk_film_act_add_fwd: null pointer");
  DAFK_REQUIRE(C % 4 == 0, DAFK_ERR_UNSUPPORTED, "dafk_film_act_add_fwd: C must be a multiple of 4 (got %d)", C);
  DAFK_REQUIRE(act == DAFK_ACT_NONE || act == DAFK_ACT_RELU || act == DAFK_ACT_LRELU, DAFK_ERR_UNSUPPORTED,
               "dafk_film_act_add_fwd: act must be NONE, RELU or LRELU");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(beta) && DAFK_ALIGNED16(res),
               DAFK_ERR_ALIGN, "dafk_film_act_add_fwd: pointers must be 16-byte aligned");
  int64_t n4 = (int64_t)B * HW * C / 4;
  film_fwd_kernel<<<bw_grid(n4, TPB), TPB, 0, as_stream(stream)>>>(x, gamma, beta, res, y, HW * C, C, n4, act, alpha);
  return check_launch("dafk_film_act_add_fwd");
}

int dafk_film_act_add_bwd(const float* dy, const float* x, const float* gamma, const float* beta, float* dx,
                          float* dgamma, float* dbeta, double* ws, int B, int64_t HW, int C, int act, float alpha,
                          void* stream);

int dafk_film_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                  double* ws, int B, int64_t HW, int C, void* stream) {
  return dafk_film_act_add_bwd(dy, x, gamma, nullptr, dx, dgamma, dbeta, ws, B, HW, C, DAFK_ACT_NONE, 0.f, stream);
}

int dafk_film_act_add_bwd(const float* dy, const float* x, const float* gamma, const float* beta, float* dx,
                          float* dgamma, float* dbeta, double* ws, int B, int64_t HW, int C, int act, float alpha,
                          void* stream) {
  DAFK_REQUIRE(B >= 0 && HW >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_film_bwd: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && ws, DAFK_ERR_BAD_ARG, "dafk_film_bwd: null pointer");
  DAFK_REQUIRE(act == DAFK_ACT_NONE || ((act == DAFK_ACT_RELU || act == DAFK_ACT_LRELU) && beta && DAFK_ALIGNED16(beta)),
               DAFK_ERR_BAD_ARG, "dafk_film_bwd: a fused activation needs beta (16-byte aligned)");
  DAFK_REQUIRE(C >= 4 && (1024 % C) == 0, DAFK_ERR_UNSUPPORTED,
               "dafk_film_bwd: C must be a power of two in [4,1024] (got %d)", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dx) && DAFK_ALIGNED16(gamma),
               DAFK_ERR_ALIGN, "dafk_film_bwd: pointers must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * C, s);
  int64_t n4 = HW * C / 4;
  int chunks = (int)((n4 + TPB - 1) / TPB);
  int cap = (kNumSMs * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  film_bwd_kernel<<<dim3(chunks, B), TPB, 2 * C * sizeof(float), s>>>(dy, x, gamma, beta, dx, ws, HW * C, C, act, alpha);
  int rc = check_launch("dafk_film_bwd");
  if (rc) return rc;
  film_bwd_finish_kernel<<<(B * C + 127) / 128, 128, 0, s>>>(ws, dgamma, dbeta, B, C);
  return check_launch("dafk_film_bwd(finish)");
}

}  // extern "C"
