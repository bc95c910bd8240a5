// MaxPooling2D(2,2), UpSampling2D(2) and nearest-neighbour resize on NHWC feature maps
// (f32 or bf16 storage).  One 4-channel vector per thread-iteration, grid-stride.
#include "common.cuh"

namespace dafk {

constexpr int TPB = 256;

template <typename T>
__global__ void __launch_bounds__(TPB) maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H,
                                                           int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1, C4 = C >> 2;
  int64_t total = (int64_t)N * Ho * Wo * C4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c4 = (int)(i % C4);
    int64_t p = i / C4;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho);
    int n = (int)(p / Ho);
    const T* base = x + (((int64_t)n * H + 2 * ho) * W + 2 * wo) * C + 4 * c4;
    float a[4], b[4], c[4], d[4], r[4];
    Vec4<T>::load(base, a);
    Vec4<T>::load(base + C, b);
    Vec4<T>::load(base + (int64_t)W * C, c);
    Vec4<T>::load(base + (int64_t)W * C + C, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    Vec4<T>::store(y + (((int64_t)n * Ho + ho) * Wo + wo) * C + 4 * c4, r);
  }
}

template <typename T>
__global__ void __launch_bounds__(TPB) maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           T* __restrict__ dx, int N, int H, int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1, C4 = C >> 2;
  int64_t total = (int64_t)N * Ho * Wo * C4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c4 = (int)(i % C4);
    int64_t p = i / C4;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho);
    int n = (int)(p / Ho);
    int64_t o00 = (((int64_t)n * H + 2 * ho) * W + 2 * wo) * C + 4 * c4;
    int64_t o01 = o00 + C, o10 = o00 + (int64_t)W * C, o11 = o10 + C;
    float a[4], b[4], c[4], d[4], g[4];
    Vec4<T>::load(x + o00, a);
    Vec4<T>::load(x + o01, b);
    Vec4<T>::load(x + o10, c);
    Vec4<T>::load(x + o11, d);
    Vec4<T>::load(dy + (((int64_t)n * Ho + ho) * Wo + wo) * C + 4 * c4, g);
    float ga[4], gb[4], gc[4], gd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // first maximum in row-major window order
      int arg = 0; float m = a[k];
      if (b[k] > m) { m = b[k]; arg = 1; }
      if (c[k] > m) { m = c[k]; arg = 2; }
      if (d[k] > m) { m = d[k]; arg = 3; }
      ga[k] = arg == 0 ? g[k] : 0.f;
      gb[k] = arg == 1 ? g[k] : 0.f;
      gc[k] = arg == 2 ? g[k] : 0.f;
      gd[k] = arg == 3 ? g[k] : 0.f;
    }
    Vec4<T>::store(dx + o00, ga);
    Vec4<T>::store(dx + o01, gb);
    Vec4<T>::store(dx + o10, gc);
    Vec4<T>::store(dx + o11, gd);
  }
}

template <typename T>
__global__ void __launch_bounds__(TPB) upsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H,
                                                            int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, C4 = C >> 2;
  int64_t total = (int64_t)N * Ho * Wo * C4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c4 = (int)(i % C4);
    int64_t p = i / C4;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho);
    int n = (int)(p / Ho);
    float v[4];
    Vec4<T>::load(x + (((int64_t)n * H + (ho >> 1)) * W + (wo >> 1)) * C + 4 * c4, v);
    Vec4<T>::store(y + 4 * i, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(TPB) upsample2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H,
                                                            int W, int C) {
  const int Wo = 2 * W, C4 = C >> 2;
  int64_t total = (int64_t)N * H * W * C4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c4 = (int)(i % C4);
    int64_t p = i / C4;
    int w = (int)(p % W); p /= W;
    int h = (int)(p % H);
    int n = (int)(p / H);
    const T* base = dy + (((int64_t)n * 2 * H + 2 * h) * Wo + 2 * w) * C + 4 * c4;
    float a[4], b[4], c[4], d[4], r[4];
    Vec4<T>::load(base, a);
    Vec4<T>::load(base + C, b);
    Vec4<T>::load(base + (int64_t)Wo * C, c);
    Vec4<T>::load(base + (int64_t)Wo * C + C, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = (a[k] + b[k]) + (c[k] + d[k]);
    Vec4<T>::store(dx + 4 * i, r);
  }
}

// ---- 16-byte vector variants (bf16: 8 channels per thread, f32: 4).  UpSampling2D reads every INPUT vector once and
// writes it to its 4 output positions (one index decomposition per 64 B of traffic instead of one per 8 B).
template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; v[4] = v[5] = v[6] = v[7] = 0.f;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct V16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
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
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// one thread per 16-byte INPUT vector; rows = N*H
template <typename T>
__global__ void __launch_bounds__(TPB) upsample2_fwd_v16_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total,
                                                                int W, int C) {
  constexpr int V = V16<T>::N;
  const int CV = C / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t orow = (int64_t)2 * W * C;      // elements of one output row
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + i * V);
    const int64_t p = i / CV;
    const int cv = (int)(i - p * CV);
    const int64_t row = p / W;
    const int w = (int)(p - row * W);
    T* o = y + (2 * row) * orow + (int64_t)(2 * w) * C + cv * V;
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + C) = v;
    *reinterpret_cast<uint4*>(o + orow) = v;
    *reinterpret_cast<uint4*>(o + orow + C) = v;
  }
}

// one thread per 16-byte OUTPUT (input-gradient) vector
template <typename T>
__global__ void __launch_bounds__(TPB) upsample2_bwd_v16_kernel(const T* __restrict__ dy, T* __restrict__ dx,
                                                                int64_t total, int W, int C) {
  constexpr int V = V16<T>::N;
  const int CV = C / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t orow = (int64_t)2 * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t p = i / CV;
    const int cv = (int)(i - p * CV);
    const int64_t row = p / W;
    const int w = (int)(p - row * W);
    const T* base = dy + (2 * row) * orow + (int64_t)(2 * w) * C + cv * V;
    float a[8], b[8], c[8], d[8], r[8];
    V16<T>::load(base, a);
    V16<T>::load(base + C, b);
    V16<T>::load(base + orow, c);
    V16<T>::load(base + orow + C, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (a[k] + b[k]) + (c[k] + d[k]);
    V16<T>::store(dx + i * V, r);
  }
}

// MaxPooling2D(2,2) with even H, W: one thread per 16-byte output vector; rows = N*Ho
template <typename T>
__global__ void __launch_bounds__(TPB) maxpool2_fwd_v16_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total,
                                                               int Wo, int C) {
  constexpr int V = V16<T>::N;
  const int CV = C / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t irow = (int64_t)2 * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t p = i / CV;
    const int cv = (int)(i - p * CV);
    const int64_t row = p / Wo;
    const int w = (int)(p - row * Wo);
    const T* base = x + (2 * row) * irow + (int64_t)(2 * w) * C + cv * V;
    float a[8], b[8], c[8], d[8], r[8];
    V16<T>::load(base, a);
    V16<T>::load(base + C, b);
    V16<T>::load(base + irow, c);
    V16<T>::load(base + irow + C, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    V16<T>::store(y + i * V, r);
  }
}

template <typename T>
__global__ void __launch_bounds__(TPB) maxpool2_bwd_v16_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                               T* __restrict__ dx, int64_t total, int Wo, int C) {
  constexpr int V = V16<T>::N;
  const int CV = C / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t irow = (int64_t)2 * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t p = i / CV;
    const int cv = (int)(i - p * CV);
    const int64_t row = p / Wo;
    const int w = (int)(p - row * Wo);
    const int64_t o00 = (2 * row) * irow + (int64_t)(2 * w) * C + cv * V;
    float a[8], b[8], c[8], d[8], g[8];
    V16<T>::load(x + o00, a);
    V16<T>::load(x + o00 + C, b);
    V16<T>::load(x + o00 + irow, c);
    V16<T>::load(x + o00 + irow + C, d);
    V16<T>::load(dy + i * V, g);
    float ga[8], gb[8], gc[8], gd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // first maximum in row-major window order
      int arg = 0; float m = a[k];
      if (b[k] > m) { m = b[k]; arg = 1; }
      if (c[k] > m) { m = c[k]; arg = 2; }
      if (d[k] > m) { m = d[k]; arg = 3; }
      ga[k] = arg == 0 ? g[k] : 0.f;
      gb[k] = arg == 1 ? g[k] : 0.f;
      gc[k] = arg == 2 ? g[k] : 0.f;
      gd[k] = arg == 3 ? g[k] : 0.f;
    }
    V16<T>::store(dx + o00, ga);
    V16<T>::store(dx + o00 + C, gb);
    V16<T>::store(dx + o00 + irow, gc);
    V16<T>::store(dx + o00 + irow + C, gd);
  }
}

__global__ void __launch_bounds__(TPB) resize_nn_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                                            int H, int W, int C, int Ho, int Wo) {
  int64_t total = (int64_t)N * Ho * Wo * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho);
    int n = (int)(p / Ho);
    int sh = min((int)(((int64_t)ho * H) / Ho), H - 1);
    int sw = min((int)(((int64_t)wo * W) / Wo), W - 1);
    y[i] = x[(((int64_t)n * H + sh) * W + sw) * C + c];
  }
}

__global__ void __launch_bounds__(TPB) resize_nn_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N,
                                                            int H, int W, int C, int Ho, int Wo) {
  int64_t total = (int64_t)N * Ho * Wo * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho);
    int n = (int)(p / Ho);
    int sh = min((int)(((int64_t)ho * H) / Ho), H - 1);
    int sw = min((int)(((int64_t)wo * W) / Wo), W - 1);
    atomicAdd(dx + (((int64_t)n * H + sh) * W + sw) * C + c, dy[i]);
  }
}

// any channel count (f32): one thread per output element.  Only reached for C % 4 != 0 (e.g. a 2-filter UNet level).
__global__ void __launch_bounds__(TPB) maxpool2_fwd_any_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                                               int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  int64_t total = (int64_t)N * Ho * Wo * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t r = i / C;
    int wo = (int)(r % Wo); r /= Wo;
    int ho = (int)(r % Ho);
    int n = (int)(r / Ho);
    const float* p = x + (((int64_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    y[i] = fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(int64_t)W * C], p[(int64_t)W * C + C]));
  }
}

}  // namespace dafk

using namespace dafk;

#define POOL_CHECKS(name)                                                                                   \
  DAFK_REQUIRE(N >= 0 && H >= 0 && W >= 0 && C > 0, DAFK_ERR_BAD_ARG, name ": bad shape");                   \
  DAFK_REQUIRE(C % 4 == 0, DAFK_ERR_UNSUPPORTED, name ": C must be a multiple of 4 (got %d)", C);           \
  DAFK_REQUIRE(dt == DAFK_F32 || dt == DAFK_BF16, DAFK_ERR_BAD_ARG, name ": bad dtype %d", dt);

extern "C" {

int dafk_maxpool2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* stream) {
  if (C > 0 && C % 4 != 0 && dt == DAFK_F32 && N >= 0 && H >= 0 && W >= 0) {
    int64_t tot = (int64_t)N * (H / 2) * (W / 2) * C;
    if (tot == 0) return DAFK_OK;
    DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "dafk_maxpool2_fwd: null pointer");
    maxpool2_fwd_any_kernel<<<bw_grid(tot, TPB), TPB, 0, as_stream(stream)>>>((const float*)x, (float*)y, N, H, W, C);
    return check_launch("dafk_maxpool2_fwd(any C)");
  }
  POOL_CHECKS("dafk_maxpool2_fwd");
  int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
  if (total == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "dafk_maxpool2_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "dafk_maxpool2_fwd: alignment");
  cudaStream_t s = as_stream(stream);
  if (dt == DAFK_BF16 && C % 8 == 0 && !(H & 1) && !(W & 1)) {
    int64_t tv = total / 2;
    maxpool2_fwd_v16_kernel<__nv_bfloat16><<<bw_grid(tv, TPB), TPB, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, tv, W / 2, C);
    return check_launch("dafk_maxpool2_fwd");
  }
  if (dt == DAFK_F32) maxpool2_fwd_kernel<float><<<bw_grid(total, TPB), TPB, 0, s>>>((const float*)x, (float*)y, N, H, W, C);
  else maxpool2_fwd_kernel<__nv_bfloat16><<<bw_grid(total, TPB), TPB, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
  return check_launch("dafk_maxpool2_fwd");
}

int dafk_maxpool2_bwd(const void* x, const void* dy, void* dx, int dt, int N, int H, int W, int C, void* stream) {
  POOL_CHECKS("dafk_maxpool2_bwd");
  int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
  if ((int64_t)N * H * W == 0) return DAFK_OK;
  DAFK_REQUIRE(x && dy && dx, DAFK_ERR_BAD_ARG, "dafk_maxpool2_bwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN, "dafk_maxpool2_bwd: alignment");
  cudaStream_t s = as_stream(stream);
  size_t esz = dt == DAFK_F32 ? 4 : 2;
  if ((H & 1) || (W & 1)) cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * esz, s);  // uncovered last row/col
  if (total == 0) return DAFK_OK;
  if (dt == DAFK_BF16 && C % 8 == 0 && !(H & 1) && !(W & 1)) {
    int64_t tv = total / 2;
    maxpool2_bwd_v16_kernel<__nv_bfloat16><<<bw_grid(tv, TPB), TPB, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, tv, W / 2, C);
    return check_launch("dafk_maxpool2_bwd");
  }
  if (dt == DAFK_F32) maxpool2_bwd_kernel<float><<<bw_grid(total, TPB), TPB, 0, s>>>((const float*)x, (const float*)dy, (float*)dx, N, H, W, C);
  else maxpool2_bwd_kernel<__nv_bfloat16><<<bw_grid(total, TPB), TPB, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, N, H, W, C);
  return check_launch("dafk_maxpool2_bwd");
}

int dafk_upsample2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* stream) {
  POOL_CHECKS("dafk_upsample2_fwd");
  int64_t total = (int64_t)N * 4 * H * W * (C / 4);
  if (total == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "dafk_upsample2_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "dafk_upsample2_fwd: alignment");
  cudaStream_t s = as_stream(stream);
  if (dt == DAFK_F32 || C % 8 == 0) {
    int64_t tv = (int64_t)N * H * W * (C / (dt == DAFK_F32 ? 4 : 8));
    if (dt == DAFK_F32) upsample2_fwd_v16_kernel<float><<<bw_grid(tv, TPB), TPB, 0, s>>>((const float*)x, (float*)y, tv, W, C);
    else upsample2_fwd_v16_kernel<__nv_bfloat16><<<bw_grid(tv, TPB), TPB, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, tv, W, C);
    return check_launch("dafk_upsample2_fwd");
  }
  if (dt == DAFK_F32) upsample2_fwd_kernel<float><<<bw_grid(total, TPB), TPB, 0, s>>>((const float*)x, (float*)y, N, H, W, C);
  else upsample2_fwd_kernel<__nv_bfloat16><<<bw_grid(total, TPB), TPB, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
  return check_launch("dafk_upsample2_fwd");
}

int dafk_upsample2_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, void* stream) {
  POOL_CHECKS("dafk_upsample2_bwd");
  int64_t total = (int64_t)N * H * W * (C / 4);
  if (total == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && dx, DAFK_ERR_BAD_ARG, "dafk_upsample2_bwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN, "dafk_upsample2_bwd: alignment");
  cudaStream_t s = as_stream(stream);
  if (dt == DAFK_BF16 && C % 8 == 0) {
    int64_t tv = total / 2;
    upsample2_bwd_v16_kernel<__nv_bfloat16><<<bw_grid(tv, TPB), TPB, 0, s>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, tv, W, C);
    return check_launch("dafk_upsample2_bwd");
  }
  if (dt == DAFK_F32) upsample2_bwd_kernel<float><<<bw_grid(total, TPB), TPB, 0, s>>>((const float*)dy, (float*)dx, N, H, W, C);
  else upsample2_bwd_kernel<__nv_bfloat16><<<bw_grid(total, TPB), TPB, 0, s>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, N, H, W, C);
  return check_launch("dafk_upsample2_bwd");
}

int dafk_resize_nn_fwd(const float* x, float* y, int N, int H, int W, int C, int Ho, int Wo, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DAFK_ERR_BAD_ARG, "dafk_resize_nn_fwd: bad shape");
  int64_t total = (int64_t)N * Ho * Wo * C;
  if (total == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "dafk_resize_nn_fwd: null pointer");
  resize_nn_fwd_kernel<<<bw_grid(total, TPB), TPB, 0, as_stream(stream)>>>(x, y, N, H, W, C, Ho, Wo);
  return check_launch("dafk_resize_nn_fwd");
}

int dafk_resize_nn_bwd(const float* dy, float* dx, int N, int H, int W, int C, int Ho, int Wo, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DAFK_ERR_BAD_ARG, "dafk_resize_nn_bwd: bad shape");
  int64_t total = (int64_t)N * Ho * Wo * C;
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && dx, DAFK_ERR_BAD_ARG, "dafk_resize_nn_bwd: null pointer");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * H * W * C, s);
  resize_nn_bwd_kernel<<<bw_grid(total, TPB), TPB, 0, s>>>(dy, dx, N, H, W, C, Ho, Wo);
  return check_launch("dafk_resize_nn_bwd");
}

}  // extern "C"
