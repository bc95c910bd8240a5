// Space-to-depth helpers that turn a stride-2 convolution into a stride-1 convolution over 2x2 pixel blocks, so
// that the narrow stride-2 layers (modality encoder 3x3 s2, modality_encoder.py:36-42; first discriminator layer
// 4x4 s2, discriminator.py:24) run on the raster-strip tcgen05 kernels of conv_nc.cu:
//   y[n, i, j, (dy*2+dx)*C + c] = x[n, 2i+dy, 2j+dx, c]          (zero beyond the image)
//   w2[a, b, (dy*2+dx)*C + c, co] = w[2a+dy, 2b+dx, c, co]       (zero beyond the kernel)
//   conv(x, w, stride 2, valid) == conv(y, w2, stride 1, valid)
#include "common.cuh"
#include <cstdlib>

namespace dafk {

template <typename TI>
__global__ void s2d_fwd_kernel(const TI* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H, int W, int C, int H2,
                               int W2) {
  const int64_t total = (int64_t)N * H2 * W2 * 4 * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c4 = (int)(i % (4 * C));
    int64_t t = i / (4 * C);
    const int j = (int)(t % W2); t /= W2;
    const int ii = (int)(t % H2);
    const int n = (int)(t / H2);
    const int q = c4 / C, c = c4 - q * C;
    const int yy = 2 * ii + (q >> 1), xx = 2 * j + (q & 1);
    float v = 0.f;
    if (yy < H && xx < W) v = to_f<TI>(x[(((int64_t)n * H + yy) * W + xx) * C + c]);
    y[i] = __float2bfloat16_rn(v);
  }
}

template <typename TI, typename TO>
__global__ void d2s_kernel(const TI* __restrict__ y, TO* __restrict__ x, int N, int H, int W, int C, int H2, int W2) {
  const int64_t total = (int64_t)N * H * W * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int xx = (int)(t % W); t /= W;
    const int yy = (int)(t % H);
    const int n = (int)(t / H);
    const int q = (yy & 1) * 2 + (xx & 1);
    x[i] = from_f<TO>(to_f<TI>(y[(((int64_t)n * H2 + (yy >> 1)) * W2 + (xx >> 1)) * (4 * C) + q * C + c]));
  }
}

// Concatenate (model_components/modality_encoder.py:34: [anatomy, image]) fused into the rearrangement: the two sources are
// read where they lie, y[n, i, j, q*(Ca+Cb) + c] = (c < Ca ? xa[.., c] : xb[.., c - Ca]); one thread per (output pixel,
// quadrant q) copies the Ca + Cb contiguous channels (32-bit index arithmetic: one division chain per Ca + Cb elements;
// the element-per-thread kernel above spends ~100 instructions of 64-bit div/mod on every element).
template <typename TA, typename TB>
__global__ void __launch_bounds__(256) s2d_cat_kernel(const TA* __restrict__ xa, int Ca, const TB* __restrict__ xb, int Cb,
                                                      __nv_bfloat16* __restrict__ y, int N, int H, int W, int H2, int W2) {
  const int64_t total = (int64_t)N * H2 * W2 * 4;
  const int C = Ca + Cb;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(t & 3);
    const int64_t pix = t >> 2;                       // (n, i, j)
    const int j = (int)(pix % W2);
    const int64_t r = pix / W2;
    const int i = (int)(r % H2);
    const int n = (int)(r / H2);
    const int yy = 2 * i + (q >> 1), xx = 2 * j + (q & 1);
    __nv_bfloat16* o = y + t * C;                     // ((n*H2 + i)*W2 + j)*4C + q*C
    if (yy < H && xx < W) {
      const int64_t src = ((int64_t)n * H + yy) * W + xx;
      const TA* pa = xa + src * Ca;
      if ((Ca & 7) == 0 && Cb == 0) {                 // single source, whole 16-byte output vectors
        for (int c = 0; c < Ca; c += 8) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = to_f<TA>(pa[c + k]);
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(o + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
        for (int c = 0; c < Ca; ++c) o[c] = __float2bfloat16_rn(to_f<TA>(pa[c]));
        const TB* pb = xb + src * Cb;
        for (int c = 0; c < Cb; ++c) o[Ca + c] = __float2bfloat16_rn(to_f<TB>(pb[c]));
      }
    } else {
      for (int c = 0; c < C; ++c) o[c] = __float2bfloat16_rn(0.f);
    }
  }
}

// its backward: the gradient of the rearranged map goes straight to the two sources (no concatenated fp32 gradient)
template <typename TI>
__global__ void __launch_bounds__(256) d2s_split_kernel(const TI* __restrict__ y, float* __restrict__ ga, int Ca,
                                                        float* __restrict__ gb, int Cb, int N, int H, int W, int H2, int W2) {
  const int64_t total = (int64_t)N * H * W;
  const int C = Ca + Cb;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % W);
    const int64_t r = t / W;
    const int yy = (int)(r % H);
    const int n = (int)(r / H);
    const int q = (yy & 1) * 2 + (xx & 1);
    const TI* src = y + ((((int64_t)n * H2 + (yy >> 1)) * W2 + (xx >> 1)) * 4 + q) * C;
    if (ga) { float* o = ga + t * Ca; for (int c = 0; c < Ca; ++c) o[c] = to_f<TI>(src[c]); }
    if (gb) { float* o = gb + t * Cb; for (int c = 0; c < Cb; ++c) o[c] = to_f<TI>(src[Ca + c]); }
  }
}

// forward (bwd == 0): w2 = rearranged w.  backward (bwd == 1): w += rearranged^T(w2)   (gradient accumulation)
__global__ void s2d_weights_kernel(float* __restrict__ w, float* __restrict__ w2, int KH, int KW, int C, int Co, int KH2,
                                   int KW2, int bwd) {
  const int total = KH2 * KW2 * 4 * C * Co;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Co;
    int t = i / Co;
    const int c4 = t % (4 * C); t /= (4 * C);
    const int b = t % KW2;
    const int a = t / KW2;
    const int q = c4 / C, c = c4 - q * C;
    const int r = 2 * a + (q >> 1), s = 2 * b + (q & 1);
    const bool in = r < KH && s < KW;
    const int64_t src = (((int64_t)r * KW + s) * C + c) * Co + co;
    if (!bwd) w2[i] = in ? w[src] : 0.f;
    else if (in) w[src] += w2[i];
  }
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_space_to_depth2(const void* x, int x_dt, void* y_bf16, int N, int H, int W, int C, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_space_to_depth2: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y_bf16, DAFK_ERR_BAD_ARG, "dafk_space_to_depth2: null pointer");
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int64_t total = (int64_t)N * H2 * W2 * 4 * C;
  cudaStream_t s = as_stream(stream);
  // one thread per (output pixel, quadrant) -- see s2d_cat_kernel; DAFK_S2D_ELEMENTWISE=1 selects the element-per-thread kernel
  static const bool elementwise = [] { const char* e = getenv("DAFK_S2D_ELEMENTWISE"); return e && atoi(e) != 0; }();
  const int gq = bw_grid(total / C, 256);
  if (!elementwise && x_dt == DAFK_F32)
    s2d_cat_kernel<float, float><<<gq, 256, 0, s>>>((const float*)x, C, nullptr, 0, (__nv_bfloat16*)y_bf16, N, H, W, H2, W2);
  else if (!elementwise && x_dt == DAFK_BF16)
    s2d_cat_kernel<__nv_bfloat16, __nv_bfloat16><<<gq, 256, 0, s>>>((const __nv_bfloat16*)x, C, nullptr, 0,
                                                                   (__nv_bfloat16*)y_bf16, N, H, W, H2, W2);
  else if (x_dt == DAFK_F32)
    s2d_fwd_kernel<float><<<bw_grid(total, 256), 256, 0, s>>>((const float*)x, (__nv_bfloat16*)y_bf16, N, H, W, C, H2, W2);
  else if (x_dt == DAFK_BF16)
    s2d_fwd_kernel<__nv_bfloat16><<<bw_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y_bf16, N, H, W,
                                                                       C, H2, W2);
  else {
    set_error("dafk_space_to_depth2: bad dtype");
    return DAFK_ERR_BAD_ARG;
  }
  return check_launch("dafk_space_to_depth2");
}

int dafk_space_to_depth2_cat(const void* xa, int xa_dt, int Ca, const void* xb, int xb_dt, int Cb, void* y_bf16, int N,
                             int H, int W, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && Ca > 0 && Cb > 0, DAFK_ERR_BAD_ARG, "dafk_space_to_depth2_cat: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(xa && xb && y_bf16, DAFK_ERR_BAD_ARG, "dafk_space_to_depth2_cat: null pointer");
  DAFK_REQUIRE((xa_dt == DAFK_F32 || xa_dt == DAFK_BF16) && (xb_dt == DAFK_F32 || xb_dt == DAFK_BF16), DAFK_ERR_BAD_ARG,
               "dafk_space_to_depth2_cat: bad dtype");
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int64_t total = (int64_t)N * H2 * W2 * 4;
  cudaStream_t s = as_stream(stream);
  const int g = bw_grid(total, 256);
  __nv_bfloat16* y = (__nv_bfloat16*)y_bf16;
  typedef __nv_bfloat16 bf;
  if (xa_dt == DAFK_F32 && xb_dt == DAFK_F32)
    s2d_cat_kernel<float, float><<<g, 256, 0, s>>>((const float*)xa, Ca, (const float*)xb, Cb, y, N, H, W, H2, W2);
  else if (xa_dt == DAFK_F32)
    s2d_cat_kernel<float, bf><<<g, 256, 0, s>>>((const float*)xa, Ca, (const bf*)xb, Cb, y, N, H, W, H2, W2);
  else if (xb_dt == DAFK_F32)
    s2d_cat_kernel<bf, float><<<g, 256, 0, s>>>((const bf*)xa, Ca, (const float*)xb, Cb, y, N, H, W, H2, W2);
  else
    s2d_cat_kernel<bf, bf><<<g, 256, 0, s>>>((const bf*)xa, Ca, (const bf*)xb, Cb, y, N, H, W, H2, W2);
  return check_launch("dafk_space_to_depth2_cat");
}

int dafk_depth_to_space2_split(const void* y, int y_dt, float* ga, int Ca, float* gb, int Cb, int N, int H, int W,
                               void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && Ca > 0 && Cb > 0, DAFK_ERR_BAD_ARG, "dafk_depth_to_space2_split: bad shape");
  if (N == 0 || (!ga && !gb)) return DAFK_OK;
  DAFK_REQUIRE(y, DAFK_ERR_BAD_ARG, "dafk_depth_to_space2_split: null pointer");
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int64_t total = (int64_t)N * H * W;
  cudaStream_t s = as_stream(stream);
  const int g = bw_grid(total, 256);
  if (y_dt == DAFK_F32) d2s_split_kernel<float><<<g, 256, 0, s>>>((const float*)y, ga, Ca, gb, Cb, N, H, W, H2, W2);
  else if (y_dt == DAFK_BF16)
    d2s_split_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)y, ga, Ca, gb, Cb, N, H, W, H2, W2);
  else { set_error("dafk_depth_to_space2_split: bad dtype"); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_depth_to_space2_split");
}

int dafk_depth_to_space2(const void* y, int y_dt, void* x, int x_dt, int N, int H, int W, int C, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_depth_to_space2: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && y, DAFK_ERR_BAD_ARG, "dafk_depth_to_space2: null pointer");
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int64_t total = (int64_t)N * H * W * C;
  cudaStream_t s = as_stream(stream);
  const int g = bw_grid(total, 256);
  if (y_dt == DAFK_F32 && x_dt == DAFK_F32) d2s_kernel<float, float><<<g, 256, 0, s>>>((const float*)y, (float*)x, N, H, W, C, H2, W2);
  else if (y_dt == DAFK_F32 && x_dt == DAFK_BF16)
    d2s_kernel<float, __nv_bfloat16><<<g, 256, 0, s>>>((const float*)y, (__nv_bfloat16*)x, N, H, W, C, H2, W2);
  else if (y_dt == DAFK_BF16 && x_dt == DAFK_F32)
    d2s_kernel<__nv_bfloat16, float><<<g, 256, 0, s>>>((const __nv_bfloat16*)y, (float*)x, N, H, W, C, H2, W2);
  else if (y_dt == DAFK_BF16 && x_dt == DAFK_BF16)
    d2s_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)y, (__nv_bfloat16*)x, N, H, W, C, H2, W2);
  else {
    set_error("dafk_depth_to_space2: bad dtype");
    return DAFK_ERR_BAD_ARG;
  }
  return check_launch("dafk_depth_to_space2");
}

int dafk_conv_s2d_weights(float* w, float* w2, int KH, int KW, int C, int Cout, int backward, void* stream) {
  DAFK_REQUIRE(w && w2 && KH > 0 && KW > 0 && C > 0 && Cout > 0, DAFK_ERR_BAD_ARG, "dafk_conv_s2d_weights: bad argument");
  const int KH2 = (KH + 1) / 2, KW2 = (KW + 1) / 2;
  const int total = KH2 * KW2 * 4 * C * Cout;
  s2d_weights_kernel<<<bw_grid(total, 256), 256, 0, as_stream(stream)>>>(w, w2, KH, KW, C, Cout, KH2, KW2, backward ? 1 : 0);
  return check_launch("dafk_conv_s2d_weights");
}

}  // extern "C"
