// Space-to-depth helpers that turn a stride-2 convolution into a stride-1 convolution over 2x2 pixel blocks, so
// that the narrow stride-2 layers (modality encoder 3x3 s2, modality_encoder.py:36-42; first discriminator layer
// 4x4 s2, discriminator.py:24) run on the raster-strip tcgen05 kernels of conv_nc.cu:
//   y[n, i, j, (dy*2+dx)*C + c] = x[n, 2i+dy, 2j+dx, c]          (zero beyond the image)
//   w2[a, b, (dy*2+dx)*C + c, co] = w[2a+dy, 2b+dx, c, co]       (zero beyond the kernel)
//   conv(x, w, stride 2, valid) == conv(y, w2, stride 1, valid)
#include "common.cuh"

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
  if (x_dt == DAFK_F32)
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
