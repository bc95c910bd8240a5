// Training-time augmentation on the device (reference model_executors/base_executor.py:37-78,103-110:
// keras ImageDataGenerator(rotation_range=20.) -> random_transform -> apply_transform, i.e.
// scipy.ndimage.affine_transform(order=1, mode='nearest') of every channel about the image centre).
// The reference rotates on the host, one slice and one channel at a time; here the batch is rotated right after its
// H2D copy: one thread per output pixel, all channels, bilinear taps with edge replication.
//   in_r = c*(i - ox) - s*(j - oy) + ox,  in_c = s*(i - ox) + c*(j - oy) + oy,  ox = H/2 + 0.5, oy = W/2 + 0.5
//   (transform_matrix_offset_center);  coordinates are clamped to the image (mode='nearest') before interpolation.
#include "common.cuh"

namespace dafk {

template <int C>
__global__ void __launch_bounds__(256) rotate_bilinear_kernel(const float* __restrict__ x, const float* __restrict__ theta,
                                                              float* __restrict__ y, int H, int W) {
  const int b = blockIdx.y;
  float sn, cs;
  sincosf(theta[b], &sn, &cs);
  const float ox = 0.5f * (float)H + 0.5f, oy = 0.5f * (float)W + 0.5f;
  const float* xb = x + (int64_t)b * H * W * C;
  float* yb = y + (int64_t)b * H * W * C;
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int i = p / W, j = p - i * W;
    const float di = (float)i - ox, dj = (float)j - oy;
    float r = cs * di - sn * dj + ox;
    float c = sn * di + cs * dj + oy;
    r = fminf(fmaxf(r, 0.f), (float)(H - 1));
    c = fminf(fmaxf(c, 0.f), (float)(W - 1));
    const int r0 = (int)floorf(r), c0 = (int)floorf(c);
    const int r1 = min(r0 + 1, H - 1), c1 = min(c0 + 1, W - 1);
    const float wr = r - (float)r0, wc = c - (float)c0;
    const float* p00 = xb + ((int64_t)r0 * W + c0) * C;
    const float* p01 = xb + ((int64_t)r0 * W + c1) * C;
    const float* p10 = xb + ((int64_t)r1 * W + c0) * C;
    const float* p11 = xb + ((int64_t)r1 * W + c1) * C;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float top = p00[k] + wc * (p01[k] - p00[k]);
      const float bot = p10[k] + wc * (p11[k] - p10[k]);
      yb[(int64_t)p * C + k] = top + wr * (bot - top);
    }
  }
}

// Executor.add_residual (model_executors/base_executor.py:83-87) on a staged batch: the reference builds the background
// channel AFTER the augmentation (dafnet_executor.py:493-494), so it is 1 wherever no rotated mask channel is exactly 1.
// m: [pixels, C], channel C-1 is rewritten from channels 0..C-2.
__global__ void __launch_bounds__(256) mask_residual_kernel(float* __restrict__ m, int64_t pixels, int C) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (int64_t)gridDim.x * blockDim.x) {
    float* q = m + p * C;
    float r = 1.f;
    for (int c = 0; c < C - 1; ++c)
      if (q[c] == 1.f) r = 0.f;
    q[C - 1] = r;
  }
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_rotate_bilinear(const float* x, const float* theta, float* y, int B, int H, int W, int C, void* stream) {
  DAFK_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_rotate_bilinear: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(x && theta && y && x != y, DAFK_ERR_BAD_ARG, "dafk_rotate_bilinear: null or aliased pointer");
  const int chunks = (H * W + 255) / 256;
  dim3 grid(chunks < 1 ? 1 : chunks, B);
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 1: rotate_bilinear_kernel<1><<<grid, 256, 0, s>>>(x, theta, y, H, W); break;
    case 2: rotate_bilinear_kernel<2><<<grid, 256, 0, s>>>(x, theta, y, H, W); break;
    case 3: rotate_bilinear_kernel<3><<<grid, 256, 0, s>>>(x, theta, y, H, W); break;
    case 4: rotate_bilinear_kernel<4><<<grid, 256, 0, s>>>(x, theta, y, H, W); break;
    case 5: rotate_bilinear_kernel<5><<<grid, 256, 0, s>>>(x, theta, y, H, W); break;
    case 8: rotate_bilinear_kernel<8><<<grid, 256, 0, s>>>(x, theta, y, H, W); break;
    default: set_error("dafk_rotate_bilinear: C must be 1..5 or 8 (got %d)", C); return DAFK_ERR_UNSUPPORTED;
  }
  return check_launch("dafk_rotate_bilinear");
}

int dafk_mask_residual(float* m, int64_t pixels, int C, void* stream) {
  DAFK_REQUIRE(pixels >= 0 && C >= 2, DAFK_ERR_BAD_ARG, "dafk_mask_residual: bad shape");
  if (pixels == 0) return DAFK_OK;
  DAFK_REQUIRE(m, DAFK_ERR_BAD_ARG, "dafk_mask_residual: null pointer");
  mask_residual_kernel<<<bw_grid(pixels, 256), 256, 0, as_stream(stream)>>>(m, pixels, C);
  return check_launch("dafk_mask_residual");
}

}  // extern "C"
