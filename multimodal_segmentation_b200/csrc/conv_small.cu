// Direct convolution kernels for the narrow layers (few input and/or output channels): the FiLM decoder's
// 8->8 3x3 stack, the first UNet / segmentor / discriminator layers, the 1x1 heads, the locnet's 5x5
// layers and the modality encoder.  These are HBM-bandwidth-bound (8-channel maps carry 32 B per pixel),
// so the design goal is: every input element is fetched from DRAM once (neighbouring threads share the
// window through L1), the whole filter bank sits in shared memory and is read as warp-wide broadcasts,
// and each thread keeps its COUT_T outputs in registers.
//   fwd        : one thread per output pixel x COUT_T channels (also the stride-1 data gradient, by
//                loading the filter bank mirrored / transposed)
//   dgrad_s2   : strided data gradient (one thread per input pixel)
//   wgrad      : one thread per (tap, ci), COUT_T accumulators in registers, pixel slabs across CTAs
#include "common.cuh"

namespace dafk {

constexpr int ST = 128;   // threads per CTA in the forward kernels

__device__ __forceinline__ float small_act(float z, int act, float alpha) {
  if (act == DAFK_ACT_RELU) return z > 0.f ? z : 0.f;
  if (act == DAFK_ACT_LRELU) return z > 0.f ? z : alpha * z;
  if (act == DAFK_ACT_TANH) return tanhf(z);
  return z;
}

struct SmallP {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo;
};

// ws layout in smem: [tap][ci][COUT_T]
// mode 0: forward weights  w[tap][ci][co0+j]                     (HWIO source with Cin_src = Cin, Cout_src = Cout)
// mode 1: stride-1 data gradient: the kernel's "input" is dy (Cin = Cout_src), its output is dx (Cout = Cin_src):
//         ws[tap][ci][j] = w_src[mirror(tap)][co0 + j][ci]
template <int COUT_T>
__global__ void __launch_bounds__(ST) conv_small_fwd_kernel(SmallP p, const float* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y, int act, float alpha, int mode) {
  extern __shared__ float ws[];
  const int taps = p.KH * p.KW;
  const int co0 = blockIdx.y * COUT_T;
  for (int e = threadIdx.x; e < taps * p.Cin * COUT_T; e += ST) {
    int j = e % COUT_T;
    int t = e / COUT_T;
    int ci = t % p.Cin;
    int tap = t / p.Cin;
    float v = 0.f;
    if (co0 + j < p.Cout) {
      if (mode == 0) v = w[((int64_t)tap * p.Cin + ci) * p.Cout + co0 + j];
      else v = w[((int64_t)(taps - 1 - tap) * p.Cout + (co0 + j)) * p.Cin + ci];
    }
    ws[e] = v;
  }
  __syncthreads();
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t m = (int64_t)blockIdx.x * ST + threadIdx.x;
  if (m >= M) return;
  const int wo = (int)(m % p.Wo);
  int64_t t = m / p.Wo;
  const int ho = (int)(t % p.Ho);
  const int n = (int)(t / p.Ho);
  const float* xb = x + (int64_t)n * p.H * p.W * p.Cin;
  float acc[COUT_T];
#pragma unroll
  for (int j = 0; j < COUT_T; ++j) acc[j] = 0.f;
  const int hi0 = ho * p.stride - p.pad, wi0 = wo * p.stride - p.pad;
  const bool vec = (p.Cin & 3) == 0;
  for (int r = 0; r < p.KH; ++r) {
    const int hi = hi0 + r;
    if (hi < 0 || hi >= p.H) continue;
    for (int q = 0; q < p.KW; ++q) {
      const int wi = wi0 + q;
      if (wi < 0 || wi >= p.W) continue;
      const float* px = xb + ((int64_t)hi * p.W + wi) * p.Cin;
      const float* pw = ws + (r * p.KW + q) * p.Cin * COUT_T;
      if (vec) {
        for (int ci = 0; ci < p.Cin; ci += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(px + ci));
          const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int j = 0; j < COUT_T; ++j) acc[j] = fmaf(xv[k], pw[(ci + k) * COUT_T + j], acc[j]);
          }
        }
      } else {
        for (int ci = 0; ci < p.Cin; ++ci) {
          const float xv = __ldg(px + ci);
#pragma unroll
          for (int j = 0; j < COUT_T; ++j) acc[j] = fmaf(xv, pw[ci * COUT_T + j], acc[j]);
        }
      }
    }
  }
  float* py = y + m * p.Cout + co0;
  if ((COUT_T & 3) == 0 && (p.Cout & 3) == 0 && co0 + COUT_T <= p.Cout) {
#pragma unroll
    for (int j = 0; j < COUT_T; j += 4) {
      float4 o;
      o.x = small_act(acc[j] + (bias ? bias[co0 + j] : 0.f), act, alpha);
      o.y = small_act(acc[j + 1] + (bias ? bias[co0 + j + 1] : 0.f), act, alpha);
      o.z = small_act(acc[j + 2] + (bias ? bias[co0 + j + 2] : 0.f), act, alpha);
      o.w = small_act(acc[j + 3] + (bias ? bias[co0 + j + 3] : 0.f), act, alpha);
      *reinterpret_cast<float4*>(py + j) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < COUT_T; ++j)
      if (co0 + j < p.Cout) py[j] = small_act(acc[j] + (bias ? bias[co0 + j] : 0.f), act, alpha);
  }
}

// strided data gradient: one thread per INPUT pixel, all Cin (<= CIN_T) outputs in registers.
// ws[tap][ci][co] = w[tap][ci][co]
template <int CIN_T>
__global__ void __launch_bounds__(ST) conv_small_dgrad_strided_kernel(SmallP p, const float* __restrict__ dy,
                                                                      const float* __restrict__ w,
                                                                      float* __restrict__ dx) {
  extern __shared__ float ws[];
  const int taps = p.KH * p.KW;
  for (int e = threadIdx.x; e < taps * p.Cin * p.Cout; e += ST) ws[e] = w[e];
  __syncthreads();
  const int64_t M = (int64_t)p.N * p.H * p.W;
  const int64_t m = (int64_t)blockIdx.x * ST + threadIdx.x;
  if (m >= M) return;
  const int wi = (int)(m % p.W);
  int64_t t = m / p.W;
  const int hi = (int)(t % p.H);
  const int n = (int)(t / p.H);
  const float* dyb = dy + (int64_t)n * p.Ho * p.Wo * p.Cout;
  float acc[CIN_T];
#pragma unroll
  for (int j = 0; j < CIN_T; ++j) acc[j] = 0.f;
  const bool vec = (p.Cout & 3) == 0;
  for (int r = 0; r < p.KH; ++r) {
    const int hh = hi + p.pad - r;
    if (hh < 0 || (hh % p.stride) != 0) continue;
    const int ho = hh / p.stride;
    if (ho >= p.Ho) continue;
    for (int q = 0; q < p.KW; ++q) {
      const int ww = wi + p.pad - q;
      if (ww < 0 || (ww % p.stride) != 0) continue;
      const int wo = ww / p.stride;
      if (wo >= p.Wo) continue;
      const float* pg = dyb + ((int64_t)ho * p.Wo + wo) * p.Cout;
      const float* pw = ws + (r * p.KW + q) * p.Cin * p.Cout;
      if (vec) {
        for (int co = 0; co < p.Cout; co += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(pg + co));
#pragma unroll
          for (int j = 0; j < CIN_T; ++j) {
            if (j < p.Cin) {
              const float4 wv = *reinterpret_cast<const float4*>(pw + j * p.Cout + co);
              acc[j] += g.x * wv.x + g.y * wv.y + g.z * wv.z + g.w * wv.w;
            }
          }
        }
      } else {
        for (int co = 0; co < p.Cout; ++co) {
          const float g = __ldg(pg + co);
#pragma unroll
          for (int j = 0; j < CIN_T; ++j)
            if (j < p.Cin) acc[j] = fmaf(g, pw[j * p.Cout + co], acc[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CIN_T; ++j)
    if (j < p.Cin) dx[m * p.Cin + j] = acc[j];
}

// weight gradient.  blockDim = (32 pixel lanes, TY tap slots); grid = (pixel slabs, ci groups, co groups).
// A thread owns one filter tap and a CIN_T x COUT_T block of dw in registers and strides over the pixels
// of its slab: per pixel it issues CIN_T/4 + COUT_T/4 128-bit loads (coalesced across the 32 lanes, which
// are consecutive pixels) for CIN_T*COUT_T FMAs.  The lanes are reduced with warp shuffles and one lane
// issues the atomics.  Tap slot 0 of ci-group 0 also reduces the bias gradient.
template <int CIN_T, int COUT_T>
__global__ void __launch_bounds__(512) conv_small_wgrad_kernel(SmallP p, const float* __restrict__ x,
                                                               const float* __restrict__ dy, float* __restrict__ dw,
                                                               float* __restrict__ db, int64_t pix_per_cta) {
  const int taps = p.KH * p.KW;
  const int ci0 = blockIdx.y * CIN_T, co0 = blockIdx.z * COUT_T;
  const int lane = threadIdx.x;
  const int64_t P = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t pbeg = (int64_t)blockIdx.x * pix_per_cta;
  const int64_t pend = min(P, pbeg + pix_per_cta);
  const bool xvec = (CIN_T & 3) == 0 && (p.Cin & 3) == 0 && ci0 + CIN_T <= p.Cin;
  const bool gvec = (COUT_T & 3) == 0 && (p.Cout & 3) == 0 && co0 + COUT_T <= p.Cout;
  const bool do_bias = db != nullptr && blockIdx.y == 0;
  for (int tap = threadIdx.y; tap < taps; tap += blockDim.y) {
    const int r = tap / p.KW, q = tap % p.KW;
    float acc[CIN_T][COUT_T];
    float bacc[COUT_T];
#pragma unroll
    for (int i = 0; i < CIN_T; ++i)
#pragma unroll
      for (int j = 0; j < COUT_T; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < COUT_T; ++j) bacc[j] = 0.f;
    for (int64_t pix = pbeg + lane; pix < pend; pix += 32) {
      const int wo = (int)(pix % p.Wo);
      const int64_t t = pix / p.Wo;
      const int ho = (int)(t % p.Ho);
      const int n = (int)(t / p.Ho);
      float g[COUT_T];
      const float* pg = dy + pix * p.Cout + co0;
      if (gvec) {
#pragma unroll
        for (int j = 0; j < COUT_T; j += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(pg + j));
          g[j] = v.x; g[j + 1] = v.y; g[j + 2] = v.z; g[j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < COUT_T; ++j) g[j] = (co0 + j < p.Cout) ? __ldg(pg + j) : 0.f;
      }
      if (do_bias && tap == 0) {
#pragma unroll
        for (int j = 0; j < COUT_T; ++j) bacc[j] += g[j];
      }
      const int hi = ho * p.stride - p.pad + r, wi = wo * p.stride - p.pad + q;
      if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
      const float* px = x + (((int64_t)n * p.H + hi) * p.W + wi) * p.Cin + ci0;
      float xv[CIN_T];
      if (xvec) {
#pragma unroll
        for (int i = 0; i < CIN_T; i += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(px + i));
          xv[i] = v.x; xv[i + 1] = v.y; xv[i + 2] = v.z; xv[i + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CIN_T; ++i) xv[i] = (ci0 + i < p.Cin) ? __ldg(px + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < CIN_T; ++i)
#pragma unroll
        for (int j = 0; j < COUT_T; ++j) acc[i][j] = fmaf(xv[i], g[j], acc[i][j]);
    }
    // reduce the 32 pixel lanes
#pragma unroll
    for (int i = 0; i < CIN_T; ++i)
#pragma unroll
      for (int j = 0; j < COUT_T; ++j) acc[i][j] = warp_sum(acc[i][j]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < CIN_T; ++i) {
        if (ci0 + i >= p.Cin) continue;
#pragma unroll
        for (int j = 0; j < COUT_T; ++j)
          if (co0 + j < p.Cout) atomicAdd(dw + ((int64_t)tap * p.Cin + ci0 + i) * p.Cout + co0 + j, acc[i][j]);
      }
    }
    if (do_bias && tap == 0) {
#pragma unroll
      for (int j = 0; j < COUT_T; ++j) bacc[j] = warp_sum(bacc[j]);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < COUT_T; ++j)
          if (co0 + j < p.Cout) atomicAdd(db + co0 + j, bacc[j]);
      }
    }
  }
}

static int small_params(const dafk_conv_desc* d, SmallP& p, const char* name) {
  DAFK_REQUIRE(d != nullptr, DAFK_ERR_BAD_ARG, "%s: null descriptor", name);
  DAFK_REQUIRE(d->N >= 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 && d->stride > 0 &&
                   d->pad >= 0,
               DAFK_ERR_BAD_ARG, "%s: bad descriptor", name);
  int Ho = (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  int Wo = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  DAFK_REQUIRE(Ho == d->Ho && Wo == d->Wo && Ho > 0 && Wo > 0, DAFK_ERR_BAD_ARG, "%s: output size mismatch", name);
  p = SmallP{d->N, d->H, d->W, d->Cin, d->Cout, d->KH, d->KW, d->stride, d->pad, d->Ho, d->Wo};
  return DAFK_OK;
}

template <int COUT_T>
static int launch_small_fwd(const SmallP& p, const float* x, const float* w, const float* bias, float* y, int act,
                            float alpha, int mode, cudaStream_t s) {
  size_t smem = sizeof(float) * (size_t)p.KH * p.KW * p.Cin * COUT_T;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv_small_fwd_kernel<COUT_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DAFK_REQUIRE(e == cudaSuccess, DAFK_ERR_UNSUPPORTED, "conv_small: filter bank of %zu bytes does not fit in shared memory", smem);
  }
  int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  dim3 grid((unsigned)((M + ST - 1) / ST), (p.Cout + COUT_T - 1) / COUT_T);
  conv_small_fwd_kernel<COUT_T><<<grid, ST, smem, s>>>(p, x, w, bias, y, act, alpha, mode);
  return check_launch("dafk_conv_small");
}

static int dispatch_small_fwd(const SmallP& p, const float* x, const float* w, const float* bias, float* y, int act,
                              float alpha, int mode, cudaStream_t s) {
  const int co = p.Cout;
  if (co == 1) return launch_small_fwd<1>(p, x, w, bias, y, act, alpha, mode, s);
  if (co <= 4) return launch_small_fwd<4>(p, x, w, bias, y, act, alpha, mode, s);
  if (co == 5) return launch_small_fwd<5>(p, x, w, bias, y, act, alpha, mode, s);
  if (co <= 8) return launch_small_fwd<8>(p, x, w, bias, y, act, alpha, mode, s);
  if (co == 9) return launch_small_fwd<9>(p, x, w, bias, y, act, alpha, mode, s);
  if (co == 20) return launch_small_fwd<20>(p, x, w, bias, y, act, alpha, mode, s);
  return launch_small_fwd<16>(p, x, w, bias, y, act, alpha, mode, s);
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_conv_small_supported(int Cin, int Cout, int KH, int KW) {
  // the filter bank [taps][Cin][COUT_T<=20] must fit in shared memory
  int64_t ct = Cout <= 20 ? Cout : 16;
  return ((int64_t)KH * KW * Cin * ct * 4 <= 160 * 1024) ? 1 : 0;
}

int dafk_conv_small_fwd(const dafk_conv_desc* d, const float* x, const float* w, const float* bias, float* y, int act,
                        float alpha, void* stream) {
  SmallP p;
  int rc = small_params(d, p, "dafk_conv_small_fwd");
  if (rc) return rc;
  if (p.N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && w && y, DAFK_ERR_BAD_ARG, "dafk_conv_small_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "dafk_conv_small_fwd: alignment");
  return dispatch_small_fwd(p, x, w, bias, y, act, alpha, 0, as_stream(stream));
}

int dafk_conv_small_dgrad(const dafk_conv_desc* d, const float* dy, const float* w, float* dx, void* stream) {
  SmallP p;
  int rc = small_params(d, p, "dafk_conv_small_dgrad");
  if (rc) return rc;
  if (p.N == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && w && dx, DAFK_ERR_BAD_ARG, "dafk_conv_small_dgrad: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN, "dafk_conv_small_dgrad: alignment");
  cudaStream_t s = as_stream(stream);
  if (p.stride == 1) {
    // dx = conv(dy, mirrored / transposed bank) with padding k-1-pad: reuse the forward kernel
    SmallP q{p.N, p.Ho, p.Wo, p.Cout, p.Cin, p.KH, p.KW, 1, p.KH - 1 - p.pad, p.H, p.W};
    DAFK_REQUIRE(p.KH == p.KW, DAFK_ERR_UNSUPPORTED, "dafk_conv_small_dgrad: square kernels only");
    return dispatch_small_fwd(q, dy, w, nullptr, dx, DAFK_ACT_NONE, 0.f, 1, s);
  }
  DAFK_REQUIRE(p.Cin <= 16, DAFK_ERR_UNSUPPORTED, "dafk_conv_small_dgrad: strided path needs Cin <= 16 (got %d)", p.Cin);
  size_t smem = sizeof(float) * (size_t)p.KH * p.KW * p.Cin * p.Cout;
  DAFK_REQUIRE(smem <= 160 * 1024, DAFK_ERR_UNSUPPORTED, "dafk_conv_small_dgrad: filter bank too large");
  int64_t M = (int64_t)p.N * p.H * p.W;
  dim3 grid((unsigned)((M + ST - 1) / ST));
#define LAUNCH_DG(T)                                                                                              \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      cudaFuncSetAttribute(conv_small_dgrad_strided_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    conv_small_dgrad_strided_kernel<T><<<grid, ST, smem, s>>>(p, dy, w, dx);                                      \
  } while (0)
  if (p.Cin == 1) LAUNCH_DG(1);
  else if (p.Cin <= 4) LAUNCH_DG(4);
  else if (p.Cin <= 9) LAUNCH_DG(9);
  else LAUNCH_DG(16);
#undef LAUNCH_DG
  return check_launch("dafk_conv_small_dgrad");
}

int dafk_conv_small_wgrad(const dafk_conv_desc* d, const float* x, const float* dy, float* dw, float* db,
                          void* stream) {
  SmallP p;
  int rc = small_params(d, p, "dafk_conv_small_wgrad");
  if (rc) return rc;
  if (p.N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_conv_small_wgrad: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy), DAFK_ERR_ALIGN, "dafk_conv_small_wgrad: alignment");
  cudaStream_t s = as_stream(stream);
  const int taps = p.KH * p.KW;
  const int ty = taps < 16 ? taps : 16;
  const int64_t P = (int64_t)p.N * p.Ho * p.Wo;
#define LAUNCH_WG(CI, CO)                                                                                \
  do {                                                                                                   \
    int gi = (p.Cin + CI - 1) / CI, go = (p.Cout + CO - 1) / CO;                                          \
    int64_t ctas = ((int64_t)kNumSMs * 4 + gi * go - 1) / (gi * go);                                      \
    int64_t per = (P + ctas - 1) / ctas;                                                                 \
    if (per < 256) per = 256;                                                                            \
    int slabs = (int)((P + per - 1) / per);                                                              \
    conv_small_wgrad_kernel<CI, CO><<<dim3(slabs, gi, go), dim3(32, ty), 0, s>>>(p, x, dy, dw, db, per);   \
  } while (0)
  // register block CIN_T x COUT_T <= 64 accumulators
  const int ci = p.Cin, co = p.Cout;
  if (ci == 1) {
    if (co >= 16) LAUNCH_WG(1, 16); else if (co >= 8) LAUNCH_WG(1, 8); else if (co >= 4) LAUNCH_WG(1, 4); else LAUNCH_WG(1, 1);
  } else if (ci <= 4 || (ci & 3) != 0) {
    // narrow or unaligned input rows (4, 5, 9 channels): 4-wide ci blocks with masking / scalar loads
    if (co >= 16) LAUNCH_WG(4, 16); else if (co >= 8) LAUNCH_WG(4, 8); else if (co >= 4) LAUNCH_WG(4, 4); else LAUNCH_WG(4, 1);
  } else {
    if (co >= 8 && (co & 7) == 0) LAUNCH_WG(8, 8);
    else if (co >= 4) LAUNCH_WG(8, 4);
    else LAUNCH_WG(8, 1);
  }
#undef LAUNCH_WG
  return check_launch("dafk_conv_small_wgrad");
}

}  // extern "C"
