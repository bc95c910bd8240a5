// General CUDA-core convolution (any kernel size / stride / padding / channel count), fp32:
// forward, data gradient and weight gradient as shared-memory tiled implicit GEMMs.
// This path serves the small-channel layers (Cin or Cout in {1,4,5,8,9,16,20}) that cannot
// feed the tensor cores; the wide 3x3 layers go through conv_tc.cu.
#include "common.cuh"
#include "norm_wide.cuh"
#include "reduce.cuh"

namespace dafk {

constexpr int CT = 256;  // threads per CTA
constexpr int BK = 16;

struct ConvP {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo;
};

__device__ __forceinline__ float apply_act(float z, int act, float alpha) {
  if (act == DAFK_ACT_RELU) return z > 0.f ? z : 0.f;
  if (act == DAFK_ACT_LRELU) return z > 0.f ? z : alpha * z;
  if (act == DAFK_ACT_TANH) return tanhf(z);
  return z;
}

template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void tile_fma(const float (*As)[BM], const float (*Bs)[BN], float (&acc)[TM][TN], int ty,
                                         int tx) {
#pragma unroll
  for (int k = 0; k < BK; ++k) {
    float a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------ forward
// GEMM: M = N*Ho*Wo pixels, Ndim = Cout, Kdim = KH*KW*Cin;  A = im2col(x), B = w[K][Cout]
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CT) conv_fwd_kernel(ConvP p, const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, int act,
                                                      float alpha) {
  static_assert((BM / TM) * (BN / TN) == CT, "tile/thread mismatch");
  __shared__ float As[BK][BM];
  __shared__ float Bs[BK][BN];
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int Kdim = p.KH * p.KW * p.Cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);

  // A loader: each thread owns A_PER pixels-slots: element e = threadIdx.x + i*CT, m = e % BM, kk = e / BM
  constexpr int A_PER = BM * BK / CT;
  int hi0[A_PER], wi0[A_PER];
  int64_t pbase[A_PER];
  bool mval[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    int e = threadIdx.x + i * CT;
    int64_t m = m0 + (e % BM);
    mval[i] = m < M;
    int64_t mm = mval[i] ? m : 0;
    int wo = (int)(mm % p.Wo);
    int64_t t = mm / p.Wo;
    int ho = (int)(t % p.Ho);
    int n = (int)(t / p.Ho);
    hi0[i] = ho * p.stride - p.pad;
    wi0[i] = wo * p.stride - p.pad;
    pbase[i] = (int64_t)n * p.H * p.W * p.Cin;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kdim; k0 += BK) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int e = threadIdx.x + i * CT;
      int kk = e / BM;
      int k = k0 + kk;
      float v = 0.f;
      if (mval[i] && k < Kdim) {
        int ci = k % p.Cin;
        int tap = k / p.Cin;
        int q = tap % p.KW, r = tap / p.KW;
        int hi = hi0[i] + r, wi = wi0[i] + q;
        if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v = __ldg(x + pbase[i] + ((int64_t)hi * p.W + wi) * p.Cin + ci);
      }
      As[kk][e % BM] = v;
    }
    for (int e = threadIdx.x; e < BK * BN; e += CT) {
      int kk = e / BN, nn = e % BN;
      int k = k0 + kk, n = n0 + nn;
      Bs[kk][nn] = (k < Kdim && n < p.Cout) ? __ldg(w + (int64_t)k * p.Cout + n) : 0.f;
    }
    __syncthreads();
    tile_fma<BM, BN, TM, TN>(As, Bs, acc, ty, tx);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < p.Cout) {
        float z = acc[i][j] + (bias ? bias[n] : 0.f);
        y[m * p.Cout + n] = apply_act(z, act, alpha);
      }
    }
  }
}

// ------------------------------------------------------------------ data gradient
// GEMM: M = N*H*W input pixels, Ndim = Cin, Kdim = KH*KW*Cout;
// A(m,k=(r,q,co)) = dy[n,(h+pad-r)/s,(w+pad-q)/s,co] where divisible and in range
// B(k,ci) = w[((r*KW+q)*Cin+ci)*Cout+co]
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CT) conv_dgrad_kernel(ConvP p, const float* __restrict__ dy,
                                                        const float* __restrict__ w, float* __restrict__ dx) {
  static_assert((BM / TM) * (BN / TN) == CT, "tile/thread mismatch");
  __shared__ float As[BK][BM];
  __shared__ float Bs[BK][BN];
  const int64_t M = (int64_t)p.N * p.H * p.W;
  const int Kdim = p.KH * p.KW * p.Cout;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);

  constexpr int A_PER = BM * BK / CT;
  int hp[A_PER], wp[A_PER];
  int64_t pbase[A_PER];
  bool mval[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    int e = threadIdx.x + i * CT;
    int64_t m = m0 + (e % BM);
    mval[i] = m < M;
    int64_t mm = mval[i] ? m : 0;
    int wi = (int)(mm % p.W);
    int64_t t = mm / p.W;
    int hi = (int)(t % p.H);
    int n = (int)(t / p.H);
    hp[i] = hi + p.pad;
    wp[i] = wi + p.pad;
    pbase[i] = (int64_t)n * p.Ho * p.Wo * p.Cout;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kdim; k0 += BK) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int e = threadIdx.x + i * CT;
      int kk = e / BM;
      int k = k0 + kk;
      float v = 0.f;
      if (mval[i] && k < Kdim) {
        int co = k % p.Cout;
        int tap = k / p.Cout;
        int q = tap % p.KW, r = tap / p.KW;
        int hh = hp[i] - r, ww = wp[i] - q;
        if (hh >= 0 && ww >= 0 && (hh % p.stride) == 0 && (ww % p.stride) == 0) {
          int ho = hh / p.stride, wo = ww / p.stride;
          if (ho < p.Ho && wo < p.Wo) v = __ldg(dy + pbase[i] + ((int64_t)ho * p.Wo + wo) * p.Cout + co);
        }
      }
      As[kk][e % BM] = v;
    }
    // B tile: consecutive threads walk k (co contiguous in memory)
    for (int e = threadIdx.x; e < BK * BN; e += CT) {
      int kk = e % BK, nn = e / BK;
      int k = k0 + kk, ci = n0 + nn;
      float v = 0.f;
      if (k < Kdim && ci < p.Cin) {
        int co = k % p.Cout;
        int tap = k / p.Cout;
        v = __ldg(w + ((int64_t)tap * p.Cin + ci) * p.Cout + co);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
    tile_fma<BM, BN, TM, TN>(As, Bs, acc, ty, tx);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < p.Cin) dx[m * p.Cin + n] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// GEMM: M = KH*KW*Cin rows of dw[K][Cout], Ndim = Cout, Kdim = pixels (split across grid.z)
// A(m=(r,q,ci), k=pixel) = x[n,ho*s-pad+r,wo*s-pad+q,ci];  B(k,co) = dy[pixel,co]
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CT) conv_wgrad_kernel(ConvP p, const float* __restrict__ x,
                                                        const float* __restrict__ dy, float* __restrict__ dw,
                                                        float* __restrict__ db, int64_t pix_per_split) {
  static_assert((BM / TM) * (BN / TN) == CT, "tile/thread mismatch");
  __shared__ float As[BK][BM];
  __shared__ float Bs[BK][BN];
  __shared__ float bsum[BN];
  const int Mdim = p.KH * p.KW * p.Cin;
  const int64_t P = (int64_t)p.N * p.Ho * p.Wo;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * pix_per_split;
  int64_t k_end = k_begin + pix_per_split;
  if (k_end > P) k_end = P;
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
  const bool do_bias = (db != nullptr) && (blockIdx.x == 0);
  if (threadIdx.x < BN) bsum[threadIdx.x] = 0.f;

  // A loader: element e = threadIdx.x + i*CT, m = e % BM (ci contiguous), kk = e / BM
  constexpr int A_PER = BM * BK / CT;
  int a_r[A_PER], a_q[A_PER], a_ci[A_PER];
  bool a_val[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    int e = threadIdx.x + i * CT;
    int m = m0 + (e % BM);
    a_val[i] = m < Mdim;
    int mm = a_val[i] ? m : 0;
    a_ci[i] = mm % p.Cin;
    int tap = mm / p.Cin;
    a_q[i] = tap % p.KW;
    a_r[i] = tap / p.KW;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int e = threadIdx.x + i * CT;
      int kk = e / BM;
      int64_t pix = k0 + kk;
      float v = 0.f;
      if (a_val[i] && pix < k_end) {
        int wo = (int)(pix % p.Wo);
        int64_t t = pix / p.Wo;
        int ho = (int)(t % p.Ho);
        int n = (int)(t / p.Ho);
        int hi = ho * p.stride - p.pad + a_r[i], wi = wo * p.stride - p.pad + a_q[i];
        if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
          v = __ldg(x + (((int64_t)n * p.H + hi) * p.W + wi) * p.Cin + a_ci[i]);
      }
      As[kk][e % BM] = v;
    }
    for (int e = threadIdx.x; e < BK * BN; e += CT) {
      int kk = e / BN, nn = e % BN;
      int64_t pix = k0 + kk;
      int n = n0 + nn;
      Bs[kk][nn] = (pix < k_end && n < p.Cout) ? __ldg(dy + pix * p.Cout + n) : 0.f;
    }
    __syncthreads();
    tile_fma<BM, BN, TM, TN>(As, Bs, acc, ty, tx);
    if (do_bias && threadIdx.x < BN) {
#pragma unroll
      for (int k = 0; k < BK; ++k) bacc += Bs[k][threadIdx.x];
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= Mdim) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < p.Cout) atomicAdd(dw + (int64_t)m * p.Cout + n, acc[i][j]);
    }
  }
  if (do_bias && threadIdx.x < BN && n0 + threadIdx.x < p.Cout) atomicAdd(db + n0 + threadIdx.x, bacc);
}

// ------------------------------------------------------------------ column sums
__global__ void __launch_bounds__(256) colsum_fast_kernel(const float* __restrict__ x, double* __restrict__ acc,
                                                          int64_t n4, int C) {
  extern __shared__ float sm[];
  float s[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldg_stream4(x + 4 * i);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
  }
  channel_reduce2<256>(s, z, C, sm, acc, acc + C);
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_generic_kernel(const T* __restrict__ x, float* __restrict__ out,
                                                             int64_t M, int C, int64_t rows_per_block) {
  extern __shared__ float sm[];  // C floats
  for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  int64_t e0 = r0 * C, e1 = r1 * C;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) atomicAdd(&sm[(int)(e % C)], to_f<T>(x[e]));
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, sm[i]);
}

__global__ void colsum_finish_kernel(const double* __restrict__ acc, float* __restrict__ out, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] += (float)acc[c];
}

static int fill_params(const dafk_conv_desc* d, ConvP& p, const char* name) {
  DAFK_REQUIRE(d != nullptr, DAFK_ERR_BAD_ARG, "%s: null descriptor", name);
  DAFK_REQUIRE(d->N >= 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 &&
                   d->stride > 0 && d->pad >= 0,
               DAFK_ERR_BAD_ARG, "%s: bad descriptor", name);
  int Ho = (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  int Wo = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  DAFK_REQUIRE(Ho == d->Ho && Wo == d->Wo && Ho > 0 && Wo > 0, DAFK_ERR_BAD_ARG,
               "%s: output size mismatch (expected %dx%d, descriptor says %dx%d)", name, Ho, Wo, d->Ho, d->Wo);
  p = ConvP{d->N, d->H, d->W, d->Cin, d->Cout, d->KH, d->KW, d->stride, d->pad, d->Ho, d->Wo};
  return DAFK_OK;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_conv2d_fwd(const dafk_conv_desc* d, const float* x, const float* w, const float* bias, float* y, int act,
                    float alpha, void* stream) {
  ConvP p;
  int rc = fill_params(d, p, "dafk_conv2d_fwd");
  if (rc) return rc;
  if (p.N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && w && y, DAFK_ERR_BAD_ARG, "dafk_conv2d_fwd: null pointer");
  DAFK_REQUIRE(act >= 0 && act <= 3, DAFK_ERR_BAD_ARG, "dafk_conv2d_fwd: bad activation");
  int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  cudaStream_t s = as_stream(stream);
  if (p.Cout > 32) {
    dim3 grid((unsigned)((M + 63) / 64), (p.Cout + 63) / 64);
    conv_fwd_kernel<64, 64, 4, 4><<<grid, CT, 0, s>>>(p, x, w, bias, y, act, alpha);
  } else if (p.Cout > 8) {
    dim3 grid((unsigned)((M + 63) / 64), (p.Cout + 31) / 32);
    conv_fwd_kernel<64, 32, 4, 2><<<grid, CT, 0, s>>>(p, x, w, bias, y, act, alpha);
  } else {
    dim3 grid((unsigned)((M + 255) / 256), 1);
    conv_fwd_kernel<256, 8, 4, 2><<<grid, CT, 0, s>>>(p, x, w, bias, y, act, alpha);
  }
  return check_launch("dafk_conv2d_fwd");
}

int dafk_conv2d_dgrad(const dafk_conv_desc* d, const float* dy, const float* w, float* dx, void* stream) {
  ConvP p;
  int rc = fill_params(d, p, "dafk_conv2d_dgrad");
  if (rc) return rc;
  if (p.N == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && w && dx, DAFK_ERR_BAD_ARG, "dafk_conv2d_dgrad: null pointer");
  int64_t M = (int64_t)p.N * p.H * p.W;
  cudaStream_t s = as_stream(stream);
  if (p.Cin > 32) {
    dim3 grid((unsigned)((M + 63) / 64), (p.Cin + 63) / 64);
    conv_dgrad_kernel<64, 64, 4, 4><<<grid, CT, 0, s>>>(p, dy, w, dx);
  } else if (p.Cin > 8) {
    dim3 grid((unsigned)((M + 63) / 64), (p.Cin + 31) / 32);
    conv_dgrad_kernel<64, 32, 4, 2><<<grid, CT, 0, s>>>(p, dy, w, dx);
  } else {
    dim3 grid((unsigned)((M + 255) / 256), 1);
    conv_dgrad_kernel<256, 8, 4, 2><<<grid, CT, 0, s>>>(p, dy, w, dx);
  }
  return check_launch("dafk_conv2d_dgrad");
}

int dafk_conv2d_wgrad(const dafk_conv_desc* d, const float* x, const float* dy, float* dw, float* db, void* stream) {
  ConvP p;
  int rc = fill_params(d, p, "dafk_conv2d_wgrad");
  if (rc) return rc;
  if (p.N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_conv2d_wgrad: null pointer");
  const int Mdim = p.KH * p.KW * p.Cin;
  const int64_t P = (int64_t)p.N * p.Ho * p.Wo;
  cudaStream_t s = as_stream(stream);
  auto splits_for = [&](int tiles) {
    int64_t want = ((int64_t)kNumSMs * 4 + tiles - 1) / tiles;
    int64_t maxs = (P + 4 * BK - 1) / (4 * BK);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return (int)want;
  };
  if (p.Cout > 32) {
    int tm = (Mdim + 63) / 64, tn = (p.Cout + 63) / 64;
    int sp = splits_for(tm * tn);
    int64_t per = ((P + sp - 1) / sp + BK - 1) / BK * BK;
    sp = (int)((P + per - 1) / per);
    conv_wgrad_kernel<64, 64, 4, 4><<<dim3(tm, tn, sp), CT, 0, s>>>(p, x, dy, dw, db, per);
  } else if (p.Cout > 8) {
    int tm = (Mdim + 63) / 64, tn = (p.Cout + 31) / 32;
    int sp = splits_for(tm * tn);
    int64_t per = ((P + sp - 1) / sp + BK - 1) / BK * BK;
    sp = (int)((P + per - 1) / per);
    conv_wgrad_kernel<64, 32, 4, 2><<<dim3(tm, tn, sp), CT, 0, s>>>(p, x, dy, dw, db, per);
  } else {
    int tm = (Mdim + 255) / 256, tn = 1;
    int sp = splits_for(tm * tn);
    int64_t per = ((P + sp - 1) / sp + BK - 1) / BK * BK;
    sp = (int)((P + per - 1) / per);
    conv_wgrad_kernel<256, 8, 4, 2><<<dim3(tm, tn, sp), CT, 0, s>>>(p, x, dy, dw, db, per);
  }
  return check_launch("dafk_conv2d_wgrad");
}

int dafk_colsum(const void* x, int x_dt, float* out, int64_t M, int C, void* stream) {
  DAFK_REQUIRE(M >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_colsum: bad shape");
  if (M == 0) return DAFK_OK;
  DAFK_REQUIRE(x && out, DAFK_ERR_BAD_ARG, "dafk_colsum: null pointer");
  cudaStream_t s = as_stream(stream);
  DAFK_REQUIRE(C <= 4096, DAFK_ERR_UNSUPPORTED, "dafk_colsum: C too large");
  if (bn_wide_ok(C) && (x_dt == DAFK_F32 || x_dt == DAFK_BF16) && DAFK_ALIGNED16(x)) {
    const int64_t n8 = M * C / 8;
    const int grid = bn_wide_grid(n8, 2, 4);
    if (x_dt == DAFK_F32) colsum_wide_kernel<float><<<grid, BW_T, 0, s>>>((const float*)x, out, n8, C);
    else colsum_wide_kernel<__nv_bfloat16><<<grid, BW_T, 0, s>>>((const __nv_bfloat16*)x, out, n8, C);
    return check_launch("dafk_colsum");
  }
  int64_t rows_per_block = (M + kNumSMs * 4 - 1) / (kNumSMs * 4);
  if (rows_per_block < 64) rows_per_block = 64;
  int blocks = (int)((M + rows_per_block - 1) / rows_per_block);
  if (x_dt == DAFK_F32) colsum_generic_kernel<float><<<blocks, 256, C * sizeof(float), s>>>((const float*)x, out, M, C, rows_per_block);
  else if (x_dt == DAFK_BF16) colsum_generic_kernel<__nv_bfloat16><<<blocks, 256, C * sizeof(float), s>>>((const __nv_bfloat16*)x, out, M, C, rows_per_block);
  else { set_error("dafk_colsum: bad dtype %d", x_dt); return DAFK_ERR_BAD_ARG; }
  return check_launch("dafk_colsum");
}

}  // extern "C"
