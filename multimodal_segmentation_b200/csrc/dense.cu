// Dense layers with a tiny batch (B <= a few dozen) and a long reduction axis
// (K up to 270 848): weight-read-bound skinny GEMMs.
//   fwd        : split-K over CTAs, W streamed once with coalesced rows, x chunk staged in smem
//   bwd_data   : one thread per k, dy staged in smem, W row walked once
//   bwd_weight : one thread per (k,n), dy staged in smem
#include "common.cuh"

namespace dafk {

constexpr int DT = 256;
constexpr int KC = 128;   // k-chunk per CTA in the forward
constexpr int MAXB = 32;  // batch rows held in registers

__global__ void dense_init_kernel(float* __restrict__ y, const float* __restrict__ bias, int B, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * N) y[i] = bias ? bias[i % N] : 0.f;
}

// grid = (k-slabs, ceil(B/MAXB)); y must be pre-initialised with the bias.
// Each CTA walks its k-slab in chunks of KC (x chunk staged in smem, W rows streamed coalesced),
// keeps MAXB accumulators per thread in registers, then reduces across the threads that share an
// output column (warp shuffles when the column count divides 32, shared-memory atomics otherwise)
// and issues ONE global atomic per output per CTA.
__global__ void __launch_bounds__(DT) dense_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       float* __restrict__ y, int B, int64_t K, int N,
                                                       int64_t k_per_cta) {
  __shared__ float xs[MAXB][KC];
  extern __shared__ float part[];   // [MAXB][ncols]
  const int b0 = blockIdx.y * MAXB;
  const int nb = min(MAXB, B - b0);
  const int64_t kbeg = (int64_t)blockIdx.x * k_per_cta;
  const int64_t kend = min(K, kbeg + k_per_cta);
  for (int nbase = 0; nbase < N; nbase += DT) {
    const int ncols = min(N - nbase, DT);
    const int groups = DT / ncols;
    const int n = threadIdx.x % ncols, g = threadIdx.x / ncols;
    const bool active = g < groups;
    float acc[MAXB];
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
    for (int64_t k0 = kbeg; k0 < kend; k0 += KC) {
      const int kc = (int)min((int64_t)KC, kend - k0);
      __syncthreads();
      for (int e = threadIdx.x; e < MAXB * KC; e += DT) {
        int b = e / KC, k = e % KC;
        xs[b][k] = (b < nb && k < kc) ? x[(int64_t)(b0 + b) * K + k0 + k] : 0.f;
      }
      __syncthreads();
      if (active) {
#pragma unroll 4
        for (int k = g; k < kc; k += groups) {
          float wv = __ldg(w + (k0 + k) * N + nbase + n);
#pragma unroll
          for (int b = 0; b < MAXB; ++b) acc[b] = fmaf(xs[b][k], wv, acc[b]);
        }
      }
    }
    // reduce over the k-groups
    for (int e = threadIdx.x; e < MAXB * ncols; e += DT) part[e] = 0.f;
    __syncthreads();
    const bool shuffle_ok = ncols < 32 && (32 % ncols) == 0;
    if (shuffle_ok) {
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        float v = active ? acc[b] : 0.f;
        for (int o = 16; o >= ncols; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[b] = v;
      }
      if ((threadIdx.x & 31) < ncols && active) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) atomicAdd(&part[b * ncols + n], acc[b]);
      }
    } else if (active) {
#pragma unroll
      for (int b = 0; b < MAXB; ++b) atomicAdd(&part[b * ncols + n], acc[b]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nb * ncols; e += DT) {
      int b = e / ncols, c = e % ncols;
      atomicAdd(y + (int64_t)(b0 + b) * N + nbase + c, part[e]);
    }
    __syncthreads();
  }
}

// Short reduction axis (K <= 2048: the FiLM gamma / beta heads 8 -> 8, z_mean / z_log_var 32 -> 8, locnet 100 -> 50):
// one thread per output, bias included, ONE launch (the split-K path needs an initialisation launch and atomics, and a
// single CTA walking K with dependent loads took 25 us for a 32 x 100 x 50 product).
__global__ void __launch_bounds__(DT) dense_fwd_shortk_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y, int B,
                                                              int K, int N) {
  const int o = blockIdx.x * DT + threadIdx.x;
  if (o >= B * N) return;
  const int b = o / N, n = o - b * N;
  const float* xr = x + (int64_t)b * K;
  const float* wc = w + n;
  float acc = bias ? bias[n] : 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) acc = fmaf(__ldg(xr + k), __ldg(wc + (int64_t)k * N), acc);
  y[o] = acc;
}

// Register-tiled forward for N % 4 == 0 (locnet Dense(100) over 48 020 features, layers/stn_spline.py:114; modality
// encoder Dense(32) over 21 632, model_components/modality_encoder.py:44): the kernel above reads one shared-memory word
// per FMA (32 broadcast LDS per weight element) and ran the 19 MB weight stream at 230 GB/s.  Here a thread owns an
// 8 (batch rows) x 4 (columns) tile: per k it reads its 8 x values as two 128-bit shared-memory loads (x staged transposed,
// [k][batch], 36-float pitch) and 4 weights as one 128-bit global load -> 32 FMAs per 3 loads.  NG = N/4 column groups x 4
// batch groups = one k-slice; 256 / (4*NG) slices walk the chunk in parallel and are reduced through shared memory.
constexpr int RT_PITCH = 36;
__global__ void __launch_bounds__(DT) dense_fwd_rt_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          float* __restrict__ y, int B, int64_t K, int N, int64_t k_per_cta) {
  __shared__ __align__(16) float xs[KC * RT_PITCH];
  extern __shared__ float part[];   // [MAXB][N]
  const int b0 = blockIdx.y * MAXB;
  const int nb = min(MAXB, B - b0);
  const int64_t kbeg = (int64_t)blockIdx.x * k_per_cta;
  const int64_t kend = min(K, kbeg + k_per_cta);
  const int NG = N >> 2;
  const int per_slice = 4 * NG;
  const int slices = DT / per_slice;
  const int sl = threadIdx.x / per_slice;
  const int r = threadIdx.x - sl * per_slice;
  const int bg = r / NG, ng = r - bg * NG;
  const bool active = sl < slices;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int e = threadIdx.x; e < MAXB * N; e += DT) part[e] = 0.f;
  for (int64_t k0 = kbeg; k0 < kend; k0 += KC) {
    const int kc = (int)min((int64_t)KC, kend - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < MAXB * KC; e += DT) {      // coalesced along k, transposed into [k][batch]
      const int b = e / KC, k = e - b * KC;
      xs[k * RT_PITCH + b] = (b < nb && k < kc) ? x[(int64_t)(b0 + b) * K + k0 + k] : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* wp = w + k0 * N + ng * 4;
      // the weight stream is the only HBM traffic: eight independent 128-bit loads in flight per thread (one per
      // iteration left each thread waiting a full DRAM latency per k: 85 us for 19 MB)
#pragma unroll 8
      for (int k = sl; k < kc; k += slices) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)k * N));
        const float4 xa = *reinterpret_cast<const float4*>(xs + k * RT_PITCH + bg * 8);
        const float4 xb = *reinterpret_cast<const float4*>(xs + k * RT_PITCH + bg * 8 + 4);
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float wj[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wj[j], acc[i][j]);
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&part[(bg * 8 + i) * N + ng * 4 + j], acc[i][j]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * N; e += DT) atomicAdd(y + (int64_t)b0 * N + e, part[e]);
}

// Register-tiled weight gradient for N % 4 == 0, K % 4 == 0: a thread owns 4 (k) x 4 (n) outputs; per batch row it reads x
// as one 128-bit global load (coalesced along k) and dy as one 128-bit shared-memory load -> 16 FMAs per 2 loads (the
// kernel below: 2 loads per FMA).
__global__ void __launch_bounds__(DT) dense_bwd_weight_rt_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 float* __restrict__ dw, float* __restrict__ db, int B,
                                                                 int64_t K, int N) {
  extern __shared__ __align__(16) float dys[];  // [B][N]
  for (int e = threadIdx.x; e < B * N; e += DT) dys[e] = dy[e];
  __syncthreads();
  if (db && blockIdx.x == 0) {
    for (int n = threadIdx.x; n < N; n += DT) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += dys[b * N + n];
      db[n] += s;
    }
  }
  const int NG = N >> 2;
  const int64_t tiles = (K >> 2) * NG;
  for (int64_t t = (int64_t)blockIdx.x * DT + threadIdx.x; t < tiles; t += (int64_t)gridDim.x * DT) {
    const int64_t kq = t / NG;
    const int ng = (int)(t - kq * NG);
    const int64_t k = kq << 2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int b = 0; b < B; ++b) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (int64_t)b * K + k));
      const float4 dv = *reinterpret_cast<const float4*>(dys + b * N + ng * 4);
      const float xi[4] = {xv.x, xv.y, xv.z, xv.w};
      const float dj[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xi[i], dj[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4* o = reinterpret_cast<float4*>(dw + (k + i) * N + ng * 4);
      float4 cur = *o;
      cur.x += acc[i][0]; cur.y += acc[i][1]; cur.z += acc[i][2]; cur.w += acc[i][3];
      *o = cur;
    }
  }
}

// Few outputs (N <= 4: the discriminators' Dense(1) over 270 848 features, models/discriminator.py:41): the op is a
// handful of long dot products per sample, bound by reading x once.  grid = (k-slabs, B); every thread streams 128-bit
// pieces of its sample's row and of the weight slab (which stays in L2 across the samples), block-reduces N partial sums
// and adds them to the bias-initialised output with one atomic per output per CTA.
template <int N>
__global__ void __launch_bounds__(DT) dense_fwd_smalln_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              float* __restrict__ y, int64_t K, int64_t k_per_cta) {
  __shared__ float red[N][DT / 32];
  const int b = blockIdx.y;
  const int64_t kbeg = (int64_t)blockIdx.x * k_per_cta;
  const int64_t kend = min(K, kbeg + k_per_cta);
  const float* xr = x + (int64_t)b * K;
  float acc[N];
#pragma unroll
  for (int n = 0; n < N; ++n) acc[n] = 0.f;
  for (int64_t k = kbeg + 4 * (int64_t)threadIdx.x; k < kend; k += 4 * DT) {      // K % 4 == 0, slabs multiples of 4
    const float4 xv = ldg_stream4(xr + k);
    const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
    if (N == 1) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
      acc[0] = fmaf(xs4[0], wv.x, fmaf(xs4[1], wv.y, fmaf(xs4[2], wv.z, fmaf(xs4[3], wv.w, acc[0]))));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = fmaf(xs4[j], __ldg(w + (k + j) * N + n), acc[n]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float v = warp_sum(acc[n]);
    if (lane == 0) red[n][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < DT / 32; ++i) t += red[threadIdx.x][i];
    atomicAdd(y + (int64_t)b * N + threadIdx.x, t);
  }
}

// grid = (ceil(K/DT), ceil(B/MAXB)); dynamic smem: MAXB*N floats
__global__ void __launch_bounds__(DT) dense_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                            float* __restrict__ dx, int B, int64_t K, int N) {
  extern __shared__ float dys[];  // [MAXB][N]
  const int b0 = blockIdx.y * MAXB;
  const int nb = min(MAXB, B - b0);
  for (int e = threadIdx.x; e < MAXB * N; e += DT) {
    int b = e / N, n = e % N;
    dys[e] = (b < nb) ? dy[(int64_t)(b0 + b) * N + n] : 0.f;
  }
  __syncthreads();
  int64_t k = (int64_t)blockIdx.x * DT + threadIdx.x;
  if (k >= K) return;
  float acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
  const float* wr = w + k * N;
#pragma unroll 4
  for (int n = 0; n < N; ++n) {
    float wv = __ldg(wr + n);
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[b] = fmaf(dys[b * N + n], wv, acc[b]);
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b)
    if (b < nb) dx[(int64_t)(b0 + b) * K + k] = acc[b];
}

// one thread per (k,n); dynamic smem: B*N floats
__global__ void __launch_bounds__(DT) dense_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dw, float* __restrict__ db, int B,
                                                              int64_t K, int N) {
  extern __shared__ float dys[];  // [B][N]
  for (int e = threadIdx.x; e < B * N; e += DT) dys[e] = dy[e];
  __syncthreads();
  if (db && blockIdx.x == 0) {
    for (int n = threadIdx.x; n < N; n += DT) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += dys[b * N + n];
      db[n] += s;
    }
  }
  int64_t total = K * N;
  int64_t stride = (int64_t)gridDim.x * DT;
  for (int64_t o = (int64_t)blockIdx.x * DT + threadIdx.x; o < total; o += stride) {
    int64_t k = o / N;
    int n = (int)(o - k * N);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(__ldg(x + (int64_t)b * K + k), dys[b * N + n], s);
    dw[o] += s;
  }
}

// wide-output variants (SPADE decoder: Dense(8 -> H*W*128/1024), model_components/decoder.py:68): dy does not fit in
// shared memory; dy rows are read coalesced from L2 instead
__global__ void __launch_bounds__(DT) dense_bwd_weight_wide_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                   float* __restrict__ dw, float* __restrict__ db, int B,
                                                                   int64_t K, int N) {
  int64_t total = K * N;
  int64_t stride = (int64_t)gridDim.x * DT;
  for (int64_t o = (int64_t)blockIdx.x * DT + threadIdx.x; o < total; o += stride) {
    int64_t k = o / N;
    int n = (int)(o - k * N);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(__ldg(x + (int64_t)b * K + k), __ldg(dy + (int64_t)b * N + n), s);
    dw[o] += s;
  }
  if (db) {
    for (int64_t n = (int64_t)blockIdx.x * DT + threadIdx.x; n < N; n += stride) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += __ldg(dy + (int64_t)b * N + n);
      db[n] += s;
    }
  }
}

// one warp per (b, k): dx[b,k] = sum_n dy[b,n] * w[k,n]
__global__ void __launch_bounds__(DT) dense_bwd_data_wide_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                 float* __restrict__ dx, int B, int64_t K, int N) {
  const int64_t warp = ((int64_t)blockIdx.x * DT + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (int64_t)B * K) return;
  const int64_t b = warp / K, k = warp - b * K;
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s = fmaf(__ldg(dy + b * N + n), __ldg(w + k * N + n), s);
  s = warp_sum(s);
  if (lane == 0) dx[b * K + k] = s;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_dense_fwd(const float* x, const float* w, const float* bias, float* y, int B, int64_t K, int Nout,
                   void* stream) {
  DAFK_REQUIRE(B >= 0 && K > 0 && Nout > 0, DAFK_ERR_BAD_ARG, "dafk_dense_fwd: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(x && w && y, DAFK_ERR_BAD_ARG, "dafk_dense_fwd: null pointer");
  cudaStream_t s = as_stream(stream);
  if (K <= 2048 && (int64_t)B * Nout >= 64) {
    dense_fwd_shortk_kernel<<<(B * Nout + DT - 1) / DT, DT, 0, s>>>(x, w, bias, y, B, (int)K, Nout);
    return check_launch("dafk_dense_fwd");
  }
  dense_init_kernel<<<(B * Nout + 255) / 256, 256, 0, s>>>(y, bias, B, Nout);
  int rc = check_launch("dafk_dense_fwd(init)");
  if (rc) return rc;
  if (Nout <= 4 && K % 4 == 0 && DAFK_ALIGNED16(x) && DAFK_ALIGNED16(w)) {
    // few long dot products per sample: x-stationary stream (see dense_fwd_smalln_kernel)
    int64_t per = (K / 4 + (4 * kNumSMs / B > 0 ? 4 * kNumSMs / B : 1) - 1) / (4 * kNumSMs / B > 0 ? 4 * kNumSMs / B : 1);
    if (per < DT) per = DT;
    const int64_t k_per = per * 4;
    dim3 g2((unsigned)((K + k_per - 1) / k_per), (unsigned)B);
    switch (Nout) {
      case 1: dense_fwd_smalln_kernel<1><<<g2, DT, 0, s>>>(x, w, y, K, k_per); break;
      case 2: dense_fwd_smalln_kernel<2><<<g2, DT, 0, s>>>(x, w, y, K, k_per); break;
      case 3: dense_fwd_smalln_kernel<3><<<g2, DT, 0, s>>>(x, w, y, K, k_per); break;
      default: dense_fwd_smalln_kernel<4><<<g2, DT, 0, s>>>(x, w, y, K, k_per); break;
    }
    return check_launch("dafk_dense_fwd");
  }
  // k-slabs: about two CTAs per SM, each slab a multiple of the chunk size
  int64_t slabs = (K + KC - 1) / KC;
  int64_t want = 2 * kNumSMs;
  int64_t k_per_cta = ((slabs + want - 1) / want) * KC;
  dim3 grid((unsigned)((K + k_per_cta - 1) / k_per_cta), (B + MAXB - 1) / MAXB);
  if (Nout % 4 == 0 && Nout >= 8 && 4 * (Nout / 4) <= DT && K >= 4 * KC && DAFK_ALIGNED16(w) &&
      sizeof(float) * MAXB * Nout + sizeof(float) * KC * RT_PITCH <= 48 * 1024) {
    dense_fwd_rt_kernel<<<grid, DT, sizeof(float) * MAXB * Nout, s>>>(x, w, y, B, K, Nout, k_per_cta);
    return check_launch("dafk_dense_fwd");
  }
  size_t smem = sizeof(float) * MAXB * (Nout < DT ? Nout : DT);
  dense_fwd_kernel<<<grid, DT, smem, s>>>(x, w, y, B, K, Nout, k_per_cta);
  return check_launch("dafk_dense_fwd");
}

int dafk_dense_bwd_data(const float* dy, const float* w, float* dx, int B, int64_t K, int Nout, void* stream) {
  DAFK_REQUIRE(B >= 0 && K > 0 && Nout > 0, DAFK_ERR_BAD_ARG, "dafk_dense_bwd_data: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && w && dx, DAFK_ERR_BAD_ARG, "dafk_dense_bwd_data: null pointer");
  size_t smem = sizeof(float) * MAXB * Nout;
  if (smem > 48 * 1024) {
    const int64_t warps = (int64_t)B * K;
    dense_bwd_data_wide_kernel<<<(unsigned)((warps * 32 + DT - 1) / DT), DT, 0, as_stream(stream)>>>(dy, w, dx, B, K, Nout);
    return check_launch("dafk_dense_bwd_data");
  }
  dim3 grid((unsigned)((K + DT - 1) / DT), (B + MAXB - 1) / MAXB);
  dense_bwd_data_kernel<<<grid, DT, smem, as_stream(stream)>>>(dy, w, dx, B, K, Nout);
  return check_launch("dafk_dense_bwd_data");
}

int dafk_dense_bwd_weight(const float* x, const float* dy, float* dw, float* db, int B, int64_t K, int Nout,
                          void* stream) {
  DAFK_REQUIRE(B >= 0 && K > 0 && Nout > 0, DAFK_ERR_BAD_ARG, "dafk_dense_bwd_weight: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_dense_bwd_weight: null pointer");
  size_t smem = sizeof(float) * B * Nout;
  int64_t total = K * Nout;
  int grid = bw_grid(total, DT, 8);
  if (smem > 48 * 1024) {
    dense_bwd_weight_wide_kernel<<<grid, DT, 0, as_stream(stream)>>>(x, dy, dw, db, B, K, Nout);
    return check_launch("dafk_dense_bwd_weight");
  }
  if (Nout % 4 == 0 && K % 4 == 0 && K >= 1024 && DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dw)) {
    const int g4 = bw_grid(total / 16, DT, 8);
    dense_bwd_weight_rt_kernel<<<g4, DT, smem, as_stream(stream)>>>(x, dy, dw, db, B, K, Nout);
    return check_launch("dafk_dense_bwd_weight");
  }
  dense_bwd_weight_kernel<<<grid, DT, smem, as_stream(stream)>>>(x, dy, dw, db, B, K, Nout);
  return check_launch("dafk_dense_bwd_weight");
}

}  // extern "C"
