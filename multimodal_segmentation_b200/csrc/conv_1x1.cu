// Pointwise (1x1) convolutions between a 64-channel bf16 feature map and a handful (<= 8) of fp32 channels: the anatomy
// head `conv_anatomy` 64 -> 8 (model_components/anatomy_encoder.py:26 + models/unet.py) and the segmentor head
// 64 -> num_masks+1 = 5 (model_components/segmentor.py:24).  At 224^2 x 32 they are pure HBM streams (128 B of bf16 in,
// 20..32 B out per pixel); a tensor-core tile would be >90 % padding.
//
// Data movement: the pixel dimension is contiguous for both tensors, so a tile of 256 pixels is ONE 1-D bulk copy per
// tensor (cp.async.bulk, TMA engine, mbarrier completion) into a 3-stage shared-memory ring; one CTA per SM keeps up to
// ~100 KB of loads in flight without spending registers on it.  The ragged tail (< 256 pixels) goes through plain-load
// variants of the same arithmetic.
//
// Arithmetic: 8 lanes share one pixel, each lane owns 8 consecutive input channels (one 128-bit shared-memory read; a
// warp instruction reads 4 whole pixels = 512 contiguous bytes, conflict free) and keeps its 8 x 8 slice of the weight
// matrix in registers.
//   forward : 64 FMAs per lane, then a 7-shuffle transpose-reduce butterfly across the 8 lanes that leaves output
//             channel (lane & 7) in each lane -> the warp stores 4 x Cout contiguous floats;
//   dgrad   : every lane reads the pixel's <= 8 output gradients (broadcast inside the 8-lane group) and writes its
//             8 input-channel gradients as one 128-bit store (a warp writes 512 contiguous bytes);
//   wgrad   : 64 register accumulators per lane over all tiles of the CTA, reduced across the 4 pixel slots of the warp
//             by shuffles, across warps in shared memory and across CTAs with one float atomic per weight per CTA.
// In tensor-core mode (`round_bf16`) the weights and the incoming output gradient are rounded to bf16 when they are
// loaded, which is what the raster-strip tcgen05 kernels (conv_nc.cu) do with the operands of every other narrow layer.
#include "tc_ptx.cuh"

namespace dafk {

constexpr int P1_T = 256;        // threads per CTA
constexpr int P1_CIN = 64;
constexpr int P1_TP = 256;       // pixels per tile
constexpr int P1_STAGES = 3;
constexpr int P1_XB = P1_TP * P1_CIN * 2;     // bytes of one bf16 feature tile
constexpr int P1_GB = P1_TP * 8 * 4;          // room for one fp32 tile of <= 8 channels

__device__ __forceinline__ void p1_unpack(const uint4& t, float (&v)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 p1_pack(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float p1_round(float v, bool r) { return r ? __bfloat162float(__float2bfloat16_rn(v)) : v; }

// lane's slice of w[Cin=64][Cout]: wr[i][co] for input channels 8*(lane&7)+i, zero beyond Cout
__device__ __forceinline__ void p1_load_w(const float* __restrict__ w, int Cout, bool rnd, float (&wr)[8][8]) {
  const int ci0 = (threadIdx.x & 7) * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int co = 0; co < 8; ++co) wr[i][co] = co < Cout ? p1_round(w[(ci0 + i) * Cout + co], rnd) : 0.f;
}

// a[co] = partial dot products of this lane -> a[0] = full output channel (lane & 7) of the lane's pixel
__device__ __forceinline__ float p1_transpose_reduce(float (&a)[8], int sub) {
  {
    const bool hi = (sub & 4) != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float send = hi ? a[k] : a[k + 4], keep = hi ? a[k + 4] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = (sub & 2) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = hi ? a[k] : a[k + 2], keep = hi ? a[k + 2] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  const bool hi = (sub & 1) != 0;
  const float send = hi ? a[0] : a[1], keep = hi ? a[1] : a[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ void p1_dot(const float (&xv)[8], const float (&wr)[8][8], float (&a)[8]) {
#pragma unroll
  for (int co = 0; co < 8; ++co) a[co] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int co = 0; co < 8; ++co) a[co] = fmaf(xv[i], wr[i][co], a[co]);
}

// ring of P1_STAGES stages; thread 0 is the producer.  which: bit 0 = feature tile, bit 1 = fp32 tile
struct P1Ring {
  uint8_t* xs;
  uint8_t* gs;
  uint64_t* full;
};
__device__ __forceinline__ P1Ring p1_ring_init(uint8_t* smem) {
  P1Ring r;
  r.xs = smem;
  r.gs = smem + P1_STAGES * P1_XB;
  r.full = reinterpret_cast<uint64_t*>(smem + P1_STAGES * (P1_XB + P1_GB));
  if (threadIdx.x == 0) {
    for (int s = 0; s < P1_STAGES; ++s) mbar_init(r.full + s, 1);
    fence_barrier_init();
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ void p1_issue(const P1Ring& r, int s, int64_t tile, const __nv_bfloat16* x, const float* g,
                                         int Cout) {
  const uint32_t gb = (uint32_t)(P1_TP * Cout * 4);
  mbar_expect_tx(r.full + s, (x ? (uint32_t)P1_XB : 0u) + (g ? gb : 0u));
  if (x) bulk_load_1d(r.xs + s * P1_XB, x + tile * (int64_t)(P1_TP * P1_CIN), P1_XB, r.full + s);
  if (g) bulk_load_1d(r.gs + s * P1_GB, g + tile * (int64_t)(P1_TP * Cout), gb, r.full + s);
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(P1_T, 1) conv1x1_c64_fwd_bulk_kernel(const __nv_bfloat16* __restrict__ x,
                                                                       const float* __restrict__ w,
                                                                       const float* __restrict__ bias,
                                                                       float* __restrict__ y, int64_t tiles, int Cout,
                                                                       int rnd) {
  extern __shared__ __align__(128) uint8_t smem[];
  P1Ring r = p1_ring_init(smem);
  if (threadIdx.x == 0) {
    for (int k = 0; k < P1_STAGES; ++k) {
      const int64_t t = blockIdx.x + (int64_t)k * gridDim.x;
      if (t < tiles) p1_issue(r, k, t, x, nullptr, Cout);
    }
  }
  float wr[8][8];
  p1_load_w(w, Cout, rnd != 0, wr);
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;       // slot 0..31
  const float bv = (bias != nullptr && sub < Cout) ? bias[sub] : 0.f;
  int it = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
    const int s = it % P1_STAGES;
    mbar_wait(r.full + s, (uint32_t)(it / P1_STAGES) & 1u);
    const uint4* xt = reinterpret_cast<const uint4*>(r.xs + s * P1_XB);
    float* yt = y + t * (int64_t)(P1_TP * Cout);
#pragma unroll 2
    for (int j = 0; j < P1_TP / 32; ++j) {
      const int pix = j * 32 + slot;
      float xv[8], a[8];
      p1_unpack(xt[pix * 8 + sub], xv);
      p1_dot(xv, wr, a);
      const float o = p1_transpose_reduce(a, sub);
      if (sub < Cout) yt[pix * Cout + sub] = o + bv;
    }
    __syncthreads();
    const int64_t tn = t + (int64_t)P1_STAGES * gridDim.x;
    if (threadIdx.x == 0 && tn < tiles) p1_issue(r, s, tn, x, nullptr, Cout);
  }
}

// pixels [p_begin, M): plain loads, warp-uniform trip count (the butterfly needs all 32 lanes)
__global__ void __launch_bounds__(P1_T) conv1x1_c64_fwd_tail_kernel(const __nv_bfloat16* __restrict__ x,
                                                                    const float* __restrict__ w,
                                                                    const float* __restrict__ bias, float* __restrict__ y,
                                                                    int64_t p_begin, int64_t M, int Cout, int rnd) {
  float wr[8][8];
  p1_load_w(w, Cout, rnd != 0, wr);
  const int sub = threadIdx.x & 7;
  const float bv = (bias != nullptr && sub < Cout) ? bias[sub] : 0.f;
  for (int64_t base = p_begin + ((int64_t)blockIdx.x * (P1_T / 32) + (threadIdx.x >> 5)) * 4; base < M;
       base += (int64_t)gridDim.x * (P1_T / 8)) {
    const int64_t p = base + ((threadIdx.x >> 3) & 3);
    float xv[8], a[8];
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (p < M) raw = *reinterpret_cast<const uint4*>(x + p * P1_CIN + sub * 8);
    p1_unpack(raw, xv);
    p1_dot(xv, wr, a);
    const float o = p1_transpose_reduce(a, sub);
    if (p < M && sub < Cout) y[p * Cout + sub] = o + bv;
  }
}

// ---------------------------------------------------------------- data gradient
__device__ __forceinline__ void p1_dgrad_pixel(const float* gsrc, int Cout, bool rnd, const float (&wr)[8][8], uint4* dst) {
  float g[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) g[co] = co < Cout ? p1_round(gsrc[co], rnd) : 0.f;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int co = 0; co < 8; ++co) acc = fmaf(g[co], wr[i][co], acc);
    o[i] = acc;
  }
  const uint4 pk = p1_pack(o);
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w));
}

__global__ void __launch_bounds__(P1_T, 1) conv1x1_c64_dgrad_bulk_kernel(const float* __restrict__ dy,
                                                                         const float* __restrict__ w,
                                                                         __nv_bfloat16* __restrict__ dx, int64_t tiles,
                                                                         int Cout, int rnd) {
  extern __shared__ __align__(128) uint8_t smem[];
  P1Ring r = p1_ring_init(smem);
  if (threadIdx.x == 0) {
    for (int k = 0; k < P1_STAGES; ++k) {
      const int64_t t = blockIdx.x + (int64_t)k * gridDim.x;
      if (t < tiles) p1_issue(r, k, t, nullptr, dy, Cout);
    }
  }
  float wr[8][8];
  p1_load_w(w, Cout, rnd != 0, wr);
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  int it = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
    const int s = it % P1_STAGES;
    mbar_wait(r.full + s, (uint32_t)(it / P1_STAGES) & 1u);
    const float* gt = reinterpret_cast<const float*>(r.gs + s * P1_GB);
    uint4* dt = reinterpret_cast<uint4*>(dx + t * (int64_t)(P1_TP * P1_CIN));
#pragma unroll 2
    for (int j = 0; j < P1_TP / 32; ++j) {
      const int pix = j * 32 + slot;
      p1_dgrad_pixel(gt + pix * Cout, Cout, rnd != 0, wr, dt + pix * 8 + sub);
    }
    __syncthreads();
    const int64_t tn = t + (int64_t)P1_STAGES * gridDim.x;
    if (threadIdx.x == 0 && tn < tiles) p1_issue(r, s, tn, nullptr, dy, Cout);
  }
}

__global__ void __launch_bounds__(P1_T) conv1x1_c64_dgrad_tail_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                      __nv_bfloat16* __restrict__ dx, int64_t p_begin,
                                                                      int64_t M, int Cout, int rnd) {
  float wr[8][8];
  p1_load_w(w, Cout, rnd != 0, wr);
  const int sub = threadIdx.x & 7;
  for (int64_t p = p_begin + (int64_t)blockIdx.x * (P1_T / 8) + (threadIdx.x >> 3); p < M;
       p += (int64_t)gridDim.x * (P1_T / 8))
    p1_dgrad_pixel(dy + p * Cout, Cout, rnd != 0, wr, reinterpret_cast<uint4*>(dx + p * P1_CIN) + sub);
}

// ---------------------------------------------------------------- weight gradient
struct P1Acc {
  float acc[8][8];
  float bs[8];
};
__device__ __forceinline__ void p1_acc_zero(P1Acc& A) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    A.bs[i] = 0.f;
#pragma unroll
    for (int co = 0; co < 8; ++co) A.acc[i][co] = 0.f;
  }
}
__device__ __forceinline__ void p1_acc_pixel(P1Acc& A, const uint4& raw, const float* gsrc, int Cout, bool rnd) {
  float xv[8], g[8];
  p1_unpack(raw, xv);
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    const float v = co < Cout ? gsrc[co] : 0.f;
    A.bs[co] += v;                       // bias gradient: the unrounded column sum
    g[co] = p1_round(v, rnd);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int co = 0; co < 8; ++co) A.acc[i][co] = fmaf(xv[i], g[co], A.acc[i][co]);
}
// warp -> CTA (shared memory) -> global (one float atomic per weight per CTA)
__device__ __forceinline__ void p1_acc_flush(P1Acc& A, float* red /* 64*8+8 floats */, float* __restrict__ dw,
                                             float* __restrict__ db, int Cout) {
  const int sub = threadIdx.x & 7, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int co = 0; co < 8; ++co) {
      float v = A.acc[i][co];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      A.acc[i][co] = v;
    }
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    float v = A.bs[co];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    A.bs[co] = v;
  }
  for (int i = threadIdx.x; i < P1_CIN * 8 + 8; i += P1_T) red[i] = 0.f;
  __syncthreads();
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int co = 0; co < 8; ++co) atomicAdd(&red[(sub * 8 + i) * 8 + co], A.acc[i][co]);
    if (lane == 0) {
#pragma unroll
      for (int co = 0; co < 8; ++co) atomicAdd(&red[P1_CIN * 8 + co], A.bs[co]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P1_CIN * 8; i += P1_T) {
    const int ci = i >> 3, co = i & 7;
    if (co < Cout) atomicAdd(dw + ci * Cout + co, red[i]);
  }
  if (db != nullptr && threadIdx.x < Cout) atomicAdd(db + threadIdx.x, red[P1_CIN * 8 + threadIdx.x]);
}

__global__ void __launch_bounds__(P1_T, 1) conv1x1_c64_wgrad_bulk_kernel(const __nv_bfloat16* __restrict__ x,
                                                                         const float* __restrict__ dy,
                                                                         float* __restrict__ dw, float* __restrict__ db,
                                                                         int64_t tiles, int Cout, int rnd) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float red[P1_CIN * 8 + 8];
  P1Ring r = p1_ring_init(smem);
  if (threadIdx.x == 0) {
    for (int k = 0; k < P1_STAGES; ++k) {
      const int64_t t = blockIdx.x + (int64_t)k * gridDim.x;
      if (t < tiles) p1_issue(r, k, t, x, dy, Cout);
    }
  }
  P1Acc A;
  p1_acc_zero(A);
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  int it = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
    const int s = it % P1_STAGES;
    mbar_wait(r.full + s, (uint32_t)(it / P1_STAGES) & 1u);
    const uint4* xt = reinterpret_cast<const uint4*>(r.xs + s * P1_XB);
    const float* gt = reinterpret_cast<const float*>(r.gs + s * P1_GB);
#pragma unroll 2
    for (int j = 0; j < P1_TP / 32; ++j) {
      const int pix = j * 32 + slot;
      p1_acc_pixel(A, xt[pix * 8 + sub], gt + pix * Cout, Cout, rnd != 0);
    }
    __syncthreads();
    const int64_t tn = t + (int64_t)P1_STAGES * gridDim.x;
    if (threadIdx.x == 0 && tn < tiles) p1_issue(r, s, tn, x, dy, Cout);
  }
  p1_acc_flush(A, red, dw, db, Cout);
}

__global__ void __launch_bounds__(P1_T) conv1x1_c64_wgrad_tail_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      const float* __restrict__ dy, float* __restrict__ dw,
                                                                      float* __restrict__ db, int64_t p_begin, int64_t M,
                                                                      int Cout, int rnd) {
  __shared__ float red[P1_CIN * 8 + 8];
  P1Acc A;
  p1_acc_zero(A);
  const int sub = threadIdx.x & 7;
  for (int64_t p = p_begin + (int64_t)blockIdx.x * (P1_T / 8) + (threadIdx.x >> 3); p < M;
       p += (int64_t)gridDim.x * (P1_T / 8))
    p1_acc_pixel(A, *reinterpret_cast<const uint4*>(x + p * P1_CIN + sub * 8), dy + p * Cout, Cout, rnd != 0);
  p1_acc_flush(A, red, dw, db, Cout);
}

constexpr size_t P1_SMEM = (size_t)P1_STAGES * (P1_XB + P1_GB) + 64;

template <typename K>
static inline bool p1_set_smem(K kernel) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P1_SMEM) == cudaSuccess;
}

}  // namespace dafk

using namespace dafk;

#define P1_CHECKS(name)                                                                                                  \
  DAFK_REQUIRE(dafk_conv1x1_supported(Cin, Cout), DAFK_ERR_UNSUPPORTED, name ": Cin must be 64 and 1 <= Cout <= 8 (got %d -> %d)", Cin, Cout); \
  DAFK_REQUIRE(M >= 0, DAFK_ERR_BAD_ARG, name ": bad size");                                                            \
  if (M == 0) return DAFK_OK;

extern "C" {

int dafk_conv1x1_supported(int Cin, int Cout) { return Cin == P1_CIN && Cout >= 1 && Cout <= 8; }

int dafk_conv1x1_fwd(const void* x, const float* w, const float* bias, float* y, int64_t M, int Cin, int Cout,
                     int round_bf16, void* stream) {
  P1_CHECKS("dafk_conv1x1_fwd");
  DAFK_REQUIRE(x && w && y, DAFK_ERR_BAD_ARG, "dafk_conv1x1_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN, "dafk_conv1x1_fwd: alignment");
  cudaStream_t s = as_stream(stream);
  const int64_t tiles = M / P1_TP;
  if (tiles > 0) {
    DAFK_REQUIRE(p1_set_smem(conv1x1_c64_fwd_bulk_kernel), DAFK_ERR_CUDA, "dafk_conv1x1_fwd: shared memory opt-in failed");
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    conv1x1_c64_fwd_bulk_kernel<<<grid, P1_T, P1_SMEM, s>>>((const __nv_bfloat16*)x, w, bias, y, tiles, Cout, round_bf16);
    int rc = check_launch("dafk_conv1x1_fwd");
    if (rc) return rc;
  }
  if (tiles * P1_TP < M) {
    conv1x1_c64_fwd_tail_kernel<<<(int)((M - tiles * P1_TP + 31) / 32), P1_T, 0, s>>>((const __nv_bfloat16*)x, w, bias, y,
                                                                                    tiles * P1_TP, M, Cout, round_bf16);
    return check_launch("dafk_conv1x1_fwd(tail)");
  }
  return DAFK_OK;
}

int dafk_conv1x1_dgrad(const float* dy, const float* w, void* dx, int64_t M, int Cin, int Cout, int round_bf16,
                       void* stream) {
  P1_CHECKS("dafk_conv1x1_dgrad");
  DAFK_REQUIRE(dy && w && dx, DAFK_ERR_BAD_ARG, "dafk_conv1x1_dgrad: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dx) && DAFK_ALIGNED16(dy), DAFK_ERR_ALIGN, "dafk_conv1x1_dgrad: alignment");
  cudaStream_t s = as_stream(stream);
  const int64_t tiles = M / P1_TP;
  if (tiles > 0) {
    DAFK_REQUIRE(p1_set_smem(conv1x1_c64_dgrad_bulk_kernel), DAFK_ERR_CUDA, "dafk_conv1x1_dgrad: shared memory opt-in failed");
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    conv1x1_c64_dgrad_bulk_kernel<<<grid, P1_T, P1_SMEM, s>>>(dy, w, (__nv_bfloat16*)dx, tiles, Cout, round_bf16);
    int rc = check_launch("dafk_conv1x1_dgrad");
    if (rc) return rc;
  }
  if (tiles * P1_TP < M) {
    conv1x1_c64_dgrad_tail_kernel<<<(int)((M - tiles * P1_TP + 31) / 32), P1_T, 0, s>>>(dy, w, (__nv_bfloat16*)dx,
                                                                                      tiles * P1_TP, M, Cout, round_bf16);
    return check_launch("dafk_conv1x1_dgrad(tail)");
  }
  return DAFK_OK;
}

int dafk_conv1x1_wgrad(const void* x, const float* dy, float* dw, float* db, int64_t M, int Cin, int Cout, int round_bf16,
                       void* stream) {
  P1_CHECKS("dafk_conv1x1_wgrad");
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_conv1x1_wgrad: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dy), DAFK_ERR_ALIGN, "dafk_conv1x1_wgrad: alignment");
  cudaStream_t s = as_stream(stream);
  const int64_t tiles = M / P1_TP;
  if (tiles > 0) {
    DAFK_REQUIRE(p1_set_smem(conv1x1_c64_wgrad_bulk_kernel), DAFK_ERR_CUDA, "dafk_conv1x1_wgrad: shared memory opt-in failed");
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    conv1x1_c64_wgrad_bulk_kernel<<<grid, P1_T, P1_SMEM, s>>>((const __nv_bfloat16*)x, dy, dw, db, tiles, Cout, round_bf16);
    int rc = check_launch("dafk_conv1x1_wgrad");
    if (rc) return rc;
  }
  if (tiles * P1_TP < M) {
    conv1x1_c64_wgrad_tail_kernel<<<1, P1_T, 0, s>>>((const __nv_bfloat16*)x, dy, dw, db, tiles * P1_TP, M, Cout, round_bf16);
    return check_launch("dafk_conv1x1_wgrad(tail)");
  }
  return DAFK_OK;
}

}  // extern "C"
