// bf16-storage variants of the element-wise passes of the FiLM decoder (model_components/decoder.py:44-54,
// layers/film.py:26-36): the fused block tail  y = res + act(x * gamma + beta)  and the activation backward.
// Arithmetic is fp32 in registers; only the feature maps are stored in bf16 (8 channels = one 16 B vector), which
// halves the HBM bytes of the decoder and lets the 8 -> 8 convolutions stage rows with cp.async.bulk
// (conv_nc.cu, BULK).  gamma / beta and their gradients stay fp32 [B, C].
//
// These kernels are only reached when the decoder keeps its activations in bf16 (engine.DEC_BF16, opt-in).
#include "common.cuh"
#include "reduce.cuh"

namespace dafk {

constexpr int DB_TPB = 256;

__device__ __forceinline__ void db_ld8(const __nv_bfloat16* p, float (&v)[8]) {
  uint32_t w[4];
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(h);
    v[2 * i + 1] = __high2float(h);
  }
}
__device__ __forceinline__ void db_st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
}
__device__ __forceinline__ void db_ld8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ float db_act(float z, int act, float alpha) {
  return (act == DAFK_ACT_LRELU) ? (z > 0.f ? z : __fmul_rn(alpha, z)) : ((act == DAFK_ACT_RELU) ? fmaxf(z, 0.f) : z);
}
__device__ __forceinline__ float db_act_grad(float z, int act, float alpha) {
  if (act == DAFK_ACT_LRELU) return z > 0.f ? 1.f : (z < 0.f ? alpha : 0.f);
  if (act == DAFK_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}

// x, res, y: [B, HW, C] bf16; gamma, beta: [B, C] f32; one 8-channel vector per thread-iteration
__global__ void __launch_bounds__(DB_TPB) film_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               const __nv_bfloat16* __restrict__ res,
                                                               __nv_bfloat16* __restrict__ y, int64_t HWC, int C,
                                                               int64_t n8, int act, float alpha) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int64_t e = i << 3;
    const int64_t b = e / HWC;
    const int c = (int)(e % C);
    float v[8], g[8], t[8];
    db_ld8(x + e, v);
    db_ld8f(gamma + b * C + c, g);
    db_ld8f(beta + b * C + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = db_act(v[k] * g[k] + t[k], act, alpha);
    if (res) {
      float r[8];
      db_ld8(res + e, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += r[k];
    }
    db_st8(y + e, v);
  }
}

// grid = (chunks, B); every thread owns a fixed group of 8 channels (C a power of two, 8 <= C <= 8 * DB_TPB).
// dx = dy * act'(x*gamma+beta) * gamma;  dgamma[b,c] = sum dy*act'*x;  dbeta[b,c] = sum dy*act'
__global__ void __launch_bounds__(DB_TPB) film_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ dx, double* __restrict__ ws,
                                                               int64_t HWC, int C, int act, float alpha) {
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const __nv_bfloat16* dyb = dy + (int64_t)b * HWC;
  const __nv_bfloat16* xb = x + (int64_t)b * HWC;
  __nv_bfloat16* dxb = dx + (int64_t)b * HWC;
  const int c = (threadIdx.x * 8) % C;
  float g[8], t[8];
  db_ld8f(gamma + (int64_t)b * C + c, g);
#pragma unroll
  for (int k = 0; k < 8; ++k) t[k] = 0.f;
  if (act != DAFK_ACT_NONE) db_ld8f(beta + (int64_t)b * C + c, t);
  float ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) ag[k] = ab[k] = 0.f;
  const int64_t n8 = HWC >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float d[8], v[8];
    db_ld8(dyb + 8 * i, d);
    db_ld8(xb + 8 * i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (act != DAFK_ACT_NONE) d[k] *= db_act_grad(v[k] * g[k] + t[k], act, alpha);
      ag[k] += d[k] * v[k];
      ab[k] += d[k];
      d[k] *= g[k];
    }
    db_st8(dxb + 8 * i, d);
  }
  // per-(b, c) reduction: lanes that share a channel group (l % G == l' % G, G = C/8 groups) -> shared atomics ->
  // one double atomic per channel per CTA
  for (int i = threadIdx.x; i < 2 * C; i += DB_TPB) sm[i] = 0.f;
  __syncthreads();
  const int G = C >> 3;
  if (G < 32) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      if (o >= G) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          ag[k] += __shfl_xor_sync(0xffffffffu, ag[k], o);
          ab[k] += __shfl_xor_sync(0xffffffffu, ab[k], o);
        }
      }
    }
  }
  if (G >= 32 || (int)(threadIdx.x & 31) < G) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&sm[c + k], ag[k]);
      atomicAdd(&sm[C + c + k], ab[k]);
    }
  }
  __syncthreads();
  double* wa = ws + (int64_t)b * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += DB_TPB) atomicAdd(wa + i, (double)sm[i]);
}

__global__ void film_bwd_bf16_finish_kernel(const double* __restrict__ ws, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * C) {
    const int b = i / C, c = i % C;
    dgamma[i] = (float)ws[(int64_t)b * 2 * C + c];
    dbeta[i] = (float)ws[(int64_t)b * 2 * C + C + c];
  }
}

// dx = dy * act'(y) with y the activation OUTPUT (same convention as dafk_act_bwd), all bf16
__global__ void __launch_bounds__(DB_TPB) act_bwd_bf16io_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ y,
                                                                __nv_bfloat16* __restrict__ dx, int64_t n8, int act,
                                                                float alpha) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float d[8], v[8];
    db_ld8(dy + 8 * i, d);
    db_ld8(y + 8 * i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (act == DAFK_ACT_RELU) d[k] = v[k] > 0.f ? d[k] : 0.f;
      else if (act == DAFK_ACT_LRELU) d[k] = v[k] > 0.f ? d[k] : (v[k] < 0.f ? alpha * d[k] : 0.f);
      else if (act == DAFK_ACT_TANH) d[k] = d[k] * (1.f - v[k] * v[k]);
    }
    db_st8(dx + 8 * i, d);
  }
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_film_act_add_fwd_bf16(const void* x, const float* gamma, const float* beta, const void* res, void* y, int B,
                               int64_t HW, int C, int act, float alpha, void* stream) {
  DAFK_REQUIRE(B >= 0 && HW >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_film_act_add_fwd_bf16: bad shape");
  if (B == 0 || HW == 0) return DAFK_OK;
  DAFK_REQUIRE(x && gamma && beta && y, DAFK_ERR_BAD_ARG, "dafk_film_act_add_fwd_bf16: null pointer");
  DAFK_REQUIRE(C % 8 == 0, DAFK_ERR_UNSUPPORTED, "dafk_film_act_add_fwd_bf16: C must be a multiple of 8 (got %d)", C);
  DAFK_REQUIRE(act == DAFK_ACT_NONE || act == DAFK_ACT_RELU || act == DAFK_ACT_LRELU, DAFK_ERR_BAD_ARG,
               "dafk_film_act_add_fwd_bf16: act must be NONE, RELU or LRELU");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(beta) && DAFK_ALIGNED16(y) &&
                   (res == nullptr || DAFK_ALIGNED16(res)),
               DAFK_ERR_ALIGN, "dafk_film_act_add_fwd_bf16: pointers must be 16-byte aligned");
  const int64_t n8 = (int64_t)B * HW * C / 8;
  film_fwd_bf16_kernel<<<bw_grid(n8, DB_TPB), DB_TPB, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, gamma, beta, (const __nv_bfloat16*)res, (__nv_bfloat16*)y, HW * C, C, n8, act, alpha);
  return check_launch("dafk_film_act_add_fwd_bf16");
}

int dafk_film_act_add_bwd_bf16(const void* dy, const void* x, const float* gamma, const float* beta, void* dx,
                               float* dgamma, float* dbeta, double* ws, int B, int64_t HW, int C, int act, float alpha,
                               void* stream) {
  DAFK_REQUIRE(B >= 0 && HW >= 0 && C > 0, DAFK_ERR_BAD_ARG, "dafk_film_act_add_bwd_bf16: bad shape");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && ws, DAFK_ERR_BAD_ARG,
               "dafk_film_act_add_bwd_bf16: null pointer");
  DAFK_REQUIRE(act == DAFK_ACT_NONE || act == DAFK_ACT_RELU || act == DAFK_ACT_LRELU, DAFK_ERR_BAD_ARG,
               "dafk_film_act_add_bwd_bf16: act must be NONE, RELU or LRELU");
  DAFK_REQUIRE(act == DAFK_ACT_NONE || (beta != nullptr && DAFK_ALIGNED16(beta)), DAFK_ERR_BAD_ARG,
               "dafk_film_act_add_bwd_bf16: a fused activation needs beta (16-byte aligned)");
  DAFK_REQUIRE(C >= 8 && C <= 8 * DB_TPB && (C & (C - 1)) == 0, DAFK_ERR_UNSUPPORTED,
               "dafk_film_act_add_bwd_bf16: C must be a power of two in [8,2048] (got %d)", C);
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(x) && DAFK_ALIGNED16(gamma) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN,
               "dafk_film_act_add_bwd_bf16: pointers must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)B * C, s);
  DAFK_REQUIRE(e == cudaSuccess, DAFK_ERR_CUDA, "dafk_film_act_add_bwd_bf16: memset failed: %s", cudaGetErrorString(e));
  const int64_t n8 = HW * C / 8;
  int chunks = (int)((n8 + DB_TPB - 1) / DB_TPB);
  const int max_chunks = (kNumSMs * 8 + B - 1) / B;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  film_bwd_bf16_kernel<<<dim3(chunks, B), DB_TPB, 2 * C * sizeof(float), s>>>(
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)dx, ws, HW * C, C, act, alpha);
  int rc = check_launch("dafk_film_act_add_bwd_bf16");
  if (rc) return rc;
  film_bwd_bf16_finish_kernel<<<(B * C + 127) / 128, 128, 0, s>>>(ws, dgamma, dbeta, B, C);
  return check_launch("dafk_film_act_add_bwd_bf16(finish)");
}

int dafk_act_bwd_bf16io(const void* dy, const void* y, void* dx, int64_t n, int act, float alpha, void* stream) {
  DAFK_REQUIRE(act >= 0 && act <= 3, DAFK_ERR_BAD_ARG, "dafk_act_bwd_bf16io: bad activation %d", act);
  DAFK_REQUIRE(n >= 0 && n % 8 == 0, DAFK_ERR_BAD_ARG, "dafk_act_bwd_bf16io: size must be a multiple of 8");
  if (n == 0) return DAFK_OK;
  DAFK_REQUIRE(dy && y && dx, DAFK_ERR_BAD_ARG, "dafk_act_bwd_bf16io: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(y) && DAFK_ALIGNED16(dx), DAFK_ERR_ALIGN,
               "dafk_act_bwd_bf16io: pointers must be 16-byte aligned");
  act_bwd_bf16io_kernel<<<bw_grid(n / 8, DB_TPB), DB_TPB, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dx, n / 8, act, alpha);
  return check_launch("dafk_act_bwd_bf16io");
}

}  // extern "C"
