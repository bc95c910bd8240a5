// Kernels of the automated-pairing trainers (reference models/dafnet.py:224-334): every candidate pair j contributes a
// PER-SAMPLE loss L_j[b] that is weighted by the Balancer's softmax output w[b,j] (model_components/balancer.py) and
// summed,  out[b] = sum_j w[b,j] * L_j[b];  Keras then takes the mean over the batch (loss `costs.ypred`).
//   * per-sample segmentation loss `make_combined_dice_bce_perbatch` (costs.py:138-143): soft Dice over the first
//     num_classes channels + 0.01 * `weighted_cross_entropy_perbatch` (costs.py:88-108), which is called with swapped
//     arguments: the class weights n_tot / (n_c + 1e-12) come from the PREDICTION summed over the whole batch, the
//     log is taken of softmax(mask) -- and the gradient therefore also flows through the class weights;
//   * per-sample reconstruction loss `mae_single_input` (costs.py:24-26): mean |x - y| over (H, W);
//   * the Balancer's Dice overlap (balancer.py:33-38) with its backward;
//   * the [B,P] combination: loss value, d loss / d w, and the per-sample coefficients coef[j,b] = weight/B * w[b,j]
//     that scale the backward passes of the per-sample losses.
// All reductions: warp shuffles -> block -> one double atomic per quantity per CTA; backward passes are elementwise.
#include "common.cuh"

namespace dafk {

constexpr int PT = 256;
constexpr int PMAXC = 8;

static inline int pb_chunks(int64_t items, int B) {
  int64_t chunks = (items + PT - 1) / PT;
  int64_t cap = ((int64_t)kNumSMs * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

// log(softmax(t)_c + 1e-12) for the C mask channels of one pixel
__device__ __forceinline__ void log_softmax_eps(const float* t, int C, float (&ls)[PMAXC]) {
  float m = t[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, t[c]);
  float e[PMAXC], s = 0.f;
#pragma unroll
  for (int c = 0; c < PMAXC; ++c) { e[c] = c < C ? expf(t[c] - m) : 0.f; s += e[c]; }
#pragma unroll
  for (int c = 0; c < PMAXC; ++c) ls[c] = c < C ? logf(e[c] / s + 1e-12f) : 0.f;
}

// ws: [B][2] (I_b, U_b) | n[C] | S[B][C]
__global__ void __launch_bounds__(PT) segloss_pb_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                            int C, int nch, double* __restrict__ ws, int B, int64_t HW) {
  __shared__ float red[2 + 2 * PMAXC][PT / 32];
  const int b = blockIdx.y;
  const float* pb = pred + (int64_t)b * HW * C;
  const float* tb = target + (int64_t)b * HW * C;
  float I = 0.f, U = 0.f, nacc[PMAXC], sacc[PMAXC];
#pragma unroll
  for (int c = 0; c < PMAXC; ++c) { nacc[c] = 0.f; sacc[c] = 0.f; }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
    float t[PMAXC], ls[PMAXC];
#pragma unroll
    for (int c = 0; c < PMAXC; ++c) t[c] = c < C ? tb[i * C + c] : 0.f;
    log_softmax_eps(t, C, ls);
#pragma unroll
    for (int c = 0; c < PMAXC; ++c) {
      if (c < C) {
        const float p = pb[i * C + c];
        if (c < nch) { I += t[c] * p; U += t[c] + p; }
        nacc[c] += p;
        sacc[c] += p * ls[c];
      }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  I = warp_sum(I); U = warp_sum(U);
  if (lane == 0) { red[0][w] = I; red[1][w] = U; }
#pragma unroll
  for (int c = 0; c < PMAXC; ++c) {
    if (c < C) {
      const float a = warp_sum(nacc[c]), s = warp_sum(sacc[c]);
      if (lane == 0) { red[2 + c][w] = a; red[2 + PMAXC + c][w] = s; }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 + 2 * PMAXC) {
    const int q = threadIdx.x;
    const bool used = q < 2 || ((q - 2) % PMAXC) < C;
    if (used) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < PT / 32; ++k) s += red[q][k];
      double* dst;
      if (q < 2) dst = ws + 2 * b + q;
      else if (q < 2 + PMAXC) dst = ws + 2 * B + (q - 2);
      else dst = ws + 2 * B + C + (int64_t)b * C + (q - 2 - PMAXC);
      atomicAdd(dst, (double)s);
    }
  }
}

__global__ void segloss_pb_finish_kernel(const double* __restrict__ ws, float* __restrict__ L, int B, int C,
                                         float lambda_bce, int64_t HW) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* n = ws + 2 * B;
  const double* S = n + C + (int64_t)b * C;
  double ntot = 0.0;
  for (int c = 0; c < C; ++c) ntot += n[c];
  double ce = 0.0;
  for (int c = 0; c < C; ++c) ce += S[c] * (ntot / (n[c] + 1e-12));
  ce = -ce / (double)HW;
  const double dice = (2.0 * ws[2 * b] + 1e-12) / (ws[2 * b + 1] + 1e-12);
  L[b] = (float)((1.0 - dice) + (double)lambda_bce * ce);
}

// d/dpred of sum_b coef[b] * L[b]
__global__ void __launch_bounds__(PT) segloss_pb_bwd_kernel(const float* __restrict__ target, int C, int nch,
                                                            float lambda_bce, const double* __restrict__ ws,
                                                            const float* __restrict__ coef, float* __restrict__ dpred,
                                                            int B, int64_t HW) {
  __shared__ float wc[PMAXC], kc[PMAXC];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    const double* n = ws + 2 * B;
    const double* S = n + C;
    double ntot = 0.0, T[PMAXC], sumT = 0.0;
    for (int c = 0; c < C; ++c) { ntot += n[c]; T[c] = 0.0; }
    for (int bb = 0; bb < B; ++bb)
      for (int c = 0; c < C; ++c) T[c] += (double)coef[bb] * S[(int64_t)bb * C + c];
    for (int c = 0; c < C; ++c) sumT += T[c] / (n[c] + 1e-12);
    for (int c = 0; c < PMAXC; ++c) {
      if (c < C) {
        wc[c] = (float)(ntot / (n[c] + 1e-12));
        kc[c] = (float)(sumT - T[c] * ntot / ((n[c] + 1e-12) * (n[c] + 1e-12)));
      } else { wc[c] = 0.f; kc[c] = 0.f; }
    }
  }
  __syncthreads();
  const double Ib = ws[2 * b], Ub = ws[2 * b + 1];
  const float den = (float)(1.0 / ((Ub + 1e-12) * (Ub + 1e-12)));
  const float uE = (float)(Ub + 1e-12), iE = (float)(2.0 * Ib + 1e-12);
  const float cb = coef[b];
  const float sb = -lambda_bce / (float)HW;
  const float* tb = target + (int64_t)b * HW * C;
  float* gb = dpred + (int64_t)b * HW * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
    float t[PMAXC], ls[PMAXC];
#pragma unroll
    for (int c = 0; c < PMAXC; ++c) t[c] = c < C ? tb[i * C + c] : 0.f;
    log_softmax_eps(t, C, ls);
#pragma unroll
    for (int c = 0; c < PMAXC; ++c) {
      if (c < C) {
        float g = sb * (cb * wc[c] * ls[c] + kc[c]);
        if (c < nch) g -= cb * (2.f * t[c] * uE - iE) * den;
        gb[i * C + c] = g;
      }
    }
  }
}

// ---------------------------------------------------------------- per-sample mean absolute error
__global__ void __launch_bounds__(PT) mae_pb_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        double* __restrict__ ws, int64_t n) {
  __shared__ float red[PT / 32];
  const int b = blockIdx.y;
  const float* p = pred + (int64_t)b * n;
  const float* t = target + (int64_t)b * n;
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += fabsf(p[i] - t[i]);
  const float s = block_sum<PT>(acc, red);
  if (threadIdx.x == 0) atomicAdd(ws + b, (double)s);
}
__global__ void mae_pb_finish_kernel(const double* __restrict__ ws, float* __restrict__ L, int B, int64_t n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) L[b] = (float)(ws[b] / (double)n);
}
__global__ void __launch_bounds__(PT) mae_pb_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const float* __restrict__ coef, float* __restrict__ dpred,
                                                        int64_t n) {
  const int b = blockIdx.y;
  const float s = coef[b] / (float)n;
  const float* p = pred + (int64_t)b * n;
  const float* t = target + (int64_t)b * n;
  float* g = dpred + (int64_t)b * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = p[i] - t[i];
    g[i] = s * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
  }
}

// ---------------------------------------------------------------- Balancer overlap backward
// dice_b = (2 I + e) / (Sa + Sb + e);  ws = [B][3] (I, Sa, Sb) left by dafk_pair_dice
__global__ void __launch_bounds__(PT) pair_dice_bwd_kernel(const float* __restrict__ a, const float* __restrict__ bb,
                                                           const double* __restrict__ ws, const float* __restrict__ g,
                                                           float* __restrict__ da, float* __restrict__ db, int64_t n) {
  const int b = blockIdx.y;
  const double I = ws[3 * b], U = ws[3 * b + 1] + ws[3 * b + 2] + 1e-12;
  const float k1 = (float)((double)g[b] * 2.0 / U);
  const float k0 = (float)((double)g[b] * (2.0 * I + 1e-12) / (U * U));
  const int64_t off = (int64_t)b * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (da) da[off + i] = k1 * bb[off + i] - k0;
    if (db) db[off + i] = k1 * a[off + i] - k0;
  }
}

// ---------------------------------------------------------------- [B,P] combination
__global__ void pair_combine_kernel(const float* __restrict__ w /*[B,P] or null*/, const float* __restrict__ L /*[P,B]*/,
                                    float weight, float* __restrict__ loss, float* __restrict__ dw /*[B,P] or null*/,
                                    float* __restrict__ coef /*[P,B]*/, int B, int P) {
  __shared__ float red[PT / 32];
  float acc = 0.f;
  const float s = weight / (float)B;
  for (int e = threadIdx.x; e < B * P; e += blockDim.x) {
    const int b = e / P, j = e - b * P;
    const float wv = w ? w[e] : 1.f;
    const float l = L[(int64_t)j * B + b];
    acc += wv * l;
    if (dw) dw[e] = s * l;
    coef[(int64_t)j * B + b] = s * wv;
  }
  const float t = block_sum<PT>(acc, red);
  if (threadIdx.x == 0 && loss) loss[0] += s * t;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int64_t dafk_segloss_pb_ws_doubles(int B, int C) { return 2 * (int64_t)B + (int64_t)C + (int64_t)B * C; }

int dafk_segloss_pb_fwd(const float* pred, const float* target, int C, int nch, float lambda_bce, double* ws, float* L,
                        int B, int64_t HW, void* stream) {
  DAFK_REQUIRE(B > 0 && HW > 0 && C > 0 && C <= PMAXC && nch > 0 && nch <= C, DAFK_ERR_BAD_ARG,
               "dafk_segloss_pb_fwd: bad shape (C=%d nch=%d)", C, nch);
  DAFK_REQUIRE(pred && target && ws && L, DAFK_ERR_BAD_ARG, "dafk_segloss_pb_fwd: null pointer");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)dafk_segloss_pb_ws_doubles(B, C), s);
  segloss_pb_fwd_kernel<<<dim3(pb_chunks(HW, B), B), PT, 0, s>>>(pred, target, C, nch, ws, B, HW);
  int rc = check_launch("dafk_segloss_pb_fwd");
  if (rc) return rc;
  segloss_pb_finish_kernel<<<(B + 127) / 128, 128, 0, s>>>(ws, L, B, C, lambda_bce, HW);
  return check_launch("dafk_segloss_pb_fwd(finish)");
}

int dafk_segloss_pb_bwd(const float* target, int C, int nch, float lambda_bce, const double* ws, const float* coef,
                        float* dpred, int B, int64_t HW, void* stream) {
  DAFK_REQUIRE(B > 0 && HW > 0 && C > 0 && C <= PMAXC && nch > 0 && nch <= C, DAFK_ERR_BAD_ARG, "dafk_segloss_pb_bwd: bad shape");
  DAFK_REQUIRE(target && ws && coef && dpred, DAFK_ERR_BAD_ARG, "dafk_segloss_pb_bwd: null pointer");
  segloss_pb_bwd_kernel<<<dim3(pb_chunks(HW, B), B), PT, 0, as_stream(stream)>>>(target, C, nch, lambda_bce, ws, coef, dpred, B, HW);
  return check_launch("dafk_segloss_pb_bwd");
}

int dafk_mae_pb_fwd(const float* pred, const float* target, double* ws, float* L, int B, int64_t n, void* stream) {
  DAFK_REQUIRE(B > 0 && n > 0 && pred && target && ws && L, DAFK_ERR_BAD_ARG, "dafk_mae_pb_fwd: bad argument");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)B, s);
  mae_pb_fwd_kernel<<<dim3(pb_chunks(n, B), B), PT, 0, s>>>(pred, target, ws, n);
  int rc = check_launch("dafk_mae_pb_fwd");
  if (rc) return rc;
  mae_pb_finish_kernel<<<(B + 127) / 128, 128, 0, s>>>(ws, L, B, n);
  return check_launch("dafk_mae_pb_fwd(finish)");
}

int dafk_mae_pb_bwd(const float* pred, const float* target, const float* coef, float* dpred, int B, int64_t n,
                    void* stream) {
  DAFK_REQUIRE(B > 0 && n > 0 && pred && target && coef && dpred, DAFK_ERR_BAD_ARG, "dafk_mae_pb_bwd: bad argument");
  mae_pb_bwd_kernel<<<dim3(pb_chunks(n, B), B), PT, 0, as_stream(stream)>>>(pred, target, coef, dpred, n);
  return check_launch("dafk_mae_pb_bwd");
}

int dafk_pair_dice_bwd(const float* a, const float* b, const double* ws, const float* g, float* da, float* db, int B,
                       int64_t HWC, void* stream) {
  DAFK_REQUIRE(B > 0 && HWC > 0 && a && b && ws && g && (da || db), DAFK_ERR_BAD_ARG, "dafk_pair_dice_bwd: bad argument");
  pair_dice_bwd_kernel<<<dim3(pb_chunks(HWC, B), B), PT, 0, as_stream(stream)>>>(a, b, ws, g, da, db, HWC);
  return check_launch("dafk_pair_dice_bwd");
}

int dafk_pair_combine(const float* w, const float* L, float weight, float* loss, float* dw, float* coef, int B, int P,
                      void* stream) {
  DAFK_REQUIRE(B > 0 && P > 0 && L && coef, DAFK_ERR_BAD_ARG, "dafk_pair_combine: bad argument");
  pair_combine_kernel<<<1, PT, 0, as_stream(stream)>>>(w, L, weight, loss, dw, coef, B, P);
  return check_launch("dafk_pair_combine");
}

}  // extern "C"
