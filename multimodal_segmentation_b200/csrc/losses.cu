// Loss kernels (costs.py): soft Dice + the swapped-argument weighted cross entropy, MAE / MSE,
// and the VAE reparameterisation + KL.  Each forward is a single pass over the prediction
// with warp-shuffle + block reductions into a small double workspace; each backward is a
// single elementwise pass.
#include "common.cuh"

namespace dafk {

constexpr int LT = 256;
constexpr int MAXC = 8;  // max channels of a segmentation output

// ws layout: [B][2] (I_b, U_b) | n[Cp] | S[Cp]
__global__ void __launch_bounds__(LT) segloss_fwd_kernel(const float* __restrict__ pred, int Cp,
                                                         const float* __restrict__ target, int Ct, int nch, int use_bce,
                                                         double* __restrict__ ws, int B, int64_t HW) {
  __shared__ float red[2 + 2 * MAXC][LT / 32];
  const int b = blockIdx.y;
  const float* pb = pred + (int64_t)b * HW * Cp;
  const float* tb = target + (int64_t)b * HW * Ct;
  float I = 0.f, U = 0.f, nacc[MAXC], sacc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { nacc[c] = 0.f; sacc[c] = 0.f; }
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < Cp) {
        float p = pb[i * Cp + c];
        float t = (c < Ct) ? tb[i * Ct + c] : 0.f;
        if (c < nch) { I += t * p; U += t + p; }
        if (use_bce) { nacc[c] += p; sacc[c] += p * logf(t + 1e-12f); }
      }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  I = warp_sum(I); U = warp_sum(U);
  if (lane == 0) { red[0][w] = I; red[1][w] = U; }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < Cp && use_bce) {
      float a = warp_sum(nacc[c]), s = warp_sum(sacc[c]);
      if (lane == 0) { red[2 + c][w] = a; red[2 + MAXC + c][w] = s; }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 + 2 * MAXC) {
    int q = threadIdx.x;
    bool used = q < 2 || (use_bce && ((q - 2) % MAXC) < Cp);
    if (used) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < LT / 32; ++k) s += red[q][k];
      double* dst;
      if (q < 2) dst = ws + 2 * b + q;
      else if (q < 2 + MAXC) dst = ws + 2 * B + (q - 2);
      else dst = ws + 2 * B + Cp + (q - 2 - MAXC);
      atomicAdd(dst, (double)s);
    }
  }
}

__global__ void segloss_finish_kernel(const double* __restrict__ ws, float weight, float* __restrict__ loss, int B,
                                      int Cp, int use_bce, float lambda_bce, int64_t HW) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double dice_loss = 0.0;
  for (int b = 0; b < B; ++b) dice_loss += 1.0 - (2.0 * ws[2 * b] + 1e-12) / (ws[2 * b + 1] + 1e-12);
  dice_loss /= (double)B;
  double bce = 0.0;
  if (use_bce) {
    const double* n = ws + 2 * B;
    const double* S = n + Cp;
    double ntot = 0.0;
    for (int c = 0; c < Cp; ++c) ntot += n[c];
    for (int c = 0; c < Cp; ++c) bce += S[c] * (ntot / (n[c] + 1e-12));
    bce = -bce / ((double)B * (double)HW);
  }
  loss[0] += weight * (float)(dice_loss + (double)lambda_bce * bce);
}

__global__ void __launch_bounds__(LT) segloss_bwd_kernel(const float* __restrict__ pred, int Cp,
                                                         const float* __restrict__ target, int Ct, int nch, int use_bce,
                                                         float lambda_bce, const double* __restrict__ ws, float weight,
                                                         float* __restrict__ dpred, int B, int64_t HW) {
  __shared__ float wc[MAXC], kc[MAXC];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    const double* n = ws + 2 * B;
    const double* S = n + Cp;
    double ntot = 0.0, sumS = 0.0;
    if (use_bce) {
      for (int c = 0; c < Cp; ++c) { ntot += n[c]; sumS += S[c] / (n[c] + 1e-12); }
    }
    for (int c = 0; c < MAXC; ++c) {
      if (c < Cp && use_bce) {
        wc[c] = (float)(ntot / (n[c] + 1e-12));
        kc[c] = (float)(sumS - S[c] * ntot / ((n[c] + 1e-12) * (n[c] + 1e-12)));
      } else { wc[c] = 0.f; kc[c] = 0.f; }
    }
  }
  __syncthreads();
  const double Ib = ws[2 * b], Ub = ws[2 * b + 1];
  const float den = (float)(1.0 / ((Ub + 1e-12) * (Ub + 1e-12)));
  const float uE = (float)(Ub + 1e-12), iE = (float)(2.0 * Ib + 1e-12);
  const float sd = -weight / (float)B;
  const float sb = -weight * lambda_bce / ((float)B * (float)HW);
  const float* pb = pred + (int64_t)b * HW * Cp;
  const float* tb = target + (int64_t)b * HW * Ct;
  float* gb = dpred + (int64_t)b * HW * Cp;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < Cp) {
        float t = (c < Ct) ? tb[i * Ct + c] : 0.f;
        float g = 0.f;
        if (c < nch) g += sd * (2.f * t * uE - iE) * den;
        if (use_bce) g += sb * (wc[c] * logf(t + 1e-12f) + kc[c]);
        gb[i * Cp + c] = g;
      }
    }
  }
  (void)pb;
}

// ---------------------------------------------------------------- mae / mse
__global__ void __launch_bounds__(LT) l1l2_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                  float cval, int kind, float scale /* weight/n */,
                                                  float* __restrict__ loss, float* __restrict__ dpred, int64_t n) {
  __shared__ float red[LT / 32];
  float acc = 0.f;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = pred[i] - (target ? target[i] : cval);
    if (kind == 0) {
      acc += fabsf(d);
      if (dpred) dpred[i] = scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    } else {
      acc += d * d;
      if (dpred) dpred[i] = scale * 2.f * d;
    }
  }
  float t = block_sum<LT>(acc, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, scale * t);
}

// ---------------------------------------------------------------- VAE
__global__ void vae_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                               float* __restrict__ z, float* __restrict__ kl, float weight, float* __restrict__ loss,
                               int B, int Z) {
  __shared__ float red[LT / 32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float k = 0.f;
    for (int j = 0; j < Z; ++j) {
      float m = mu[b * Z + j], l = lv[b * Z + j];
      if (z) z[b * Z + j] = m + expf(0.5f * l) * eps[b * Z + j];
      k += 1.f + l - m * m - expf(l);
    }
    k *= -0.5f;
    if (kl) kl[b] = k;
    acc += k;
  }
  float t = block_sum<LT>(acc, red);
  if (threadIdx.x == 0 && loss) loss[0] += weight * t / (float)B;
}

__global__ void vae_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                               const float* __restrict__ dz, float weight, float* __restrict__ dmu,
                               float* __restrict__ dlv, int B, int Z) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Z) return;
  float m = mu[i], l = lv[i];
  float g = dz ? dz[i] : 0.f;
  float wk = weight / (float)B;
  dmu[i] = g + wk * m;
  dlv[i] = g * 0.5f * expf(0.5f * l) * eps[i] + wk * 0.5f * (expf(l) - 1.f);
}

static inline int chunks_for(int64_t HW, int B) {
  int64_t chunks = (HW + LT - 1) / LT;
  int64_t cap = ((int64_t)kNumSMs * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int64_t dafk_segloss_ws_doubles(int B, int C) { return 2 * (int64_t)B + 2 * (int64_t)C; }

int dafk_segloss_fwd(const float* pred, int Cp, const float* target, int Ct, int nch, int use_bce, float lambda_bce,
                     double* ws, int B, int64_t HW, void* stream) {
  (void)lambda_bce;
  DAFK_REQUIRE(B > 0 && HW > 0 && Cp > 0 && Cp <= MAXC && Ct > 0 && nch > 0 && nch <= Cp && nch <= Ct, DAFK_ERR_BAD_ARG,
               "dafk_segloss_fwd: bad shape (Cp=%d Ct=%d nch=%d)", Cp, Ct, nch);
  DAFK_REQUIRE(!use_bce || Ct == Cp, DAFK_ERR_BAD_ARG,
               "dafk_segloss_fwd: the weighted cross entropy needs target and prediction with the same channels");
  DAFK_REQUIRE(pred && target && ws, DAFK_ERR_BAD_ARG, "dafk_segloss_fwd: null pointer");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * (2 * B + 2 * Cp), s);
  segloss_fwd_kernel<<<dim3(chunks_for(HW, B), B), LT, 0, s>>>(pred, Cp, target, Ct, nch, use_bce, ws, B, HW);
  return check_launch("dafk_segloss_fwd");
}

int dafk_segloss_finish(const double* ws, float weight, float* loss, int B, int Cp, int nch, int use_bce,
                        float lambda_bce, int64_t HW, void* stream) {
  (void)nch;
  DAFK_REQUIRE(ws && loss && B > 0, DAFK_ERR_BAD_ARG, "dafk_segloss_finish: bad argument");
  segloss_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(ws, weight, loss, B, Cp, use_bce, lambda_bce, HW);
  return check_launch("dafk_segloss_finish");
}

int dafk_segloss_bwd(const float* pred, int Cp, const float* target, int Ct, int nch, int use_bce, float lambda_bce,
                     const double* ws, float weight, float* dpred, int B, int64_t HW, void* stream) {
  DAFK_REQUIRE(B > 0 && HW > 0 && Cp > 0 && Cp <= MAXC && Ct > 0 && nch > 0 && nch <= Cp && nch <= Ct, DAFK_ERR_BAD_ARG,
               "dafk_segloss_bwd: bad shape");
  DAFK_REQUIRE(!use_bce || Ct == Cp, DAFK_ERR_BAD_ARG, "dafk_segloss_bwd: channel mismatch for the cross entropy");
  DAFK_REQUIRE(pred && target && ws && dpred, DAFK_ERR_BAD_ARG, "dafk_segloss_bwd: null pointer");
  segloss_bwd_kernel<<<dim3(chunks_for(HW, B), B), LT, 0, as_stream(stream)>>>(pred, Cp, target, Ct, nch, use_bce,
                                                                              lambda_bce, ws, weight, dpred, B, HW);
  return check_launch("dafk_segloss_bwd");
}

int dafk_l1l2_loss(const float* pred, const float* target, float cval, int kind, float weight, float* loss,
                   float* dpred, int64_t n, void* stream) {
  DAFK_REQUIRE(n > 0 && pred && (kind == 0 || kind == 1), DAFK_ERR_BAD_ARG, "dafk_l1l2_loss: bad argument");
  l1l2_kernel<<<bw_grid(n, LT, 4), LT, 0, as_stream(stream)>>>(pred, target, cval, kind, weight / (float)n, loss, dpred, n);
  return check_launch("dafk_l1l2_loss");
}

int dafk_vae_fwd(const float* mu, const float* logvar, const float* eps, float* z, float* kl, float weight,
                 float* loss, int B, int Z, void* stream) {
  DAFK_REQUIRE(B > 0 && Z > 0 && mu && logvar, DAFK_ERR_BAD_ARG, "dafk_vae_fwd: bad argument");
  DAFK_REQUIRE(!z || eps, DAFK_ERR_BAD_ARG, "dafk_vae_fwd: eps required when z is requested");
  vae_fwd_kernel<<<1, LT, 0, as_stream(stream)>>>(mu, logvar, eps, z, kl, weight, loss, B, Z);
  return check_launch("dafk_vae_fwd");
}

int dafk_vae_bwd(const float* mu, const float* logvar, const float* eps, const float* dz, float weight, float* dmu,
                 float* dlogvar, int B, int Z, void* stream) {
  DAFK_REQUIRE(B > 0 && Z > 0 && mu && logvar && eps && dmu && dlogvar, DAFK_ERR_BAD_ARG, "dafk_vae_bwd: bad argument");
  vae_bwd_kernel<<<(B * Z + 127) / 128, 128, 0, as_stream(stream)>>>(mu, logvar, eps, dz, weight, dmu, dlogvar, B, Z);
  return check_launch("dafk_vae_bwd");
}

}  // extern "C"
