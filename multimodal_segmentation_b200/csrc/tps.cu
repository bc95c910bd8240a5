// Thin-plate-spline spatial transformer:
//   * general batched polyharmonic solve (one warp per (n+3)x(n+3) system, LU with partial
//     pivoting in shared memory) + evaluation  -- layers/interpolate_spline.py
//   * fast path for ThinPlateSpline2D(inverse=False): constant LHS => w = Winv.theta,
//     v = v_id + Vinv.theta; fused coordinate evaluation + bilinear gather forward, and a
//     backward that scatters into the volume with vector atomics and reduces the
//     control-point gradient as a small smem GEMM per CTA  -- layers/stn_spline.py:55-67 and
//     tf.contrib.resampler.
#include "common.cuh"
#include <math.h>
#include <vector>

namespace dafk {

constexpr float TPS_EPS = 1e-10f;  // interpolate_spline.py:26

__device__ __forceinline__ float phi_dev(float r, int order) {
  // interpolate_spline.py:182-209
  if (order == 1) return sqrtf(fmaxf(r, TPS_EPS));
  if (order == 2) return 0.5f * r * logf(fmaxf(r, TPS_EPS));
  if (order == 4) return 0.5f * r * r * logf(fmaxf(r, TPS_EPS));
  float rc = fmaxf(r, TPS_EPS);
  if ((order & 1) == 0) return 0.5f * powf(rc, 0.5f * order) * logf(rc);
  return powf(rc, 0.5f * order);
}

// ------------------------------------------------------------------ general solve
constexpr int SYS_MAX = 32;
constexpr int RHS_MAX = 8;

// one warp per batch element; lane = row
__global__ void __launch_bounds__(32) tps_solve_kernel(const float* __restrict__ pts, const float* __restrict__ vals,
                                                       float* __restrict__ w_out, float* __restrict__ v_out, int n,
                                                       int k, int order, float reg) {
  __shared__ float A[SYS_MAX][SYS_MAX + RHS_MAX + 1];
  __shared__ float c[SYS_MAX][2];
  __shared__ float xs[SYS_MAX][RHS_MAX];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int d = 2, sz = n + d + 1;
  const float* cb = pts + (int64_t)b * n * d;
  const float* fb = vals + (int64_t)b * n * k;
  if (lane < n) { c[lane][0] = cb[lane * 2]; c[lane][1] = cb[lane * 2 + 1]; }
  __syncwarp();
  // build [A B; B^T 0 | f; 0]   (interpolate_spline.py:112-137)
  if (lane < sz) {
    for (int j = 0; j < sz + k; ++j) {
      float v = 0.f;
      if (lane < n) {
        if (j < n) {
          // _pairwise_squared_distance_matrix: |xi|^2 - 2 xi.xj + |xj|^2
          float xx = c[lane][0] * c[lane][0] + c[lane][1] * c[lane][1];
          float yy = c[j][0] * c[j][0] + c[j][1] * c[j][1];
          float xy = c[lane][0] * c[j][0] + c[lane][1] * c[j][1];
          float r = xx - 2.f * xy + yy;
          v = phi_dev(r, order);
          if (j == lane && reg > 0.f) v += reg;
        } else if (j < n + d) {
          v = c[lane][j - n];
        } else if (j == n + d) {
          v = 1.f;
        } else {
          v = fb[lane * k + (j - sz)];
        }
      } else {
        int r = lane - n;  // rows of B^T
        if (j < n) v = (r < d) ? c[j][r] : 1.f;
      }
      A[lane][j] = v;
    }
  }
  __syncwarp();
  // LU with partial pivoting (tf.matrix_solve)
  for (int col = 0; col < sz; ++col) {
    float mag = (lane >= col && lane < sz) ? fabsf(A[lane][col]) : -1.f;
    int arg = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, mag, o);
      int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mag || (om == mag && oa < arg)) { mag = om; arg = oa; }
    }
    if (arg != col) {
      for (int j = lane; j < sz + k; j += 32) {
        float t = A[col][j]; A[col][j] = A[arg][j]; A[arg][j] = t;
      }
    }
    __syncwarp();
    float piv = A[col][col];
    if (lane > col && lane < sz) {
      float f = A[lane][col] / piv;
      for (int j = col + 1; j < sz + k; ++j) A[lane][j] -= f * A[col][j];
      A[lane][col] = 0.f;
    }
    __syncwarp();
  }
  // back substitution, lane = rhs column
  if (lane < k) {
    for (int r = sz - 1; r >= 0; --r) {
      float s = A[r][sz + lane];
      for (int j = r + 1; j < sz; ++j) s -= A[r][j] * xs[j][lane];
      xs[r][lane] = s / A[r][r];
    }
  }
  __syncwarp();
  for (int e = lane; e < n * k; e += 32) w_out[(int64_t)b * n * k + e] = xs[e / k][e % k];
  for (int e = lane; e < (d + 1) * k; e += 32) v_out[(int64_t)b * (d + 1) * k + e] = xs[n + e / k][e % k];
}

// out[b,m,:] = phi(|q-c|^2) w + [q,1] v       (interpolate_spline.py:150-179)
__global__ void __launch_bounds__(256) tps_apply_kernel(const float* __restrict__ query, const float* __restrict__ pts,
                                                        const float* __restrict__ w, const float* __restrict__ v,
                                                        float* __restrict__ out, int64_t m, int n, int k, int order,
                                                        int query_batched) {
  extern __shared__ float sm[];  // c[n][2], w[n][k], v[3][k]
  float* cs = sm;
  float* ws = cs + 2 * n;
  float* vs = ws + n * k;
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < 2 * n; e += blockDim.x) cs[e] = pts[(int64_t)b * n * 2 + e];
  for (int e = threadIdx.x; e < n * k; e += blockDim.x) ws[e] = w[(int64_t)b * n * k + e];
  for (int e = threadIdx.x; e < 3 * k; e += blockDim.x) vs[e] = v[(int64_t)b * 3 * k + e];
  __syncthreads();
  const float* qb = query + (query_batched ? (int64_t)b * m * 2 : 0);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    float q0 = qb[2 * i], q1 = qb[2 * i + 1];
    float qq = q0 * q0 + q1 * q1;
    float acc[RHS_MAX];
#pragma unroll
    for (int j = 0; j < RHS_MAX; ++j) acc[j] = 0.f;
    for (int t = 0; t < n; ++t) {
      float c0 = cs[2 * t], c1 = cs[2 * t + 1];
      float r = qq - 2.f * (q0 * c0 + q1 * c1) + (c0 * c0 + c1 * c1);
      float p = phi_dev(r, order);
#pragma unroll
      for (int j = 0; j < RHS_MAX; ++j)
        if (j < k) acc[j] = fmaf(p, ws[t * k + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < RHS_MAX; ++j)
      if (j < k) out[((int64_t)b * m + i) * k + j] = acc[j] + (q0 * vs[j] + q1 * vs[k + j] + vs[2 * k + j]);
  }
}

// ------------------------------------------------------------------ fused fast path
constexpr int TPS_MAXN = 32;   // control points
constexpr int TPS_BS = 2;      // samples per CTA (grid.y splits the batch): 3136 CTAs at B=32, 224^2
constexpr int TPS_T = 256;

struct Bilin {
  bool valid;
  int fx, fy;
  float dx, dy;
};
__device__ __forceinline__ Bilin bilin_setup(float x, float y, int H, int W) {
  Bilin r;
  r.valid = (x > -1.f) && (y > -1.f) && (x < (float)W) && (y < (float)H);
  float flx = floorf(x), fly = floorf(y);
  r.fx = (int)flx; r.fy = (int)fly;
  r.dx = (flx + 1.f) - x;
  r.dy = (fly + 1.f) - y;
  return r;
}

// per-sample spline coefficients in smem: coef[s][j][2], j<n: w ; j=n..n+2: v   ((row,col) components)
// `consts` points to the CTA's shared-memory copy of the constants (tps_stage_consts): the 25-term dot products
// below would otherwise be chains of dependent L2 loads in the prologue of every CTA
__device__ __forceinline__ void tps_stage_consts(const float* __restrict__ consts, int n, float* cst) {
  for (int e = threadIdx.x; e < n * (n + 5); e += blockDim.x) cst[e] = consts[e];
  __syncthreads();
}
__device__ __forceinline__ void tps_coefs(const float* __restrict__ theta, const float* consts, int n,
                                          int b0, int nb, float* coef /*[TPS_BS][TPS_MAXN+3][2]*/) {
  const float* Winv = consts + 2 * n;
  const float* Vinv = Winv + n * n;
  const int rows = n + 3;
  for (int e = threadIdx.x; e < nb * rows * 2; e += blockDim.x) {
    int comp = e & 1;
    int j = (e >> 1) % rows;
    int s = (e >> 1) / rows;
    const float* th = theta + (int64_t)(b0 + s) * n * 2;
    const float* row = (j < n) ? (Winv + j * n) : (Vinv + (j - n) * n);
    float acc = 0.f;
    for (int t = 0; t < n; ++t) acc = fmaf(row[t], th[2 * t + comp], acc);
    // identity affine part: v = [[1,0],[0,1],[0,0]]
    if (j == n + comp) acc += 1.f;
    coef[(s * (TPS_MAXN + 3) + j) * 2 + comp] = acc;
  }
}

template <int C>
__global__ void __launch_bounds__(TPS_T) tps_warp_fwd_kernel(const float* __restrict__ vol, const float* __restrict__ theta,
                                                             const float* __restrict__ consts, float* __restrict__ out,
                                                             float* __restrict__ locs, int B, int H, int W, int n) {
  __shared__ float cst[TPS_MAXN * (TPS_MAXN + 5)];
  __shared__ float coef[TPS_BS * (TPS_MAXN + 3) * 2];
  const float* cs = cst;
  const int b0 = blockIdx.y * TPS_BS;
  const int nb = min(TPS_BS, B - b0);
  tps_stage_consts(consts, n, cst);
  tps_coefs(theta, cst, n, b0, nb, coef);
  __syncthreads();
  const int HW = H * W;
  const int m = blockIdx.x * TPS_T + threadIdx.x;
  if (m >= HW) return;
  const int row = m / W, col = m - row * W;
  // nDgrid(normalise=True): float32(row/(H-1)), float32(col/(W-1))   (stn_spline.py:70-91)
  const float q0 = (float)((double)row / (double)(H - 1));
  const float q1 = (float)((double)col / (double)(W - 1));
  const float qq = q0 * q0 + q1 * q1;
  float ph[TPS_MAXN];
#pragma unroll
  for (int t = 0; t < TPS_MAXN; ++t) {
    if (t < n) {
      float c0 = cs[2 * t], c1 = cs[2 * t + 1];
      float r = qq - 2.f * (q0 * c0 + q1 * c1) + (c0 * c0 + c1 * c1);
      ph[t] = 0.5f * r * logf(fmaxf(r, TPS_EPS));
    } else {
      ph[t] = 0.f;
    }
  }
  for (int s = 0; s < nb; ++s) {
    const float* cf = coef + s * (TPS_MAXN + 3) * 2;
    float lr = 0.f, lc = 0.f;
#pragma unroll
    for (int t = 0; t < TPS_MAXN; ++t) {
      if (t < n) { lr = fmaf(ph[t], cf[2 * t], lr); lc = fmaf(ph[t], cf[2 * t + 1], lc); }
    }
    lr += q0 * cf[2 * n] + q1 * cf[2 * (n + 1)] + cf[2 * (n + 2)];
    lc += q0 * cf[2 * n + 1] + q1 * cf[2 * (n + 1) + 1] + cf[2 * (n + 2) + 1];
    // reverse (row,col)->(x,y) and scale to pixels   (stn_spline.py:61-64)
    const float x = lc * (float)(W - 1);
    const float y = lr * (float)(H - 1);
    const int b = b0 + s;
    if (locs) { locs[((int64_t)b * HW + m) * 2] = x; locs[((int64_t)b * HW + m) * 2 + 1] = y; }
    Bilin bl = bilin_setup(x, y, H, W);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (bl.valid) {
      const float* vb = vol + (int64_t)b * HW * C;
      const int cx = bl.fx + 1, cy = bl.fy + 1;
      const float w00 = bl.dx * bl.dy, w11 = (1.f - bl.dx) * (1.f - bl.dy);
      const float w01 = bl.dx * (1.f - bl.dy), w10 = (1.f - bl.dx) * bl.dy;   // (fx,cy), (cx,fy)
      const bool fxin = bl.fx >= 0 && bl.fx <= W - 1, cxin = cx >= 0 && cx <= W - 1;
      const bool fyin = bl.fy >= 0 && bl.fy <= H - 1, cyin = cy >= 0 && cy <= H - 1;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        float4 v;
        if (fxin && fyin) { v = *reinterpret_cast<const float4*>(vb + ((int64_t)bl.fy * W + bl.fx) * C + c);
          acc[c] += w00 * v.x; acc[c + 1] += w00 * v.y; acc[c + 2] += w00 * v.z; acc[c + 3] += w00 * v.w; }
        if (cxin && cyin) { v = *reinterpret_cast<const float4*>(vb + ((int64_t)cy * W + cx) * C + c);
          acc[c] += w11 * v.x; acc[c + 1] += w11 * v.y; acc[c + 2] += w11 * v.z; acc[c + 3] += w11 * v.w; }
        if (fxin && cyin) { v = *reinterpret_cast<const float4*>(vb + ((int64_t)cy * W + bl.fx) * C + c);
          acc[c] += w01 * v.x; acc[c + 1] += w01 * v.y; acc[c + 2] += w01 * v.z; acc[c + 3] += w01 * v.w; }
        if (cxin && fyin) { v = *reinterpret_cast<const float4*>(vb + ((int64_t)bl.fy * W + cx) * C + c);
          acc[c] += w10 * v.x; acc[c + 1] += w10 * v.y; acc[c + 2] += w10 * v.z; acc[c + 3] += w10 * v.w; }
      }
    }
    float* ob = out + ((int64_t)b * HW + m) * C;
#pragma unroll
    for (int c = 0; c < C; c += 4) stg_stream4(ob + c, make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
  }
}

// phi(|q - c_t|^2) for every output pixel q and control point t depends only on the geometry (H, W, control grid): it
// is tabulated once per geometry, [n][HW] so that a warp reads 128 contiguous bytes per control point, with exactly the
// arithmetic of the in-kernel evaluation above.  The 25 logf per pixel were ~2/3 of the forward kernel's instructions.
__global__ void __launch_bounds__(TPS_T) tps_phi_table_kernel(const float* __restrict__ consts, float* __restrict__ tab,
                                                              int H, int W, int n) {
  const int HW = H * W;
  const int m = blockIdx.x * TPS_T + threadIdx.x;
  if (m >= HW) return;
  const int row = m / W, col = m - row * W;
  const float q0 = (float)((double)row / (double)(H - 1));
  const float q1 = (float)((double)col / (double)(W - 1));
  const float qq = q0 * q0 + q1 * q1;
  for (int t = 0; t < n; ++t) {
    const float c0 = consts[2 * t], c1 = consts[2 * t + 1];
    const float r = qq - 2.f * (q0 * c0 + q1 * c1) + (c0 * c0 + c1 * c1);
    tab[(int64_t)t * HW + m] = 0.5f * r * logf(fmaxf(r, TPS_EPS));
  }
  // rows n, n+1: the normalised pixel coordinates themselves (the affine part of the spline reads them; computing them in
  // the warp kernels costs an integer division and two double-precision divisions per thread)
  tab[(int64_t)n * HW + m] = q0;
  tab[(int64_t)(n + 1) * HW + m] = q1;
}

constexpr int TPS_BS_TAB = 1;      // samples per thread when phi comes from the table (1: most thread-level parallelism, fewest registers)

// per-sample spline coefficients coef[b][j][2] (j < n: w, j = n..n+2: v), computed ONCE per call.  In the kernels above
// every CTA recomputes them for its samples (25-term dot products whose theta operand comes from global memory); with
// a few hundred CTAs per sample that prologue, not the warp itself, was most of the run time.
__global__ void tps_coef_kernel(const float* __restrict__ theta, const float* __restrict__ consts, float* __restrict__ coef,
                                int B, int n) {
  const int rows = n + 3;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * rows * 2) return;
  const int comp = e & 1;
  const int j = (e >> 1) % rows;
  const int b = (e >> 1) / rows;
  const float* Winv = consts + 2 * n;
  const float* Vinv = Winv + n * n;
  const float* th = theta + (int64_t)b * n * 2;
  const float* row = (j < n) ? (Winv + j * n) : (Vinv + (j - n) * n);
  float acc = 0.f;
  for (int t = 0; t < n; ++t) acc = fmaf(row[t], th[2 * t + comp], acc);
  if (j == n + comp) acc += 1.f;                    // identity affine part, as in tps_coefs
  coef[e] = acc;
}

// NCP > 0: the control-point count is a compile-time constant (25 for the 5x5 grid every configuration uses): the
// 32-iteration predicated loops of the generic kernel (NCP = 0) made it instruction-bound -- ncu: 33.5 M warp
// instructions, 66 % issue-active, 50 us at 32 x 224^2 x 8; same arithmetic in the same order, bit-identical output.
template <int C, int NCP>
__global__ void __launch_bounds__(TPS_T) tps_warp_fwd_tab_kernel(const float* __restrict__ vol,
                                                                 const float* __restrict__ coef_g,
                                                                 const float* __restrict__ tab, float* __restrict__ out,
                                                                 float* __restrict__ locs, int B, int H, int W, int n) {
  __shared__ float coef[TPS_BS_TAB * (TPS_MAXN + 3) * 2];
  const int b0 = blockIdx.y * TPS_BS_TAB;
  const int nb = min(TPS_BS_TAB, B - b0);
  const int rows2 = (n + 3) * 2;
  for (int e = threadIdx.x; e < nb * rows2; e += TPS_T) {
    const int s = e / rows2, r = e - s * rows2;
    coef[s * (TPS_MAXN + 3) * 2 + r] = coef_g[(int64_t)(b0 + s) * rows2 + r];
  }
  __syncthreads();
  const int HW = H * W;
  const int m = blockIdx.x * TPS_T + threadIdx.x;
  if (m >= HW) return;
  constexpr int NT = NCP > 0 ? NCP : TPS_MAXN;
  float ph[NT];
  const float* tp = tab + m;
#pragma unroll
  for (int t = 0; t < NT; ++t) ph[t] = (NCP > 0 || t < n) ? __ldg(tp + (size_t)t * (size_t)HW) : 0.f;
  const float q0 = __ldg(tp + (size_t)n * (size_t)HW);            // row / (H - 1), col / (W - 1): rows n, n+1 of the table
  const float q1 = __ldg(tp + (size_t)(n + 1) * (size_t)HW);
#pragma unroll 2
  for (int s = 0; s < nb; ++s) {
    const float* cf = coef + s * (TPS_MAXN + 3) * 2;
    float lr = 0.f, lc = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (NCP > 0 || t < n) { lr = fmaf(ph[t], cf[2 * t], lr); lc = fmaf(ph[t], cf[2 * t + 1], lc); }
    }
    lr += q0 * cf[2 * n] + q1 * cf[2 * (n + 1)] + cf[2 * (n + 2)];
    lc += q0 * cf[2 * n + 1] + q1 * cf[2 * (n + 1) + 1] + cf[2 * (n + 2) + 1];
    const float x = lc * (float)(W - 1);
    const float y = lr * (float)(H - 1);
    const int b = b0 + s;
    if (locs) { locs[((int64_t)b * HW + m) * 2] = x; locs[((int64_t)b * HW + m) * 2 + 1] = y; }
    Bilin bl = bilin_setup(x, y, H, W);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (bl.valid) {
      // branch-free corners: an out-of-range corner is read at a clamped (in-range) address with weight 0 -- the same
      // sums in the same order as four guarded blocks (adding 0 * v leaves a finite accumulator unchanged); 32-bit
      // offsets inside one image
      const float* vb = vol + (int64_t)b * HW * C;
      const int cx = bl.fx + 1, cy = bl.fy + 1;
      const bool fxin = bl.fx >= 0 && bl.fx <= W - 1, cxin = cx >= 0 && cx <= W - 1;
      const bool fyin = bl.fy >= 0 && bl.fy <= H - 1, cyin = cy >= 0 && cy <= H - 1;
      const float w00 = (fxin && fyin) ? bl.dx * bl.dy : 0.f, w11 = (cxin && cyin) ? (1.f - bl.dx) * (1.f - bl.dy) : 0.f;
      const float w01 = (fxin && cyin) ? bl.dx * (1.f - bl.dy) : 0.f, w10 = (cxin && fyin) ? (1.f - bl.dx) * bl.dy : 0.f;
      const int fxc = min(max(bl.fx, 0), W - 1), cxc = min(max(cx, 0), W - 1);
      const int fyc = min(max(bl.fy, 0), H - 1), cyc = min(max(cy, 0), H - 1);
      const int o00 = (fyc * W + fxc) * C, o11 = (cyc * W + cxc) * C, o01 = (cyc * W + fxc) * C, o10 = (fyc * W + cxc) * C;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 v00 = *reinterpret_cast<const float4*>(vb + o00 + c);
        const float4 v11 = *reinterpret_cast<const float4*>(vb + o11 + c);
        const float4 v01 = *reinterpret_cast<const float4*>(vb + o01 + c);
        const float4 v10 = *reinterpret_cast<const float4*>(vb + o10 + c);
        acc[c] += w00 * v00.x; acc[c + 1] += w00 * v00.y; acc[c + 2] += w00 * v00.z; acc[c + 3] += w00 * v00.w;
        acc[c] += w11 * v11.x; acc[c + 1] += w11 * v11.y; acc[c + 2] += w11 * v11.z; acc[c + 3] += w11 * v11.w;
        acc[c] += w01 * v01.x; acc[c + 1] += w01 * v01.y; acc[c + 2] += w01 * v01.z; acc[c + 3] += w01 * v01.w;
        acc[c] += w10 * v10.x; acc[c + 1] += w10 * v10.y; acc[c + 2] += w10 * v10.z; acc[c + 3] += w10 * v10.w;
      }
    }
    float* ob = out + ((int64_t)b * HW + m) * C;
#pragma unroll
    for (int c = 0; c < C; c += 4) stg_stream4(ob + c, make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
  }
}

// backward.  smem: Phi[(n+3)][257] and D[2*TPS_BS][257]; G[b][j][2] += Phi^T D
// tab != nullptr: phi and the normalised pixel coordinates come from the geometry's table (tps_phi_table_kernel: the
// same values bit for bit) instead of 25 logf + an integer and two double divisions per thread; NCP > 0: compile-time
// control-point count (see the forward kernel).
template <int C, int NCP>
__global__ void __launch_bounds__(TPS_T) tps_warp_bwd_kernel(const float* __restrict__ vol, const float* __restrict__ theta,
                                                             const float* __restrict__ consts, const float* __restrict__ dout,
                                                             float* __restrict__ dvol, double* __restrict__ G, int B, int H,
                                                             int W, int n, const float* __restrict__ tab) {
  extern __shared__ float dyn[];
  constexpr int LD = TPS_T + 1;
  constexpr int NT = NCP > 0 ? NCP : TPS_MAXN;
  float* Phi = dyn;                          // [(TPS_MAXN+3)][LD]
  float* D = Phi + (TPS_MAXN + 3) * LD;      // [2*TPS_BS][LD]
  __shared__ float cst[TPS_MAXN * (TPS_MAXN + 5)];
  __shared__ float coef[TPS_BS * (TPS_MAXN + 3) * 2];
  const float* cs = cst;
  const int b0 = blockIdx.y * TPS_BS;
  const int nb = min(TPS_BS, B - b0);
  tps_stage_consts(consts, n, cst);
  tps_coefs(theta, cst, n, b0, nb, coef);
  __syncthreads();
  const int HW = H * W;
  const int m = blockIdx.x * TPS_T + threadIdx.x;
  const bool live = m < HW;
  float q0, q1;
  float ph[NT];
  if (tab != nullptr) {
    const float* tp = tab + (live ? m : 0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      ph[t] = (live && (NCP > 0 || t < n)) ? __ldg(tp + (size_t)t * (size_t)HW) : 0.f;
      if (NCP > 0 || t < n) Phi[t * LD + threadIdx.x] = ph[t];
    }
    q0 = live ? __ldg(tp + (size_t)n * (size_t)HW) : 0.f;
    q1 = live ? __ldg(tp + (size_t)(n + 1) * (size_t)HW) : 0.f;
  } else {
    const int row = live ? m / W : 0, col = live ? m - row * W : 0;
    q0 = (float)((double)row / (double)(H - 1));
    q1 = (float)((double)col / (double)(W - 1));
    const float qq = q0 * q0 + q1 * q1;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (NCP > 0 || t < n) {
        float c0 = cs[2 * t], c1 = cs[2 * t + 1];
        float r = qq - 2.f * (q0 * c0 + q1 * c1) + (c0 * c0 + c1 * c1);
        ph[t] = live ? 0.5f * r * logf(fmaxf(r, TPS_EPS)) : 0.f;
        Phi[t * LD + threadIdx.x] = ph[t];
      } else {
        ph[t] = 0.f;
      }
    }
  }
  Phi[(n + 0) * LD + threadIdx.x] = live ? q0 : 0.f;
  Phi[(n + 1) * LD + threadIdx.x] = live ? q1 : 0.f;
  Phi[(n + 2) * LD + threadIdx.x] = live ? 1.f : 0.f;

  for (int s = 0; s < nb; ++s) {
    float dlr = 0.f, dlc = 0.f;
    if (live) {
      const float* cf = coef + s * (TPS_MAXN + 3) * 2;
      float lr = 0.f, lc = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (NCP > 0 || t < n) { lr = fmaf(ph[t], cf[2 * t], lr); lc = fmaf(ph[t], cf[2 * t + 1], lc); }
      }
      lr += q0 * cf[2 * n] + q1 * cf[2 * (n + 1)] + cf[2 * (n + 2)];
      lc += q0 * cf[2 * n + 1] + q1 * cf[2 * (n + 1) + 1] + cf[2 * (n + 2) + 1];
      const float x = lc * (float)(W - 1);
      const float y = lr * (float)(H - 1);
      const int b = b0 + s;
      Bilin bl = bilin_setup(x, y, H, W);
      if (bl.valid) {
        const float* vb = vol + (int64_t)b * HW * C;
        float* gb = dvol ? dvol + (int64_t)b * HW * C : nullptr;
        const float* go = dout + ((int64_t)b * HW + m) * C;
        const int cx = bl.fx + 1, cy = bl.fy + 1;
        const float w00 = bl.dx * bl.dy, w11 = (1.f - bl.dx) * (1.f - bl.dy);
        const float w01 = bl.dx * (1.f - bl.dy), w10 = (1.f - bl.dx) * bl.dy;
        const bool fxin = bl.fx >= 0 && bl.fx <= W - 1, cxin = cx >= 0 && cx <= W - 1;
        const bool fyin = bl.fy >= 0 && bl.fy <= H - 1, cyin = cy >= 0 && cy <= H - 1;
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          float4 g = ldg_stream4(go + c);
          float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 v00 = z, v11 = z, v01 = z, v10 = z;
          const int64_t o00 = ((int64_t)bl.fy * W + bl.fx) * C + c, o11 = ((int64_t)cy * W + cx) * C + c;
          const int64_t o01 = ((int64_t)cy * W + bl.fx) * C + c, o10 = ((int64_t)bl.fy * W + cx) * C + c;
          if (fxin && fyin) v00 = *reinterpret_cast<const float4*>(vb + o00);
          if (cxin && cyin) v11 = *reinterpret_cast<const float4*>(vb + o11);
          if (fxin && cyin) v01 = *reinterpret_cast<const float4*>(vb + o01);
          if (cxin && fyin) v10 = *reinterpret_cast<const float4*>(vb + o10);
          // d/dx = dy*(D(cx,fy)-D(fx,fy)) + (1-dy)*(D(cx,cy)-D(fx,cy))
          gx += g.x * (bl.dy * (v10.x - v00.x) + (1.f - bl.dy) * (v11.x - v01.x));
          gx += g.y * (bl.dy * (v10.y - v00.y) + (1.f - bl.dy) * (v11.y - v01.y));
          gx += g.z * (bl.dy * (v10.z - v00.z) + (1.f - bl.dy) * (v11.z - v01.z));
          gx += g.w * (bl.dy * (v10.w - v00.w) + (1.f - bl.dy) * (v11.w - v01.w));
          // d/dy = dx*(D(fx,cy)-D(fx,fy)) + (1-dx)*(D(cx,cy)-D(cx,fy))
          gy += g.x * (bl.dx * (v01.x - v00.x) + (1.f - bl.dx) * (v11.x - v10.x));
          gy += g.y * (bl.dx * (v01.y - v00.y) + (1.f - bl.dx) * (v11.y - v10.y));
          gy += g.z * (bl.dx * (v01.z - v00.z) + (1.f - bl.dx) * (v11.z - v10.z));
          gy += g.w * (bl.dx * (v01.w - v00.w) + (1.f - bl.dx) * (v11.w - v10.w));
          if (gb) {
            if (fxin && fyin) atomicAdd(reinterpret_cast<float4*>(gb + o00), make_float4(w00 * g.x, w00 * g.y, w00 * g.z, w00 * g.w));
            if (cxin && cyin) atomicAdd(reinterpret_cast<float4*>(gb + o11), make_float4(w11 * g.x, w11 * g.y, w11 * g.z, w11 * g.w));
            if (fxin && cyin) atomicAdd(reinterpret_cast<float4*>(gb + o01), make_float4(w01 * g.x, w01 * g.y, w01 * g.z, w01 * g.w));
            if (cxin && fyin) atomicAdd(reinterpret_cast<float4*>(gb + o10), make_float4(w10 * g.x, w10 * g.y, w10 * g.z, w10 * g.w));
          }
        }
        dlc = gx * (float)(W - 1);
        dlr = gy * (float)(H - 1);
      }
    }
    D[(2 * s) * LD + threadIdx.x] = dlr;
    D[(2 * s + 1) * LD + threadIdx.x] = dlc;
  }
  __syncthreads();
  // G[b0+s][j][comp] += sum_p Phi[j][p]*D[2s+comp][p]
  const int rows = n + 3;
  for (int o = threadIdx.x; o < rows * 2 * nb; o += TPS_T) {
    int j = o % rows;
    int sc = o / rows;  // 2*s+comp
    const float* pr = Phi + j * LD;
    const float* dr = D + sc * LD;
    float acc = 0.f;
#pragma unroll 8
    for (int p = 0; p < TPS_T; ++p) acc = fmaf(pr[p], dr[p], acc);
    int s = sc >> 1, comp = sc & 1;
    atomicAdd(G + ((int64_t)(b0 + s) * rows + j) * 2 + comp, (double)acc);
  }
}

// dtheta[b,t,comp] = sum_j Winv[j,t] G[b,j,comp] + sum_i Vinv[i,t] G[b,n+i,comp]
__global__ void tps_dtheta_kernel(const double* __restrict__ G, const float* __restrict__ consts,
                                  float* __restrict__ dtheta, int B, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n * 2) return;
  int comp = e & 1;
  int t = (e >> 1) % n;
  int b = (e >> 1) / n;
  const float* Winv = consts + 2 * n;
  const float* Vinv = Winv + n * n;
  const double* g = G + (int64_t)b * (n + 3) * 2;
  double acc = 0.0;
  for (int j = 0; j < n; ++j) acc += (double)Winv[j * n + t] * g[2 * j + comp];
  for (int i = 0; i < 3; ++i) acc += (double)Vinv[i * n + t] * g[2 * (n + i) + comp];
  dtheta[e] = (float)acc;
}

template <int C>
__global__ void __launch_bounds__(256) resampler_fwd_kernel(const float* __restrict__ vol, const float* __restrict__ warp,
                                                            float* __restrict__ out, int H, int W, int64_t m) {
  const int b = blockIdx.y;
  const float* vb = vol + (int64_t)b * H * W * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    float x = warp[((int64_t)b * m + i) * 2], y = warp[((int64_t)b * m + i) * 2 + 1];
    Bilin bl = bilin_setup(x, y, H, W);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (bl.valid) {
      const int cx = bl.fx + 1, cy = bl.fy + 1;
      const float w00 = bl.dx * bl.dy, w11 = (1.f - bl.dx) * (1.f - bl.dy);
      const float w01 = bl.dx * (1.f - bl.dy), w10 = (1.f - bl.dx) * bl.dy;
      const bool fxin = bl.fx >= 0 && bl.fx <= W - 1, cxin = cx >= 0 && cx <= W - 1;
      const bool fyin = bl.fy >= 0 && bl.fy <= H - 1, cyin = cy >= 0 && cy <= H - 1;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (fxin && fyin) acc[c] += w00 * vb[((int64_t)bl.fy * W + bl.fx) * C + c];
        if (cxin && cyin) acc[c] += w11 * vb[((int64_t)cy * W + cx) * C + c];
        if (fxin && cyin) acc[c] += w01 * vb[((int64_t)cy * W + bl.fx) * C + c];
        if (cxin && fyin) acc[c] += w10 * vb[((int64_t)bl.fy * W + cx) * C + c];
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[((int64_t)b * m + i) * C + c] = acc[c];
  }
}

// ------------------------------------------------------------------ host: fp64 constants
static double phi2_host(double r) { return 0.5 * r * log(r > 1e-10 ? r : 1e-10); }

// Gauss-Jordan inverse with partial pivoting, fp64
static bool invert(std::vector<double>& a, int n) {
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    double best = fabs(a[(size_t)col * n + col]);
    for (int r = col + 1; r < n; ++r)
      if (fabs(a[(size_t)r * n + col]) > best) { best = fabs(a[(size_t)r * n + col]); piv = r; }
    if (best == 0.0) return false;
    if (piv != col)
      for (int j = 0; j < n; ++j) {
        std::swap(a[(size_t)col * n + j], a[(size_t)piv * n + j]);
        std::swap(inv[(size_t)col * n + j], inv[(size_t)piv * n + j]);
      }
    double d = a[(size_t)col * n + col];
    for (int j = 0; j < n; ++j) { a[(size_t)col * n + j] /= d; inv[(size_t)col * n + j] /= d; }
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      double f = a[(size_t)r * n + col];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        a[(size_t)r * n + j] -= f * a[(size_t)col * n + j];
        inv[(size_t)r * n + j] -= f * inv[(size_t)col * n + j];
      }
    }
  }
  a.swap(inv);
  return true;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_tps_solve_batched(const float* train_points, const float* train_values, float* w_out, float* v_out, int B,
                           int n, int k, int order, float reg, void* stream) {
  DAFK_REQUIRE(B >= 0 && n > 0 && k > 0, DAFK_ERR_BAD_ARG, "dafk_tps_solve_batched: bad shape");
  DAFK_REQUIRE(n + 3 <= SYS_MAX && k <= RHS_MAX, DAFK_ERR_UNSUPPORTED,
               "dafk_tps_solve_batched: need n+3 <= %d and k <= %d", SYS_MAX, RHS_MAX);
  DAFK_REQUIRE(order >= 1, DAFK_ERR_BAD_ARG, "dafk_tps_solve_batched: bad order");
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(train_points && train_values && w_out && v_out, DAFK_ERR_BAD_ARG, "dafk_tps_solve_batched: null pointer");
  tps_solve_kernel<<<B, 32, 0, as_stream(stream)>>>(train_points, train_values, w_out, v_out, n, k, order, reg);
  return check_launch("dafk_tps_solve_batched");
}

int dafk_tps_apply(const float* query, const float* train_points, const float* w, const float* v, float* out, int B,
                   int64_t m, int n, int k, int order, int query_batched, void* stream) {
  DAFK_REQUIRE(B >= 0 && m >= 0 && n > 0 && k > 0 && k <= RHS_MAX, DAFK_ERR_BAD_ARG, "dafk_tps_apply: bad shape");
  if (B == 0 || m == 0) return DAFK_OK;
  DAFK_REQUIRE(query && train_points && w && v && out, DAFK_ERR_BAD_ARG, "dafk_tps_apply: null pointer");
  size_t smem = sizeof(float) * (2 * n + n * k + 3 * k);
  int chunks = (int)((m + 255) / 256);
  int cap = (kNumSMs * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  tps_apply_kernel<<<dim3(chunks, B), 256, smem, as_stream(stream)>>>(query, train_points, w, v, out, m, n, k, order,
                                                                      query_batched);
  return check_launch("dafk_tps_apply");
}

int dafk_tps_consts_floats(int n_cp) { return n_cp * (n_cp + 5); }

int dafk_tps_build_constants(int cp_h, int cp_w, float* consts_host) {
  DAFK_REQUIRE(cp_h > 1 && cp_w > 1 && consts_host, DAFK_ERR_BAD_ARG, "dafk_tps_build_constants: bad argument");
  const int n = cp_h * cp_w;
  DAFK_REQUIRE(n <= TPS_MAXN, DAFK_ERR_UNSUPPORTED, "dafk_tps_build_constants: at most %d control points", TPS_MAXN);
  const int sz = n + 3;
  // control points as float32(nDgrid) promoted to double: the reference grid is a float32 tensor
  std::vector<double> c((size_t)n * 2);
  for (int i = 0; i < cp_h; ++i)
    for (int j = 0; j < cp_w; ++j) {
      c[(size_t)(i * cp_w + j) * 2] = (double)(float)((double)i / (double)(cp_h - 1));
      c[(size_t)(i * cp_w + j) * 2 + 1] = (double)(float)((double)j / (double)(cp_w - 1));
    }
  std::vector<double> L((size_t)sz * sz, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      double d0 = c[2 * i] - c[2 * j], d1 = c[2 * i + 1] - c[2 * j + 1];
      L[(size_t)i * sz + j] = phi2_host(d0 * d0 + d1 * d1);
    }
    L[(size_t)i * sz + n] = c[2 * i];
    L[(size_t)i * sz + n + 1] = c[2 * i + 1];
    L[(size_t)i * sz + n + 2] = 1.0;
    L[(size_t)n * sz + i] = c[2 * i];
    L[(size_t)(n + 1) * sz + i] = c[2 * i + 1];
    L[(size_t)(n + 2) * sz + i] = 1.0;
  }
  DAFK_REQUIRE(invert(L, sz), DAFK_ERR_BAD_ARG, "dafk_tps_build_constants: singular system");
  for (int i = 0; i < 2 * n; ++i) consts_host[i] = (float)c[i];
  float* Winv = consts_host + 2 * n;
  float* Vinv = Winv + n * n;
  for (int j = 0; j < n; ++j)
    for (int t = 0; t < n; ++t) Winv[j * n + t] = (float)L[(size_t)j * sz + t];
  for (int i = 0; i < 3; ++i)
    for (int t = 0; t < n; ++t) Vinv[i * n + t] = (float)L[(size_t)(n + i) * sz + t];
  return DAFK_OK;
}

int dafk_tps_warp_fwd(const float* vol, const float* theta, const float* consts, float* out, float* locs, int B,
                      int H, int W, int C, int n_cp, void* stream) {
  DAFK_REQUIRE(B >= 0 && H > 1 && W > 1 && C > 0 && n_cp > 0, DAFK_ERR_BAD_ARG, "dafk_tps_warp_fwd: bad shape");
  DAFK_REQUIRE(n_cp <= TPS_MAXN, DAFK_ERR_UNSUPPORTED, "dafk_tps_warp_fwd: at most %d control points", TPS_MAXN);
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(vol && theta && consts && out, DAFK_ERR_BAD_ARG, "dafk_tps_warp_fwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(vol) && DAFK_ALIGNED16(out), DAFK_ERR_ALIGN, "dafk_tps_warp_fwd: alignment");
  dim3 grid((H * W + TPS_T - 1) / TPS_T, (B + TPS_BS - 1) / TPS_BS);
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 4: tps_warp_fwd_kernel<4><<<grid, TPS_T, 0, s>>>(vol, theta, consts, out, locs, B, H, W, n_cp); break;
    case 8: tps_warp_fwd_kernel<8><<<grid, TPS_T, 0, s>>>(vol, theta, consts, out, locs, B, H, W, n_cp); break;
    case 16: tps_warp_fwd_kernel<16><<<grid, TPS_T, 0, s>>>(vol, theta, consts, out, locs, B, H, W, n_cp); break;
    default: set_error("dafk_tps_warp_fwd: C must be 4, 8 or 16 (got %d)", C); return DAFK_ERR_UNSUPPORTED;
  }
  return check_launch("dafk_tps_warp_fwd");
}

int64_t dafk_tps_phi_table_floats(int H, int W, int n_cp) { return (int64_t)H * W * (n_cp + 2); }

int dafk_tps_phi_table(const float* consts, float* table, int H, int W, int n_cp, void* stream) {
  DAFK_REQUIRE(H > 1 && W > 1 && n_cp > 0 && n_cp <= TPS_MAXN, DAFK_ERR_BAD_ARG, "dafk_tps_phi_table: bad shape");
  DAFK_REQUIRE(consts && table, DAFK_ERR_BAD_ARG, "dafk_tps_phi_table: null pointer");
  tps_phi_table_kernel<<<(H * W + TPS_T - 1) / TPS_T, TPS_T, 0, as_stream(stream)>>>(consts, table, H, W, n_cp);
  return check_launch("dafk_tps_phi_table");
}

int dafk_tps_warp_fwd_tab(const float* vol, const float* theta, const float* consts, const float* phi_table,
                          float* coef_ws, float* out, float* locs, int B, int H, int W, int C, int n_cp, void* stream) {
  DAFK_REQUIRE(B >= 0 && H > 1 && W > 1 && C > 0 && n_cp > 0, DAFK_ERR_BAD_ARG, "dafk_tps_warp_fwd_tab: bad shape");
  DAFK_REQUIRE(n_cp <= TPS_MAXN, DAFK_ERR_UNSUPPORTED, "dafk_tps_warp_fwd_tab: at most %d control points", TPS_MAXN);
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(vol && theta && consts && phi_table && coef_ws && out, DAFK_ERR_BAD_ARG, "dafk_tps_warp_fwd_tab: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(vol) && DAFK_ALIGNED16(out), DAFK_ERR_ALIGN, "dafk_tps_warp_fwd_tab: alignment");
  DAFK_REQUIRE(C == 4 || C == 8 || C == 16, DAFK_ERR_UNSUPPORTED, "dafk_tps_warp_fwd_tab: C must be 4, 8 or 16 (got %d)", C);
  cudaStream_t s = as_stream(stream);
  const int ncoef = B * (n_cp + 3) * 2;
  tps_coef_kernel<<<(ncoef + 127) / 128, 128, 0, s>>>(theta, consts, coef_ws, B, n_cp);
  int rc = check_launch("dafk_tps_warp_fwd_tab(coef)");
  if (rc) return rc;
  dim3 grid((H * W + TPS_T - 1) / TPS_T, (B + TPS_BS_TAB - 1) / TPS_BS_TAB);
  if (n_cp == 25 && C == 8) {        // the shipped geometry: 5 x 5 control points on the 8-channel anatomy
    tps_warp_fwd_tab_kernel<8, 25><<<grid, TPS_T, 0, s>>>(vol, coef_ws, phi_table, out, locs, B, H, W, n_cp);
    return check_launch("dafk_tps_warp_fwd_tab");
  }
  switch (C) {
    case 4: tps_warp_fwd_tab_kernel<4, 0><<<grid, TPS_T, 0, s>>>(vol, coef_ws, phi_table, out, locs, B, H, W, n_cp); break;
    case 8: tps_warp_fwd_tab_kernel<8, 0><<<grid, TPS_T, 0, s>>>(vol, coef_ws, phi_table, out, locs, B, H, W, n_cp); break;
    case 16: tps_warp_fwd_tab_kernel<16, 0><<<grid, TPS_T, 0, s>>>(vol, coef_ws, phi_table, out, locs, B, H, W, n_cp); break;
    default: set_error("dafk_tps_warp_fwd_tab: C must be 4, 8 or 16 (got %d)", C); return DAFK_ERR_UNSUPPORTED;
  }
  return check_launch("dafk_tps_warp_fwd_tab");
}

static int tps_warp_bwd_impl(const float* vol, const float* theta, const float* consts, const float* phi_table,
                             const float* dout, float* dvol, float* dtheta, double* ws, int B, int H, int W, int C, int n_cp,
                             void* stream) {
  DAFK_REQUIRE(B >= 0 && H > 1 && W > 1 && C > 0 && n_cp > 0, DAFK_ERR_BAD_ARG, "dafk_tps_warp_bwd: bad shape");
  DAFK_REQUIRE(n_cp <= TPS_MAXN, DAFK_ERR_UNSUPPORTED, "dafk_tps_warp_bwd: at most %d control points", TPS_MAXN);
  if (B == 0) return DAFK_OK;
  DAFK_REQUIRE(vol && theta && consts && dout && dtheta && ws, DAFK_ERR_BAD_ARG, "dafk_tps_warp_bwd: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(vol) && DAFK_ALIGNED16(dout) && DAFK_ALIGNED16(dvol), DAFK_ERR_ALIGN, "dafk_tps_warp_bwd: alignment");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)B * (n_cp + 3) * 2, s);
  dim3 grid((H * W + TPS_T - 1) / TPS_T, (B + TPS_BS - 1) / TPS_BS);
  size_t smem = sizeof(float) * (size_t)((TPS_MAXN + 3) + 2 * TPS_BS) * (TPS_T + 1);
#define TPS_BWD(CC, NN)                                                                                            \
  do {                                                                                                             \
    cudaFuncSetAttribute(tps_warp_bwd_kernel<CC, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    tps_warp_bwd_kernel<CC, NN><<<grid, TPS_T, smem, s>>>(vol, theta, consts, dout, dvol, ws, B, H, W, n_cp, phi_table); \
  } while (0)
  if (C == 8 && n_cp == 25) TPS_BWD(8, 25);            // the shipped geometry
  else if (C == 4) TPS_BWD(4, 0);
  else if (C == 8) TPS_BWD(8, 0);
  else if (C == 16) TPS_BWD(16, 0);
  else { set_error("dafk_tps_warp_bwd: C must be 4, 8 or 16 (got %d)", C); return DAFK_ERR_UNSUPPORTED; }
#undef TPS_BWD
  int rc = check_launch("dafk_tps_warp_bwd");
  if (rc) return rc;
  int total = B * n_cp * 2;
  tps_dtheta_kernel<<<(total + 127) / 128, 128, 0, s>>>(ws, consts, dtheta, B, n_cp);
  return check_launch("dafk_tps_warp_bwd(dtheta)");
}

int dafk_tps_warp_bwd(const float* vol, const float* theta, const float* consts, const float* dout, float* dvol,
                      float* dtheta, double* ws, int B, int H, int W, int C, int n_cp, void* stream) {
  return tps_warp_bwd_impl(vol, theta, consts, nullptr, dout, dvol, dtheta, ws, B, H, W, C, n_cp, stream);
}

int dafk_tps_warp_bwd_tab(const float* vol, const float* theta, const float* consts, const float* phi_table,
                          const float* dout, float* dvol, float* dtheta, double* ws, int B, int H, int W, int C, int n_cp,
                          void* stream) {
  DAFK_REQUIRE(phi_table, DAFK_ERR_BAD_ARG, "dafk_tps_warp_bwd_tab: null table");
  return tps_warp_bwd_impl(vol, theta, consts, phi_table, dout, dvol, dtheta, ws, B, H, W, C, n_cp, stream);
}

int dafk_resampler_fwd(const float* vol, const float* warp, float* out, int B, int H, int W, int C, int64_t m,
                       void* stream) {
  DAFK_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && m >= 0, DAFK_ERR_BAD_ARG, "dafk_resampler_fwd: bad shape");
  if (B == 0 || m == 0) return DAFK_OK;
  DAFK_REQUIRE(vol && warp && out, DAFK_ERR_BAD_ARG, "dafk_resampler_fwd: null pointer");
  int chunks = (int)((m + 255) / 256);
  int cap = (kNumSMs * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  dim3 grid(chunks, B);
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 1: resampler_fwd_kernel<1><<<grid, 256, 0, s>>>(vol, warp, out, H, W, m); break;
    case 4: resampler_fwd_kernel<4><<<grid, 256, 0, s>>>(vol, warp, out, H, W, m); break;
    case 8: resampler_fwd_kernel<8><<<grid, 256, 0, s>>>(vol, warp, out, H, W, m); break;
    case 16: resampler_fwd_kernel<16><<<grid, 256, 0, s>>>(vol, warp, out, H, W, m); break;
    default: set_error("dafk_resampler_fwd: C must be 1, 4, 8 or 16 (got %d)", C); return DAFK_ERR_UNSUPPORTED;
  }
  return check_launch("dafk_resampler_fwd");
}

}  // extern "C"
