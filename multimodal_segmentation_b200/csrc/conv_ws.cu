// Narrow-channel convolutions as HBM-bandwidth kernels ("warp-strip"): every warp of a CTA is producer, tensor-core
// user and writer in turn, several CTAs per SM hide each other's memory latency.
//
// The layers served here -- the FiLM decoder's 8 -> 8 stack (model_components/decoder.py:44-54), the first layers of the
// segmentor / UNet / discriminators (8 -> 64, 1 -> 64, space-to-depth 4 -> 64), the modality encoder's strided stack and
// the locnet's 5x5 layers (layers/stn_spline.py:106-112) -- move 16 .. 300 bytes per pixel and need 0.1 .. 5 kFLOP per
// pixel: they are bound by HBM, not by the tensor pipe.  The tcgen05 raster-strip kernels (conv_nc.cu) keep ONE CTA per
// SM with 4 epilogue warps behind a TMEM round trip and measured 15-65 % of the copy bandwidth (profiles/r1_bench_nc.txt):
// the per-tile chain TMEM load -> index arithmetic -> activation -> store on one warp per scheduler bounded them.  Here the
// accumulators never leave the register file:
//
//   * the R + KH - 1 input rows of a strip are contiguous in global memory: ONE cp.async.bulk (TMA engine, mbarrier byte
//     count) brings them into a raw shared-memory stage, one or two strips AHEAD of the arithmetic, so DRAM latency is
//     never exposed to the warps and costs no registers;
//   * all warps then rewrite the raw stage as bf16 channel-group planes [cg][row * P + col][8 ch = 16 B] (P = W + 2*pad:
//     the zero halo columns are part of the raster; fp32 -> bf16 and, for a data / weight gradient, the activation
//     backward dy * act'(y) happen in this shared-to-shared step) and hand the raw stage back to the TMA engine;
//   * filter tap (r,q) of channel group cg is the SAME shared memory shifted by (r*P + q) positions, so the A fragment of
//     mma.sync.m16n8k16 for 16 consecutive raster positions and two (tap, group) slices is ONE ldmatrix.x4 whose 32 row
//     addresses are computed per lane -- im2col is never materialised;
//   * B fragments (weights) are built once per CTA in shared memory in register order (one 8-byte load per mma);
//   * outputs leave straight from the accumulator registers: bias, activation, dtype conversion, 8-byte stores that
//     cover whole 32-byte sectors per quad.
//   forward / stride-1 data gradient:  y[pos, co] = act(sum_e A_e[pos, 0:8] . W[e][0:8][co] + b[co])
//   weight gradient:  dW[e][ci][co] = sum_pos X[pos + off_e][ci] * dY[pos][co]  (A and B both through ldmatrix.trans, the
//     reduction runs over raster positions), fp32 accumulators live in registers over all strips of a CTA.
#include "tc_ptx.cuh"
#include <mutex>

namespace dafk {

constexpr int WS_THREADS = 256;
constexpr int WS_WARPS = WS_THREADS / 32;
constexpr int WS_U = 2;          // staging: (pixel, channel group) units per lane per work item

struct WsP {
  int N, H, W, Cin, Cout;          // kernel view: input [N,H,W,Cin] -> output [N,Ho,Wo,Cout]
  int KH, KW, pad, Ho, Wo;
  int P, R, RS, CG, NT, J, E;      // raster pitch, output rows / strip, input rows / strip, input channel groups,
                                   // output n-tiles of 8, k-steps of 16, (tap, group) slices
  int plane;                       // positions per channel-group plane
  int strips_per_img, total_strips;
  int mode;                        // 0: weights as stored (forward); 1: mirrored + transposed (data gradient)
  int wCin, wCout;                 // dimensions of the HWIO weight tensor in global memory
  int act; float alpha;            // epilogue activation
  int x_dt, y_dt;
  int gact, ga_dt; float galpha;   // staging: x := x * act'(ya) (activation backward fused into a data gradient)
  int ipr;                         // staging work items (32 * WS_U units) per image row
  int S, stage_bytes, raw_a_off;   // raw (TMA) stages: count, bytes per stage, offset of the ya rows inside a stage
  uint32_t rowbytes, rowbytes_a;   // bytes of one image row of x / ya
  uint32_t magicP, magicCG, magic_ipr, magic_spi;   // ceil(2^32 / d): u / d == umulhi(u, magic) for u, d < 2^16
};

__host__ __device__ inline uint32_t ws_magic(int d) { return (uint32_t)((0x100000000ULL + (uint64_t)d - 1) / (uint64_t)d); }
// u / d for u, d < 2^16 (d == 1: the magic number does not fit 32 bits)
__device__ __forceinline__ int ws_div(int u, uint32_t magic, int d) { return d == 1 ? u : (int)__umulhi((uint32_t)u, magic); }

template <int ACT>
__device__ __forceinline__ float ws_act(float z, float alpha) {
  if (ACT == DAFK_ACT_RELU) return fmaxf(z, 0.f);
  if (ACT == DAFK_ACT_LRELU) return z > 0.f ? z : alpha * z;
  if (ACT == DAFK_ACT_TANH) return tanhf(z);
  return z;
}
// derivative of the activation expressed through its OUTPUT y (Keras 2.1.6: 0 at exactly 0 for relu / lrelu)
__device__ __forceinline__ float ws_dact(float y, int act, float alpha) {
  if (act == DAFK_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == DAFK_ACT_LRELU) return y > 0.f ? 1.f : (y < 0.f ? alpha : 0.f);
  if (act == DAFK_ACT_TANH) return 1.f - y * y;
  return 1.f;
}

// 8 consecutive channels at shared-memory address `a` as floats.  FAST: the channel count is a multiple of 8 (whole 16 /
// 32-byte vectors); otherwise nvalid of the 8 exist and are fetched one by one (C = 1, 4, 20, 36: small layers)
template <typename T, bool FAST>
__device__ __forceinline__ void ws_lds8(uint32_t a, int nvalid, float (&v)[8]) {
  if (FAST) {
    if (sizeof(T) == 4) {
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a + 16u));
    } else {
      uint32_t w[4];
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        v[2 * i] = __low2float(h);
        v[2 * i + 1] = __high2float(h);
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = 0.f;
      if (c < nvalid) {
        if (sizeof(T) == 4) {
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[c]) : "r"(a + 4u * c));
        } else {
          unsigned short h;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(a + 2u * c));
          v[c] = __uint_as_float((uint32_t)h << 16);
        }
      }
    }
  }
}
__device__ __forceinline__ uint4 ws_pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Raw stage -> raster: rows [iy0, iy0 + rows) of one image sit in shared memory exactly as in global memory
// ([row][pixel][C], element type TX, raw row r at raw + r * wcols * C); they become bf16 planes [g][row * P + col0 + px][16 B].
// A work item = 32 * WS_U consecutive (pixel, channel group) units of one row, items go round-robin over the warps; all
// index arithmetic is per item (warp-uniform) or a multiply-high.  Rows outside the image were not fetched and are
// written as zeros.  GA: x := x * act'(ya), ya (type TA, same shape) in a second raw buffer.
// Deliberately NOT inlined: every (dtype, fused activation, vector width) variant is its own function with its own
// register allocation -- inlined side by side behind a runtime dispatch ptxas moved the values in flight to local memory
// -- and the call happens once per strip.
template <typename TX, bool GA, typename TA, bool FAST>
__device__ __noinline__ void ws_convert(uint32_t raw, uint32_t raw_a, int gact, float galpha, uint8_t* planes, int plane_pos,
                                        int Himg, int wcols, int C, int CG, uint32_t magicCG, int iy0, int rows, int P, int col0,
                                        int ipr, uint32_t magic_ipr, int warp, int lane) {
  const int units_row = wcols * CG;
  const int items = rows * ipr;
  for (int item = warp; item < items; item += WS_WARPS) {
    const int row = ws_div(item, magic_ipr, ipr);
    const int ub = (item - row * ipr) * (32 * WS_U) + lane;
    const int iy = iy0 + row;
    uint8_t* drow = planes + (size_t)(row * P + col0) * 16;
    if (iy >= 0 && iy < Himg) {
      const int roff = row * wcols * C;
      float v[WS_U][8];
      int dst[WS_U];
#pragma unroll
      for (int k = 0; k < WS_U; ++k) {
        const int u = ub + 32 * k;
        dst[k] = -1;
        if (u < units_row) {
          const int px = ws_div(u, magicCG, CG);
          const int g = u - px * CG;
          dst[k] = g * plane_pos + px;
          const int off = roff + px * C + g * 8;
          const int nv = min(8, C - g * 8);
          ws_lds8<TX, FAST>(raw + (uint32_t)off * (uint32_t)sizeof(TX), nv, v[k]);
          if (GA) {
            float a[8];
            ws_lds8<TA, FAST>(raw_a + (uint32_t)off * (uint32_t)sizeof(TA), nv, a);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[k][c] *= ws_dact(a[c], gact, galpha);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < WS_U; ++k)
        if (dst[k] >= 0) *reinterpret_cast<uint4*>(drow + (size_t)dst[k] * 16) = ws_pack8(v[k]);
    } else {
#pragma unroll
      for (int k = 0; k < WS_U; ++k) {
        const int u = ub + 32 * k;
        if (u < units_row) {
          const int px = ws_div(u, magicCG, CG);
          const int g = u - px * CG;
          *reinterpret_cast<uint4*>(drow + (size_t)(g * plane_pos + px) * 16) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
}

// dtype / fused-activation dispatch (warp-uniform branches OUTSIDE the copy loops)
__device__ __forceinline__ void ws_convert_any(uint32_t raw, int dt, uint32_t raw_a, int ya_dt, int gact, float galpha,
                                               uint8_t* planes, int plane_pos, int Himg, int wcols, int C, int CG,
                                               uint32_t magicCG, int iy0, int rows, int P, int col0, int ipr, uint32_t magic_ipr,
                                               int warp, int lane) {
  typedef __nv_bfloat16 bf;
#define WS_ST2(TX, GA, TA, FAST)                                                                                            \
  ws_convert<TX, GA, TA, FAST>(raw, raw_a, gact, galpha, planes, plane_pos, Himg, wcols, C, CG, magicCG, iy0, rows, P, col0, \
                               ipr, magic_ipr, warp, lane)
#define WS_ST(TX, GA, TA)                                      \
  do {                                                         \
    if ((C & 7) == 0) WS_ST2(TX, GA, TA, true);                \
    else WS_ST2(TX, GA, TA, false);                            \
  } while (0)
  if (gact == DAFK_ACT_NONE) {
    if (dt == DAFK_F32) WS_ST(float, false, float); else WS_ST(bf, false, float);
  } else if (dt == DAFK_F32) {
    if (ya_dt == DAFK_F32) WS_ST(float, true, float); else WS_ST(float, true, bf);
  } else {
    if (ya_dt == DAFK_F32) WS_ST(bf, true, float); else WS_ST(bf, true, bf);
  }
#undef WS_ST2
#undef WS_ST
}

// One elected thread: fetch rows [iy0, iy0 + rows) of an image (clipped to the image) with ONE bulk copy into a raw stage.
// Rows of an NHWC image are contiguous, so the clipped row range is a single byte range.  Returns the bytes requested.
__device__ __forceinline__ uint32_t ws_fetch(uint8_t* raw, const void* src, int64_t img_bytes_off, int Himg, uint32_t rowbytes,
                                             int iy0, int rows, uint64_t* bar) {
  const int first = max(iy0, 0), last = min(iy0 + rows, Himg);
  if (last <= first) return 0u;
  const uint32_t bytes = (uint32_t)(last - first) * rowbytes;
  bulk_load_1d(raw + (size_t)(first - iy0) * rowbytes, reinterpret_cast<const uint8_t*>(src) + img_bytes_off + (int64_t)first * rowbytes,
               bytes, bar);
  return bytes;
}
__device__ __forceinline__ uint32_t ws_fetch_bytes(int Himg, uint32_t rowbytes, int iy0, int rows) {
  const int first = max(iy0, 0), last = min(iy0 + rows, Himg);
  return last > first ? (uint32_t)(last - first) * rowbytes : 0u;
}

__device__ __forceinline__ void ws_ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ws_ldsm4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ws_ldsm2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
// D (16x8, f32) += A (16x16, bf16, row) * B (16x8, bf16, col)
__device__ __forceinline__ void ws_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// kernel-view weight element (slice e = (cg, r, q), channel c of the group, output channel n)
__device__ __forceinline__ float ws_weight(const float* __restrict__ w, const float* __restrict__ scale, const WsP& p,
                                           int e, int c, int n) {
  if (e >= p.E) return 0.f;
  const int q = e % p.KW;
  const int t = e / p.KW;
  const int r = t % p.KH;
  const int cg = t / p.KH;
  const int ck = cg * 8 + c;
  if (ck >= p.Cin || n >= p.Cout) return 0.f;
  if (p.mode == 0) {
    const float v = w[(((int64_t)r * p.KW + q) * p.wCin + ck) * p.wCout + n];
    return scale ? v * scale[n] : v;
  }
  // data gradient: kernel-view input channels are the layer's outputs; taps mirrored
  return w[(((int64_t)(p.KH - 1 - r) * p.KW + (p.KW - 1 - q)) * p.wCin + n) * p.wCout + ck];
}

// rows g and g + 8 of each 16-position block of a pass, channels nt*8 + 2t, 2t + 1: bias, activation, conversion, store
template <int NT, int MB, int ACT, typename TY>
__device__ __forceinline__ void ws_epilogue(const float (&acc)[MB][NT][4], const float (&breg)[NT][2], TY* __restrict__ ystrip,
                                            int c0, int chunks, int rows_here, int P, uint32_t magicP, int Wo, int Cout, int g,
                                            int t, float alpha) {
  const bool even = (Cout & 1) == 0;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = (c0 + mb) * 16 + g + 8 * h;
      const int orow = (int)__umulhi((uint32_t)m, magicP);
      const int ocol = m - orow * P;
      if ((c0 + mb) < chunks && orow < rows_here && ocol < Wo) {
        TY* o = ystrip + (orow * Wo + ocol) * Cout + 2 * t;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int ch = nt * 8 + 2 * t;
          const float v0 = ws_act<ACT>(acc[mb][nt][2 * h] + breg[nt][0], alpha);
          const float v1 = ws_act<ACT>(acc[mb][nt][2 * h + 1] + breg[nt][1], alpha);
          if (even) {
            if (ch < Cout) {
              if (sizeof(TY) == 4) {
                *reinterpret_cast<float2*>(o + nt * 8) = make_float2(v0, v1);
              } else {
                __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                *reinterpret_cast<__nv_bfloat162*>(o + nt * 8) = hh;
              }
            }
          } else {
            if (ch < Cout) o[nt * 8] = from_f<TY>(v0);
            if (ch + 1 < Cout) o[nt * 8 + 1] = from_f<TY>(v1);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward / stride-1 data gradient.  smem: [Wf: J*NT*32 x 8 B][slice offsets: 2*J u32][raster planes]
// NT = output n-tiles of 8 channels; MB = 16-position blocks a warp processes per pass (register blocking of M: every
// B fragment is used MB times); JR > 0: J <= JR and the B fragments + slice offsets of the whole filter live in registers
// (3x3 / 2x2 layers with up to 8 input and 16 output channels: the FiLM decoder)
// ---------------------------------------------------------------------------------------------
template <int NT, int MB, int JR>
__global__ void __launch_bounds__(WS_THREADS, 2) conv_ws_fwd_kernel(WsP p, const void* __restrict__ x,
                                                                    const void* __restrict__ ya, const float* __restrict__ w,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ bias, void* __restrict__ y) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint2* s_wf = reinterpret_cast<uint2*>(smem);
  uint32_t* s_off = reinterpret_cast<uint32_t*>(s_wf + (size_t)p.J * NT * 32);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_off + 2 * p.J) + 15) & ~(uintptr_t)15);
  uint8_t* s_x = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_bar + 4) + 127) & ~(uintptr_t)127);
  uint8_t* s_raw = s_x + (size_t)p.CG * p.plane * 16;        // S raw stages, 128-byte aligned (plane is a multiple of 8)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- one-time setup: B fragments in register order, slice offsets, zeroed raster (halo columns stay zero)
  for (int i = tid; i < p.J * NT * 32; i += WS_THREADS) {
    const int l = i & 31;
    const int nt = (i >> 5) % NT;
    const int j = (i >> 5) / NT;
    const int gg = l >> 2, tt = l & 3;
    const int n = nt * 8 + gg;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(ws_weight(w, scale, p, 2 * j, 2 * tt, n), ws_weight(w, scale, p, 2 * j, 2 * tt + 1, n));
    __nv_bfloat162 b1 = __floats2bfloat162_rn(ws_weight(w, scale, p, 2 * j + 1, 2 * tt, n), ws_weight(w, scale, p, 2 * j + 1, 2 * tt + 1, n));
    s_wf[i] = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
  }
  for (int e = tid; e < 2 * p.J; e += WS_THREADS) {
    const int ee = e < p.E ? e : p.E - 1;          // the padding slice of an odd E reads a valid address (zero weights)
    const int q = ee % p.KW;
    const int tt = ee / p.KW;
    const int r = tt % p.KH;
    const int cg = tt / p.KH;
    s_off[e] = (uint32_t)((cg * p.plane + r * p.P + q) * 16);
  }
  for (int i = tid; i < p.CG * p.plane; i += WS_THREADS) reinterpret_cast<uint4*>(s_x)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) mbar_init(s_bar + i, 1);
    fence_barrier_init();
  }
  float breg[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int ch = nt * 8 + 2 * t;
    breg[nt][0] = (bias != nullptr && ch < p.Cout) ? __ldg(bias + ch) : 0.f;
    breg[nt][1] = (bias != nullptr && ch + 1 < p.Cout) ? __ldg(bias + ch + 1) : 0.f;
  }
  __syncthreads();

  const uint32_t sx_addr = (uint32_t)__cvta_generic_to_shared(s_x);
  const int lm = lane >> 3, lr = lane & 7;               // ldmatrix: this lane supplies row lr of matrix lm
  const uint32_t lane_row = (uint32_t)(((lm & 1) * 8 + lr) * 16);
  const int lsel = lm >> 1;                              // matrices 0,1: first slice of the k-step; 2,3: second slice
  const int J = p.J, P = p.P, Wo = p.Wo, Cout = p.Cout, R = p.R, Ho = p.Ho, spi = p.strips_per_img;
  const uint32_t magicP = p.magicP;
  constexpr int JRN = JR > 0 ? JR : 1;
  uint32_t offr[JRN];
  uint2 wfr[JRN][NT];
  if (JR > 0) {
#pragma unroll
    for (int j = 0; j < JRN; ++j) {
      offr[j] = j < J ? s_off[2 * j + lsel] : 0u;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) wfr[j][nt] = j < J ? s_wf[(j * NT + nt) * 32 + lane] : make_uint2(0u, 0u);
    }
  }

  // ---- TMA pipeline: strip i of this CTA lives in raw stage i % S; one thread issues, everybody waits on the mbarrier
  const int my_strips = (int)blockIdx.x < p.total_strips ? (p.total_strips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int S = p.S;
  const int64_t img_x = (int64_t)p.H * p.rowbytes, img_a = (int64_t)p.H * p.rowbytes_a;
  auto issue = [&](int i) {
    const int s2 = (int)blockIdx.x + i * (int)gridDim.x;
    const int n2 = ws_div(s2, p.magic_spi, spi);
    const int iy0 = (s2 - n2 * spi) * R - p.pad;
    uint64_t* bar = s_bar + (i % S);
    uint8_t* raw = s_raw + (size_t)(i % S) * p.stage_bytes;
    uint32_t bytes = ws_fetch_bytes(p.H, p.rowbytes, iy0, p.RS);
    if (p.gact != DAFK_ACT_NONE) bytes += ws_fetch_bytes(p.H, p.rowbytes_a, iy0, p.RS);
    mbar_expect_tx(bar, bytes);
    ws_fetch(raw, x, n2 * img_x, p.H, p.rowbytes, iy0, p.RS, bar);
    if (p.gact != DAFK_ACT_NONE) ws_fetch(raw + p.raw_a_off, ya, n2 * img_a, p.H, p.rowbytes_a, iy0, p.RS, bar);
  };
  if (tid == 0)
    for (int i = 0; i < S && i < my_strips; ++i) issue(i);

  for (int i = 0; i < my_strips; ++i) {
    const int s = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = ws_div(s, p.magic_spi, spi);
    const int y0 = (s - n * spi) * R;
    const int rows_here = min(R, Ho - y0);
    const int st = i % S;
    mbar_wait(s_bar + st, (uint32_t)(i / S) & 1u);
    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw + (size_t)st * p.stage_bytes);
    ws_convert_any(raw_addr, p.x_dt, raw_addr + (uint32_t)p.raw_a_off, p.ga_dt, p.gact, p.galpha, s_x, p.plane, p.H, p.W, p.Cin,
                   p.CG, p.magicCG, y0 - p.pad, p.RS, P, p.pad, p.ipr, p.magic_ipr, warp, lane);
    __syncthreads();                       // raster complete; the raw stage is free again
    if (tid == 0 && i + S < my_strips) {
      fence_proxy_async();                 // the generic-proxy reads of the stage are ordered before the TMA engine's writes
      issue(i + S);
    }
    const int chunks = (rows_here * P + 15) >> 4;
    const int64_t ybase = ((int64_t)n * Ho + y0) * Wo * Cout;
    for (int c0 = warp * MB; c0 < chunks; c0 += WS_WARPS * MB) {
      float acc[MB][NT][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mb][nt][i] = 0.f;
      const uint32_t a_base = sx_addr + (uint32_t)(c0 * 256) + lane_row;
      if (JR > 0) {
#pragma unroll
        for (int j = 0; j < JRN; ++j) {
          if (j < J) {
            uint32_t a[MB][4];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) ws_ldsm4(a_base + offr[j] + (uint32_t)(mb * 256), a[mb]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) ws_mma(acc[mb][nt], a[mb], wfr[j][nt].x, wfr[j][nt].y);
          }
        }
      } else {
        const uint2* wf = s_wf + lane;
        const uint32_t* so = s_off + lsel;
#pragma unroll 2
        for (int j = 0; j < J; ++j) {
          const uint32_t off = so[2 * j];
          uint32_t a[MB][4];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) ws_ldsm4(a_base + off + (uint32_t)(mb * 256), a[mb]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint2 b = wf[(j * NT + nt) * 32];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) ws_mma(acc[mb][nt], a[mb], b.x, b.y);
          }
        }
      }
      // ---- epilogue (activation and output dtype: warp-uniform dispatch outside the store loops)
#define WS_EPI(ACT, TY)                                                                                                  \
  ws_epilogue<NT, MB, ACT, TY>(acc, breg, reinterpret_cast<TY*>(y) + ybase, c0, chunks, rows_here, P, magicP, Wo, Cout, g, t, \
                               p.alpha)
      if (p.y_dt == DAFK_F32) {
        switch (p.act) {
          case DAFK_ACT_RELU: WS_EPI(DAFK_ACT_RELU, float); break;
          case DAFK_ACT_LRELU: WS_EPI(DAFK_ACT_LRELU, float); break;
          case DAFK_ACT_TANH: WS_EPI(DAFK_ACT_TANH, float); break;
          default: WS_EPI(DAFK_ACT_NONE, float); break;
        }
      } else {
        switch (p.act) {
          case DAFK_ACT_RELU: WS_EPI(DAFK_ACT_RELU, __nv_bfloat16); break;
          case DAFK_ACT_LRELU: WS_EPI(DAFK_ACT_LRELU, __nv_bfloat16); break;
          case DAFK_ACT_TANH: WS_EPI(DAFK_ACT_TANH, __nv_bfloat16); break;
          default: WS_EPI(DAFK_ACT_NONE, __nv_bfloat16); break;
        }
      }
#undef WS_EPI
    }
    __syncthreads();          // every warp is done with this raster before the next strip overwrites it
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient (+ bias gradient).  M = (slice, channel) rows in m-tiles of two slices, N = output channels, K = raster
// positions.  A warp owns MTW x NTW accumulator tiles (registers) for its share of the positions of every strip the CTA
// walks; at the end the position groups are summed through shared memory and added to dW with fp32 atomics.
// The bias gradient is one more slice whose A rows all point at a constant row (1, 0, ..., 0): its accumulator row 0 is
// the column sum of dY.
// smem: [slice offsets 2*MT u32][ones row 16 B][X planes][dY planes]  (reduction scratch aliases the planes)
// ---------------------------------------------------------------------------------------------
struct WsWgP {
  int N, H, W, Cin, Cout;
  int KH, KW, pad, Ho, Wo;
  int P, R, RS, CG, COG, E, ET, MT, NT;   // ET = E + 1 (bias slice) when db is wanted
  int planeX, planeY;
  int strips_per_img, total_strips;
  int x_dt, dy_dt;
  int gact, ga_dt; float galpha;          // dy := dy * act'(ya)
  int mblocks, nblocks, kgroups;          // warp grid: (m block, n block) x position groups
  int want_db;
  int iprX, iprY;
  uint32_t magicCG, magicCOG, magic_iprX, magic_iprY, magic_spi;
  int S, stage_bytes, raw_dy_off, raw_a_off;     // raw (TMA) stages: [x rows | dy rows | ya rows]
  uint32_t rowbytes_x, rowbytes_dy, rowbytes_a;
};

template <int MTW, int NTW>
__global__ void __launch_bounds__(WS_THREADS, 2) conv_ws_wgrad_kernel(WsWgP p, const void* __restrict__ x,
                                                                      const void* __restrict__ dy,
                                                                      const void* __restrict__ ya, float* __restrict__ dw,
                                                                      float* __restrict__ db) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint32_t* s_off = reinterpret_cast<uint32_t*>(smem);
  uint8_t* s_ones = smem + ((2 * p.MT * 4 + 127) & ~127);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_ones + 64);
  uint8_t* s_x = s_ones + 128;
  uint8_t* s_y = s_x + (size_t)p.CG * p.planeX * 16;
  uint8_t* s_raw = s_y + (size_t)p.COG * p.planeY * 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr uint32_t ONES = 0xFFFFFFFFu;

  for (int e = tid; e < 2 * p.MT; e += WS_THREADS) {
    uint32_t off;
    if (e < p.E) {
      const int q = e % p.KW;
      const int tt = e / p.KW;
      const int r = tt % p.KH;
      const int cg = tt / p.KH;
      off = (uint32_t)((cg * p.planeX + r * p.P + q) * 16);
    } else if (e == p.E && p.want_db) {
      off = ONES;
    } else {
      off = 0;                                   // padding slice: its rows are never written out
    }
    s_off[e] = off;
  }
  if (tid < 8) reinterpret_cast<__nv_bfloat16*>(s_ones)[tid] = __float2bfloat16_rn(tid == 0 ? 1.f : 0.f);
  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) mbar_init(s_bar + i, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < p.CG * p.planeX + p.COG * p.planeY; i += WS_THREADS) reinterpret_cast<uint4*>(s_x)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  // this warp's accumulator block and position group
  const int blocks = p.mblocks * p.nblocks;
  const int blk = warp % blocks, kg = warp / blocks;
  const bool active = kg < p.kgroups;
  const int mt0 = (blk / p.nblocks) * MTW, nt0 = (blk % p.nblocks) * NTW;
  float acc[MTW][NTW][4];
#pragma unroll
  for (int a = 0; a < MTW; ++a)
#pragma unroll
    for (int b = 0; b < NTW; ++b)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[a][b][i] = 0.f;

  const uint32_t sx_addr = (uint32_t)__cvta_generic_to_shared(s_x);
  const uint32_t sy_addr = (uint32_t)__cvta_generic_to_shared(s_y);
  const uint32_t ones_addr = (uint32_t)__cvta_generic_to_shared(s_ones);
  const int lm = lane >> 3, lr = lane & 7;
  // A (16 x 16) = [slice 2mt ch 0-7 | slice 2mt+1 ch 0-7] x 16 positions, stored as rows = positions: .trans.
  // matrices: 0 = (first slice, pos 0-7), 1 = (second slice, pos 0-7), 2 = (first, pos 8-15), 3 = (second, pos 8-15)
  const int a_sel = lm & 1;
  const uint32_t a_row = (uint32_t)(((lm >> 1) * 8 + lr) * 16);
  // B (16 x 8) per n-tile: matrices 0 = pos 0-7, 1 = pos 8-15 of channel group og (lanes 16-31: addresses ignored)
  const uint32_t b_row = (uint32_t)((((lm & 1) * 8) + lr) * 16);
  // per-lane A addresses of this warp's m-tiles (ONES: the constant row) and B plane addresses, fixed for the whole kernel
  uint32_t a_off[MTW];
  bool a_ones[MTW];
#pragma unroll
  for (int ma = 0; ma < MTW; ++ma) {
    const int mt = min(mt0 + ma, p.MT - 1);
    const uint32_t off = s_off[2 * mt + a_sel];
    a_ones[ma] = off == ONES;
    a_off[ma] = a_ones[ma] ? ones_addr : sx_addr + off + a_row;
  }
  uint32_t b_off[NTW];
#pragma unroll
  for (int nb = 0; nb < NTW; ++nb) b_off[nb] = sy_addr + (uint32_t)(min(nt0 + nb, p.COG - 1) * p.planeY * 16) + b_row;
  const int MT = p.MT, P = p.P, R = p.R, Ho = p.Ho, spi = p.strips_per_img, kgroups = p.kgroups;

  // ---- TMA pipeline: strip i of this CTA lives in raw stage i % S ([x rows | dy rows | ya rows])
  const int my_strips = (int)blockIdx.x < p.total_strips ? (p.total_strips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int S = p.S;
  const int64_t img_x = (int64_t)p.H * p.rowbytes_x, img_dy = (int64_t)Ho * p.rowbytes_dy, img_a = (int64_t)Ho * p.rowbytes_a;
  auto issue = [&](int i) {
    const int s2 = (int)blockIdx.x + i * (int)gridDim.x;
    const int n2 = ws_div(s2, p.magic_spi, spi);
    const int yy = (s2 - n2 * spi) * R;
    uint64_t* bar = s_bar + (i % S);
    uint8_t* raw = s_raw + (size_t)(i % S) * p.stage_bytes;
    uint32_t bytes = ws_fetch_bytes(p.H, p.rowbytes_x, yy - p.pad, p.RS) + ws_fetch_bytes(Ho, p.rowbytes_dy, yy, R);
    if (p.gact != DAFK_ACT_NONE) bytes += ws_fetch_bytes(Ho, p.rowbytes_a, yy, R);
    mbar_expect_tx(bar, bytes);
    ws_fetch(raw, x, n2 * img_x, p.H, p.rowbytes_x, yy - p.pad, p.RS, bar);
    ws_fetch(raw + p.raw_dy_off, dy, n2 * img_dy, Ho, p.rowbytes_dy, yy, R, bar);
    if (p.gact != DAFK_ACT_NONE) ws_fetch(raw + p.raw_a_off, ya, n2 * img_a, Ho, p.rowbytes_a, yy, R, bar);
  };
  if (tid == 0)
    for (int i = 0; i < S && i < my_strips; ++i) issue(i);

  for (int i = 0; i < my_strips; ++i) {
    const int s = (int)blockIdx.x + i * (int)gridDim.x;
    const int n = ws_div(s, p.magic_spi, spi);
    const int y0 = (s - n * spi) * R;
    const int rows_here = min(R, Ho - y0);
    const int st = i % S;
    mbar_wait(s_bar + st, (uint32_t)(i / S) & 1u);
    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw + (size_t)st * p.stage_bytes);
    ws_convert_any(raw_addr, p.x_dt, 0u, 0, DAFK_ACT_NONE, 0.f, s_x, p.planeX, p.H, p.W, p.Cin, p.CG, p.magicCG, y0 - p.pad, p.RS, P,
                   p.pad, p.iprX, p.magic_iprX, warp, lane);
    // dY rows sit at raster column 0 .. Wo-1 (columns Wo .. P-1 keep the zeros of the setup: halo outputs contribute nothing);
    // rows past the image bottom are written as zeros by the conversion routine
    ws_convert_any(raw_addr + (uint32_t)p.raw_dy_off, p.dy_dt, raw_addr + (uint32_t)p.raw_a_off, p.ga_dt, p.gact, p.galpha, s_y,
                   p.planeY, Ho, p.Wo, p.Cout, p.COG, p.magicCOG, y0, R, P, 0, p.iprY, p.magic_iprY, warp, lane);
    __syncthreads();                       // rasters complete; the raw stage is free again
    if (tid == 0 && i + S < my_strips) {
      fence_proxy_async();
      issue(i + S);
    }
    if (active) {
      const int chunks = (rows_here * P + 15) >> 4;
      for (int c = kg; c < chunks; c += kgroups) {
        const uint32_t coff = (uint32_t)(c * 256);
        uint32_t b[NTW][2];
#pragma unroll
        for (int nb = 0; nb < NTW; ++nb) ws_ldsm2_t(b_off[nb] + coff, b[nb]);
#pragma unroll
        for (int ma = 0; ma < MTW; ++ma) {
          if (mt0 + ma < MT) {
            uint32_t a[4];
            ws_ldsm4_t(a_ones[ma] ? a_off[ma] : a_off[ma] + coff, a);
            // ldmatrix.trans register order (m0: first slice / pos 0-7, m1: second slice / pos 0-7, m2, m3: pos 8-15)
            // = mma A order (a0: rows 0-7 k 0-7, a1: rows 8-15 k 0-7, a2: rows 0-7 k 8-15, a3: rows 8-15 k 8-15)
#pragma unroll
            for (int nb = 0; nb < NTW; ++nb) ws_mma(acc[ma][nb], a, b[nb][0], b[nb][1]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- reduce the position groups through shared memory (aliases the raster), then atomics into dW / db
  float* scratch = reinterpret_cast<float*>(s_x);       // [MT*16][NT*8]
  const int ncols = p.NT * 8;
  for (int k = 0; k < p.kgroups; ++k) {
    if (active && kg == k) {
#pragma unroll
      for (int ma = 0; ma < MTW; ++ma) {
        const int mt = mt0 + ma;
        if (mt >= p.MT) continue;
#pragma unroll
        for (int nb = 0; nb < NTW; ++nb) {
          const int nt = nt0 + nb;
          if (nt >= p.NT) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = mt * 16 + g + 8 * (i >> 1), col = nt * 8 + 2 * t + (i & 1);
            float* d = scratch + row * ncols + col;
            *d = (k == 0 ? 0.f : *d) + acc[ma][nb][i];
          }
        }
      }
    }
    __syncthreads();
  }
  const int rows_total = p.MT * 16;
  for (int i = tid; i < rows_total * ncols; i += WS_THREADS) {
    const int row = i / ncols, co = i - row * ncols;
    const int e = row >> 3, c = row & 7;
    if (co >= p.Cout) continue;
    const float v = scratch[i];
    if (e < p.E) {
      const int q = e % p.KW;
      const int tt = e / p.KW;
      const int r = tt % p.KH;
      const int cg = tt / p.KH;
      const int ci = cg * 8 + c;
      if (ci < p.Cin) atomicAdd(dw + (((int64_t)r * p.KW + q) * p.Cin + ci) * p.Cout + co, v);
    } else if (e == p.E && p.want_db && c == 0) {
      atomicAdd(db + co, v);
    }
  }
}

static inline int ws_round_up(int a, int b) { return (a + b - 1) / b * b; }
static const size_t kWsSmemMax = 108 * 1024;     // per CTA: at least two CTAs per SM (2 x 109 KB of the 227 KB)

static inline int ws_fwd_mb(int NT) { return NT <= 2 ? 4 : 2; }

static bool ws_fwd_geom(WsP& p, size_t& smem, int& ctas_per_sm) {
  p.P = p.W + 2 * p.pad;
  p.CG = (p.Cin + 7) / 8;
  p.E = p.CG * p.KH * p.KW;
  p.J = (p.E + 1) / 2;
  p.NT = (p.Cout + 7) / 8;
  if (p.NT > 8 || p.CG > 8 || p.P >= 32768 || p.W * p.CG >= 32768) return false;
  if (p.NT > 5) p.NT = 8;                      // instantiated n-tile counts: 1, 2, 3, 4, 5, 8
  const int mb = ws_fwd_mb(p.NT);
  const size_t fixed = (size_t)p.J * p.NT * 256 + (size_t)2 * p.J * 4 + 128 + 512;
  int best = 0;
  double best_cost = 0;
  size_t best_smem = 0;
  int best_occ = 1;
  for (int R = 1; R <= 32; R *= 2) {
    if (R > 1 && R / 2 >= p.Ho) break;
    const int plane = ws_round_up(ws_round_up(R * p.P, 16 * mb) + (p.KH - 1) * p.P + p.KW + 8, 8);
    if (plane * 16 >= (1 << 20)) continue;
    const size_t sm = fixed + (size_t)p.CG * plane * 16;
    if (sm > kWsSmemMax) continue;
    int occ = (int)((220 * 1024) / (sm + 1024));
    if (occ > 4) occ = 4;
    if (occ < 1) continue;
    // bytes moved per output row: (R + KH - 1) / R input rows (the halo rows come from L2, count them at half price);
    // a single resident CTA cannot overlap its staging with anybody's arithmetic
    const double cost = ((double)(R + 0.5 * (p.KH - 1)) / R) * (occ >= 2 ? 1.0 : 1.3);
    if (!best || cost < best_cost) { best = R; best_cost = cost; best_smem = sm; best_occ = occ; }
  }
  if (!best) return false;
  p.R = best;
  p.RS = p.R + p.KH - 1;
  p.plane = ws_round_up(ws_round_up(p.R * p.P, 16 * mb) + (p.KH - 1) * p.P + p.KW + 8, 8);
  p.strips_per_img = (p.Ho + p.R - 1) / p.R;
  p.total_strips = p.N * p.strips_per_img;
  p.ipr = (p.W * p.CG + 32 * WS_U - 1) / (32 * WS_U);
  p.magicP = ws_magic(p.P);
  p.magicCG = ws_magic(p.CG);
  p.magic_ipr = ws_magic(p.ipr);
  p.magic_spi = ws_magic(p.strips_per_img);
  smem = best_smem;
  ctas_per_sm = best_occ;
  return true;
}

struct WsWgCfg { int mtw, ntw; };

static bool ws_wg_pick(int MT, int NT, WsWgCfg& cfg, int& mblocks, int& nblocks) {
  static const WsWgCfg cands[] = {{5, 1}, {5, 2}, {2, 8}, {4, 4}, {4, 3}, {5, 3}};
  int best = -1, best_warps = 0, best_pad = 0;
  for (int i = 0; i < 6; ++i) {
    const int mb = (MT + cands[i].mtw - 1) / cands[i].mtw, nb = (NT + cands[i].ntw - 1) / cands[i].ntw;
    if (mb * nb > WS_WARPS) continue;
    const int kg = WS_WARPS / (mb * nb);
    const int warps = kg * mb * nb;
    const int pad = mb * cands[i].mtw * nb * cands[i].ntw - MT * NT;     // wasted accumulator tiles
    if (best < 0 || warps > best_warps || (warps == best_warps && pad < best_pad)) {
      best = i; best_warps = warps; best_pad = pad;
    }
  }
  if (best < 0) return false;
  cfg = cands[best];
  mblocks = (MT + cfg.mtw - 1) / cfg.mtw;
  nblocks = (NT + cfg.ntw - 1) / cfg.ntw;
  return true;
}

static bool ws_wg_geom(WsWgP& p, size_t& smem, WsWgCfg& cfg, int& occ_out) {
  p.P = p.W + 2 * p.pad;
  p.CG = (p.Cin + 7) / 8;
  p.COG = (p.Cout + 7) / 8;
  p.NT = p.COG;
  p.E = p.CG * p.KH * p.KW;
  p.ET = p.E + (p.want_db ? 1 : 0);
  p.MT = (p.ET + 1) / 2;
  if (p.NT > 8 || p.CG > 8 || p.W * p.CG >= 32768 || (p.W + 2 * p.pad) >= 32768) return false;
  if (!ws_wg_pick(p.MT, p.NT, cfg, p.mblocks, p.nblocks)) return false;
  p.kgroups = WS_WARPS / (p.mblocks * p.nblocks);
  const size_t fixed = (size_t)ws_round_up(2 * p.MT * 4, 128) + 128 + 512;
  const size_t scratch = (size_t)p.MT * 16 * p.NT * 8 * 4;
  int best = 0;
  size_t best_smem = 0;
  int best_occ = 1;
  double best_cost = 0;
  for (int R = 1; R <= 32; R *= 2) {
    if (R > 1 && R / 2 >= p.Ho) break;
    const int kpos = ws_round_up(R * p.P, 16);
    const int planeX = ws_round_up(kpos + (p.KH - 1) * p.P + p.KW + 8, 8);
    const int planeY = ws_round_up(kpos + 8, 8);
    size_t sm = (size_t)p.CG * planeX * 16 + (size_t)p.COG * planeY * 16;
    if (sm < scratch) sm = scratch;
    sm += fixed;
    if (sm > kWsSmemMax) continue;
    int occ = (int)((220 * 1024) / (sm + 1024));
    if (occ > 3) occ = 3;
    const double cost = ((double)(R + 0.5 * (p.KH - 1)) / R) * (occ >= 2 ? 1.0 : 1.3);
    if (!best || cost < best_cost) { best = R; best_cost = cost; best_smem = sm; best_occ = occ; }
  }
  if (!best) return false;
  p.R = best;
  p.RS = p.R + p.KH - 1;
  const int kpos = ws_round_up(p.R * p.P, 16);
  p.planeX = ws_round_up(kpos + (p.KH - 1) * p.P + p.KW + 8, 8);
  p.planeY = ws_round_up(kpos + 8, 8);
  p.strips_per_img = (p.Ho + p.R - 1) / p.R;
  p.total_strips = p.N * p.strips_per_img;
  p.iprX = (p.W * p.CG + 32 * WS_U - 1) / (32 * WS_U);
  p.iprY = (p.Wo * p.COG + 32 * WS_U - 1) / (32 * WS_U);
  p.magicCG = ws_magic(p.CG);
  p.magicCOG = ws_magic(p.COG);
  p.magic_iprX = ws_magic(p.iprX);
  p.magic_iprY = ws_magic(p.iprY);
  p.magic_spi = ws_magic(p.strips_per_img);
  smem = best_smem;
  occ_out = best_occ;
  return true;
}

template <typename K>
static int ws_set_smem(K kernel, const char* name) {
  static std::mutex mu;
  static const void* done[64];
  static int ndone = 0;
  std::lock_guard<std::mutex> lk(mu);
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < ndone; ++i)
    if (done[i] == key) return DAFK_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemMax + 1024);
  DAFK_REQUIRE(e == cudaSuccess, DAFK_ERR_CUDA, "%s: cudaFuncSetAttribute failed: %s", name, cudaGetErrorString(e));
  if (ndone < 64) done[ndone++] = key;
  return DAFK_OK;
}

template <int NT, int MB, int JR>
static int ws_launch_fwd(const WsP& p, size_t smem, int grid, const void* x, const void* ya, const float* w, const float* scale,
                         const float* bias, void* y, cudaStream_t s) {
  int rc = ws_set_smem(conv_ws_fwd_kernel<NT, MB, JR>, "dafk_conv_ws_fwd");
  if (rc) return rc;
  conv_ws_fwd_kernel<NT, MB, JR><<<grid, WS_THREADS, smem, s>>>(p, x, ya, w, scale, bias, y);
  return check_launch("dafk_conv_ws_fwd");
}

template <int MTW, int NTW>
static int ws_launch_wg(const WsWgP& p, size_t smem, int grid, const void* x, const void* dy, const void* ya, float* dw,
                        float* db, cudaStream_t s) {
  int rc = ws_set_smem(conv_ws_wgrad_kernel<MTW, NTW>, "dafk_conv_ws_wgrad");
  if (rc) return rc;
  conv_ws_wgrad_kernel<MTW, NTW><<<grid, WS_THREADS, smem, s>>>(p, x, dy, ya, dw, db);
  return check_launch("dafk_conv_ws_wgrad");
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_conv_ws_supported(int Cin, int Cout, int KH, int KW, int W, int pad, int kind) {
  if (Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || W <= 0 || pad < 0) return 0;
  if (kind == 2) {
    WsWgP p{};
    p.H = p.Ho = 1 << 20; p.W = W; p.Wo = W + 2 * pad - KW + 1; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
    p.want_db = 1;
    size_t smem; WsWgCfg cfg; int occ;
    return (p.Wo > 0 && ws_wg_geom(p, smem, cfg, occ)) ? 1 : 0;
  }
  WsP p{};
  p.H = p.Ho = 1 << 20; p.W = W; p.Wo = W + 2 * pad - KW + 1; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  size_t smem; int occ;
  return (p.Wo > 0 && ws_fwd_geom(p, smem, occ)) ? 1 : 0;
}

int dafk_conv_ws_fwd(const void* x, int x_dt, const void* ya, int ya_dt, int gact, float galpha, const float* w_hwio,
                     int wCin, int wCout, int mode, const float* scale, const float* bias, void* y, int y_dt, int N, int H,
                     int W, int Cin, int Cout, int KH, int KW, int pad, int act, float alpha, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && pad >= 0, DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_fwd: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && w_hwio && y, DAFK_ERR_BAD_ARG, "dafk_conv_ws_fwd: null pointer");
  DAFK_REQUIRE((x_dt == DAFK_F32 || x_dt == DAFK_BF16) && (y_dt == DAFK_F32 || y_dt == DAFK_BF16), DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_fwd: bad dtype");
  DAFK_REQUIRE(mode == 0 || mode == 1, DAFK_ERR_BAD_ARG, "dafk_conv_ws_fwd: mode must be 0 (forward) or 1 (data gradient)");
  DAFK_REQUIRE(mode == 0 ? (wCin == Cin && wCout == Cout) : (wCin == Cout && wCout == Cin), DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_fwd: weight tensor [%d,%d] does not match Cin=%d Cout=%d mode=%d", wCin, wCout, Cin, Cout, mode);
  DAFK_REQUIRE(gact == DAFK_ACT_NONE || (ya != nullptr && (ya_dt == DAFK_F32 || ya_dt == DAFK_BF16)), DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_fwd: the fused activation backward needs the activation output");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(y) && DAFK_ALIGNED16(ya), DAFK_ERR_ALIGN,
               "dafk_conv_ws_fwd: pointers must be 16-byte aligned");
  WsP p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  p.Ho = H + 2 * pad - KH + 1;
  p.Wo = W + 2 * pad - KW + 1;
  DAFK_REQUIRE(p.Ho > 0 && p.Wo > 0, DAFK_ERR_BAD_ARG, "dafk_conv_ws_fwd: empty output");
  p.mode = mode; p.wCin = wCin; p.wCout = wCout; p.act = act; p.alpha = alpha; p.x_dt = x_dt; p.y_dt = y_dt;
  p.gact = gact; p.ga_dt = ya_dt; p.galpha = galpha;
  size_t smem; int occ;
  DAFK_REQUIRE(ws_fwd_geom(p, smem, occ), DAFK_ERR_UNSUPPORTED,
               "dafk_conv_ws_fwd: geometry does not fit (Cin=%d Cout=%d k=%dx%d W=%d)", Cin, Cout, KH, KW, W);
  int grid = kNumSMs * occ;
  if (grid > p.total_strips) grid = p.total_strips;
  cudaStream_t s = as_stream(stream);
  switch (p.NT) {
    case 1:
      if (p.J <= 5) return ws_launch_fwd<1, 4, 5>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
      return ws_launch_fwd<1, 4, 0>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
    case 2:
      if (p.J <= 5) return ws_launch_fwd<2, 4, 5>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
      return ws_launch_fwd<2, 4, 0>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
    case 3: return ws_launch_fwd<3, 2, 0>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
    case 4: return ws_launch_fwd<4, 2, 0>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
    case 5: return ws_launch_fwd<5, 2, 0>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
    case 8: return ws_launch_fwd<8, 2, 0>(p, smem, grid, x, ya, w_hwio, scale, bias, y, s);
  }
  set_error("dafk_conv_ws_fwd: unsupported Cout %d", Cout);
  return DAFK_ERR_UNSUPPORTED;
}

int dafk_conv_ws_wgrad(const void* x, int x_dt, const void* dy, int dy_dt, const void* ya, int ya_dt, int gact, float galpha,
                       float* dw, float* db, int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && pad >= 0, DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_wgrad: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_conv_ws_wgrad: null pointer");
  DAFK_REQUIRE((x_dt == DAFK_F32 || x_dt == DAFK_BF16) && (dy_dt == DAFK_F32 || dy_dt == DAFK_BF16), DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_wgrad: bad dtype");
  DAFK_REQUIRE(gact == DAFK_ACT_NONE || (ya != nullptr && (ya_dt == DAFK_F32 || ya_dt == DAFK_BF16)), DAFK_ERR_BAD_ARG,
               "dafk_conv_ws_wgrad: the fused activation backward needs the activation output");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(ya), DAFK_ERR_ALIGN,
               "dafk_conv_ws_wgrad: pointers must be 16-byte aligned");
  WsWgP p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  p.Ho = H + 2 * pad - KH + 1;
  p.Wo = W + 2 * pad - KW + 1;
  DAFK_REQUIRE(p.Ho > 0 && p.Wo > 0, DAFK_ERR_BAD_ARG, "dafk_conv_ws_wgrad: empty output");
  p.x_dt = x_dt; p.dy_dt = dy_dt; p.gact = gact; p.ga_dt = ya_dt; p.galpha = galpha;
  p.want_db = db != nullptr ? 1 : 0;
  size_t smem; WsWgCfg cfg; int occ;
  DAFK_REQUIRE(ws_wg_geom(p, smem, cfg, occ), DAFK_ERR_UNSUPPORTED,
               "dafk_conv_ws_wgrad: geometry does not fit (Cin=%d Cout=%d k=%dx%d W=%d)", Cin, Cout, KH, KW, W);
  int grid = kNumSMs * occ;
  if (grid > p.total_strips) grid = p.total_strips;
  cudaStream_t s = as_stream(stream);
  if (cfg.mtw == 5 && cfg.ntw == 1) return ws_launch_wg<5, 1>(p, smem, grid, x, dy, ya, dw, db, s);
  if (cfg.mtw == 5 && cfg.ntw == 2) return ws_launch_wg<5, 2>(p, smem, grid, x, dy, ya, dw, db, s);
  if (cfg.mtw == 2 && cfg.ntw == 8) return ws_launch_wg<2, 8>(p, smem, grid, x, dy, ya, dw, db, s);
  if (cfg.mtw == 4 && cfg.ntw == 4) return ws_launch_wg<4, 4>(p, smem, grid, x, dy, ya, dw, db, s);
  if (cfg.mtw == 4 && cfg.ntw == 3) return ws_launch_wg<4, 3>(p, smem, grid, x, dy, ya, dw, db, s);
  return ws_launch_wg<5, 3>(p, smem, grid, x, dy, ya, dw, db, s);
}

}  // extern "C"
