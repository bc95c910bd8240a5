// Narrow-channel convolutions on tcgen05 ("raster-strip implicit GEMM") for sm_100a.
//
// The FiLM decoder's 8->8 3x3 stack (model_components/decoder.py:44-54), the first layers of the segmentor,
// UNet and discriminators (8->64, 1->64), and the locnet's 5x5 layers (layers/stn_spline.py:106-112) have so
// few input channels that one pixel is only 16..48 bytes of bf16.  A 128B-swizzled TMA box would be mostly
// padding, and CUDA-core direct convolution is FMA/LDS-bound at ~5% of the machine.  Here the tensor cores
// are fed from an UN-SWIZZLED raster copy of the input strip instead:
//
//   * a CTA stages R + KH - 1 input rows of one image into shared memory as channel-group planes
//     [cg][row * P + col][8 ch bf16 = 16 B], P = W + 2*pad (the zero halo columns are part of the raster);
//     any dtype conversion (fp32 -> bf16) happens in this staging step, each input byte is read once;
//   * the GEMM M index is the raster position itself: rows of the A operand for filter tap (r,q) are the
//     raster shifted by (r*P + q) positions, i.e. the SAME shared memory with a different descriptor start
//     address -- im2col is never materialised.  8 consecutive positions x 16 B form exactly one un-swizzled
//     UMMA core matrix, so one tcgen05.mma (K = 16) consumes two (tap, channel-group) slices at once, the
//     K-half stride (descriptor LBO) being the byte distance between the two slices;
//   * outputs that fall on halo columns are computed and discarded (2/P of the tile).
//
//   forward / stride-1 data gradient:  D[pos, co] = sum_e A_e[pos, 0:8] . Wp[e][co][0:8]      (M=128, N=Cout)
//   weight gradient:  dW[r][q][ci][co] = sum_pos X[pos + r*P + q][ci] . dY[pos][co]; X is the MN-major A
//       operand whose 8 "M groups" are the taps q = 0..7 of one filter row (group stride = 16 B = one
//       position), dY the MN-major B operand, K = positions; accumulated in TMEM over all strips of a CTA.
#include "tc_ptx.cuh"
#include <atomic>
#include <mutex>

namespace dafk {


struct NcFwdP {
  int N, H, W, Cin, Cout;
  int KH, KW, pad, Ho, Wo;
  int P, R, RS, CG, E, Npad, T, G;   // pitch, rows/strip, input rows/strip, channel groups, K entries (even),
                                     // padded Cout, 128-position tiles per strip, tiles per TMEM group
  int plane;                         // positions per channel-group plane
  int S;                             // raster stages in the producer -> MMA ring
  int strips_per_img, total_strips;
  int x_dt, y_dt, act;
  float alpha;
  int dbg;
  int Ca, Cb;                        // two concatenated sources (Cb > 0): channels [0,Ca) from x, [Ca,Ca+Cb) from xb; Ca % 8 == 0
  int raw_slots, raw_slot_bytes;     // RAW mode: ring of row segments filled by the TMA engine (0 slots = not used)
  int seg_px, nseg;                  //           pixels per segment, segments per image row
};

struct NcWgP {
  int N, H, W, Cin, Cout;
  int KH, KW, pad, Ho, Wo;
  int P, R, RS, CG, COG, N8;         // N8 = Cout rounded up to 8 (UMMA N)
  int planeX, planeY, chunks;        // positions per plane; K chunks of 16 positions per strip
  int S;                             // stages in the producer -> MMA ring
  int strips_per_img, total_strips;
  int x_dt, dy_dt;
  int tmem_cols;
  int fold;                          // Cout <= 8: the KH filter rows are folded into the MMA's N dimension
  int yoff;                          // fold: positions of zero halo in front of the dY rows = (KH-1)*P
  int Ca, Cb;                        // two concatenated input sources, as in NcFwdP
  int raw_slots, raw_slot_bytes;     // RAW mode: ring of row segments (X rows and dY rows share it)
  int seg_px, nseg, seg_px_y, nseg_y;
};

// load 8 consecutive channels starting at p (nvalid of them exist) as floats
template <typename T>
__device__ __forceinline__ void nc_load8(const T* __restrict__ p, int nvalid, bool vec, float (&v)[8]);

template <>
__device__ __forceinline__ void nc_load8<float>(const float* __restrict__ p, int nvalid, bool vec, float (&v)[8]) {
  if (vec) {   // channel count multiple of 4 and base 16B aligned
    if (nvalid >= 4) {
      float4 a = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else { v[0] = v[1] = v[2] = v[3] = 0.f; }
    if (nvalid >= 8) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else { v[4] = v[5] = v[6] = v[7] = 0.f; }
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < nvalid ? __ldg(p + c) : 0.f;
  }
}
template <>
__device__ __forceinline__ void nc_load8<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, int nvalid, bool vec,
                                                         float (&v)[8]) {
  if (vec && nvalid >= 8) {   // channel count multiple of 8
    uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] = __low2float(h);
      v[2 * i + 1] = __high2float(h);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < nvalid ? __bfloat162float(p[c]) : 0.f;
  }
}

__device__ __forceinline__ uint4 nc_pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float nc_act(float z, int act, float alpha) {
  if (act == DAFK_ACT_RELU) return z > 0.f ? z : 0.f;
  if (act == DAFK_ACT_LRELU) return z > 0.f ? z : alpha * z;
  if (act == DAFK_ACT_TANH) return tanhf(z);
  return z;
}

constexpr int NC_PROD = 256;     // producer threads (warps 0..7)
constexpr int NC_U_FWD = 10;     // (pixel, channel-group) units in flight per producer thread, forward kernel
constexpr int NC_U_WG = 6;       // same, weight-gradient kernel (two tensors are staged per strip)

// Stage `rows` image rows [iy0, iy0+rows) of image n into the raster planes.  Unit of work = (row, pixel, channel
// group): 8 channels = 16..32 contiguous bytes in global memory, one 16 B shared-memory store.  The units of the whole
// strip are flattened row-major, so consecutive producer threads read consecutive global addresses; each thread
// issues NC_U independent loads before it converts the first one (one latency exposure per NC_U units).
// `col0` is the raster column of image column 0 (= pad for inputs, 0 for output gradients).  With BSUM the
// per-thread channel sums are accumulated (bias gradient): every thread then stays on ONE channel group,
// which needs the thread count to be a multiple of `groups` (nthr = (NC_PROD / groups) * groups).
// Two concatenated sources (srcb != nullptr; the locnet's Concatenate([s1, s2]), layers/stn_spline.py:104): the first
// `C` channels (a whole number of 8-channel groups) come from src, the other Cb from srcb -- the concatenation is never
// materialised, a channel group simply picks its base pointer.
template <typename T, bool BSUM, int NC_U>
__device__ __forceinline__ void nc_stage_rows(const T* __restrict__ src, uint8_t* planes, int plane_pos, int n, int Himg,
                                              int wcols, int C, int groups, int iy0, int rows, int P, int col0, int ptid,
                                              int nthr, float (&bsum)[8], const T* __restrict__ srcb = nullptr, int Cb = 0) {
  if (ptid >= nthr) return;
  const int ga = srcb != nullptr ? C >> 3 : groups;          // channel groups that live in the first source
  const bool vec = (sizeof(T) == 4) ? (((C | Cb) & 3) == 0) : (((C | Cb) & 7) == 0);
  const T* imgb = srcb != nullptr ? srcb + (int64_t)n * Himg * wcols * Cb : nullptr;
  const int units_row = wcols * groups;
  const int total = rows * units_row;
  const int gshift = (groups & (groups - 1)) == 0 ? 31 - __clz(groups) : -1;
  const uint32_t magic_row = (uint32_t)((0x100000000ULL + (uint64_t)units_row - 1) / (uint64_t)units_row);
  const T* img = src + (int64_t)n * Himg * wcols * C;
  for (int u0 = ptid; u0 < total; u0 += nthr * NC_U) {
    float v[NC_U][8];
    int dst[NC_U];
#pragma unroll
    for (int k = 0; k < NC_U; ++k) {
      const int u = u0 + k * nthr;
      dst[k] = -1;
      if (u < total) {
        const int row = (int)__umulhi((uint32_t)u, magic_row);
        const int ur = u - row * units_row;
        int cg, px;
        if (gshift >= 0) { cg = ur & (groups - 1); px = ur >> gshift; }
        else { px = ur / groups; cg = ur - px * groups; }
        const int iy = iy0 + row;
        dst[k] = cg * plane_pos + row * P + col0 + px;
        if (iy >= 0 && iy < Himg) {
          if (cg < ga) nc_load8<T>(img + ((int64_t)iy * wcols + px) * C + cg * 8, min(8, C - cg * 8), vec, v[k]);
          else nc_load8<T>(imgb + ((int64_t)iy * wcols + px) * Cb + (cg - ga) * 8, min(8, Cb - (cg - ga) * 8), vec, v[k]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[k][c] = 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NC_U; ++k) {
      if (dst[k] >= 0) {
        *reinterpret_cast<uint4*>(planes + (size_t)dst[k] * 16) = nc_pack8(v[k]);
        if (BSUM) {
#pragma unroll
          for (int c = 0; c < 8; ++c) bsum[c] += v[k][c];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RAW staging: the rows of a strip are brought into shared memory AS THEY LIE in global memory (fp32 or bf16, C channels
// per pixel) by 1-D cp.async.bulk copies, one per row segment, into a ring of slots; converter warps turn each segment
// into bf16 raster units (smem -> smem).  The register-staged path above keeps at most NC_U loads per thread in flight and
// pays a full DRAM latency per batch (8 -> 8 @ 192 x 224^2: 4.1 TB/s forward, 2.4 TB/s weight gradient); here the bytes in
// flight are the ring (>= 8 segments of up to 8 KB per SM, issued by one lane, no registers), and the issuer runs ahead of
// the converters across strip boundaries.  A segment = seg_px pixels of one image row (a whole row when it fits 8 KB).
// Both sides walk the same (row, segment) sequence; rows outside the image are not copied, the converters zero them.
// ---------------------------------------------------------------------------------------------
constexpr int NC_RAW_MAX_SLOTS = 12;     // a multiple of the converter warps, see nc_raw_convert
constexpr int NC_RAW_SEG_MAX = 8192;
constexpr int NC_RAW_CONV = 192;      // converter threads (six warps)

struct NcRing {
  uint8_t* base;
  uint64_t* full;     // [slots] count 1 + transaction bytes
  uint64_t* empty;    // [slots] count 1: the converter warp that owns the segment
  int slots, slot_bytes;
  int slot;           // position of this thread's walk
  uint32_t ph;
  int turn;           // converter warp that owns the segment at `slot` (segments go round-robin over the warps)
};

template <typename T>
__device__ __forceinline__ void nc_raw_issue(NcRing& r, const T* __restrict__ src, int n, int Himg, int wcols, int C, int iy0,
                                             int rows, int seg_px, int nseg) {
  const size_t px_bytes = (size_t)C * sizeof(T);
  const size_t row_bytes = (size_t)wcols * px_bytes;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(src) + (size_t)n * Himg * row_bytes;
  for (int row = 0; row < rows; ++row) {
    const int iy = iy0 + row;
    if (iy < 0 || iy >= Himg) continue;
    for (int sg = 0; sg < nseg; ++sg) {
      const int px0 = sg * seg_px;
      const uint32_t bytes = (uint32_t)(min(seg_px, wcols - px0) * px_bytes);
      mbar_wait(r.empty + r.slot, r.ph ^ 1u);
      mbar_expect_tx(r.full + r.slot, bytes);
      bulk_load_1d(r.base + (size_t)r.slot * r.slot_bytes, img + (size_t)iy * row_bytes + (size_t)px0 * px_bytes, bytes,
                   r.full + r.slot);
      if (++r.slot == r.slots) { r.slot = 0; r.ph ^= 1u; }
    }
  }
}

// 8 channels of a staged pixel (shared memory) as floats; bf16 rows need C % 8 == 0, fp32 rows take any C (vector reads
// when C % 4 == 0: a group is then whole or holds exactly 4 valid channels)
__device__ __forceinline__ void nc_raw_load8(const float* p, int nvalid, bool vec, float (&v)[8]) {
  if (!vec) {        // channel counts that are not multiples of 4 (the 1-channel images): scalar reads
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < nvalid ? p[c] : 0.f;
    return;
  }
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  if (nvalid >= 8) {
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else { v[4] = v[5] = v[6] = v[7] = 0.f; }
}

// Converter side, called by every converter warp (cw = its index among the ncw of them).  A segment belongs to ONE warp
// and, the slot count being a multiple of ncw (nc_raw_slots), so does every slot:
// the warps work on different segments at the same time, and nothing but that warp's arrival frees the slot (a segment
// shared by all warps cost a barrier round trip of all 192 threads per image row: the load path alone took 117 us for
// 8 -> 8 @ 192 x 224^2 against 93 us with register staging).  `lanes` of the warp convert (32, or a multiple of `groups`
// when the per-lane channel sums of the bias gradient need every lane to stay on one channel group).
template <typename T, bool BSUM>
__device__ __forceinline__ void nc_raw_convert(NcRing& r, uint8_t* planes, int plane_pos, int Himg, int wcols, int C, int groups,
                                               int iy0, int rows, int P, int col0, int cw, int ncw, int lanes, int seg_px,
                                               int nseg, int lane, float (&bsum)[8]) {
  const int gshift = (groups & (groups - 1)) == 0 ? 31 - __clz(groups) : -1;
  for (int row = 0; row < rows; ++row) {
    const int iy = iy0 + row;
    uint8_t* prow = planes + (size_t)(row * P + col0) * 16;
    if (iy < 0 || iy >= Himg) {      // the raster of this stage still holds a row of an earlier strip
      for (int u = cw * 32 + lane; u < wcols * groups; u += ncw * 32) {
        int cg, px;
        if (gshift >= 0) { cg = u & (groups - 1); px = u >> gshift; }
        else { px = u / groups; cg = u - px * groups; }
        *reinterpret_cast<uint4*>(prow + ((size_t)cg * plane_pos + px) * 16) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    for (int sg = 0; sg < nseg; ++sg) {
      if (r.turn == cw) {
        const int px0 = sg * seg_px;
        const int units = min(seg_px, wcols - px0) * groups;
        mbar_wait(r.full + r.slot, r.ph);
        const T* raw = reinterpret_cast<const T*>(r.base + (size_t)r.slot * r.slot_bytes);
        if (lane < lanes) {
#pragma unroll 4
          for (int u = lane; u < units; u += lanes) {
            int cg, px;
            if (gshift >= 0) { cg = u & (groups - 1); px = u >> gshift; }
            else { px = u / groups; cg = u - px * groups; }
            uint4 packed;
            if (sizeof(T) == 4) {
              float v[8];
              nc_raw_load8(reinterpret_cast<const float*>(raw) + px * C + cg * 8, C - cg * 8, (C & 3) == 0, v);
              packed = nc_pack8(v);
              if (BSUM) {
#pragma unroll
                for (int c = 0; c < 8; ++c) bsum[c] += v[c];
              }
            } else {
              packed = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(raw) + px * C + cg * 8);
              if (BSUM) {
                const uint32_t w[4] = {packed.x, packed.y, packed.z, packed.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
                  bsum[2 * i] += __low2float(h);
                  bsum[2 * i + 1] += __high2float(h);
                }
              }
            }
            *reinterpret_cast<uint4*>(prow + ((size_t)cg * plane_pos + px0 + px) * 16) = packed;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty + r.slot);
      }
      if (++r.slot == r.slots) { r.slot = 0; r.ph ^= 1u; }
      if (++r.turn == ncw) r.turn = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward / stride-1 data gradient.  Warp-specialised persistent CTA (one per SM):
//   warps 0-7  producers : global -> bf16 raster planes of stage s           (full[s] / empty[s] ring)
//   warp  8    MMA issuer: per 128-position tile E/2 tcgen05.mma into one of two TMEM buffers
//   warps 9-12 epilogue  : TMEM -> registers -> bias/activation -> global    (tfull[b] / tempty[b])
// smem: [wp: E*Npad*16][S stages x CG*plane*16][descA: E/2 u64][descB: E/2 u64][bias: Npad f32][barriers]
// ---------------------------------------------------------------------------------------------
constexpr int NC_FWD_THREADS = NC_PROD + 32 + 128;
constexpr int NC_MAX_STAGES = 4;

//
// BULK (bf16 input with exactly 8 channels, i.e. one pixel = one 16 B raster position): an image row IS a raster row,
// so warp 0 alone stages a strip with one 1-D cp.async.bulk per row (TMA engine, mbarrier byte count; rows outside the
// image are zeroed by the warp), the halo columns keep the zeros of the one-time setup, and the freed warps 4-7 become a
// second epilogue group: group A (warps 9-12) drains TMEM buffer 0, group B (warps 4-7) buffer 1.
// BULK runs 12 warps (three per scheduler, 168 registers each): 0 = copies, 1 = MMA issuer, 2-3 idle, 4-11 epilogue.
//
// MODE 2 (default for staged inputs; DAFK_NC_L12=0 selects MODE 0): MODE 0's role layout has 13 warps, i.e. four on one
// scheduler, which caps every thread at 128 registers (ptxas spills 16 B).  Twelve warps -- 7 producers, MMA issuer,
// 4 epilogue -- get 168: the same code then needs 148 registers and no stack.  Bit-identical outputs; measured on B200
// (32 x 224^2): 8 -> 8 forward 32.0 -> 30.9 us, 64 -> 8 data gradient 344 -> 324 us.
// (A setmaxnreg split into four warpgroups with an eight-tile lock-step epilogue was measured too: slower, 34 / 200 us
// against 32 / 149 us on 8 -> 8 / 8 -> 64, and removed.)
constexpr int NC_BULK_THREADS = 384;
//
// MODE 3 (RAW): MODE 2's layout with the staging split in two -- warp 0 issues the bulk copies of the row segments, warps 1-6
// convert them into the raster (nc_raw_issue / nc_raw_convert above).
constexpr int NC_MODE_DEFAULT = 0, NC_MODE_BULK = 1, NC_MODE_L12 = 2, NC_MODE_RAW = 3;

template <typename TX, int MODE>
__global__ void __launch_bounds__(MODE == NC_MODE_DEFAULT ? NC_FWD_THREADS : NC_BULK_THREADS, 1)
conv_nc_fwd_kernel(NcFwdP p, const TX* __restrict__ x, const __nv_bfloat16* __restrict__ wp,
                   const float* __restrict__ bias, void* __restrict__ y, const TX* __restrict__ xb = nullptr) {
  constexpr bool BULK = MODE == NC_MODE_BULK;
  constexpr bool RAW = MODE == NC_MODE_RAW;
  constexpr int THREADS = MODE == NC_MODE_DEFAULT ? NC_FWD_THREADS : NC_BULK_THREADS;
  constexpr int PROD = (MODE == NC_MODE_L12 || RAW) ? 224 : NC_PROD;         // staging threads (not BULK)
  constexpr int MMA_WARP = BULK ? 1 : PROD / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int w_bytes = p.E * p.Npad * 16;
  const int st_bytes = p.CG * p.plane * 16;
  uint8_t* s_w = smem;
  uint8_t* s_x = smem + w_bytes;
  uint64_t* s_descA = reinterpret_cast<uint64_t*>(s_x + (size_t)p.S * st_bytes);
  uint64_t* s_descB = s_descA + p.E / 2;
  float* s_bias = reinterpret_cast<float*>(s_descB + p.E / 2);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_bias + p.Npad);
  uint64_t* empty = full + NC_MAX_STAGES;
  uint64_t* tfull = empty + NC_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* rfull = reinterpret_cast<uint64_t*>(tmem_slot + 2);            // RAW: segment ring barriers + ring
  uint64_t* rempty = rfull + NC_RAW_MAX_SLOTS;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rempty + NC_RAW_MAX_SLOTS) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  // one-time setup: weights, bias, zeroed rasters (halo columns and slack stay zero for the CTA's lifetime)
  for (int i = tid; i < w_bytes / 16; i += THREADS)
    reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(wp) + i);
  for (int i = tid; i < p.S * st_bytes / 16; i += THREADS) reinterpret_cast<uint4*>(s_x)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < p.Npad; i += THREADS) s_bias[i] = (bias != nullptr && i < p.Cout) ? bias[i] : 0.f;
  const int Ereal = p.CG * p.KH * p.KW;
  for (int j = tid; j < p.E / 2; j += THREADS) {
    auto off = [&](int e) {
      int q = e % p.KW;
      int t = e / p.KW;
      int r = t % p.KH;
      int cg = t / p.KH;
      return (uint32_t)((cg * p.plane + r * p.P + q) * 16);
    };
    const uint32_t o0 = off(2 * j);
    const uint32_t lbo = (2 * j + 1 < Ereal) ? off(2 * j + 1) - o0 : 16u;
    // K-major, un-swizzled: LBO = distance between the two K halves (the two (tap, group) slices),
    // SBO = distance between 8-row groups (8 raster positions / 8 output channels = 128 B)
    s_descA[j] = make_smem_desc_ns(smem_u32(s_x) + o0, lbo, 128);
    s_descB[j] = make_smem_desc_ns(smem_u32(s_w) + (uint32_t)(2 * j * p.Npad * 16), (uint32_t)p.Npad * 16, 128);
  }
  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) { mbar_init(full + i, BULK ? 1 : (RAW ? NC_RAW_CONV : PROD)); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
    if (RAW)
      for (int i = 0; i < p.raw_slots; ++i) { mbar_init(rfull + i, 1); mbar_init(rempty + i, 1); }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (BULK && warp != MMA_WARP && warp < 4) {
    // ===================== bulk-copy producer (warp 0; warps 2-3 idle) =====================
    if (warp == 0) {
      const uint32_t row_bytes = (uint32_t)p.W * 16u;
      int it = 0;
      for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
        const int st = it % p.S;
        const uint32_t ph = (uint32_t)(it / p.S) & 1u;
        const int n = s / p.strips_per_img;
        const int iy0 = (s - n * p.strips_per_img) * p.R - p.pad;
        uint8_t* planes = s_x + (size_t)st * st_bytes;
        mbar_wait(empty + st, ph ^ 1u);
        // rows outside the image: their shared memory still holds a row of the strip that used this stage before
        int nvalid = 0;
        for (int row = 0; row < p.RS; ++row) {
          const int iy = iy0 + row;
          if (iy < 0 || iy >= p.H) {
            uint4* d = reinterpret_cast<uint4*>(planes + (size_t)(row * p.P + p.pad) * 16);
            for (int i = lane; i < p.W; i += 32) d[i] = make_uint4(0, 0, 0, 0);
          } else {
            ++nvalid;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(full + st, (uint32_t)nvalid * row_bytes);
          const uint8_t* img = reinterpret_cast<const uint8_t*>(x) + (size_t)n * p.H * row_bytes;
          for (int row = 0; row < p.RS; ++row) {
            const int iy = iy0 + row;
            if (iy >= 0 && iy < p.H)
              bulk_load_1d(planes + (size_t)(row * p.P + p.pad) * 16, img + (size_t)iy * row_bytes, row_bytes, full + st);
          }
        }
        __syncwarp();
      }
    }
  } else if (RAW && warp == 0) {
    // ===================== RAW: bulk-copy issuer (one lane walks every strip of this CTA, bounded only by the ring) =====
    if (lane == 0) {
      NcRing r{ring, rfull, rempty, p.raw_slots, p.raw_slot_bytes, 0, 0u, 0};
      for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x) {
        const int n = s / p.strips_per_img;
        const int y0 = (s - n * p.strips_per_img) * p.R;
        nc_raw_issue<TX>(r, x, n, p.H, p.W, p.Cin, y0 - p.pad, p.RS, p.seg_px, p.nseg);
      }
    }
    __syncwarp();
  } else if (RAW && warp < MMA_WARP) {
    // ===================== RAW: converters (warps 1-6) =====================
    float dummy[8];
    NcRing r{ring, rfull, rempty, p.raw_slots, p.raw_slot_bytes, 0, 0u, 0};
    int it = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
      const int st = it % p.S;
      const uint32_t ph = (uint32_t)(it / p.S) & 1u;
      const int n = s / p.strips_per_img;
      const int y0 = (s - n * p.strips_per_img) * p.R;
      mbar_wait(empty + st, ph ^ 1u);
      nc_raw_convert<TX, false>(r, s_x + (size_t)st * st_bytes, p.plane, p.H, p.W, p.Cin, p.CG, y0 - p.pad, p.RS, p.P, p.pad,
                                warp - 1, NC_RAW_CONV / 32, 32, p.seg_px, p.nseg, lane, dummy);
      fence_proxy_async();
      mbar_arrive(full + st);
    }
  } else if (!BULK && warp < MMA_WARP) {
    // ===================== producers =====================
    float dummy[8];
    int it = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
      const int st = it % p.S;
      const uint32_t ph = (uint32_t)(it / p.S) & 1u;
      const int n = s / p.strips_per_img;
      const int y0 = (s - n * p.strips_per_img) * p.R;
      mbar_wait(empty + st, ph ^ 1u);
      if (!(p.dbg & 4))
      nc_stage_rows<TX, false, NC_U_FWD>(x, s_x + (size_t)st * st_bytes, p.plane, n, p.H, p.W, p.Cb > 0 ? p.Ca : p.Cin, p.CG,
                                 y0 - p.pad, p.RS, p.P, p.pad, tid, PROD, dummy, xb, p.Cb);
      fence_proxy_async();
      mbar_arrive(full + st);
    }
  } else if (warp == MMA_WARP) {
    {
    // ===================== MMA issuer =====================
    // warp-uniform control flow (all lanes walk the loops, one elected lane issues) keeps the operands in
    // uniform registers; the (tap, group) descriptor pairs come from the table built above
    const uint32_t idesc = make_idesc(128, p.Npad, 0, 0);
    const uint32_t leader = elect_one();
    const uint32_t tmem_acc = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int nj = (p.dbg & 1) ? 1 : p.E / 2;
    const int G = p.G, Npad = p.Npad, S = p.S;
    int it = 0, gc = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
      const int st = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      const int n = s / p.strips_per_img;
      const int y0 = (s - n * p.strips_per_img) * p.R;
      const int rows_here = min(p.R, p.Ho - y0);
      const int tiles_here = (rows_here * p.P + 127) / 128;
      mbar_wait(full + st, ph);
      tc_fence_after();
      const uint64_t st_off = (uint64_t)((uint32_t)(st * st_bytes) >> 4);
      for (int t0 = 0; t0 < tiles_here; t0 += G, ++gc) {
        const uint32_t b = (uint32_t)gc & 1u;
        mbar_wait(tempty + b, (((uint32_t)gc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const int gt = min(G, tiles_here - t0);
        if (leader) {
          for (int j = 0; j < nj; ++j) {
            // the same (tap, group) pair for every tile of the group: descriptors are loaded once per pair
            const uint64_t da = s_descA[j] + st_off + (uint64_t)(t0 * 128);
            const uint64_t db = s_descB[j];
            const uint32_t acc = j > 0 ? 1u : 0u;
            for (int tt = 0; tt < gt; ++tt)
              umma_bf16(tmem_acc + b * 256u + (uint32_t)(tt * Npad), da + (uint64_t)(tt * 128), db, idesc, acc);
          }
          umma_commit(tfull + b);
        }
        __syncwarp();
      }
      if (leader) umma_commit(empty + st);   // the raster of this stage has been consumed once these MMAs retire
      __syncwarp();
    }
    }
  } else {
    // ===================== epilogue =====================
    const int q4 = warp & 3;
    const int egrp = warp >= 8 ? 0 : 1;   // BULK: which TMEM buffer this epilogue group drains
    const uint32_t magicP = (uint32_t)((0x100000000ULL + (uint64_t)p.P - 1) / (uint64_t)p.P);
    int gc = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x) {
      const int n = s / p.strips_per_img;
      const int y0 = (s - n * p.strips_per_img) * p.R;
      const int rows_here = min(p.R, p.Ho - y0);
      const int tiles_here = (rows_here * p.P + 127) / 128;
      for (int t0 = 0; t0 < tiles_here; t0 += p.G, ++gc) {
        const int b = gc & 1;
        if (BULK && b != egrp) continue;
        mbar_wait(tfull + b, (uint32_t)(gc >> 1) & 1u);
        tc_fence_after();
        if (p.dbg & 8) {          // diagnostic: accumulators are dropped
          tc_fence_before();
          mbar_arrive(tempty + b);
          continue;
        }
        const int gt = min(p.G, tiles_here - t0);
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(b * 256);
        // TMEM loads are issued EB tiles ahead of their use (one wait per batch instead of one per tile)
        constexpr int EB = 4;   // MODE 2: 148 registers, no spills; EB = 6 / 8 spill even at 168
        for (int tb = 0; tb < gt; tb += EB) {
          for (int c0 = 0; c0 < p.Cout; c0 += 8) {
            uint32_t v[EB][8];
#pragma unroll
            for (int e = 0; e < EB; ++e)
              if (tb + e < gt) tmem_ld8(tbase + (uint32_t)((tb + e) * p.Npad + c0), v[e]);
            tmem_ld_wait();
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            // the EB tiles of the batch are processed in lock step, branch-free up to the (predicated) stores: a single
            // epilogue warp per scheduler runs one dependent chain at ~1 instruction per 5 cycles, four independent
            // chains hide that latency
            bool ok[EB];
            int64_t obase[EB];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
              const int m = (t0 + tb + e) * 128 + q4 * 32 + lane;
              const int orow = (int)__umulhi((uint32_t)m, magicP);
              const int ocol = m - orow * p.P;
              ok[e] = (tb + e < gt) && orow < rows_here && ocol < p.Wo && !(p.dbg & 2);
              obase[e] = (((int64_t)n * p.Ho + y0 + orow) * p.Wo + ocol) * p.Cout + c0;
            }
            float f[EB][8];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[e][j] = __uint_as_float(v[e][j]) + bb[j];
            }
            if (p.act == DAFK_ACT_LRELU) {
#pragma unroll
              for (int e = 0; e < EB; ++e)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[e][j] = f[e][j] > 0.f ? f[e][j] : p.alpha * f[e][j];
            } else if (p.act == DAFK_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < EB; ++e)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[e][j] = fmaxf(f[e][j], 0.f);
            } else if (p.act == DAFK_ACT_TANH) {
#pragma unroll
              for (int e = 0; e < EB; ++e)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[e][j] = tanhf(f[e][j]);
            }
            if (p.y_dt == DAFK_F32 && (p.Cout & 3) == 0) {
              const bool second = c0 + 4 < p.Cout;
#pragma unroll
              for (int e = 0; e < EB; ++e) {
                if (ok[e]) {
                  float* o = reinterpret_cast<float*>(y) + obase[e];
                  *reinterpret_cast<float4*>(o) = make_float4(f[e][0], f[e][1], f[e][2], f[e][3]);
                  if (second) *reinterpret_cast<float4*>(o + 4) = make_float4(f[e][4], f[e][5], f[e][6], f[e][7]);
                }
              }
            } else if (p.y_dt == DAFK_BF16 && (p.Cout & 7) == 0) {
#pragma unroll
              for (int e = 0; e < EB; ++e)
                if (ok[e]) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + obase[e]) = nc_pack8(f[e]);
            } else {
#pragma unroll
              for (int e = 0; e < EB; ++e) {
                if (ok[e]) {
                  if (p.y_dt == DAFK_F32) {
                    float* o = reinterpret_cast<float*>(y) + obase[e];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (c0 + j < p.Cout) o[j] = f[e][j];
                  } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(y) + obase[e];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (c0 + j < p.Cout) o[j] = __float2bfloat16_rn(f[e][j]);
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty + b);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient (+ bias gradient).  Warp-specialised persistent CTA:
//   warps 0-7 producers: X rows and dY rows of a strip -> raster planes of stage s;   warp 8: MMA issuer.
//   The KH*CG accumulators [64 x N8] stay in TMEM over all strips of the CTA; the producers drain them at
//   the end with fp32 atomics into the HWIO gradient.
// smem: [S stages x (CG*planeX + COG*planeY)*16][bsum scratch: 128*8 f32][barriers]
// ---------------------------------------------------------------------------------------------
// Role layouts: staged (RAW = false): warps 0-7 producers, warp 8 MMA issuer (288 threads).  RAW: warp 0 issues the bulk
// copies, warps 1-6 convert, warp 7 is the MMA issuer (256 threads = two warps per scheduler).
constexpr int NC_WG_PROD = 256;
constexpr int NC_WG_THREADS = NC_WG_PROD + 32;
constexpr int NC_WG_RAW_THREADS = 256;

template <typename TX, typename TY, bool RAW>
__global__ void __launch_bounds__(RAW ? NC_WG_RAW_THREADS : NC_WG_THREADS, 1)
conv_nc_wgrad_kernel(NcWgP p, const TX* __restrict__ x, const TY* __restrict__ dy, float* __restrict__ dw,
                     float* __restrict__ db, const TX* __restrict__ xb = nullptr) {
  constexpr int THREADS = RAW ? NC_WG_RAW_THREADS : NC_WG_THREADS;
  constexpr int MMA_WARP = RAW ? 7 : NC_WG_PROD / 32;
  constexpr int NSTAGE = RAW ? NC_RAW_CONV : NC_WG_PROD;       // threads that fill the raster
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int x_bytes = p.CG * p.planeX * 16;
  const int y_bytes = p.COG * p.planeY * 16;
  const int st_bytes = x_bytes + y_bytes;
  float* s_bsum = reinterpret_cast<float*>(smem + (size_t)p.S * st_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_bsum + NC_WG_PROD * 8);
  uint64_t* empty = full + NC_MAX_STAGES;
  uint64_t* done = empty + NC_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint64_t* rfull = reinterpret_cast<uint64_t*>(tmem_slot + 2);
  uint64_t* rempty = rfull + NC_RAW_MAX_SLOTS;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rempty + NC_RAW_MAX_SLOTS) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  for (int i = tid; i < p.S * st_bytes / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) { mbar_init(full + i, NSTAGE); mbar_init(empty + i, 1); }
    mbar_init(done, 1);
    if (RAW)
      for (int i = 0; i < p.raw_slots; ++i) { mbar_init(rfull + i, 1); mbar_init(rempty + i, 1); }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool any = (int)blockIdx.x < p.total_strips;
  const int nthr_y = (NSTAGE / p.COG) * p.COG;
  const int raw_lanes_y = (32 / p.COG) * p.COG;     // RAW: lanes of a converter warp that take part in a dY segment (COG <= 32)

  if (warp < MMA_WARP) {
    float bsum[8], dummy[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bsum[c] = 0.f;
    const int stid = RAW ? tid - 32 : tid;          // index among the threads that fill the raster (RAW: warp 0 is the issuer)
    if (RAW && warp == 0) {
      if (lane == 0) {
        NcRing r{ring, rfull, rempty, p.raw_slots, p.raw_slot_bytes, 0, 0u, 0};
        for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x) {
          const int n = s / p.strips_per_img;
          const int y0 = (s - n * p.strips_per_img) * p.R;
          nc_raw_issue<TX>(r, x, n, p.H, p.W, p.Cin, y0 - p.pad, p.RS, p.seg_px, p.nseg);
          nc_raw_issue<TY>(r, dy, n, p.Ho, p.Wo, p.Cout, y0, p.R, p.seg_px_y, p.nseg_y);
        }
      }
      __syncwarp();
    } else if (RAW) {
      NcRing r{ring, rfull, rempty, p.raw_slots, p.raw_slot_bytes, 0, 0u, 0};
      int it = 0;
      for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
        const int st = it % p.S;
        const uint32_t ph = (uint32_t)(it / p.S) & 1u;
        const int n = s / p.strips_per_img;
        const int y0 = (s - n * p.strips_per_img) * p.R;
        uint8_t* sx = smem + (size_t)st * st_bytes;
        uint8_t* sy = sx + x_bytes + (size_t)p.yoff * 16;
        mbar_wait(empty + st, ph ^ 1u);
        nc_raw_convert<TX, false>(r, sx, p.planeX, p.H, p.W, p.Cin, p.CG, y0 - p.pad, p.RS, p.P, p.pad, warp - 1,
                                  NC_RAW_CONV / 32, 32, p.seg_px, p.nseg, lane, dummy);
        if (db != nullptr)
          nc_raw_convert<TY, true>(r, sy, p.planeY, p.Ho, p.Wo, p.Cout, p.COG, y0, p.R, p.P, 0, warp - 1, NC_RAW_CONV / 32,
                                   raw_lanes_y, p.seg_px_y, p.nseg_y, lane, bsum);
        else
          nc_raw_convert<TY, false>(r, sy, p.planeY, p.Ho, p.Wo, p.Cout, p.COG, y0, p.R, p.P, 0, warp - 1, NC_RAW_CONV / 32, 32,
                                    p.seg_px_y, p.nseg_y, lane, dummy);
        fence_proxy_async();
        mbar_arrive(full + st);
      }
    } else {
    int it = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
      const int st = it % p.S;
      const uint32_t ph = (uint32_t)(it / p.S) & 1u;
      const int n = s / p.strips_per_img;
      const int y0 = (s - n * p.strips_per_img) * p.R;
      uint8_t* sx = smem + (size_t)st * st_bytes;
      mbar_wait(empty + st, ph ^ 1u);
      nc_stage_rows<TX, false, NC_U_WG>(x, sx, p.planeX, n, p.H, p.W, p.Cb > 0 ? p.Ca : p.Cin, p.CG, y0 - p.pad, p.RS, p.P, p.pad, tid,
                                        NC_WG_PROD, dummy, xb, p.Cb);
      uint8_t* sy = sx + x_bytes + (size_t)p.yoff * 16;      // fold: (KH-1) zero rows stay in front of (and behind) the dY rows
      if (db != nullptr)
        nc_stage_rows<TY, true, NC_U_WG>(dy, sy, p.planeY, n, p.Ho, p.Wo, p.Cout, p.COG, y0, p.R, p.P, 0, tid, nthr_y, bsum);
      else
        nc_stage_rows<TY, false, NC_U_WG>(dy, sy, p.planeY, n, p.Ho, p.Wo, p.Cout, p.COG, y0, p.R, p.P, 0, tid, NC_WG_PROD, dummy);
      fence_proxy_async();
      mbar_arrive(full + st);
    }
    }
    if (any) {
      // bias gradient: per-thread partial sums -> scratch -> one atomic per channel per CTA
      if (db != nullptr && !(RAW && warp == 0)) {
#pragma unroll
        for (int c = 0; c < 8; ++c) s_bsum[stid * 8 + c] = (RAW ? lane < raw_lanes_y : stid < nthr_y) ? bsum[c] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(NSTAGE) : "memory");
        if (stid < p.Cout) {
          const int cog = stid >> 3, c = stid & 7;
          float acc = 0.f;
          if (RAW) {       // lane l of every converter warp summed channel group l % COG
            for (int wv = 0; wv < NC_RAW_CONV / 32; ++wv)
              for (int l = cog; l < raw_lanes_y; l += p.COG) acc += s_bsum[(wv * 32 + l) * 8 + c];
          } else {
            for (int t = cog; t < nthr_y; t += p.COG) acc += s_bsum[t * 8 + c];
          }
          atomicAdd(db + stid, acc);
        }
      }
      mbar_wait(done, 0);
      tc_fence_after();
      if (warp < 4) {
      // accumulator a = (r, cg) holds rows i = q*8 + c (tap q of filter row r, channel cg*8+c), columns = co.
      // M = 64 accumulators keep rows 16*j .. 16*j+15 in TMEM lanes 32*j .. 32*j+15.
      const int q4 = warp & 3;
      const int i = q4 * 16 + lane;
      const int q = i >> 3, c = i & 7;
      const int nacc = p.KH * p.CG;
      for (int a = 0; a < nacc; ++a) {
        // unfolded: accumulator a = (r, cg), columns = co.  folded (N8 == 8): accumulator cg, column group g holds the
        // filter row r = KH-1-g (the dY raster shifted by g rows)
        int r, cg, col;
        if (p.fold) { cg = a / p.KH; const int gq = a - cg * p.KH; r = p.KH - 1 - gq; col = a * 8; }
        else { r = a / p.CG; cg = a - r * p.CG; col = a * p.N8; }
        const int ci = cg * 8 + c;
        const bool ok = lane < 16 && q < p.KW && ci < p.Cin;
        for (int c0 = 0; c0 < p.N8; c0 += 8) {
          uint32_t v[8];
          tmem_ld8(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(col + c0), v);
          tmem_ld_wait();
          if (ok) {
            float* o = dw + ((int64_t)(r * p.KW + q) * p.Cin + ci) * p.Cout + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < p.Cout) atomicAdd(o + j, __uint_as_float(v[j]));
          }
        }
      }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(64, p.N8, 1, 1);
    const uint32_t leader = elect_one();
    const uint32_t tmem_acc = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int KH = p.KH, CG = p.CG, N8 = p.N8, chunks = p.chunks, S = p.S;
    // MN-major, un-swizzled: LBO = distance between groups of 8 K rows (8 positions = 128 B), SBO = distance
    // between groups of 8 MN elements (dY: the next channel-group plane; X: the next tap of the filter row =
    // the next raster position = 16 B)
    const uint64_t descY = make_smem_desc_ns(smem_u32(smem) + (uint32_t)x_bytes, 128, (uint32_t)p.planeY * 16u);
    const uint64_t descX = make_smem_desc_ns(smem_u32(smem), 128, 16);
    int it = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x, ++it) {
      const int st = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(full + st, ph);
      tc_fence_after();
      if (leader && p.fold) {
        // one accumulator per channel group: M = 64 (8 taps of a filter row x 8 channels), N = KH*8: column group g is
        // the dY raster started g rows LATER in its zero-framed plane (SBO = P*16 B), i.e. dY shifted back by
        // (KH-1-g) rows = filter row r = KH-1-g.  K runs over the RS = R+KH-1 input rows: KH*R/(R+KH-1) fewer MMAs.
        const uint64_t st_off = (uint64_t)((uint32_t)(st * st_bytes) >> 4);
        const uint32_t idf = make_idesc(64, KH * 8, 1, 1);
        const uint64_t dyf = make_smem_desc_ns(smem_u32(smem) + (uint32_t)x_bytes, 128, (uint32_t)p.P * 16u) + st_off;
        for (int cg = 0; cg < CG; ++cg) {
          const uint64_t dx = descX + st_off + (uint64_t)(cg * p.planeX);
          const uint32_t d_col = tmem_acc + (uint32_t)(cg * KH * 8);
          if (it == 0) umma_bf16(d_col, dx, dyf, idf, 0u);
          else umma_bf16(d_col, dx, dyf, idf, 1u);
          for (int c = 1; c < chunks; ++c) umma_bf16(d_col, dx + (uint64_t)(c * 16), dyf + (uint64_t)(c * 16), idf, 1u);
        }
        umma_commit(empty + st);
      } else if (leader) {
        const uint64_t st_off = (uint64_t)((uint32_t)(st * st_bytes) >> 4);
        int a = 0;
        for (int r = 0; r < KH; ++r) {
          for (int cg = 0; cg < CG; ++cg, ++a) {
            const uint64_t dx = descX + st_off + (uint64_t)(cg * p.planeX + r * p.P);
            const uint64_t dyd = descY + st_off;
            const uint32_t d_col = tmem_acc + (uint32_t)(a * N8);
            if (it == 0) umma_bf16(d_col, dx, dyd, idesc, 0u);
            else umma_bf16(d_col, dx, dyd, idesc, 1u);
            for (int c = 1; c < chunks; ++c)     // 16 positions x 16 B = 16 units per K chunk
              umma_bf16(d_col, dx + (uint64_t)(c * 16), dyd + (uint64_t)(c * 16), idesc, 1u);
          }
        }
        umma_commit(empty + st);
      }
      __syncwarp();
    }
    if (any && leader) umma_commit(done);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: HWIO f32 -> bf16 [e = (cg*KH + r)*KW + q][Npad][8]; E entries (zero padded)
//   mode 0 (forward):        Wp[e][n][c] = w[r][q][cg*8+c][n]                 (Cin_k = Cin,  Cout_k = Cout)
//   mode 1 (data gradient):  Wp[e][n][c] = w[KH-1-r][KW-1-q][n][cg*8+c]       (Cin_k = Cout, Cout_k = Cin)
// ---------------------------------------------------------------------------------------------
__global__ void pack_nc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int KH, int KW, int Cin,
                               int Cout, int mode, int E, int Npad, const float* __restrict__ scale) {
  const int Ck = mode == 0 ? Cin : Cout;    // the kernel's reduction channels
  const int Nk = mode == 0 ? Cout : Cin;    // the kernel's output channels
  const int CG = (Ck + 7) / 8;
  const int total = E * Npad * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 7;
    const int nn = (i >> 3) % Npad;
    const int e = (i >> 3) / Npad;
    float v = 0.f;
    if (e < CG * KH * KW) {
      const int q = e % KW;
      const int t = e / KW;
      const int r = t % KH;
      const int cg = t / KH;
      const int ck = cg * 8 + c;
      if (ck < Ck && nn < Nk) {
        if (mode == 0) v = w[(((int64_t)r * KW + q) * Cin + ck) * Cout + nn] * (scale ? scale[nn] : 1.f);   // folded BatchNorm
        else v = w[(((int64_t)(KH - 1 - r) * KW + (KW - 1 - q)) * Cin + nn) * Cout + ck];
      }
    }
    wp[i] = __float2bfloat16_rn(v);
  }
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Geometry: rows per strip R and ring depth S.  One persistent CTA per SM walks the strips round-robin, so the
// cost of a choice is  rounds(R) * bytes moved per strip; the candidates must fit S >= 2 stages in shared memory.
static const size_t kNcSmemMax = 200 * 1024;

// Tuning knobs of the geometry search, read per call (diagnostic; scripts/bench_nc_sweep.py): DAFK_NC_R / DAFK_NC_WG_R
// force the rows per strip, DAFK_NC_S caps the ring depth, DAFK_NC_FIXED is the per-strip fixed cost in byte
// equivalents, DAFK_NC_ANYR=1 lets the search use every R in 1..32 instead of the powers of two.
struct NcTune { int fwd_R, wg_R, max_S, any_R; double fixed; int s1; };
// rows per strip the search may pick: powers of two and 1.5 x powers of two (any_R = 2: every R up to 32)
static bool nc_r_candidate(int R, int any_R) {
  if (any_R >= 2 || (R & (R - 1)) == 0) return true;
  return any_R == 1 && (R == 6 || R == 12 || R == 24);
}
static NcTune nc_tune() {
  NcTune t{0, 0, 3, 1, 20000.0, 1};
  if (const char* e = getenv("DAFK_NC_R")) t.fwd_R = atoi(e);
  if (const char* e = getenv("DAFK_NC_WG_R")) t.wg_R = atoi(e);
  if (const char* e = getenv("DAFK_NC_S")) { t.max_S = atoi(e); if (t.max_S < 2) t.max_S = 2; if (t.max_S > NC_MAX_STAGES) t.max_S = NC_MAX_STAGES; }
  if (const char* e = getenv("DAFK_NC_ANYR")) t.any_R = atoi(e);
  if (const char* e = getenv("DAFK_NC_FIXED")) t.fixed = atof(e);
  if (const char* e = getenv("DAFK_NC_S1")) t.s1 = atoi(e);
  return t;
}

// RAW staging of a [.., W, C] tensor of `esz`-byte elements: segment = whole row if it fits NC_RAW_SEG_MAX, else the largest
// pixel count whose bytes are a multiple of 16.  False when the layout does not allow 16-byte bulk copies / vector reads.
static bool nc_raw_seg(int W, int C, int esz, int& seg_px, int& nseg, int& seg_bytes) {
  const int px_bytes = C * esz;
  if (esz != 4 && C % 8 != 0) return false;
  if (((int64_t)W * px_bytes) % 16 != 0) return false;
  if ((int64_t)W * px_bytes <= NC_RAW_SEG_MAX) { seg_px = W; nseg = 1; seg_bytes = W * px_bytes; return true; }
  int m = 1;
  while ((m * px_bytes) % 16 != 0) m <<= 1;      // esz = 4: at most 4 pixels
  seg_px = (NC_RAW_SEG_MAX / px_bytes) / m * m;
  if (seg_px <= 0) return false;
  nseg = (W + seg_px - 1) / seg_px;
  seg_bytes = seg_px * px_bytes;
  return true;
}
static const size_t kNcSmemMaxRaw = 220 * 1024;
// Ring slots: a multiple of the converter warps, so that segment g (slot g % slots, owner warp g % warps) gives every slot
// ONE owner.  A slot's "full" barrier is waited on by parity: a warp that waits for wrap k of a slot must have seen wrap
// k - 1 complete, which only holds if it consumed wrap k - 1 itself -- with 8 or 16 slots and 6 warps a warp could reach its
// wait while another warp's earlier segment of the same slot was still in flight (bulk copies complete out of order), pass
// at once on the stale parity, convert garbage and release the slot (caught by scripts/stress_nc.py as a trapped launch).
static int nc_raw_slots(int fit) {
  const int w = NC_RAW_CONV / 32;
  const int slots = fit / w * w;
  return slots > NC_RAW_MAX_SLOTS ? NC_RAW_MAX_SLOTS : slots;
}
// RAW is selected by default for maps of at least 64 MB staged per launch whose segments are at least 2 KB (measured on
// B200, gpurun_out/r2t_sweep.txt: 8 -> 8 @ 192 x 224^2 forward 147 -> 126 us, weight gradient 256 -> 205 us, 8 -> 64 weight
// gradient 201 -> 144 us; small maps (56^2, 110^2) and 896-byte rows were 10 - 40 % slower)
static const int64_t kNcRawMinBytes = 64ll << 20;
static const int kNcRawMinSeg = 2048;
static const int kNcRawReserveSlots = 6;

static bool nc_fwd_geom(NcFwdP& p, size_t& smem, bool raw = false, double* cost_out = nullptr, int* seg_bytes_out = nullptr) {
  int slot_bytes = 0;
  p.raw_slots = 0;
  if (raw) {
    int sb;
    if (p.Cb > 0 || !nc_raw_seg(p.W, p.Cin, p.x_dt == DAFK_F32 ? 4 : 2, p.seg_px, p.nseg, sb)) return false;
    slot_bytes = round_up(sb, 128);
    if (seg_bytes_out) *seg_bytes_out = sb;
  }
  const size_t cap = raw ? kNcSmemMaxRaw - 512 - (size_t)kNcRawReserveSlots * slot_bytes : kNcSmemMax;
  p.P = p.W + 2 * p.pad;
  p.CG = (p.Cin + 7) / 8;
  p.E = round_up(p.CG * p.KH * p.KW, 2);
  p.Npad = round_up(p.Cout, 16);
  if (p.Npad > 256) return false;
  p.G = 256 / p.Npad;
  const size_t fixed = (size_t)p.E * p.Npad * 16 + (size_t)p.E * 16 + (size_t)p.Npad * 4 + 256 + 256;
  double best_cost = 0;
  int best = 0, best_S = 0;
  size_t best_smem = 0;
  const NcTune tune = nc_tune();
  for (int R = 1; R <= 32; ++R) {
    if (tune.fwd_R > 0 ? R != tune.fwd_R : !nc_r_candidate(R, tune.any_R)) continue;
    if (R > 1 && R / 2 >= p.Ho && tune.fwd_R <= 0) break;
    const int T = (R * p.P + 127) / 128;
    const int plane = round_up(128 * T + (p.KH - 1) * p.P + p.KW + 8, 8);
    const size_t stage = (size_t)p.CG * plane * 16;
    // RAW may run on ONE raster stage (the ring keeps the loads going while the MMAs read the raster; conversion and MMAs
    // of successive strips then alternate, costed at 1.5 x): wide inputs (64 channels = 8 planes) otherwise only fit
    // R = 1, i.e. every input row staged KH times
    const int min_S = (raw && tune.s1) ? 1 : 2;
    if (cap < fixed + min_S * stage) continue;
    int S = (int)((cap - fixed) / stage);
    if (S > tune.max_S) S = tune.max_S;
    if (raw && S > 2) S = 2;          // the issuer runs ahead through the ring, not through raster stages
    const int64_t strips = (int64_t)p.N * ((p.Ho + R - 1) / R);
    const int64_t rounds = (strips + kNumSMs - 1) / kNumSMs;
    // (rounds + 1): filling and draining the CTA's pipeline costs about one strip (one strip per CTA has no overlap at all)
    const double in_b = p.x_dt == DAFK_F32 ? 4.0 : 2.0;
    const double cost = (double)(rounds + 1) * (S == 1 ? 1.5 : 1.0) *
                        ((double)R * p.Wo * p.Cout * 4.0 + (double)(R + p.KH - 1) * p.W * p.Cin * in_b + tune.fixed);
    if (!best || cost < best_cost) {
      best = R; best_S = S; best_cost = cost; best_smem = fixed + S * stage;
    }
  }
  if (!best) return false;
  if (cost_out) *cost_out = best_cost;
  p.R = best;
  p.S = best_S;
  p.RS = p.R + p.KH - 1;
  p.T = (p.R * p.P + 127) / 128;
  p.plane = round_up(128 * p.T + (p.KH - 1) * p.P + p.KW + 8, 8);
  p.strips_per_img = (p.Ho + p.R - 1) / p.R;
  p.total_strips = p.N * p.strips_per_img;
  smem = best_smem;
  if (raw) {
    p.raw_slots = nc_raw_slots((int)((kNcSmemMaxRaw - 512 - best_smem) / slot_bytes));
    p.raw_slot_bytes = slot_bytes;
    smem = best_smem + 512 + (size_t)p.raw_slots * slot_bytes;
  }
  return true;
}

static bool nc_wg_geom(NcWgP& p, size_t& smem, bool raw = false, double* cost_out = nullptr, int* seg_bytes_out = nullptr) {
  int slot_bytes = 0;
  p.raw_slots = 0;
  if (raw) {
    int sbx, sby;
    if (p.Cb > 0 || p.Cout > NC_RAW_CONV || !nc_raw_seg(p.W, p.Cin, p.x_dt == DAFK_F32 ? 4 : 2, p.seg_px, p.nseg, sbx)) return false;
    p.Wo = p.W + 2 * p.pad - p.KW + 1;
    if (!nc_raw_seg(p.Wo, p.Cout, p.dy_dt == DAFK_F32 ? 4 : 2, p.seg_px_y, p.nseg_y, sby)) return false;
    slot_bytes = round_up(sbx > sby ? sbx : sby, 128);
    if (seg_bytes_out) *seg_bytes_out = sbx > sby ? sbx : sby;   // one wide tensor is enough (1 -> 64: 896-byte image rows)
  }
  const size_t cap = raw ? kNcSmemMaxRaw - 512 - (size_t)kNcRawReserveSlots * slot_bytes : kNcSmemMax;
  p.P = p.W + 2 * p.pad;
  p.CG = (p.Cin + 7) / 8;
  p.COG = (p.Cout + 7) / 8;
  p.N8 = p.COG * 8;
  if (p.KW > 8 || p.N8 > 256 || p.COG > NC_WG_PROD) return false;
  const int cols = p.KH * p.CG * p.N8;
  if (cols > 512) return false;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const size_t fixed = (size_t)NC_WG_PROD * 8 * 4 + 256 + 256;
  {
    static std::atomic<int> fold_ok{-2};       // lazily read once; a race only repeats the getenv
    if (fold_ok == -2) { const char* e = getenv("DAFK_NC_WG_FOLD"); fold_ok = e ? atoi(e) : 1; }
    p.fold = (fold_ok && p.COG == 1 && p.KH > 1 && p.KH * 8 <= 256) ? 1 : 0;
  }
  p.yoff = p.fold ? (p.KH - 1) * p.P : 0;
  double best_cost = 0;
  int best = 0, best_S = 0;
  size_t best_smem = 0;
  const NcTune tune = nc_tune();
  for (int R = 1; R <= 32; ++R) {
    if (tune.wg_R > 0 ? R != tune.wg_R : !nc_r_candidate(R, tune.any_R)) continue;
    if (R > 1 && R / 2 >= p.Ho && tune.wg_R <= 0) break;
    // fold: K covers the R+KH-1 input rows; the dY plane carries KH-1 zero rows on either side of its R rows
    const int kpos = round_up((p.fold ? R + p.KH - 1 : R) * p.P, 16);
    const int planeX = kpos + (p.fold ? 0 : (p.KH - 1) * p.P) + 16;
    const int planeYc = p.fold ? kpos + (p.KH - 1) * p.P + 16 : kpos;
    const size_t stage = (size_t)p.CG * planeX * 16 + (size_t)p.COG * planeYc * 16;
    if (cap < fixed + 2 * stage) continue;
    int S = (int)((cap - fixed) / stage);
    if (S > tune.max_S) S = tune.max_S;
    if (raw && S > 2) S = 2;
    if (S < 2) continue;
    const int64_t strips = (int64_t)p.N * ((p.Ho + R - 1) / R);
    const int64_t rounds = (strips + kNumSMs - 1) / kNumSMs;
    const double cost = (double)(rounds + 1) * ((double)R * p.Wo * p.Cout * 2.0 + (double)(R + p.KH - 1) * p.W * p.Cin * 2.0 + tune.fixed);
    if (!best || cost < best_cost) {
      best = R; best_S = S; best_cost = cost; best_smem = fixed + S * stage;
    }
  }
  if (!best) return false;
  if (cost_out) *cost_out = best_cost;
  p.R = best;
  p.S = best_S;
  p.RS = p.R + p.KH - 1;
  const int kpos = round_up((p.fold ? p.R + p.KH - 1 : p.R) * p.P, 16);
  p.chunks = kpos / 16;
  p.planeX = kpos + (p.fold ? 0 : (p.KH - 1) * p.P) + 16;
  p.planeY = p.fold ? kpos + (p.KH - 1) * p.P + 16 : kpos;
  p.strips_per_img = (p.Ho + p.R - 1) / p.R;
  p.total_strips = p.N * p.strips_per_img;
  smem = best_smem;
  if (raw) {
    p.raw_slots = nc_raw_slots((int)((kNcSmemMaxRaw - 512 - best_smem) / slot_bytes));
    p.raw_slot_bytes = slot_bytes;
    smem = best_smem + 512 + (size_t)p.raw_slots * slot_bytes;
  }
  return true;
}

// opt every kernel instantiation in to the full 227 KB of dynamic shared memory once (not a stream operation, so it is
// done at the first launch and never again -- later launches may be inside a CUDA-graph capture)
template <typename K>
static int nc_set_smem(K kernel, size_t smem, const char* name) {
  static std::mutex mu;
  static const void* done[32];
  static int ndone = 0;
  std::lock_guard<std::mutex> lk(mu);
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < ndone; ++i)
    if (done[i] == key) return DAFK_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  DAFK_REQUIRE(e == cudaSuccess, DAFK_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu bytes) failed: %s", name, smem,
               cudaGetErrorString(e));
  if (ndone < 32) done[ndone++] = key;
  return DAFK_OK;
}

// Geometry + staging mode of a forward / data-gradient launch (p holds the shape and dtypes).  RAW staging (rows by
// cp.async.bulk + converter warps): DAFK_NC_RAW unset = where it measured faster (large maps whose RAW geometry costs no more
// than the staged one); 1 = wherever the layout allows it; 0 = never (read per call so that a test can compare both kernels
// in one process).
static bool nc_fwd_select(NcFwdP& p, size_t& smem, bool& raw) {
  int raw_mode = -1;
  { const char* e = getenv("DAFK_NC_RAW"); raw_mode = e ? atoi(e) : -1; }
  NcFwdP pn = p;
  size_t smem_n = 0;
  double cost_r = 0, cost_n = 0;
  int seg_bytes = 0;
  const bool ok_n = nc_fwd_geom(pn, smem_n, false, &cost_n);
  const bool ok_r = raw_mode != 0 && nc_fwd_geom(p, smem, true, &cost_r, &seg_bytes);
  const int64_t in_bytes = (int64_t)p.N * p.H * p.W * p.Cin * (p.x_dt == DAFK_F32 ? 4 : 2);
  raw = ok_r && (raw_mode > 0 || !ok_n || (in_bytes >= kNcRawMinBytes && seg_bytes >= kNcRawMinSeg && cost_r <= 1.15 * cost_n));
  if (!raw) { p = pn; smem = smem_n; }
  return raw || ok_n;
}

// Geometry + staging mode of a weight-gradient launch (p holds the shape and dtypes).  DAFK_NC_RAW: unset = RAW where it
// measured faster, 1 = wherever the layout allows it, 0 = never.
static bool nc_wg_select(NcWgP& p, size_t& smem, bool& raw) {
  int raw_mode = -1;
  { const char* e = getenv("DAFK_NC_RAW"); raw_mode = e ? atoi(e) : -1; }
  NcWgP pn = p;
  size_t smem_n = 0;
  double cost_r = 0, cost_n = 0;
  int seg_bytes = 0;
  const bool ok_n = nc_wg_geom(pn, smem_n, false, &cost_n);
  const bool ok_r = raw_mode != 0 && nc_wg_geom(p, smem, true, &cost_r, &seg_bytes);
  const int64_t in_bytes = (int64_t)p.N * p.H * p.W * p.Cin * (p.x_dt == DAFK_F32 ? 4 : 2) +
                           (int64_t)p.N * p.Ho * p.Wo * p.Cout * (p.dy_dt == DAFK_F32 ? 4 : 2);
  raw = ok_r && (raw_mode > 0 || !ok_n || (in_bytes >= kNcRawMinBytes && seg_bytes >= kNcRawMinSeg && cost_r <= 1.15 * cost_n));
  if (!raw) { p = pn; smem = smem_n; }
  return raw || ok_n;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_conv_nc_supported(int Cin, int Cout, int KH, int KW, int W, int pad, int kind) {
  if (kind == 2) {
    NcWgP p{};
    p.H = p.Ho = 1 << 20; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
    size_t smem;
    return nc_wg_geom(p, smem) ? 1 : 0;
  }
  NcFwdP p{};
  p.H = p.Ho = 1 << 20; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  size_t smem;
  return nc_fwd_geom(p, smem) ? 1 : 0;
}

int dafk_conv_nc_wgrad_stages_raw(int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, int x_dt, int dy_dt) {
  NcWgP p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  p.Ho = H + 2 * pad - KH + 1;
  p.Wo = W + 2 * pad - KW + 1;
  if (N <= 0 || p.Ho <= 0 || p.Wo <= 0) return 0;
  p.x_dt = x_dt; p.dy_dt = dy_dt;
  size_t smem;
  bool raw = false;
  return nc_wg_select(p, smem, raw) && raw ? 1 : 0;
}

int dafk_conv_nc_plan(int kind, int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, int x_dt, int dy_dt,
                      int64_t* plan) {
  DAFK_REQUIRE(plan != nullptr && (kind == 0 || kind == 2), DAFK_ERR_BAD_ARG, "dafk_conv_nc_plan: bad argument");
  size_t smem = 0;
  bool raw = false;
  if (kind == 2) {
    NcWgP p{};
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
    p.Ho = H + 2 * pad - KH + 1;
    p.Wo = W + 2 * pad - KW + 1;
    p.x_dt = x_dt; p.dy_dt = dy_dt;
    if (N <= 0 || p.Ho <= 0 || p.Wo <= 0 || !nc_wg_select(p, smem, raw)) return DAFK_ERR_UNSUPPORTED;
    const int64_t v[10] = {raw, p.R, p.S, p.raw_slots, p.raw_slot_bytes, p.seg_px, p.nseg, p.seg_px_y, p.nseg_y, (int64_t)smem};
    for (int i = 0; i < 10; ++i) plan[i] = v[i];
    return DAFK_OK;
  }
  NcFwdP p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  p.Ho = H + 2 * pad - KH + 1;
  p.Wo = W + 2 * pad - KW + 1;
  p.x_dt = x_dt; p.y_dt = dy_dt;
  p.Ca = Cin;
  if (N <= 0 || p.Ho <= 0 || p.Wo <= 0 || !nc_fwd_select(p, smem, raw)) return DAFK_ERR_UNSUPPORTED;
  const int64_t v[10] = {raw, p.R, p.S, p.raw_slots, p.raw_slot_bytes, p.seg_px, p.nseg, 0, 0, (int64_t)smem};
  for (int i = 0; i < 10; ++i) plan[i] = v[i];
  return DAFK_OK;
}

int64_t dafk_conv_nc_packed_elems(int Cin_k, int Cout_k, int KH, int KW) {
  int CG = (Cin_k + 7) / 8;
  int E = round_up(CG * KH * KW, 2);
  return (int64_t)E * round_up(Cout_k, 16) * 8;
}

int dafk_pack_conv_nc(const float* w_hwio, void* wp, int KH, int KW, int Cin, int Cout, int mode, void* stream) {
  DAFK_REQUIRE(w_hwio && wp && KH > 0 && KW > 0 && Cin > 0 && Cout > 0 && (mode == 0 || mode == 1), DAFK_ERR_BAD_ARG,
               "dafk_pack_conv_nc: bad argument");
  const int Ck = mode == 0 ? Cin : Cout, Nk = mode == 0 ? Cout : Cin;
  const int CG = (Ck + 7) / 8;
  const int E = round_up(CG * KH * KW, 2), Npad = round_up(Nk, 16);
  const int total = E * Npad * 8;
  pack_nc_kernel<<<bw_grid(total, 256), 256, 0, as_stream(stream)>>>(w_hwio, (__nv_bfloat16*)wp, KH, KW, Cin, Cout, mode,
                                                                     E, Npad, nullptr);
  return check_launch("dafk_pack_conv_nc");
}

int dafk_pack_conv_nc_scaled(const float* w_hwio, const float* scale, void* wp, int KH, int KW, int Cin, int Cout,
                             void* stream) {
  DAFK_REQUIRE(w_hwio && scale && wp && KH > 0 && KW > 0 && Cin > 0 && Cout > 0, DAFK_ERR_BAD_ARG,
               "dafk_pack_conv_nc_scaled: bad argument");
  const int CG = (Cin + 7) / 8;
  const int E = round_up(CG * KH * KW, 2), Npad = round_up(Cout, 16);
  const int total = E * Npad * 8;
  pack_nc_kernel<<<bw_grid(total, 256), 256, 0, as_stream(stream)>>>(w_hwio, (__nv_bfloat16*)wp, KH, KW, Cin, Cout, 0, E,
                                                                     Npad, scale);
  return check_launch("dafk_pack_conv_nc_scaled");
}

static int conv_nc_fwd_impl(const void* x, const void* xb, int Ca, int Cb, int x_dt, const void* wp, const float* bias, void* y,
                           int y_dt, int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, int act, float alpha,
                           void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && pad >= 0, DAFK_ERR_BAD_ARG,
               "dafk_conv_nc_fwd: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && wp && y, DAFK_ERR_BAD_ARG, "dafk_conv_nc_fwd: null pointer");
  DAFK_REQUIRE((x_dt == DAFK_F32 || x_dt == DAFK_BF16) && (y_dt == DAFK_F32 || y_dt == DAFK_BF16), DAFK_ERR_BAD_ARG,
               "dafk_conv_nc_fwd: bad dtype");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(wp) && DAFK_ALIGNED16(y), DAFK_ERR_ALIGN,
               "dafk_conv_nc_fwd: pointers must be 16-byte aligned");
  NcFwdP p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  p.Ho = H + 2 * pad - KH + 1;
  p.Wo = W + 2 * pad - KW + 1;
  DAFK_REQUIRE(p.Ho > 0 && p.Wo > 0, DAFK_ERR_BAD_ARG, "dafk_conv_nc_fwd: empty output");
  p.x_dt = x_dt; p.y_dt = y_dt; p.act = act; p.alpha = alpha;
  p.Ca = Ca; p.Cb = Cb;
  { const char* e = getenv("DAFK_NC_DEBUG"); p.dbg = e ? atoi(e) : 0; }
  size_t smem;
  bool raw = false;
  DAFK_REQUIRE(nc_fwd_select(p, smem, raw), DAFK_ERR_UNSUPPORTED,
               "dafk_conv_nc_fwd: geometry does not fit (Cin=%d Cout=%d k=%dx%d W=%d)", Cin, Cout, KH, KW, W);
  // one persistent CTA per SM (it owns all 512 TMEM columns): request more than half of the shared memory
  if (smem < 120 * 1024) smem = 120 * 1024;
  int grid = kNumSMs < p.total_strips ? kNumSMs : p.total_strips;
  cudaStream_t s = as_stream(stream);
  int rc;
  // opt-in (DAFK_NC_BULK=1) until the decoder keeps its activations in bf16; read per call so that a test can compare
  // both kernels in one process.  Measured at 32 x 224^2: 8 -> 8 37.1 -> 24.9 us, 8 -> 64 105 -> 123 us (slower).
  int bulk_ok = 0;
  { const char* e = getenv("DAFK_NC_BULK"); bulk_ok = e ? atoi(e) : 0; }
  int l12 = 1;
  { const char* e = getenv("DAFK_NC_L12"); l12 = e ? atoi(e) : 1; }
  if (raw && x_dt == DAFK_F32) {
    rc = nc_set_smem(conv_nc_fwd_kernel<float, NC_MODE_RAW>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<float, NC_MODE_RAW><<<grid, NC_BULK_THREADS, smem, s>>>(p, (const float*)x, (const __nv_bfloat16*)wp,
                                                                              bias, y);
  } else if (raw) {
    rc = nc_set_smem(conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_RAW>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_RAW><<<grid, NC_BULK_THREADS, smem, s>>>(p, (const __nv_bfloat16*)x,
                                                                                      (const __nv_bfloat16*)wp, bias, y);
  } else if (l12 && x_dt == DAFK_F32) {
    rc = nc_set_smem(conv_nc_fwd_kernel<float, NC_MODE_L12>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<float, NC_MODE_L12><<<grid, NC_BULK_THREADS, smem, s>>>(p, (const float*)x, (const __nv_bfloat16*)wp,
                                                                              bias, y, (const float*)xb);
  } else if (l12 && !(bulk_ok && Cin == 8 && Cb == 0)) {
    rc = nc_set_smem(conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_L12>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_L12><<<grid, NC_BULK_THREADS, smem, s>>>(p, (const __nv_bfloat16*)x,
                                                                                      (const __nv_bfloat16*)wp, bias, y,
                                                                                      (const __nv_bfloat16*)xb);
  } else if (x_dt == DAFK_F32) {
    rc = nc_set_smem(conv_nc_fwd_kernel<float, NC_MODE_DEFAULT>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<float, NC_MODE_DEFAULT><<<grid, NC_FWD_THREADS, smem, s>>>(p, (const float*)x, (const __nv_bfloat16*)wp, bias, y,
                                                                                 (const float*)xb);
  } else if (bulk_ok && Cin == 8 && Cb == 0) {
    // one pixel = 16 B = one raster position: rows are staged by the TMA engine, two epilogue groups
    rc = nc_set_smem(conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_BULK>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_BULK><<<grid, NC_BULK_THREADS, smem, s>>>(p, (const __nv_bfloat16*)x,
                                                                              (const __nv_bfloat16*)wp, bias, y);
  } else {
    rc = nc_set_smem(conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_DEFAULT>, smem, "dafk_conv_nc_fwd");
    if (rc) return rc;
    conv_nc_fwd_kernel<__nv_bfloat16, NC_MODE_DEFAULT><<<grid, NC_FWD_THREADS, smem, s>>>(p, (const __nv_bfloat16*)x,
                                                                               (const __nv_bfloat16*)wp, bias, y,
                                                                               (const __nv_bfloat16*)xb);
  }
  return check_launch("dafk_conv_nc_fwd");
}

int dafk_conv_nc_fwd(const void* x, int x_dt, const void* wp, const float* bias, void* y, int y_dt, int N, int H,
                     int W, int Cin, int Cout, int KH, int KW, int pad, int act, float alpha, void* stream) {
  return conv_nc_fwd_impl(x, nullptr, Cin, 0, x_dt, wp, bias, y, y_dt, N, H, W, Cin, Cout, KH, KW, pad, act, alpha, stream);
}

int dafk_conv_nc_fwd_cat(const void* xa, int Ca, const void* xb, int Cb, int x_dt, const void* wp, const float* bias,
                         void* y, int y_dt, int N, int H, int W, int Cout, int KH, int KW, int pad, int act, float alpha,
                         void* stream) {
  DAFK_REQUIRE(Ca > 0 && Cb > 0 && Ca % 8 == 0, DAFK_ERR_UNSUPPORTED,
               "dafk_conv_nc_fwd_cat: the first source must hold a whole number of 8-channel groups (Ca=%d Cb=%d)", Ca, Cb);
  DAFK_REQUIRE(xb && DAFK_ALIGNED16(xb), DAFK_ERR_BAD_ARG, "dafk_conv_nc_fwd_cat: second source missing or misaligned");
  return conv_nc_fwd_impl(xa, xb, Ca, Cb, x_dt, wp, bias, y, y_dt, N, H, W, Ca + Cb, Cout, KH, KW, pad, act, alpha, stream);
}

static int conv_nc_wgrad_impl(const void* x, const void* xb, int Ca, int Cb, int x_dt, const void* dy, int dy_dt, float* dw,
                              float* db, int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, void* stream) {
  DAFK_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && pad >= 0, DAFK_ERR_BAD_ARG,
               "dafk_conv_nc_wgrad: bad shape");
  if (N == 0) return DAFK_OK;
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_conv_nc_wgrad: null pointer");
  DAFK_REQUIRE((x_dt == DAFK_F32 || x_dt == DAFK_BF16) && (dy_dt == DAFK_F32 || dy_dt == DAFK_BF16), DAFK_ERR_BAD_ARG,
               "dafk_conv_nc_wgrad: bad dtype");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dy), DAFK_ERR_ALIGN, "dafk_conv_nc_wgrad: pointers must be 16-byte aligned");
  NcWgP p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad;
  p.Ho = H + 2 * pad - KH + 1;
  p.Wo = W + 2 * pad - KW + 1;
  DAFK_REQUIRE(p.Ho > 0 && p.Wo > 0, DAFK_ERR_BAD_ARG, "dafk_conv_nc_wgrad: empty output");
  p.x_dt = x_dt; p.dy_dt = dy_dt;
  p.Ca = Ca; p.Cb = Cb;
  size_t smem;
  bool raw = false;
  DAFK_REQUIRE(nc_wg_select(p, smem, raw), DAFK_ERR_UNSUPPORTED,
               "dafk_conv_nc_wgrad: geometry does not fit (Cin=%d Cout=%d k=%dx%d W=%d)", Cin, Cout, KH, KW, W);
  if (smem < 120 * 1024) smem = 120 * 1024;     // one persistent CTA per SM
  int grid = kNumSMs < p.total_strips ? kNumSMs : p.total_strips;
  cudaStream_t s = as_stream(stream);
  int rc;
#define NC_WG(TX, TY)                                                                                   \
  do {                                                                                                  \
    if (raw) {                                                                                          \
      rc = nc_set_smem(conv_nc_wgrad_kernel<TX, TY, true>, smem, "dafk_conv_nc_wgrad");                   \
      if (rc) return rc;                                                                                \
      conv_nc_wgrad_kernel<TX, TY, true><<<grid, NC_WG_RAW_THREADS, smem, s>>>(p, (const TX*)x, (const TY*)dy, dw, db); \
    } else {                                                                                            \
      rc = nc_set_smem(conv_nc_wgrad_kernel<TX, TY, false>, smem, "dafk_conv_nc_wgrad");                  \
      if (rc) return rc;                                                                                \
      conv_nc_wgrad_kernel<TX, TY, false><<<grid, NC_WG_THREADS, smem, s>>>(p, (const TX*)x, (const TY*)dy, dw, db, (const TX*)xb); \
    }                                                                                                   \
  } while (0)
  if (x_dt == DAFK_F32 && dy_dt == DAFK_F32) NC_WG(float, float);
  else if (x_dt == DAFK_F32) NC_WG(float, __nv_bfloat16);
  else if (dy_dt == DAFK_F32) NC_WG(__nv_bfloat16, float);
  else NC_WG(__nv_bfloat16, __nv_bfloat16);
#undef NC_WG
  return check_launch("dafk_conv_nc_wgrad");
}

int dafk_conv_nc_wgrad(const void* x, int x_dt, const void* dy, int dy_dt, float* dw, float* db, int N, int H, int W,
                       int Cin, int Cout, int KH, int KW, int pad, void* stream) {
  return conv_nc_wgrad_impl(x, nullptr, Cin, 0, x_dt, dy, dy_dt, dw, db, N, H, W, Cin, Cout, KH, KW, pad, stream);
}

int dafk_conv_nc_wgrad_cat(const void* xa, int Ca, const void* xb, int Cb, int x_dt, const void* dy, int dy_dt, float* dw,
                           float* db, int N, int H, int W, int Cout, int KH, int KW, int pad, void* stream) {
  DAFK_REQUIRE(Ca > 0 && Cb > 0 && Ca % 8 == 0, DAFK_ERR_UNSUPPORTED,
               "dafk_conv_nc_wgrad_cat: the first source must hold a whole number of 8-channel groups (Ca=%d Cb=%d)", Ca, Cb);
  DAFK_REQUIRE(xb && DAFK_ALIGNED16(xb), DAFK_ERR_BAD_ARG, "dafk_conv_nc_wgrad_cat: second source missing or misaligned");
  return conv_nc_wgrad_impl(xa, xb, Ca, Cb, x_dt, dy, dy_dt, dw, db, N, H, W, Ca + Cb, Cout, KH, KW, pad, stream);
}

}  // extern "C"
