// Weight gradient of the stride-1 3x3 'same' convolutions with few channels per pixel (64 / 128: the full- and
// half-resolution UNet and segmentor layers, models/unet.py:95,99, model_components/segmentor.py:18) on tcgen05.
//
// The general kernel (conv_tc.cu: one CTA per filter tap) re-reads the 128-pixel X and dY boxes once per tap and is
// bound by the L2 -> SM bandwidth when Cin = Cout = 64 (32 KB of operands per 0.5 M MACs; 294 TFLOP/s measured).
// Here a CTA owns ALL NINE taps of a (64 ci x 64 co) block:
//   * per 16 x 8 pixel tile it loads the dY box once (16 KB) and ONE haloed X tile, 18 x 10 pixels (22.5 KB), both
//     128B-swizzled by TMA; OOB zero fill is the convolution's padding;
//   * dW[tap][ci][co] = sum_pixels X[pixel + tap][ci] * dY[pixel][co]: X is the MN-major A operand, dY the MN-major B
//     operand, K = 16 pixels = one tile row.  Tap (r,q) is the same shared memory with the A descriptor started
//     ((y+r)*18 + q) pixels (128 B each) later -- the 128B swizzle is a function of the absolute shared-memory address,
//     so a start in the middle of a swizzle atom reads what TMA wrote (as in conv_tc_halo_kernel);
//   * TWO taps share one MMA: M = 128 = [tap a: 64 ci | tap b: 64 ci], the second half being `LBO` bytes after the
//     first (the descriptor's 64-element MN-block stride) = the byte distance between the two taps.  Nine taps = five
//     M=128 x N=64 accumulators (the tenth half is ignored) = 320 TMEM columns, kept over all tiles of the CTA;
//   * split-K over pixel tiles; the epilogue adds the CTA's partial with 128-bit vector atomics.
// Operand bytes per tile fall from 9 x 32 KB to 38.5 KB, the MMAs run at M = 128.
//   warp 0: TMA producer    warps 1, 6: MMA issuers (warp-uniform, one elected lane each)    warps 2-5: epilogue
#include "tc_ptx.cuh"
#include <atomic>

namespace dafk {

constexpr int WH_THREADS = 224;                      // warp 0 TMA, warps 1 and 6 MMA issue, warps 2-5 epilogue
constexpr int WH_MMA2_WARP = 6;
constexpr int WH_TW = 16, WH_TH = 8;                 // dY tile (pixels of one image)
constexpr int WH_PW = WH_TW + 2, WH_PH = WH_TH + 2;  // haloed X tile
constexpr int WH_DY_BYTES = WH_TW * WH_TH * 128;     // 16 KB
constexpr int WH_X_BYTES = WH_PW * WH_PH * 128;      // 23 040 B
constexpr int WH_X_PAD = 24 * 1024;                  // stage slot (1024 B aligned; slack for the ignored tenth half)
constexpr int WH_STAGE = WH_DY_BYTES + WH_X_PAD;
constexpr int WH_STAGES = 4;
constexpr int WH_SMEM = WH_STAGES * WH_STAGE + 1024 + 256;
constexpr uint32_t WH_TMEM_COLS = 512;               // 5 accumulators x 64 columns -> next power of two

__global__ void __launch_bounds__(WH_THREADS, 1) conv_tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                           const __grid_constant__ CUtensorMap tmDY,
                                                                           float* __restrict__ dw, int cin_off,
                                                                           int cin_total, int Cout, int co_blocks,
                                                                           int tiles_x, int tiles_y, int total_tiles,
                                                                           int tiles_per_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WH_STAGE);
  uint64_t* empty_bar = full_bar + WH_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WH_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int unit = blockIdx.y;
  const int cob = unit % co_blocks, cib = unit / co_blocks;
  const int t_begin = blockIdx.x * tiles_per_split;
  const int t_end = min(total_tiles, t_begin + tiles_per_split);
  const int num_tiles = t_end - t_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    // one tcgen05.commit from each of the two issuing warps
    for (int s = 0; s < WH_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 2); }
    mbar_init(tmem_full_bar, 2);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, WH_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_tiles > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < num_tiles; ++it) {
          int t = t_begin + it;
          const int txi = t % tiles_x; t /= tiles_x;
          const int tyi = t % tiles_y; t /= tiles_y;
          const int x0 = txi * WH_TW, y0 = tyi * WH_TH, img = t;
          const int s = it % WH_STAGES;
          mbar_wait(empty_bar + s, (((uint32_t)(it / WH_STAGES)) & 1u) ^ 1u);
          mbar_expect_tx(full_bar + s, WH_DY_BYTES + WH_X_BYTES);
          uint8_t* st = smem + s * WH_STAGE;
          tma_load_4d(st, &tmDY, full_bar + s, cob * 64, x0, y0, img);
          tma_load_4d(st + WH_DY_BYTES, &tmX, full_bar + s, cib * 64, x0 - 1, y0 - 1, img);
        }
      }
    } else if (warp == 1 || warp == WH_MMA2_WARP) {
      // two issuers: a single thread needs ~30-40 cycles per tcgen05.mma (descriptor arithmetic, predicate, issue) and a
      // 128 x 64 x 16 MMA runs 48; warp 1 owns accumulators 0, 2, 4 and warp 6 accumulators 1, 3
      constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);     // A (X) and B (dY) both MN-major
      const bool first = warp == 1;
      const uint32_t leader = elect_one();
      const uint32_t tmem_acc = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t s0 = smem_u32(smem);
      for (int it = 0; it < num_tiles; ++it) {
        const int s = it % WH_STAGES;
        mbar_wait(full_bar + s, ((uint32_t)(it / WH_STAGES)) & 1u);
        tc_fence_after();
        if (leader) {
          const uint32_t dy_s = s0 + (uint32_t)(s * WH_STAGE);
          const uint32_t x_s = dy_s + (uint32_t)WH_DY_BYTES;
#pragma unroll 1
          for (int y = 0; y < WH_TH; ++y) {
            // B: tile row y of dY = 16 pixels = two 8-pixel groups 1024 B apart
            const uint64_t db = make_smem_desc(dy_s + (uint32_t)(y * WH_TW * 128), 16 * 1024, 1024);
            const uint32_t acc = (it > 0 || y > 0) ? 1u : 0u;
            // A: tap pairs (0,0)+(0,1) | (0,2)+(1,0) | (1,1)+(1,2) | (2,0)+(2,1) | (2,2)+ignored
            const uint32_t row0 = x_s + (uint32_t)((y * WH_PW) * 128);
            if (first) {
              umma_bf16(tmem_acc + 0u, make_smem_desc(row0 + (0 * WH_PW + 0) * 128, 128, 1024), db, idesc, acc);
              umma_bf16(tmem_acc + 128u, make_smem_desc(row0 + (1 * WH_PW + 1) * 128, 128, 1024), db, idesc, acc);
              umma_bf16(tmem_acc + 256u, make_smem_desc(row0 + (2 * WH_PW + 2) * 128, 128, 1024), db, idesc, acc);
            } else {
              umma_bf16(tmem_acc + 64u, make_smem_desc(row0 + (0 * WH_PW + 2) * 128, (WH_PW - 2) * 128, 1024), db, idesc, acc);
              umma_bf16(tmem_acc + 192u, make_smem_desc(row0 + (2 * WH_PW + 0) * 128, 128, 1024), db, idesc, acc);
            }
          }
          umma_commit(empty_bar + s);
        }
        __syncwarp();
      }
      if (leader) umma_commit(tmem_full_bar);
      __syncwarp();
    } else if (warp < WH_MMA2_WARP) {
      const int q4 = warp & 3;
      const int row = q4 * 32 + lane;                 // accumulator row = TMEM lane: [second tap of the pair][ci]
      const int half = row >> 6, ci = row & 63;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
        const int tap = 2 * j + half;
        float* dst = dw + ((int64_t)tap * cin_total + cin_off + cib * 64 + ci) * Cout + cob * 64;
#pragma unroll 1
        for (int c = 0; c < 64; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(j * 64 + c), v);
          tmem_ld_wait();
          if (tap < 9) {
#pragma unroll
            for (int k = 0; k < 16; k += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + c + k),
                        make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]),
                                    __uint_as_float(v[k + 3])));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WH_TMEM_COLS);
  }
}

typedef CUresult (*PFN_encodeTiledWH)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiledWH wh_get_encode() {
  static std::atomic<PFN_encodeTiledWH> fn{nullptr};   // a race only repeats the lookup
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiledWH>(p);
  }
  return fn;
}

// NHWC bf16 map with a (64 ch, bw, bh, 1) box, 128B swizzle, zero fill outside the image
static int wh_make_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int bw, int bh) {
  PFN_encodeTiledWH enc = wh_get_encode();
  DAFK_REQUIRE(enc != nullptr, DAFK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAFK_REQUIRE(r == CUDA_SUCCESS, DAFK_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad halo) failed with %d", (int)r);
  return DAFK_OK;
}

}  // namespace dafk

using namespace dafk;

extern "C" {

int dafk_conv3x3_tc_wgrad_halo_supported(int Cin, int Cout) { return Cin > 0 && Cout > 0 && Cin % 64 == 0 && Cout % 64 == 0; }

int dafk_conv3x3_tc_wgrad_halo(const void* x, int Cin, int cin_off, int cin_total, const void* dy, int Cout, float* dw,
                               int N, int H, int W, void* stream) {
  DAFK_REQUIRE(N > 0 && H > 0 && W > 0 && cin_off >= 0 && cin_off + Cin <= cin_total, DAFK_ERR_BAD_ARG,
               "dafk_conv3x3_tc_wgrad_halo: bad shape");
  DAFK_REQUIRE(dafk_conv3x3_tc_wgrad_halo_supported(Cin, Cout), DAFK_ERR_UNSUPPORTED,
               "dafk_conv3x3_tc_wgrad_halo: channels must be multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  DAFK_REQUIRE(x && dy && dw, DAFK_ERR_BAD_ARG, "dafk_conv3x3_tc_wgrad_halo: null pointer");
  DAFK_REQUIRE(DAFK_ALIGNED16(x) && DAFK_ALIGNED16(dy) && DAFK_ALIGNED16(dw), DAFK_ERR_ALIGN,
               "dafk_conv3x3_tc_wgrad_halo: pointers must be 16-byte aligned");
  CUtensorMap mx, mdy;
  int rc = wh_make_map(&mx, x, N, H, W, Cin, WH_PW, WH_PH);
  if (rc) return rc;
  rc = wh_make_map(&mdy, dy, N, H, W, Cout, WH_TW, WH_TH);
  if (rc) return rc;
  static std::atomic<bool> configured{false};   // idempotent one-time attribute set: a race only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM);
    DAFK_REQUIRE(e == cudaSuccess, DAFK_ERR_CUDA, "cudaFuncSetAttribute(conv_tc_wgrad_halo) failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles_x = (W + WH_TW - 1) / WH_TW, tiles_y = (H + WH_TH - 1) / WH_TH;
  const int64_t total64 = (int64_t)tiles_x * tiles_y * N;
  DAFK_REQUIRE(total64 < (1LL << 31), DAFK_ERR_UNSUPPORTED, "dafk_conv3x3_tc_wgrad_halo: too many tiles");
  const int total = (int)total64;
  const int co_blocks = Cout / 64, units = co_blocks * (Cin / 64);
  int want = (kNumSMs + units - 1) / units;
  if (want > total) want = total;
  if (want < 1) want = 1;
  const int per = (total + want - 1) / want;
  const int splits = (total + per - 1) / per;
  conv_tc_wgrad_halo_kernel<<<dim3(splits, units), WH_THREADS, WH_SMEM, as_stream(stream)>>>(
      mx, mdy, dw, cin_off, cin_total, Cout, co_blocks, tiles_x, tiles_y, total, per);
  return check_launch("dafk_conv3x3_tc_wgrad_halo");
}

}  // extern "C"
