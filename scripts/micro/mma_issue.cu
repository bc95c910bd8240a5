// micro-benchmark: cycles per tcgen05.mma for different shapes / operand layouts / accumulator patterns
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../multimodal_segmentation_b200/csrc mma_issue.cu -o mma_issue
#include "tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace dafk;
namespace dafk { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

struct Cfg { int M, N, swz, mn_major, nacc, nmma, accum_first; };

__global__ void __launch_bounds__(128) k(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm = slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = make_idesc(c.M, c.N, c.mn_major, c.mn_major);
    uint32_t a = smem_u32(smem), b = a + 32 * 1024;
    uint64_t da, db;
    if (c.swz) { da = make_smem_desc(a, c.mn_major ? 16384 : 16, 1024); db = make_smem_desc(b, c.mn_major ? 16384 : 16, 1024); }
    else if (c.mn_major) { da = make_smem_desc_ns(a, 128, 16); db = make_smem_desc_ns(b, 128, 4096); }
    else { da = make_smem_desc_ns(a, 2048, 128); db = make_smem_desc_ns(b, 2048, 128); }
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      const uint32_t N = (uint32_t)c.N;
      if (c.nacc == 1) {
        umma_bf16(tm, da, db, idesc, 0u);
#pragma unroll 1
        for (int i = 1; i < c.nmma; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) umma_bf16(tm, da, db, idesc, 1u);
        }
      } else if (c.nacc == 2) {
#pragma unroll 1
        for (int i = 0; i < c.nmma; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) umma_bf16(tm + (u & 1) * N, da, db, idesc, i > 0 ? 1u : 0u);
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < c.nmma; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) umma_bf16(tm + (u & 3) * N, da, db, idesc, i > 0 ? 1u : 0u);
        }
      }
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}


__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_wait_asm(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128) k2(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
    const uint32_t idesc = make_idesc(c.M, c.N, c.mn_major, c.mn_major);
    const uint32_t a = __shfl_sync(0xffffffffu, smem_u32(smem), 0), b = a + 32 * 1024;
    const uint64_t da = make_smem_desc(a, 16, 1024), db = make_smem_desc(b, 16, 1024);
    const uint32_t leader = elect_one();
    uint32_t phase = 0;
    const uint32_t N = (uint32_t)c.N;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < c.nmma; i += 8) {
        if (leader) {
#pragma unroll
          for (int u = 0; u < 8; ++u) umma_bf16(tm + (u & 1) * N, da + (uint64_t)(u * 2), db + (uint64_t)(u * 2), idesc, i > 0 ? 1u : 0u);
        }
        __syncwarp();
      }
      long long t1 = clock64();
      if (leader) umma_commit(&bar);
      __syncwarp();
      mbar_wait_asm(&bar, phase);
      phase ^= 1;
      long long t2 = clock64();
      if (leader) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(__shfl_sync(0xffffffffu, slot, 0), 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64); long long h[6];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  Cfg cfgs[] = {
    // M, N, swz, mn, nacc, nmma, accum_first
    {128, 16, 0, 0, 1, 256, 0}, {128, 16, 0, 0, 2, 256, 0}, {128, 16, 0, 0, 4, 256, 0},
    {128, 16, 1, 0, 1, 256, 0}, {128, 16, 1, 0, 4, 256, 0},
    {128, 64, 0, 0, 1, 256, 0}, {128, 64, 1, 0, 1, 256, 0}, {128, 64, 1, 0, 4, 256, 0},
    {128, 128, 0, 0, 1, 256, 0}, {128, 128, 1, 0, 1, 256, 0}, {128, 128, 1, 0, 2, 256, 0}, {128, 128, 1, 0, 4, 256, 0},
    {128, 256, 1, 0, 1, 256, 0}, {128, 256, 1, 0, 2, 256, 0}, {128, 256, 0, 0, 1, 256, 0},
    {64, 8, 0, 1, 1, 256, 0}, {64, 8, 0, 1, 4, 256, 0}, {64, 64, 0, 1, 1, 256, 0}, {64, 64, 0, 1, 4, 256, 0},
    {64, 64, 1, 1, 1, 256, 0}, {128, 128, 1, 1, 1, 256, 0}, {128, 128, 1, 1, 2, 256, 0},
    {64, 8, 0, 0, 1, 256, 0}, {64, 64, 0, 0, 1, 256, 0}, {64, 64, 1, 0, 1, 256, 0},
  };
  for (auto& c : cfgs) {
    k<<<1, 128, 100 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d: CUDA error %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d swz=%d mn=%d nacc=%d : issue %6.1f cyc/mma, total %6.1f cyc/mma (floor %d)\n", c.M, c.N, c.swz, c.mn_major,
           c.nacc, h[4] / (double)c.nmma, h[5] / (double)c.nmma, (c.M > 64 ? c.M : 128) * c.N / 256);
  }
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  Cfg c2[] = {{128, 16, 1, 0, 2, 256, 0}, {128, 64, 1, 0, 2, 256, 0}, {128, 128, 1, 0, 2, 256, 0}, {128, 256, 1, 0, 2, 256, 0}};
  for (auto& c : c2) {
    k2<<<1, 128, 100 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k2 M=%d N=%d: CUDA error %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("uniform style M=%3d N=%3d : issue %6.1f cyc/mma, total %6.1f cyc/mma\n", c.M, c.N, h[4] / (double)c.nmma, h[5] / (double)c.nmma);
  }
  return 0;
}
