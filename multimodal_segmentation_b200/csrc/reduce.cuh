// Per-channel two-quantity reduction shared by FiLM backward, BatchNorm statistics and
// BatchNorm backward.  Each thread owns a fixed group of 4 consecutive channels
// (c0 = (threadIdx.x*4) % C, guaranteed by 1024 % C == 0 and 256-thread blocks), so the
// reduction is: warp shuffles across lanes that share a group -> shared-memory atomics
// across warps -> one double atomicAdd per channel per block into global memory.
#pragma once
#include "common.cuh"

namespace dafk {

template <int THREADS>
__device__ __forceinline__ void channel_reduce2(float (&a)[4], float (&b)[4], int C, float* sm /* 2*C floats */,
                                                double* __restrict__ out_a, double* __restrict__ out_b) {
  const int G = C >> 2;  // channel groups (power of two)
  for (int i = threadIdx.x; i < 2 * C; i += THREADS) sm[i] = 0.f;
  __syncthreads();
  // lanes l and l' share a channel group iff l % G == l' % G (when G <= 32)
  if (G < 32) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      if (o >= G) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
          b[k] += __shfl_xor_sync(0xffffffffu, b[k], o);
        }
      }
    }
  }
  const int lane = threadIdx.x & 31;
  if (G >= 32 || lane < G) {
    const int c0 = (threadIdx.x * 4) % C;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&sm[c0 + k], a[k]);
      atomicAdd(&sm[C + c0 + k], b[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += THREADS) {
    atomicAdd(out_a + i, (double)sm[i]);
    atomicAdd(out_b + i, (double)sm[C + i]);
  }
}

}  // namespace dafk
