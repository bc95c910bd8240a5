// Error plumbing and bookkeeping of the dafk library.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

namespace dafk {

static thread_local char g_err[512] = "no error";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky-less error so later calls report their own
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return DAFK_ERR_CUDA;
  }
  return DAFK_OK;
}

}  // namespace dafk

extern "C" {

const char* dafk_last_error_string(void) { return dafk::g_err; }
int dafk_version(void) { return 100; }
int64_t dafk_launch_count(void) { return dafk::g_launches.load(); }

int dafk_memset_zero(void* p, int64_t bytes, void* stream) {
  DAFK_REQUIRE(bytes >= 0, DAFK_ERR_BAD_ARG, "dafk_memset_zero: negative size");
  if (bytes == 0) return DAFK_OK;
  DAFK_REQUIRE(p != nullptr, DAFK_ERR_BAD_ARG, "dafk_memset_zero: null pointer");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, dafk::as_stream(stream));
  DAFK_REQUIRE(e == cudaSuccess, DAFK_ERR_CUDA, "dafk_memset_zero: %s", cudaGetErrorString(e));
  return DAFK_OK;
}

}  // extern "C"
