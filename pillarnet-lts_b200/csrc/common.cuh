// Shared device/host helpers for libpillarnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/pillarnet_b200.h"

#define PN_DIVUP(a, b) (((a) + (b) - 1) / (b))

// Every entry point returns one of the PN_* codes; kernels are launched on the caller's stream and a
// launch failure is reported (never exit(), never throw across the ABI).
#define PN_CHECK_LAUNCH()                                  \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return pn_detail::fail(e__);   \
    pn_detail::count_launch();                             \
  } while (0)

#define PN_CUDA(call)                                      \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return pn_detail::fail(e__);   \
  } while (0)

#define PN_REQUIRE(cond)                                   \
  do {                                                     \
    if (!(cond)) return PN_ERR_INVALID_ARG;                \
  } while (0)

namespace pn_detail {
int fail(cudaError_t e);     // records the CUDA error string, returns PN_ERR_CUDA
int sm_count();              // cached multiProcessorCount of the current device
void count_launch();         // bumps the process-wide kernel-launch counter (pn_launch_count)

// cudaFuncSetAttribute(..MaxDynamicSharedMemorySize..) is a per-DEVICE function attribute: a process that drives several
// devices must opt in on each of them.  `static PerDeviceOnce once; if (once.need()) cudaFuncSetAttribute(...)`.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Number of 32-bit occupancy words for a (B,H,W) raster.
__host__ __device__ inline long long n_words(long long cells) { return (cells + 31) >> 5; }
}  // namespace pn_detail

// ---- occupancy-bitmask rank lookup ------------------------------------------------------------
// A raster of cells is described by 32-bit occupancy words plus the exclusive popcount prefix of
// those words; the rank of an occupied cell in ascending cell order is then
//   prefix[cell >> 5] + popc(words[cell >> 5] & lower_bits(cell & 31)).
// This reproduces the reference's cumsum-over-mask ordering (pillar_utils.py:43-45) in 1 bit/cell.
__device__ __forceinline__ int pn_rank_of(const uint32_t* __restrict__ words,
                                          const int* __restrict__ prefix, int cell) {
  const uint32_t w = __ldg(words + (cell >> 5));
  const uint32_t bit = 1u << (cell & 31);
  if (!(w & bit)) return -1;
  return __ldg(prefix + (cell >> 5)) + __popc(w & (bit - 1u));
}
