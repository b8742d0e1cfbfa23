// Occupancy-bitmask scan: exclusive popcount prefix per 32-bit word, total count, and emission of
// [b,y,x] coordinates in ascending cell order.  Shared by pillarization and the strided rulebook.
//
// This is the 1-bit/cell replacement for the reference's B*H*W-wide bool mask + int32 cumsum
// (det3d/ops/pillar_ops/pillar_utils.py:38-48, pillar_ops_gpu.cu:60-78).
#pragma once
#include "common.cuh"

namespace pn_detail {

constexpr int kScanThreads = 256;
// One word (32 cells) per thread: occupied cells cluster spatially, so giving a thread several words
// serialises the coordinate emission of a dense region in a few threads (measured: 65 us -> see profiles/).
constexpr int kWordsPerThread = 1;
constexpr int kScanTile = kScanThreads * kWordsPerThread;  // 256 words = 8192 cells per CTA

inline int scan_tiles(long long n_words) { return (int)PN_DIVUP(n_words, (long long)kScanTile); }
// scratch: the single-pass scan's state = [ticket, status of tile 0 .. n_tiles-1] (32-bit words); the size formula
// predates it and stays, callers have allocated by it
inline size_t scan_scratch_bytes(long long n_words) {
  return sizeof(int) * 2 * (size_t)(scan_tiles(n_words) + 1);
}
inline int scan_state_words(long long n_words) { return scan_tiles(n_words) + 1; }

// words/prefix: n_words entries. coords: (m_cap,3) or nullptr. cells_per_frame = H*W.
// One launch (chained scan with decoupled look-back).  `scratch` holds the scan state and MUST be zero when the
// kernel starts: state_is_zero = true when the caller's preceding kernel on `stream` cleared scan_state_words()
// words of it (zero_scan_state below), false to have a memset node issued here.
int mask_scan_emit(const uint32_t* words, int* prefix, long long n_words, int cells_per_frame,
                   int W, int* coords, int m_cap, int* num_out, void* scratch,
                   size_t scratch_bytes, cudaStream_t stream, bool state_is_zero = false);

// Several independent masks in one launch (the strided rulebook levels): block b works on job j with
// block_begin[j] <= b < block_begin[j+1].  Every job's state must be zero.
struct ScanJob {
  const uint32_t* words;
  long long n_words;
  int n_tiles;
  uint32_t* state;
  int cells_per_frame, W;
  int* prefix;
  int* coords;
  int m_cap;
  int* num_out;
};
constexpr int kMaxScanJobs = 4;
struct ScanJobs {
  ScanJob job[kMaxScanJobs];
  int block_begin[kMaxScanJobs + 1];
  int n_jobs;
};
int mask_scan_emit_multi(const ScanJobs& jobs, cudaStream_t stream);

// for the kernel that runs right before the scan: all threads of the grid call it
__device__ __forceinline__ void zero_scan_state(int* state, int n_state) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_state; i += gridDim.x * blockDim.x) state[i] = 0;
}

}  // namespace pn_detail
