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
// scratch: tile sums + tile offsets
inline size_t scan_scratch_bytes(long long n_words) {
  return sizeof(int) * 2 * (size_t)(scan_tiles(n_words) + 1);
}

// words/prefix: n_words entries. coords: (m_cap,3) or nullptr. cells_per_frame = H*W.
int mask_scan_emit(const uint32_t* words, int* prefix, long long n_words, int cells_per_frame,
                   int W, int* coords, int m_cap, int* num_out, void* scratch,
                   size_t scratch_bytes, cudaStream_t stream);

}  // namespace pn_detail
