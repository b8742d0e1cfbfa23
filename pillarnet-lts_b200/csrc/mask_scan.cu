#include "mask_scan.cuh"

#include <atomic>
#include <mutex>
#include <string>

namespace pn_detail {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int fail(cudaError_t e) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
  return PN_ERR_CUDA;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached = n;
  return n;
}

const char* last_error_cstr() { return g_last_error.c_str(); }

// ---- kernels ---------------------------------------------------------------------------------

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // 256 threads: warp shuffle scan + one smem hop.
  __shared__ int warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int warp_off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    int s = warp_sums[w];
    if (w < warp) warp_off += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return warp_off + incl - v;
}

// Single-pass scan + emission.  Tile = 256 words; tiles are handed out by an atomic ticket (a CTA holding ticket t
// knows every earlier tile is already running, so waiting on them cannot deadlock), each publishes its word-count
// aggregate, then its inclusive prefix once warp 0 has looked back over its predecessors 32 at a time
// (status word = flag << 30 | value; value < 2^30 cells).  Replaces three launches (tile sums, scan of the sums,
// emission): beside the persistent conv CTAs every extra dependent launch of the rulebook chain cost 5-15 us.
constexpr int kSinglePassMaxTiles = 1024;
constexpr uint32_t kFlagAgg = 1u << 30, kFlagPre = 2u << 30, kFlagMask = 3u << 30, kValMask = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool LOOKBACK>
__device__ __forceinline__ void scan_emit_tile(const uint32_t* __restrict__ words, long long n_words, int n_tiles,
                                               uint32_t* __restrict__ state, int cells_per_frame, int W,
                                               int* __restrict__ prefix, int* __restrict__ coords, int m_cap,
                                               int* __restrict__ num_out) {
  __shared__ int s_tile, s_excl;
  if (LOOKBACK) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(state, 1u);
    __syncthreads();
  }
  // !LOOKBACK: `state` holds the exclusive offsets of the tiles (large masks, see mask_scan_emit)
  const int tile = LOOKBACK ? s_tile : (int)blockIdx.x;
  const long long w = (long long)tile * kScanTile + threadIdx.x;
  uint32_t bits = w < n_words ? __ldg(words + w) : 0u;
  int tot;
  const int ex = block_exclusive_scan(__popc(bits), &tot);
  if (!LOOKBACK) {
    if (threadIdx.x == 0) s_excl = (int)state[tile];
  } else if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t* st = state + 1;
    int excl = 0;
    if (tile > 0) {
      if (lane == 0) st_volatile_u32(st + tile, kFlagAgg | (uint32_t)tot);
      int base = tile - 1;
      while (true) {
        const int idx = base - lane;
        uint32_t v = kFlagPre;                 // before tile 0: an empty prefix
        if (idx >= 0) {
          do { v = ld_volatile_u32(st + idx); } while ((v & kFlagMask) == 0u);
        }
        const unsigned is_p = __ballot_sync(0xffffffffu, (v & kFlagPre) != 0u);
        const int upto = is_p ? __ffs(is_p) - 1 : 31;      // nearest predecessor that knows its whole prefix
        int val = lane <= upto ? (int)(v & kValMask) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        excl += val;
        if (is_p) break;
        base -= 32;
      }
    }
    if (lane == 0) {
      st_volatile_u32(st + tile, kFlagPre | (uint32_t)(excl + tot));
      s_excl = excl;
      if (tile == n_tiles - 1) *num_out = excl + tot;
    }
  }
  __syncthreads();
  int run = s_excl + ex;
  if (w < n_words) prefix[w] = run;
  if (coords == nullptr || bits == 0u) return;
  // all cells of a word share most of the decomposition: the word never straddles more than 2 rows
  const int cell0 = (int)(w * 32);
  int b = cell0 / cells_per_frame;
  int r = cell0 - b * cells_per_frame;
  int y = r / W;
  int x = r - y * W;
  int prev = 0;
  while (bits) {
    const int bit = __ffs(bits) - 1;
    bits &= bits - 1;
    x += bit - prev;
    prev = bit;
    while (x >= W) { x -= W; if (++y * W >= cells_per_frame) { y = 0; ++b; } }
    if (run < m_cap) {
      int* o = coords + 3ll * run;
      o[0] = b;
      o[1] = y;
      o[2] = x;
    }
    ++run;
  }
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_emit(const uint32_t* __restrict__ words, long long n_words, int n_tiles, uint32_t* __restrict__ state,
            int cells_per_frame, int W, int* __restrict__ prefix, int* __restrict__ coords, int m_cap,
            int* __restrict__ num_out) {
  scan_emit_tile<true>(words, n_words, n_tiles, state, cells_per_frame, W, prefix, coords, m_cap, num_out);
}

// Large masks (thousands of tiles: batches of 8+ frames): with 1 KB tiles the look-back chain's hop latency adds up
// (measured at 4050 tiles: pillarize 137 -> 160 us), so they keep the three-launch form: per-tile sums, one CTA
// scanning the sums, emission from the precomputed offsets.
__global__ void __launch_bounds__(kScanThreads)
k_tile_sums(const uint32_t* __restrict__ words, long long n_words, int* __restrict__ tile_sums) {
  const long long w = (long long)blockIdx.x * kScanTile + threadIdx.x;
  const int c = w < n_words ? __popc(__ldg(words + w)) : 0;
  int tot;
  (void)block_exclusive_scan(c, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k_tile_offsets(const int* __restrict__ tile_sums, int n_tiles, int* __restrict__ tile_offsets,
               int* __restrict__ num_out) {
  int carry = 0;
  for (int base = 0; base < n_tiles; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? tile_sums[i] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, &tot);
    if (i < n_tiles) tile_offsets[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *num_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
k_emit(const uint32_t* __restrict__ words, long long n_words, int n_tiles, const int* __restrict__ tile_offsets,
       int cells_per_frame, int W, int* __restrict__ prefix, int* __restrict__ coords, int m_cap) {
  scan_emit_tile<false>(words, n_words, n_tiles, reinterpret_cast<uint32_t*>(const_cast<int*>(tile_offsets)),
                        cells_per_frame, W, prefix, coords, m_cap, nullptr);
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_emit_multi(const __grid_constant__ ScanJobs J) {
  int j = 0;
#pragma unroll
  for (int i = 1; i < kMaxScanJobs; ++i)
    if (i < J.n_jobs && (int)blockIdx.x >= J.block_begin[i]) j = i;
  const ScanJob& q = J.job[j];
  scan_emit_tile<true>(q.words, q.n_words, q.n_tiles, q.state, q.cells_per_frame, q.W, q.prefix, q.coords, q.m_cap,
                       q.num_out);
}

int mask_scan_emit_multi(const ScanJobs& jobs, cudaStream_t stream) {
  if (jobs.n_jobs < 1 || jobs.n_jobs > kMaxScanJobs) return PN_ERR_INVALID_ARG;
  k_scan_emit_multi<<<jobs.block_begin[jobs.n_jobs], kScanThreads, 0, stream>>>(jobs);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int mask_scan_emit(const uint32_t* words, int* prefix, long long n_words, int cells_per_frame,
                   int W, int* coords, int m_cap, int* num_out, void* scratch,
                   size_t scratch_bytes, cudaStream_t stream, bool state_is_zero) {
  if (n_words <= 0) return PN_ERR_INVALID_ARG;
  if (scratch_bytes < scan_scratch_bytes(n_words)) return PN_ERR_WORKSPACE;
  const int n_tiles = scan_tiles(n_words);
  if (n_tiles > kSinglePassMaxTiles) {
    int* tile_sums = reinterpret_cast<int*>(scratch);
    int* tile_offsets = tile_sums + (n_tiles + 1);
    k_tile_sums<<<n_tiles, kScanThreads, 0, stream>>>(words, n_words, tile_sums);
    PN_CHECK_LAUNCH();
    k_tile_offsets<<<1, kScanThreads, 0, stream>>>(tile_sums, n_tiles, tile_offsets, num_out);
    PN_CHECK_LAUNCH();
    k_emit<<<n_tiles, kScanThreads, 0, stream>>>(words, n_words, n_tiles, tile_offsets, cells_per_frame, W, prefix,
                                                 coords, m_cap);
    PN_CHECK_LAUNCH();
    return PN_OK;
  }
  if (!state_is_zero) {
    PN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int) * (size_t)scan_state_words(n_words), stream));
  }
  k_scan_emit<<<n_tiles, kScanThreads, 0, stream>>>(words, n_words, n_tiles, reinterpret_cast<uint32_t*>(scratch),
                                                    cells_per_frame, W, prefix, coords, m_cap, num_out);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // namespace pn_detail

extern "C" {
int pn_abi_version(void) { return 1; }
const char* pn_last_error(void) { return pn_detail::last_error_cstr(); }
int pn_device_sm_count(void) { return pn_detail::sm_count(); }
long long pn_launch_count(void) { return pn_detail::launches(); }
size_t pn_mask_words(int n_frames, int H, int W) {
  return (size_t)pn_detail::n_words((long long)n_frames * H * W);
}
}
