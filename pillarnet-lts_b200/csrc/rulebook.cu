// Rulebook construction for the sparse 2D backbone (replaces spconv's indice-pair generation used by
// SubMConv2d / SparseConv2d at det3d/models/backbones/base.py:38-63 and PillarResNet.py:87,95,103).
//
// Every active-site set is an occupancy bitmask + popcount prefix (mask_scan.cuh), the same structure
// pillarization produces, so "which row lives at (b,y,x)?" is two L2-resident loads and a popcount.
// Rulebooks are emitted output-stationary: nbr[o, ky*3+kx] = input row or -1, which is what the
// gather-GEMM kernels consume.  Row order everywhere is ascending b*H*W + y*W + x.
#include "common.cuh"
#include "mask_scan.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_subm_nbr(const uint32_t* __restrict__ words, const int* __restrict__ prefix,
           const int* __restrict__ coords, const int* __restrict__ num_rows, int m_cap, int H,
           int W, int* __restrict__ nbr) {
  const int n = min(*num_rows, m_cap);
  const long long total = (long long)n * 9;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(t / 9), k = (int)(t - (long long)o * 9);
    const int b = __ldg(coords + 3 * o), y = __ldg(coords + 3 * o + 1), x = __ldg(coords + 3 * o + 2);
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    int r = -1;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) r = pn_rank_of(words, prefix, (b * H + yy) * W + xx);
    nbr[t] = r;
  }
}

// Occupancy of the strided level, computed densely from the input bitmask: output (oy,ox) of a 3x3/s2/p1 conv is
// active iff any input cell of rows 2oy-1..2oy+1, columns 2ox-1..2ox+1 is.  No memset, no atomics, work independent
// of the number of active sites (the first version had every active input atomicOr its <= 4 outputs into a cleared
// mask: two launches), and it needs only the previous level's *mask*, so the masks of all levels can be chained.
// One thread per output WORD: for each run of its cells that lies in one output row it pulls the 2*len input bits of
// the three input rows (unaligned 64-bit extracts), ORs the rows, smears each bit onto its neighbours and keeps the
// even positions — ~100 integer instructions per 32 cells (the per-cell version cost 15x that at batch 8).
// Also clears the state of the scan that follows.
__device__ __forceinline__ unsigned long long extract_bits64(const uint32_t* __restrict__ words, long long n_words,
                                                             long long bit, int n) {
  // n (1..64) bits starting at linear bit offset `bit`; bits past the array read as 0
  const long long wi = bit >> 5;
  const int sh = (int)(bit & 31);
  const unsigned long long w0 = wi < n_words ? __ldg(words + wi) : 0u;
  const unsigned long long w1 = wi + 1 < n_words ? __ldg(words + wi + 1) : 0u;
  unsigned long long v = (w0 | (w1 << 32)) >> sh;
  if (sh > 0 && sh + n > 64) {
    const unsigned long long w2 = wi + 2 < n_words ? __ldg(words + wi + 2) : 0u;
    v |= w2 << (64 - sh);
  }
  return n >= 64 ? v : (v & ((1ull << n) - 1ull));
}

__device__ __forceinline__ uint32_t even_bits(unsigned long long x) {
  x &= 0x5555555555555555ull;
  x = (x | (x >> 1)) & 0x3333333333333333ull;
  x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
  x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
  x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
  return (uint32_t)x;
}

__global__ void __launch_bounds__(256)
k_down_mask(const uint32_t* __restrict__ in_words, long long n_in_words, int n_frames, int H, int W, int Ho, int Wo,
            uint32_t* __restrict__ out_words, long long n_out_words, int* __restrict__ scan_state, int n_state) {
  pn_detail::zero_scan_state(scan_state, n_state);
  const long long cells = (long long)n_frames * Ho * Wo;
  for (long long ow = (long long)blockIdx.x * blockDim.x + threadIdx.x; ow < n_out_words;
       ow += (long long)gridDim.x * blockDim.x) {
    long long c = ow * 32;
    const long long c_end = min(c + 32, cells);
    uint32_t word = 0u;
    while (c < c_end) {
      const int b = (int)(c / ((long long)Ho * Wo));
      const int r = (int)(c - (long long)b * Ho * Wo);
      const int oy = r / Wo, xa = r - oy * Wo;
      const int len = (int)min((long long)(Wo - xa), c_end - c);      // cells of this run (one output row)
      const int n = min(2 * len, W - 2 * xa);                          // input columns 2xa .. 2xa+n-1 exist
      unsigned long long main_bits = 0ull, left = 0ull;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = 2 * oy - 1 + ky;
        if (yy < 0 || yy >= H) continue;
        const long long row0 = ((long long)b * H + yy) * W;
        main_bits |= extract_bits64(in_words, n_in_words, row0 + 2 * xa, n);
        if (xa > 0) {
          const long long lb = row0 + 2 * xa - 1;
          left |= (__ldg(in_words + (lb >> 5)) >> (lb & 31)) & 1u;
        }
      }
      const unsigned long long t = main_bits | (main_bits >> 1) | (main_bits << 1) | left;
      uint32_t seg = even_bits(t);
      if (len < 32) seg &= (1u << len) - 1u;
      word |= seg << (int)(c - ow * 32);
      c += len;
    }
    out_words[ow] = word;
  }
}

__global__ void __launch_bounds__(256)
k_down_nbr(const uint32_t* __restrict__ in_words, const int* __restrict__ in_prefix, int H, int W,
           const int* __restrict__ out_coords, const int* __restrict__ out_num_rows, int out_m_cap,
           int* __restrict__ nbr) {
  const int n = min(*out_num_rows, out_m_cap);
  const long long total = (long long)n * 9;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(t / 9), k = (int)(t - (long long)o * 9);
    const int b = __ldg(out_coords + 3 * o), oy = __ldg(out_coords + 3 * o + 1),
              ox = __ldg(out_coords + 3 * o + 2);
    const int yy = 2 * oy - 1 + k / 3, xx = 2 * ox - 1 + k % 3;
    int r = -1;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      r = pn_rank_of(in_words, in_prefix, (b * H + yy) * W + xx);
    nbr[t] = r;
  }
}

// Non-overlapping strided conv, kernel = stride = s (spconv SparseConv2d(k=s, stride=s, padding=0), the lateral
// layers of the second stage: det3d/models/second_stage/bev_interpolation.py:66-72): output cell (oy, ox) of the
// (H / s, W / s) grid is active iff any input cell of its s x s block is; tap k = ky*s + kx reads (s*oy + ky, s*ox + kx).
__global__ void __launch_bounds__(256)
k_block_mask(const uint32_t* __restrict__ in_words, int n_frames, int H, int W, int s, int Ho, int Wo,
             uint32_t* __restrict__ out_words, long long n_out_words, int* __restrict__ scan_state, int n_state) {
  pn_detail::zero_scan_state(scan_state, n_state);
  const long long cells = (long long)n_frames * Ho * Wo;
  // thread = output cell, the warp's ballot is the output word (grid covers whole words: blockDim % 32 == 0)
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_out_words * 32;
       c += (long long)gridDim.x * blockDim.x) {
    bool on = false;
    if (c < cells) {
      const int b = (int)(c / ((long long)Ho * Wo));
      const int r = (int)(c - (long long)b * Ho * Wo);
      const int oy = r / Wo, ox = r - oy * Wo;
      for (int ky = 0; ky < s && !on; ++ky) {
        const long long base = ((long long)b * H + (long long)s * oy + ky) * W + (long long)s * ox;
        for (int kx = 0; kx < s; ++kx) {
          const long long bit = base + kx;
          if ((__ldg(in_words + (bit >> 5)) >> (bit & 31)) & 1u) { on = true; break; }
        }
      }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0) out_words[c >> 5] = word;
  }
}

__global__ void __launch_bounds__(256)
k_block_nbr(const uint32_t* __restrict__ in_words, const int* __restrict__ in_prefix, int H, int W, int s,
            const int* __restrict__ out_coords, const int* __restrict__ out_num_rows, int out_m_cap,
            int* __restrict__ nbr) {
  const int n = min(*out_num_rows, out_m_cap);
  const int taps = s * s;
  const long long total = (long long)n * taps;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(t / taps), k = (int)(t - (long long)o * taps);
    const int b = __ldg(out_coords + 3 * o), oy = __ldg(out_coords + 3 * o + 1), ox = __ldg(out_coords + 3 * o + 2);
    const int yy = s * oy + k / s, xx = s * ox + k % s;
    nbr[t] = pn_rank_of(in_words, in_prefix, (b * H + yy) * W + xx);
  }
}

// Both neighbour tables of every strided level in one launch: for output row o of level l, tap k reads
// (2oy-1+ky, 2ox-1+kx) of level l-1 (strided conv) and (oy+ky-1, ox+kx-1) of level l (submanifold convs).
struct NbrLevel {
  const uint32_t* in_words;  const int* in_prefix;  int H_in, W_in;
  const uint32_t* words;     const int* prefix;     int H, W;
  const int* coords;         const int* num_rows;   int m_cap;
  int* nbr_down;             int* nbr_subm;
};
struct NbrLevels {
  NbrLevel lv[PN_MAX_RULEBOOK_LEVELS];
  int n_levels;
};

__global__ void __launch_bounds__(256)
k_pyramid_nbr(const __grid_constant__ NbrLevels L) {
  for (int l = 0; l < L.n_levels; ++l) {
    const NbrLevel& q = L.lv[l];
    const int n = min(*q.num_rows, q.m_cap);
    const long long total = (long long)n * 9;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
      const int o = (int)(t / 9), k = (int)(t - (long long)o * 9);
      const int b = __ldg(q.coords + 3 * o), oy = __ldg(q.coords + 3 * o + 1), ox = __ldg(q.coords + 3 * o + 2);
      const int ky = k / 3, kx = k - ky * 3;
      int yy = 2 * oy - 1 + ky, xx = 2 * ox - 1 + kx;
      int r = -1;
      if (yy >= 0 && yy < q.H_in && xx >= 0 && xx < q.W_in)
        r = pn_rank_of(q.in_words, q.in_prefix, (b * q.H_in + yy) * q.W_in + xx);
      q.nbr_down[t] = r;
      yy = oy + ky - 1, xx = ox + kx - 1;
      r = -1;
      if (yy >= 0 && yy < q.H && xx >= 0 && xx < q.W) r = pn_rank_of(q.words, q.prefix, (b * q.H + yy) * q.W + xx);
      q.nbr_subm[t] = r;
    }
  }
}

// Dense NHWC gather tables (static per shape).  pi/po = 1 when the input/output rows index a
// zero-padded (H+2, W+2) map.
__global__ void __launch_bounds__(256)
k_dense_nbr_conv3(int n_frames, int H, int W, int stride, int Ho, int Wo, int pi, int po,
                  int* __restrict__ nbr) {
  const int Hop = Ho + 2 * po, Wop = Wo + 2 * po, Hip = H + 2 * pi, Wip = W + 2 * pi;
  const long long total = (long long)n_frames * Hop * Wop * 9;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long o = t / 9;
    const int k = (int)(t - o * 9);
    const int ox = (int)(o % Wop) - po;
    const int oy = (int)((o / Wop) % Hop) - po;
    const int b = (int)(o / ((long long)Wop * Hop));
    int v = -1;
    if (ox >= 0 && ox < Wo && oy >= 0 && oy < Ho) {
      const int yy = oy * stride - 1 + k / 3, xx = ox * stride - 1 + k % 3;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = (b * Hip + yy + pi) * Wip + xx + pi;
    }
    nbr[t] = v;
  }
}

__global__ void __launch_bounds__(256)
k_dense_nbr_deconv2(int n_frames, int H, int W, int pi, int po, int* __restrict__ nbr) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int Hop = Ho + 2 * po, Wop = Wo + 2 * po, Hip = H + 2 * pi, Wip = W + 2 * pi;
  const long long total = (long long)n_frames * Hop * Wop * 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long o = t >> 2;
    const int k = (int)(t & 3);
    const int ox = (int)(o % Wop) - po;
    const int oy = (int)((o / Wop) % Hop) - po;
    const int b = (int)(o / ((long long)Wop * Hop));
    int v = -1;
    if (ox >= 0 && ox < Wo && oy >= 0 && oy < Ho) {
      const int tap = (oy & 1) * 2 + (ox & 1);
      if (k == tap) v = (b * Hip + (oy >> 1) + pi) * Wip + (ox >> 1) + pi;
    }
    nbr[t] = v;
  }
}

inline int grid_for(long long work, int threads) {
  const int sms = pn_detail::sm_count();
  long long g = PN_DIVUP(work, (long long)threads);
  // Grids are sized from capacities (the live row count is on the device) and the kernels are grid-stride loops:
  // 8 blocks of 256 threads fill every thread slot of an SM, and a rulebook built on the side stream has to find room beside
  // the persistent conv CTAs of the main stream — thousands of empty blocks queued for 49 us there.
  const long long cap = (long long)(sms > 0 ? sms : 148) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

int pn_rulebook_subm3x3(const uint32_t* occ_words, const int* word_prefix, const int* coords,
                        const int* num_rows, int m_cap, int H, int W, int* nbr,
                        pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(occ_words && word_prefix && coords && num_rows && nbr && H > 0 && W > 0 && m_cap >= 0);
  if (m_cap == 0) return PN_OK;
  k_subm_nbr<<<grid_for((long long)m_cap * 9, 256), 256, 0, stream>>>(occ_words, word_prefix, coords,
                                                                     num_rows, m_cap, H, W, nbr);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

size_t pn_rulebook_down_scratch_bytes(int n_frames, int H_out, int W_out) {
  return pn_detail::scan_scratch_bytes(pn_detail::n_words((long long)n_frames * H_out * W_out));
}

int pn_rulebook_down3x3s2(const uint32_t* in_words, const int* in_prefix, const int* in_coords,
                          const int* in_num_rows, int in_m_cap, int n_frames, int H_in, int W_in,
                          uint32_t* out_words, int* out_prefix, int* out_coords, int* out_num_rows,
                          int out_m_cap, int* nbr, void* scratch, size_t scratch_bytes,
                          pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in_words && in_prefix && in_coords && in_num_rows && out_words && out_prefix &&
             out_coords && out_num_rows && nbr && scratch);
  PN_REQUIRE(n_frames >= 1 && H_in > 0 && W_in > 0 && in_m_cap >= 0 && out_m_cap >= 0);
  const int Ho = (H_in + 2 - 3) / 2 + 1, Wo = (W_in + 2 - 3) / 2 + 1;
  const long long nw = pn_detail::n_words((long long)n_frames * Ho * Wo);
  (void)in_coords; (void)in_num_rows; (void)in_m_cap;   // the occupancy comes from the input bitmask alone
  if (scratch_bytes < pn_detail::scan_scratch_bytes(nw)) return PN_ERR_WORKSPACE;
  k_down_mask<<<grid_for(nw, 256), 256, 0, stream>>>(in_words, pn_detail::n_words((long long)n_frames * H_in * W_in),
                                                    n_frames, H_in, W_in, Ho, Wo, out_words, nw,
                                                    reinterpret_cast<int*>(scratch), pn_detail::scan_state_words(nw));
  PN_CHECK_LAUNCH();
  int rc = pn_detail::mask_scan_emit(out_words, out_prefix, nw, Ho * Wo, Wo, out_coords, out_m_cap,
                                     out_num_rows, scratch, scratch_bytes, stream, /*state_is_zero=*/true);
  if (rc != PN_OK) return rc;
  if (out_m_cap > 0) {
    k_down_nbr<<<grid_for((long long)out_m_cap * 9, 256), 256, 0, stream>>>(
        in_words, in_prefix, H_in, W_in, out_coords, out_num_rows, out_m_cap, nbr);
    PN_CHECK_LAUNCH();
  }
  return PN_OK;
}

int pn_rulebook_block(const uint32_t* in_words, const int* in_prefix, int n_frames, int H_in, int W_in, int s,
                      uint32_t* out_words, int* out_prefix, int* out_coords, int* out_num_rows, int out_m_cap,
                      int* nbr, void* scratch, size_t scratch_bytes, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in_words && in_prefix && out_words && out_prefix && out_coords && out_num_rows && nbr && scratch);
  PN_REQUIRE(n_frames >= 1 && s >= 1 && s <= 8 && H_in >= s && W_in >= s && out_m_cap >= 0);
  const int Ho = H_in / s, Wo = W_in / s;       // floor((H - s) / s) + 1
  const long long nw = pn_detail::n_words((long long)n_frames * Ho * Wo);
  if (scratch_bytes < pn_detail::scan_scratch_bytes(nw)) return PN_ERR_WORKSPACE;
  k_block_mask<<<grid_for(nw * 32, 256), 256, 0, stream>>>(in_words, n_frames, H_in, W_in, s, Ho, Wo, out_words, nw,
                                                          reinterpret_cast<int*>(scratch), pn_detail::scan_state_words(nw));
  PN_CHECK_LAUNCH();
  int rc = pn_detail::mask_scan_emit(out_words, out_prefix, nw, Ho * Wo, Wo, out_coords, out_m_cap, out_num_rows,
                                     scratch, scratch_bytes, stream, /*state_is_zero=*/true);
  if (rc != PN_OK) return rc;
  if (out_m_cap > 0) {
    k_block_nbr<<<grid_for((long long)out_m_cap * s * s, 256), 256, 0, stream>>>(in_words, in_prefix, H_in, W_in, s,
                                                                               out_coords, out_num_rows, out_m_cap, nbr);
    PN_CHECK_LAUNCH();
  }
  return PN_OK;
}

size_t pn_rulebook_pyramid_scratch_bytes(int n_frames, int H0, int W0, int n_levels) {
  size_t bytes = 0;
  int H = H0, W = W0;
  for (int l = 0; l < n_levels; ++l) {
    H = (H + 2 - 3) / 2 + 1;
    W = (W + 2 - 3) / 2 + 1;
    bytes += pn_detail::scan_scratch_bytes(pn_detail::n_words((long long)n_frames * H * W));
  }
  return bytes;
}

int pn_rulebook_pyramid3x3s2(const uint32_t* words0, const int* prefix0, int n_frames, int H0, int W0,
                             int n_levels, const pn_rulebook_level* levels, void* scratch,
                             size_t scratch_bytes, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(words0 && prefix0 && levels && scratch && n_frames >= 1 && H0 > 0 && W0 > 0);
  PN_REQUIRE(n_levels >= 1 && n_levels <= PN_MAX_RULEBOOK_LEVELS);
  if (scratch_bytes < pn_rulebook_pyramid_scratch_bytes(n_frames, H0, W0, n_levels)) return PN_ERR_WORKSPACE;
  pn_detail::ScanJobs jobs = {};
  NbrLevels nl = {};
  jobs.n_jobs = nl.n_levels = n_levels;
  int state_words = 0;
  {
    int H = H0, W = W0;
    for (int l = 0; l < n_levels; ++l) {
      H = (H + 2 - 3) / 2 + 1;
      W = (W + 2 - 3) / 2 + 1;
      state_words += (int)(pn_detail::scan_scratch_bytes(pn_detail::n_words((long long)n_frames * H * W)) / sizeof(int));
    }
  }
  int* state = reinterpret_cast<int*>(scratch);
  const uint32_t* in_words = words0;
  const int* in_prefix = prefix0;
  int H = H0, W = W0, blocks = 0;
  long long nbr_items = 0;
  for (int l = 0; l < n_levels; ++l) {
    const pn_rulebook_level& lv = levels[l];
    PN_REQUIRE(lv.words && lv.prefix && lv.coords && lv.num_rows && lv.nbr_down && lv.nbr_subm && lv.m_cap >= 1);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long nw = pn_detail::n_words((long long)n_frames * Ho * Wo);
    // the first mask kernel clears the scan state of every level
    k_down_mask<<<grid_for(nw, 256), 256, 0, stream>>>(in_words, pn_detail::n_words((long long)n_frames * H * W),
                                                      n_frames, H, W, Ho, Wo, lv.words, nw,
                                                      reinterpret_cast<int*>(scratch), l == 0 ? state_words : 0);
    PN_CHECK_LAUNCH();
    pn_detail::ScanJob& j = jobs.job[l];
    j.words = lv.words;  j.n_words = nw;  j.n_tiles = pn_detail::scan_tiles(nw);
    j.state = reinterpret_cast<uint32_t*>(state);
    j.cells_per_frame = Ho * Wo;  j.W = Wo;  j.prefix = lv.prefix;  j.coords = lv.coords;  j.m_cap = lv.m_cap;
    j.num_out = lv.num_rows;
    jobs.block_begin[l] = blocks;
    blocks += j.n_tiles;
    state += pn_detail::scan_scratch_bytes(nw) / sizeof(int);
    NbrLevel& q = nl.lv[l];
    q.in_words = in_words;  q.in_prefix = in_prefix;  q.H_in = H;  q.W_in = W;
    q.words = lv.words;  q.prefix = lv.prefix;  q.H = Ho;  q.W = Wo;
    q.coords = lv.coords;  q.num_rows = lv.num_rows;  q.m_cap = lv.m_cap;
    q.nbr_down = lv.nbr_down;  q.nbr_subm = lv.nbr_subm;
    nbr_items += (long long)lv.m_cap * 9;
    in_words = lv.words;  in_prefix = lv.prefix;  H = Ho;  W = Wo;
  }
  jobs.block_begin[n_levels] = blocks;
  int rc = pn_detail::mask_scan_emit_multi(jobs, stream);
  if (rc != PN_OK) return rc;
  k_pyramid_nbr<<<grid_for(nbr_items, 256), 256, 0, stream>>>(nl);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_dense_nbr_table(int mode, int n_frames, int H_in, int W_in, int stride, int pad_flags, int* nbr,
                       pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(nbr && n_frames >= 1 && H_in > 0 && W_in > 0);
  const int pi = pad_flags & 1, po = (pad_flags >> 1) & 1;
  if (mode == 0) {
    PN_REQUIRE(stride == 1 || stride == 2);
    const int Ho = (H_in + 2 - 3) / stride + 1, Wo = (W_in + 2 - 3) / stride + 1;
    const long long total = (long long)n_frames * (Ho + 2 * po) * (Wo + 2 * po) * 9;
    k_dense_nbr_conv3<<<grid_for(total, 256), 256, 0, stream>>>(n_frames, H_in, W_in, stride, Ho, Wo, pi, po, nbr);
  } else if (mode == 1) {
    const long long total = (long long)n_frames * (2 * H_in + 2 * po) * (2 * W_in + 2 * po) * 4;
    k_dense_nbr_deconv2<<<grid_for(total, 256), 256, 0, stream>>>(n_frames, H_in, W_in, pi, po, nbr);
  } else {
    return PN_ERR_INVALID_ARG;
  }
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
