// Gather-GEMM convolution on 5th-gen tensor cores (PN_IMPL_TCGEN05) for sm_100a.
//
//   out[o, n] = act( (sum_t sum_c in[nbr[o,t], c] * W[n, t*cin + c]) * scale[n] + shift[n] + res[o,n] )
//
// Replaces spconv's implicit-GEMM SubMConv2d / SparseConv2d (det3d/models/backbones/base.py:38-63,
// PillarResNet.py:87,95,103) and the cuDNN 3x3 / transposed-2x2 convs of the dense BEV neck and
// head (necks/rpn.py:147-207, bbox_heads/center_head.py:27-35,101-112).  Only the dense channel
// contraction touches the tensor cores; the gather stays a gather.
//
// Structure (one persistent CTA per SM, 672 threads, warp-specialised):
//   warps 0-15  A producers: gather 128 activation rows x 64 channels (bf16, 128 B per row) per K-chunk with
//               16-byte cp.async into a 128B-swizzled K-major tile (rows that are missing in the rulebook are
//               zero-filled by cp.async src-size 0) and signal the stage's mbarrier with
//               cp.async.mbarrier.arrive.noinc; thread 0 also issues the TMA load of the weight tile
//               (BLOCK_N x 64, SWIZZLE_128B tensor map) and arms the barrier with the TMA byte count.  Rulebook
//               rows travel global -> registers -> shared two tiles ahead.  (Opt-in PN_CONV_TMA_GATHER=1: the
//               activations through TMA gather4 issued by 32 lanes spread over these warps; measured slower.)
//   warp 20     allocates TMEM (2 x BLOCK_N fp32 columns: double-buffered accumulator); the warp walks the issue
//               loop converged and one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16, kind::f16,
//               bf16 x bf16 -> fp32) and tcgen05.commit to release smem stages / publish the accumulator.
//   warps 16-19 epilogue: tcgen05.ld 32 lanes x 32 columns, fused scale/shift (BN+bias), residual, ReLU, convert,
//               row-contiguous stores (or the scattering store of the GEMM-form transposed conv); overlaps with
//               the next tile's MMAs.
// K = taps*cin is walked in 64-element chunks (zero-padded weights), so a chunk may straddle taps
// (cin = 32) — each 16-byte piece belongs to exactly one tap because cin % 8 == 0.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;       // bf16 elements = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * 128;
constexpr int kProducerThreads = 512;   // 16 warps: in-flight cp.async bytes scale with the number of issuing warps
constexpr int kEpilogueThreads = 128;
constexpr int kProducerWarps = kProducerThreads / 32;
constexpr int kEpilogueWarp0 = kProducerWarps;        // 4 epilogue warps; index % 4 == TMEM lane quarter
constexpr int kMmaWarp = kProducerWarps + 4;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kMaxTaps = 9;

using namespace pn_tc;

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = BN.
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct KArgs {
  const __nv_bfloat16* in;
  int in_ld;
  const int* nbr;
  int taps;
  int n_chunks;  // k_pad / 64
  const float* scale;
  const float* shift;
  const void* residual;
  int res_ld;
  void* out;
  int out_f32;   // 1: float output, 0: bf16
  int out_ld;
  int out_coff;
  int relu;
  const int* num_rows;
  int rows_cap;
  int cin;
  int cout;
  int out_hp, out_wp;  // != 0: zero the border rows of a padded output map
  int in_rows;         // allocated rows of `in` (TMA gather path: index >= in_rows reads zeros)
  int cin_shift;       // log2(cin) when cin is a power of two, else -1
  int dc_cout, dc_hp_in, dc_wp_in;   // transposed 2x2/s2 conv as one GEMM (pn_conv_args.deconv_*): 0 = off
  int tma_store;       // bf16 output rows through shared memory + TMA tensor stores (tmap_o is valid)
  unsigned long long* dbg;   // development aid (PN_CONV_TIMELINE=1): per-CTA %globaltimer milestones
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait-time accounting (SM clocks) for the timeline: where each role of the pipeline stalls
#define PN_WAIT_T0() const long long _w0 = P.dbg ? clock64() : 0
#define PN_WAIT_ACC(var) do { if (P.dbg) var += clock64() - _w0; } while (0)
#define PN_WAIT_OUT(slot, var) do { if (P.dbg) P.dbg[blockIdx.x * 16 + slot] = (unsigned long long)(var); } while (0)
#define PN_DBG(slot)                                                        \
  do {                                                                      \
    if (P.dbg) P.dbg[blockIdx.x * 16 + slot] = gtime_ns();                   \
  } while (0)

template <int BN, int STAGES>
struct Smem {
  alignas(1024) uint8_t a[STAGES][A_STAGE_BYTES];
  alignas(1024) uint8_t b[STAGES][BN * 128];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  int nbr[3][BLOCK_M * kMaxTaps];   // rulebook rows of the current tile and the next two (ring)
  alignas(16) float2 ss[BN];   // {scale, shift} per output column, read two columns per LDS.128 (a broadcast LDS.32
                               // costs the shared-memory pipe a full wavefront)
  // Epilogue staging for TMA tensor stores (see conv_dense_tc.cu): 32 rows x 32 bf16 columns (2 KB, SWIZZLE_64B),
  // two buffers per epilogue warp.  Here the direct row-per-thread stores cost twice: their 32 lines per warp
  // instruction occupy the same LSU that the producers' gathers are bound by.
  alignas(1024) uint8_t stage_out[BN >= 32 ? 4 * 2 * 2048 : 16];
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <int BN>
constexpr int tmem_cols() {
  return 2 * BN < 32 ? 32 : 2 * BN;
}

template <int BN, int STAGES, bool TMA_A>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
          const __grid_constant__ CUtensorMap tmap_o, const KArgs P) {
  extern __shared__ uint8_t smem_raw[];
  using S = Smem<BN, STAGES>;
  S& sm = *reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TCOLS = tmem_cols<BN>();
  if (threadIdx.x == 0) PN_DBG(0);
  pdl_launch_dependents();   // our successor may be scheduled; it waits for this grid's completion itself
  // Barriers and TMEM do not depend on the predecessor: set up while its last CTAs are still running.  Everything
  // after the wait reads its results (live row count, rulebook, activations) or overwrites buffers it may still read.
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&sm.full[s], TMA_A ? 1 : kProducerThreads + 1);
        mbar_init(&sm.empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sm.tmem_full[i], 1);
        mbar_init(&sm.tmem_empty[i], kEpilogueThreads);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(&sm.tmem_base);
  }
  pdl_wait();
  const int rows = P.num_rows ? min(*P.num_rows, P.rows_cap) : P.rows_cap;
  const int n_n_tiles = (P.cout + BN - 1) / BN;
  // Balanced schedule: every CTA owns an equal, contiguous share of the rows and walks it in 128-row tiles (the
  // last one partial).  The kernel is bound by the gathers, whose cost is proportional to live rows, not by the
  // MMAs: with round-robin 128-row tiles, 235 tiles on 148 SMs made 87 CTAs work twice as long as the rest.
  // Only when one N tile covers cout: with several, a CTA would gather its rows and stream ALL the weights once per
  // N tile (measured: 30 -> 36 us on the 256->256 stage), so those layers keep round-robin (row tile, N tile) units.
  const bool balanced = n_n_tiles == 1;
  const int share = max(64, (((rows + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7);
  const int row_begin = balanced ? min(rows, (int)blockIdx.x * share) : 0;
  const int row_end = balanced ? min(rows, row_begin + share) : rows;
  const int tiles_all = ((rows + BLOCK_M - 1) / BLOCK_M) * n_n_tiles;
  // this CTA's k-th unit is global tile tile0 + k*tstep; n_tiles = number of units it owns
  const int tile0 = balanced ? 0 : (int)blockIdx.x, tstep = balanced ? 1 : (int)gridDim.x;
  const int n_tiles = balanced ? (row_end - row_begin + BLOCK_M - 1) / BLOCK_M
                               : (tiles_all > tile0 ? (tiles_all - tile0 + tstep - 1) / tstep : 0);

  // Rulebook rows travel global -> registers -> shared two tiles ahead of their use (two register sets, three
  // shared buffers): short-K layers (32/64 channels, < 1 us per tile) would otherwise expose the load latency per tile.
  constexpr int kNbrPerThread = (BLOCK_M * kMaxTaps + kProducerThreads - 1) / kProducerThreads;
  const int tid = threadIdx.x;
  const int nbr_elems = BLOCK_M * P.taps;
  auto fetch_nbr = [&](int tile, int (&regs)[kNbrPerThread]) {
    const int m_tile = (tile0 + tile * tstep) / n_n_tiles;   // `tile` = local unit index
    const int row0 = row_begin + m_tile * BLOCK_M;
#pragma unroll
    for (int q = 0; q < kNbrPerThread; ++q) {
      const int i = tid + q * kProducerThreads;
      int src = -1;
      if (i < nbr_elems && tile < n_tiles) {
        const int r = i / P.taps, t = i - r * P.taps;
        const int row = row0 + r;
        if (row < row_end) src = P.nbr ? __ldg(P.nbr + (long long)row * P.taps + t) : row;
      }
      regs[q] = src;
    }
  };
  int nbr_r0[kNbrPerThread], nbr_r1[kNbrPerThread];
  int ra0[kMaxTaps], rb0[kMaxTaps];
  // TMA gather4 path: warp w feeds rows 8w..8w+7 of a tile; lane l < 8 keeps row 8w+l's entries (one per tap),
  // lanes 0 and 4 collect four rows' entries by shuffle and issue one gather4 each per chunk.
  auto fetch_rows = [&](int tile, int (&r)[kMaxTaps]) {
    const int m_tile = (tile0 + tile * tstep) / n_n_tiles;
    const int row = row_begin + m_tile * BLOCK_M + warp * 8 + (lane & 7);
    const bool ok = tile < n_tiles && row < row_end && lane < 8;
#pragma unroll
    for (int u = 0; u < kMaxTaps; ++u) {
      int v = -1;
      if (ok && u < P.taps) v = P.nbr ? __ldg(P.nbr + (long long)row * P.taps + u) : row;
      r[u] = v >= 0 ? v : P.in_rows;     // outside the tensor => TMA zero-fills the row
    }
  };
  if (warp < kProducerWarps) {
    if constexpr (TMA_A) {
      fetch_rows(0, ra0);
    } else {
      fetch_nbr(0, nbr_r0);
      fetch_nbr(1, nbr_r1);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  if (threadIdx.x == 0) PN_DBG(1);

  if (warp < kProducerWarps) {
    // ===================== A producers (+ weight TMA) =====================
    const int piece = tid & 7, rg = tid >> 3;          // rg in [0,64): rows rg + 64*i, i < 2
    const int k_total = P.taps * P.cin;
    // swizzled destination of this thread's 16-byte piece inside a stage (row & 7 == rg & 7 for all i)
    const uint32_t dst_off = (uint32_t)rg * 128u + (uint32_t)((piece ^ (rg & 7)) << 4);
    const uint32_t in_ld_bytes = (uint32_t)P.in_ld * 2u;
    const char* in_bytes = reinterpret_cast<const char*>(P.in);
    uint32_t g = 0;
    long long w_empty = 0, w_bar = 0;
    if constexpr (TMA_A) {
      const int cps = P.cin >> 6;
      uint32_t s = 0, ph = 0;
      const bool issuer = (lane & 27) == 0;          // lanes 0 and 4
      const uint32_t a_off = (uint32_t)(warp * 2 + (lane >> 2)) * 512u;
      auto run_tile_tma = [&](int tile, const int (&r)[kMaxTaps]) {
        const int gt = tile0 + tile * tstep;
        const int n_tile = gt - (gt / n_n_tiles) * n_n_tiles;
        int kc = 0;
#pragma unroll
        for (int u = 0; u < kMaxTaps; ++u) {
          if (u < P.taps) {
            const int base = lane & 4;
            const int i0 = __shfl_sync(0xffffffffu, r[u], base), i1 = __shfl_sync(0xffffffffu, r[u], base + 1);
            const int i2 = __shfl_sync(0xffffffffu, r[u], base + 2), i3 = __shfl_sync(0xffffffffu, r[u], base + 3);
            for (int j = 0; j < cps; ++j, ++kc) {
              { PN_WAIT_T0(); mbar_wait(&sm.empty[s], ph ^ 1u); PN_WAIT_ACC(w_empty); }
              if (tid == 0) {
                mbar_arrive_expect_tx(&sm.full[s], A_STAGE_BYTES + BN * 128);
                tma_load_2d(smem_u32(sm.b[s]), &tmap_w, kc * BLOCK_K, n_tile * BN, &sm.full[s]);
              }
              if (issuer) tma_gather4(smem_u32(sm.a[s]) + a_off, &tmap_a, j * BLOCK_K, i0, i1, i2, i3, &sm.full[s]);
              if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
          }
        }
      };
      for (int tile = 0; tile < n_tiles; tile += 2) {
        fetch_rows(tile + 1, rb0);
        run_tile_tma(tile, ra0);
        if (tile + 1 < n_tiles) {
          fetch_rows(tile + 2, ra0);
          run_tile_tma(tile + 1, rb0);
        }
      }
      if (tid == 0) { PN_WAIT_OUT(8, w_empty); PN_WAIT_OUT(9, w_bar); }
    } else {
    auto park_nbr = [&](int buf, const int (&regs)[kNbrPerThread]) {
#pragma unroll
      for (int q = 0; q < kNbrPerThread; ++q) {
        const int i = tid + q * kProducerThreads;
        if (i < nbr_elems) sm.nbr[buf][i] = regs[q];
      }
    };
    park_nbr(0, nbr_r0);
    named_bar_sync(1, kProducerThreads);
    // one tile: rows of tile+2 start travelling into `rf`; at the end `rp` (tile+1, fetched a tile ago) is parked
    auto run_tile = [&](int tile, int (&rf)[kNbrPerThread], const int (&rp)[kNbrPerThread]) {
      const int gt = tile0 + tile * tstep;
      const int m_tile = gt / n_n_tiles, n_tile = gt - m_tile * n_n_tiles;
      (void)m_tile;
      const int* s_nbr = sm.nbr[tile % 3];
      fetch_nbr(tile + 2, rf);
      {
      for (int kc = 0; kc < P.n_chunks; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          { PN_WAIT_T0(); mbar_wait(&sm.empty[s], ph ^ 1u); PN_WAIT_ACC(w_empty); }
          if (tid == 0) {
            mbar_arrive_expect_tx(&sm.full[s], BN * 128);
            tma_load_2d(smem_u32(sm.b[s]), &tmap_w, kc * BLOCK_K, n_tile * BN, &sm.full[s]);
          }
          const int k = kc * BLOCK_K + piece * 8;
          int t, c;
          if (P.cin_shift >= 0) { t = k >> P.cin_shift; c = k & (P.cin - 1); }
          else { t = k / P.cin; c = k - t * P.cin; }
          const bool k_ok = k < k_total;
          const uint32_t dst = smem_u32(sm.a[s]) + dst_off;
          const int* nb = s_nbr + rg * P.taps + t;
          const uint32_t coff = (uint32_t)c * 2u;
#pragma unroll
          for (int i = 0; i < BLOCK_M / (kProducerThreads / 8); ++i) {
            const int src = k_ok ? nb[i * (kProducerThreads / 8) * P.taps] : -1;
            // 32-bit byte offset (the host checks in_rows*in_ld*2 < 4 GiB); missing neighbour => zero fill
            const uint32_t off = src >= 0 ? (uint32_t)src * in_ld_bytes + coff : 0u;
            cp_async16(dst + (uint32_t)i * ((uint32_t)(kProducerThreads / 8) * 128u), in_bytes + off,
                       src >= 0 ? 16u : 0u);
          }
            cp_async_mbar_arrive_noinc(&sm.full[s]);
        }
      }
      { PN_WAIT_T0();
      park_nbr((tile + 1) % 3, rp);          // read from the tile after this barrier on; its old content (tile-2) is dead
      named_bar_sync(1, kProducerThreads);
      PN_WAIT_ACC(w_bar); }
    };
    for (int tile = 0; tile < n_tiles; tile += 2) {
      run_tile(tile, nbr_r0, nbr_r1);
      if (tile + 1 < n_tiles) run_tile(tile + 1, nbr_r1, nbr_r0);
    }
    if (tid == 0) { PN_WAIT_OUT(8, w_empty); PN_WAIT_OUT(9, w_bar); }
    if constexpr (!TMA_A) cp_async_wait_all();   // nothing of this CTA's may still be in flight at exit
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop converged (waits, stage counters and descriptors are warp-uniform and live in
    // uniform registers) and one elected lane issues the tcgen05 instructions: entering with a single active lane
    // made the compiler wrap every UTCHMMA in an ELECT / broadcast / retry sequence, ~90 dependent instructions
    // per 64-channel chunk against the 180-256 clk its four MMAs take (see conv_dense_tc.cu).
    {
      const bool issuer = elect_one();
      constexpr uint32_t idesc = make_idesc<BN>();
      const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(sm.a[0]));
      const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(sm.b[0]));
      constexpr uint32_t kAStep = (uint32_t)A_STAGE_BYTES >> 4, kBStep = (uint32_t)(BN * 128) >> 4;
      uint32_t s = 0, ph = 0, tcount = 0;
      bool first_chunk = true;
      long long w_full = 0, w_tempty = 0;
      for (int tile = 0; tile < n_tiles; ++tile, ++tcount) {
        const uint32_t acc = tcount & 1u, acc_ph = (tcount >> 1) & 1u;
        { PN_WAIT_T0(); mbar_wait(&sm.tmem_empty[acc], acc_ph ^ 1u); PN_WAIT_ACC(w_tempty); }
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < P.n_chunks; ++kc) {
          { PN_WAIT_T0(); mbar_wait(&sm.full[s], ph); if (!first_chunk) PN_WAIT_ACC(w_full); }
          tcgen05_fence_after();
          if (first_chunk) { if (issuer) PN_DBG(2); first_chunk = false; }
          const uint64_t a_desc = a_desc0 + (uint64_t)(s * kAStep), b_desc = b_desc0 + (uint64_t)(s * kBStep);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
            }
            umma_commit(&sm.empty[s]);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (issuer) {
          umma_commit(&sm.tmem_full[acc]);
          PN_DBG(3);
        }
        __syncwarp();
      }
      if (issuer) { PN_WAIT_OUT(10, w_full); PN_WAIT_OUT(11, w_tempty); }
    }
  } else {
    // ===================== epilogue =====================
    const int e = warp - kEpilogueWarp0;
    const int etid = threadIdx.x - kEpilogueWarp0 * 32;
    uint32_t tcount = 0, stage_k = 0;
    long long w_ss = 0, w_tfull = 0, w_body = 0;
    for (int tile = 0; tile < n_tiles; ++tile, ++tcount) {
      const int gt = tile0 + tile * tstep;
      const int m_tile = gt / n_n_tiles, n_tile = gt - m_tile * n_n_tiles;
      const int n0 = n_tile * BN;
      const uint32_t acc = tcount & 1u, acc_ph = (tcount >> 1) & 1u;
      const long long _e0 = P.dbg ? clock64() : 0;
      named_bar_sync(2, kEpilogueThreads);
      for (int i = etid; i < BN; i += kEpilogueThreads) {
        const int n = n0 + i;
        sm.ss[i] = make_float2((n < P.cout && P.scale) ? __ldg(P.scale + n) : 1.f,
                               (n < P.cout && P.shift) ? __ldg(P.shift + n) : 0.f);
      }
      named_bar_sync(2, kEpilogueThreads);
      const long long _e1 = P.dbg ? clock64() : 0;
      mbar_wait_relaxed(&sm.tmem_full[acc], acc_ph);
      tcgen05_fence_after();
      const long long _e2 = P.dbg ? clock64() : 0;
      if (etid == 0) PN_DBG(4);
      int row = row_begin + m_tile * BLOCK_M + e * 32 + lane;
      bool row_ok = row < row_end;
      // TMA-stored box: the warp's 32 rows must all be this CTA's live rows (a share ends on a multiple of 8 rows,
      // so only the warp that straddles the end falls back to direct stores)
      const bool box_ok = BN >= 32 && P.tma_store != 0 && row - lane + 32 <= row_end;
      bool border = false;
      int dc_col = 0;   // deconv mode: first output column of this N tile within its tap
      if (P.dc_cout > 0) {
        // `row` is an input pixel of the padded map, the N tile lies inside one tap (BN divides dc_cout): the result
        // belongs to output pixel (2py-1+dy, 2px-1+dx) — every pixel of the padded output map, its zero border
        // included, is produced exactly once; positions that fall outside it are dropped
        const int tap = n0 / P.dc_cout;
        dc_col = tap * P.dc_cout;
        const int per = P.dc_hp_in * P.dc_wp_in;
        const int b = row / per, q = row - b * per;
        const int py = q / P.dc_wp_in, px = q - py * P.dc_wp_in;
        const int oy = 2 * py - 1 + (tap >> 1), ox = 2 * px - 1 + (tap & 1);
        row_ok = row_ok && oy >= 0 && oy < P.out_hp && ox >= 0 && ox < P.out_wp;
        border = ox == 0 || ox == P.out_wp - 1 || oy == 0 || oy == P.out_hp - 1;
        row = (b * P.out_hp + oy) * P.out_wp + ox;
      } else if (P.out_wp > 0) {
        const int q = row % (P.out_hp * P.out_wp);
        const int y = q / P.out_wp, x = q - y * P.out_wp;
        border = x == 0 || x == P.out_wp - 1 || y == 0 || y == P.out_hp - 1;
      }
      constexpr int CH = BN < 32 ? 16 : 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        if (n0 + c0 >= P.cout) break;  // warp-uniform
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * BN + c0;
        if (CH == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_wait_ld();
        if (row_ok) {
          const int nvalid = min(CH, P.cout - (n0 + c0));
          float f[CH];
#pragma unroll
          for (int j = 0; j < CH; j += 2) {
            const float4 s2 = *reinterpret_cast<const float4*>(&sm.ss[c0 + j]);      // c0, j even: 16-byte aligned
            f[j] = fmaf(__uint_as_float(v[j]), s2.x, s2.y);
            f[j + 1] = fmaf(__uint_as_float(v[j + 1]), s2.z, s2.w);
          }
          const long long ooff = (long long)row * P.out_ld + P.out_coff + (n0 - dc_col) + c0;
          if (P.out_f32) {
            float* op = reinterpret_cast<float*>(P.out) + ooff;
            if (P.residual) {
              const float* rp = reinterpret_cast<const float*>(P.residual) + (long long)row * P.res_ld + n0 + c0;
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (j < nvalid) f[j] += rp[j];
            }
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (border) {
#pragma unroll
              for (int j = 0; j < CH; ++j) f[j] = 0.f;
            }
            if (nvalid == CH && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < CH; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (j < nvalid) op[j] = f[j];
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(P.out) + ooff;
            if (P.residual) {
              const __nv_bfloat16* rp =
                  reinterpret_cast<const __nv_bfloat16*>(P.residual) + (long long)row * P.res_ld + n0 + c0;
              if (nvalid == CH && ((reinterpret_cast<uintptr_t>(rp) & 15u) == 0)) {
#pragma unroll
                for (int j = 0; j < CH; j += 8) {
                  const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
                  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float2 ff = __bfloat1622float2(h[u]);
                    f[j + 2 * u] += ff.x;
                    f[j + 2 * u + 1] += ff.y;
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                  if (j < nvalid) f[j] += __bfloat162float(rp[j]);
              }
            }
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (border) {
#pragma unroll
              for (int j = 0; j < CH; ++j) f[j] = 0.f;
            }
            if (CH == 32 && box_ok && nvalid == CH) {
              // all 32 lanes are here (box_ok: the warp's 32 rows are this CTA's live rows); lane = row of the box
              uint8_t* stg = sm.stage_out + (e * 2 + (int)(stage_k & 1u)) * 2048;
              ++stage_k;
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box before last was read
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 q;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
                for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(f[8 * j + 2 * u], f[8 * j + 2 * u + 1]);
                // SWIZZLE_64B: 16-byte chunk index ^ address bits 7-8 = (row >> 1) & 3 for 64-byte rows
                *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = q;
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmap_o, smem_u32(stg), P.out_coff + n0 + c0, row - lane);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            } else if (nvalid == CH && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                uint4 q;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
                for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(f[j + 2 * u], f[j + 2 * u + 1]);
                *reinterpret_cast<uint4*>(op + j) = q;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (j < nvalid) op[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&sm.tmem_empty[acc]);
      if (P.dbg) { const long long _e3 = clock64(); w_ss += _e1 - _e0; w_tfull += _e2 - _e1; w_body += _e3 - _e2; }
      if (etid == 0) PN_DBG(5);
    }
    if (etid == 0) { PN_WAIT_OUT(12, w_ss); PN_WAIT_OUT(13, w_tfull); PN_WAIT_OUT(14, w_body); PN_WAIT_OUT(15, n_tiles); }
    if (BN >= 32 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged boxes are out
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
  if (threadIdx.x == 0) PN_DBG(6);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long rows;
  int cols, ld, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.rows * 1000003u ^ (size_t)k.cols * 10007u ^
           (size_t)k.ld * 131u ^ (size_t)k.box_rows;
  }
};

// bf16 row-major (rows, cols) matrix with row stride ld; box = 64 columns x box_rows, SWIZZLE_128B.
// Weight tiles use box_rows = BLOCK_N; the gather4 activation map uses box_rows = 1.
// Maps are pure functions of (pointer, shape, box): cached (encoding costs ~1 us).
int get_map(const void* base, long long rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{base, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return PN_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PN_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PN_ERR_CUDA;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return PN_OK;
}

// bf16 output matrix (rows, cols) with row stride ld for the epilogue's TMA stores: box 32 cols x 32 rows, SWIZZLE_64B
int get_map_out(const void* base, long long rows, int cols, int ld, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{base, rows, cols, ld, -32};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return PN_OK; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PN_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PN_ERR_CUDA;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return PN_OK;
}

template <int BN, int STAGES, bool TMA_A>
int launch(const CUtensorMap& map_w, const CUtensorMap& map_a, const CUtensorMap& map_o, const KArgs& ka, int grid,
           cudaStream_t stream) {
  constexpr size_t smem = sizeof(Smem<BN, STAGES>) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static pn_detail::PerDeviceOnce once;
  if (once.need())
    PN_CUDA(cudaFuncSetAttribute(k_conv_tc<BN, STAGES, TMA_A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  static const bool timeline = [] { const char* e = getenv("PN_CONV_TIMELINE"); return e && e[0] == '1'; }();
  static unsigned long long* dbg_buf = nullptr;
  KArgs ka_dbg = ka;
  if (timeline) {
    if (!dbg_buf) PN_CUDA(cudaMalloc(&dbg_buf, 16 * 1024 * sizeof(unsigned long long)));
    PN_CUDA(cudaMemsetAsync(dbg_buf, 0, 16 * 1024 * sizeof(unsigned long long), stream));
    PN_CUDA(cudaStreamSynchronize(stream));
    ka_dbg.dbg = dbg_buf;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  PN_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<BN, STAGES, TMA_A>, map_w, map_a, map_o, ka_dbg));
  PN_CHECK_LAUNCH();
  if (timeline) {
    PN_CUDA(cudaStreamSynchronize(stream));
    static unsigned long long t[16 * 1024];
    PN_CUDA(cudaMemcpy(t, dbg_buf, sizeof(t), cudaMemcpyDeviceToHost));
    unsigned long long t_min = ~0ull, t_max = 0;
    int n = 0;
    for (int c = 0; c < grid; ++c) {
      if (t[c * 16 + 3] == 0) continue;   // CTA without tiles
      ++n;
      if (t[c * 16] < t_min) t_min = t[c * 16];
      if (t[c * 16 + 6] > t_max) t_max = t[c * 16 + 6];
    }
    double s_first = 0, s_mma = 0, s_epi = 0, s_tot = 0, m_mma = 0, m_tot = 0, s_setup = 0, s_w[8] = {0};
    for (int c = 0; c < grid; ++c) {
      const unsigned long long* q = t + c * 16;
      if (q[3] == 0) continue;
      const double fi = (double)(q[2] - q[1]), mm = (double)(q[3] - q[2]), ep = (double)(q[5] - q[4]),
                   to = (double)(q[6] - q[0]);
      s_setup += (double)(q[1] - q[0]); s_first += fi; s_mma += mm; s_epi += ep; s_tot += to;
      if (mm > m_mma) m_mma = mm;
      if (to > m_tot) m_tot = to;
      for (int k = 0; k < 8; ++k) s_w[k] += (double)q[8 + k];
    }
    if (n > 0)
      fprintf(stderr, "[conv_tc<%d,%d> cin %d cout %d taps %d rows_cap %d grid %d busy %d] span %.1f us | setup avg %.1f | "
                      "first operands avg %.1f | mma phase avg %.1f max %.1f | last epilogue avg %.1f | CTA total avg %.1f max %.1f\n",
              BN, STAGES, ka.cin, ka.cout, ka.taps, ka.rows_cap, grid, n, (t_max - t_min) / 1e3, s_setup / n / 1e3,
              s_first / n / 1e3, s_mma / n / 1e3, m_mma / 1e3, s_epi / n / 1e3, s_tot / n / 1e3, m_tot / 1e3);
    if (n > 0)
      fprintf(stderr, "    stalls per CTA (kclk): producer empty %.1f, producer tile barrier %.1f | mma full %.1f, mma tmem_empty %.1f | "
                      "epilogue scale/shift %.1f, tmem_full %.1f, body %.1f | tiles %.1f\n",
              s_w[0] / n / 1e3, s_w[1] / n / 1e3, s_w[2] / n / 1e3, s_w[3] / n / 1e3, s_w[4] / n / 1e3, s_w[5] / n / 1e3,
              s_w[6] / n / 1e3, s_w[7] / n);
  }
  return PN_OK;
}

}  // namespace

namespace pn_detail {

int conv_win(const pn_conv_args* a, cudaStream_t stream);   // conv_win_tc.cu

int conv_tcgen05(const pn_conv_args* a, cudaStream_t stream) {
  if (a->in_dtype != PN_BF16) return PN_ERR_UNSUPPORTED;
  if (a->nbr_kind == PN_NBR_SUBM_SORTED) {
    // raster-sorted submanifold rulebook: window-staged kernel (each input row fetched once per tile and kernel row)
    const int rc = conv_win(a, stream);
    if (rc != PN_ERR_UNSUPPORTED) return rc;
  }
  if (a->taps > kMaxTaps) return PN_ERR_UNSUPPORTED;
  if (a->cin % 8 != 0 || a->in_ld % 8 != 0 || (reinterpret_cast<uintptr_t>(a->in) & 15u) != 0)
    return PN_ERR_UNSUPPORTED;
  if (a->k_pad % BLOCK_K != 0 || (reinterpret_cast<uintptr_t>(a->weight) & 15u) != 0) return PN_ERR_UNSUPPORTED;
  int bn = a->cout <= 16 ? 16 : a->cout <= 32 ? 32 : a->cout <= 64 ? 64 : a->cout <= 128 ? 128 : 256;
  const bool deconv = a->deconv_cout != 0;
  if (deconv) {
    if (a->taps != 1 || a->nbr != nullptr || a->cout != 4 * a->deconv_cout || a->residual != nullptr ||
        a->deconv_cout % 16 != 0 || a->deconv_hp_in <= 0 || a->deconv_wp_in <= 0 || a->out_hp <= 0 || a->out_wp <= 0)
      return PN_ERR_INVALID_ARG;
    while (a->deconv_cout % bn != 0) bn >>= 1;   // an N tile must lie inside one tap
  }
  {
    // Tile-shape choice.  The kernel is bound by L2->SM bytes (A tile 128 rows + B tile bn rows per
    // K chunk), so pick the bn that minimises waves x bytes-per-tile; with few row tiles (deep,
    // low-resolution layers) a narrower N tile lets every SM stream only its slice of the weights.
    // The live row count is on the device: rows_hint (expected rows) stands in for it when given.
    const int sms_ = sm_count();
    const bool estimated = a->rows_hint > 0;
    const long long rows_est = estimated ? a->rows_hint : a->rows_cap;
    if (bn > 64) {
      long long best_cost = -1;
      int best = bn;
      for (int cand = bn; cand >= 64; cand >>= 1) {
        if (deconv && a->deconv_cout % cand != 0) continue;
        const int n_n = PN_DIVUP(a->cout, cand);
        long long units;   // 128-row tiles the busiest CTA walks
        if (n_n == 1) {
          // one N tile: CTAs own equal contiguous row shares (see the kernel), the cost is smooth in the row count
          const long long share = PN_DIVUP(rows_est, (long long)sms_);
          units = PN_DIVUP(share < 64 ? 64 : share, (long long)BLOCK_M);
        } else {
          // several N tiles: round-robin (row tile, N tile) units, whose cost jumps at every multiple of the SM
          // count.  An estimated row count gets 35 % headroom: stage 4 of nuScenes frames varies between 7.7 k and
          // 10.3 k rows, and 162 units on 148 SMs cost two full waves (31 us instead of 17 us per conv).
          const long long rows_m = estimated ? rows_est + (rows_est * 35) / 100 : rows_est;
          units = PN_DIVUP(PN_DIVUP(rows_m, (long long)BLOCK_M) * n_n, (long long)sms_);
        }
        const long long cost = units * (BLOCK_M + cand);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = cand; }
      }
      bn = best;
    }
  }
  CUtensorMap map, map_a;
  int rc = get_map(a->weight, a->cout, a->k_pad, a->k_pad, bn, &map);
  if (rc != PN_OK) return rc;
  // activations through TMA gather4 when a 64-channel chunk never straddles taps and the allocation
  // size of `in` is known; otherwise 16-byte cp.async gathers
  // Measured on B200: 32 gather4 ops per 16 KB chunk are slower than 1024 cp.async pieces (57 vs 42 us
  // on the 128-channel stage), so the TMA gather path stays opt-in (PN_CONV_TMA_GATHER=1).
  static const bool tma_gather_enabled = [] {
    const char* e = getenv("PN_CONV_TMA_GATHER");
    return e && e[0] == '1';
  }();
  const bool tma_a = tma_gather_enabled && a->cin % BLOCK_K == 0 && a->in_rows > 0;
  if (tma_a) {
    rc = get_map(a->in, a->in_rows, a->cin, a->in_ld, 1, &map_a);
    if (rc != PN_OK) return rc;
  } else {
    map_a = map;
  }
  KArgs ka;
  ka.in = reinterpret_cast<const __nv_bfloat16*>(a->in);
  ka.in_ld = a->in_ld;
  ka.nbr = a->nbr;
  ka.taps = a->taps;
  ka.n_chunks = a->k_pad / BLOCK_K;
  ka.scale = a->scale;
  ka.shift = a->shift;
  ka.residual = a->residual;
  ka.res_ld = a->res_ld;
  ka.out = a->out;
  ka.out_f32 = a->out_dtype == PN_F32;
  ka.out_ld = a->out_ld;
  ka.out_coff = a->out_coff;
  ka.relu = a->relu;
  ka.num_rows = a->num_rows;
  ka.rows_cap = a->rows_cap;
  ka.cin = a->cin;
  ka.cout = a->cout;
  ka.out_hp = a->out_hp;
  ka.out_wp = a->out_wp;
  ka.in_rows = a->in_rows;
  ka.cin_shift = -1;
  ka.dc_cout = a->deconv_cout;
  ka.dc_hp_in = a->deconv_hp_in;
  ka.dc_wp_in = a->deconv_wp_in;
  ka.dbg = nullptr;
  // epilogue through TMA tensor stores: bf16 rows written in place (no row remap), whole 32-column chunks
  static const bool tma_store_enabled = [] { const char* e = getenv("PN_CONV_TMA_STORE"); return !(e && e[0] == '0'); }();
  CUtensorMap map_o = map;
  ka.tma_store = 0;
  if (tma_store_enabled && a->out_dtype == PN_BF16 && !deconv && a->cout % 32 == 0 && a->out_coff % 8 == 0 &&
      a->out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15u) == 0 && a->rows_cap > 0 &&
      get_map_out(a->out, a->rows_cap, a->out_ld, a->out_ld, &map_o) == PN_OK)
    ka.tma_store = 1;
  for (int sft = 3; sft < 16; ++sft)
    if ((1 << sft) == a->cin) ka.cin_shift = sft;
  if (a->in_rows > 0 && (long long)a->in_rows * a->in_ld * 2 >= (1ll << 32)) return PN_ERR_UNSUPPORTED;
  const int sms = sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  // every CTA handles all N tiles of its row share (>= 64 rows), so the grid depends on the rows only
  const long long shares_cap = PN_DIVUP(a->cout, bn) == 1
                                   ? PN_DIVUP((long long)a->rows_cap, 64ll)
                                   : (long long)PN_DIVUP(a->rows_cap, BLOCK_M) * PN_DIVUP(a->cout, bn);
  const int grid = (int)(shares_cap < sms ? (shares_cap < 1 ? 1 : shares_cap) : sms);
  if (tma_a) {
    switch (bn) {
      case 16: return launch<16, 8, true>(map, map_a, map_o, ka, grid, stream);
      case 32: return launch<32, 8, true>(map, map_a, map_o, ka, grid, stream);
      case 64: return launch<64, 8, true>(map, map_a, map_o, ka, grid, stream);
      case 128: return launch<128, 6, true>(map, map_a, map_o, ka, grid, stream);
      default: return launch<256, 4, true>(map, map_a, map_o, ka, grid, stream);
    }
  }
  switch (bn) {
    // Few stages on purpose: pipeline depth beyond ~4 chunks bought nothing (measured), while the shared
    // memory left to L1 lets the .ca gathers hit on the 3x reuse of activation rows between the taps of
    // horizontally adjacent outputs.
    case 16: return launch<16, 6, false>(map, map_a, map_o, ka, grid, stream);
    case 32: return launch<32, 6, false>(map, map_a, map_o, ka, grid, stream);
    case 64: return launch<64, 5, false>(map, map_a, map_o, ka, grid, stream);
    case 128: return launch<128, 4, false>(map, map_a, map_o, ka, grid, stream);
    default: return launch<256, 4, false>(map, map_a, map_o, ka, grid, stream);
  }
}

}  // namespace pn_detail
