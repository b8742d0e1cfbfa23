// Window-staged submanifold 3x3 convolution on 5th-gen tensor cores (sm_100a): every input row is fetched ONCE per
// output tile and kernel row, by TMA, instead of once per tap by the LSU.
//
//   out[o, n] = act( (sum_t sum_c in[nbr[o,t], c] * W[n, t*cin + c]) * scale[n] + shift[n] + res[o,n] ),  t = ky*3+kx
//
// Replaces spconv's SubMConv2d (det3d/models/backbones/base.py:38-52,145-213) for the site sets this library
// produces: rows sorted in raster order (pn_pillarize / pn_rulebook_*), rulebook output-stationary.
//
// Why (round-1 measurements, profiles/r1_conv_tc_stalls_nusc18.txt): the gather kernel (conv_tcgen05.cu) feeds the
// tensor core through 16-byte cp.async gathers, one per (output row, tap, 16 B); the LSU charges ~8 + 2 clk per
// distinct 128-byte line per warp instruction, which caps the feed at 23-28 B/clk/SM — 600-800 clk per (tap,
// 64-channel chunk) against the 90-256 clk its MMAs take — and every input row is re-fetched by up to nine taps.
//
// What this kernel does instead.  In raster-sorted order the neighbours of 128 CONSECUTIVE output rows under one kernel
// row ky are a CONTIGUOUS run of input rows (the outputs' x-neighbours one raster line up, on, or below; measured
// span 128-190 rows on nuScenes/Waymo-shaped frames).  Per unit = (output tile, 64-channel chunk, ky):
//   * one TMA box load stages the window rows [lo_ky, lo_ky + S) x chunk into shared memory (no LSU issue cost);
//   * 12 builder warps (thread = output row, warp = row quarter x ky) read the window row that holds their neighbour
//     (KU/8 LDS.128, conflict-free on the swizzled window) and write it to THEIR LANE OF TENSOR MEMORY with
//     tcgen05.st; absent neighbours are zeros from registers, the rare neighbours outside the window come straight
//     from global memory (always correct, whatever the rulebook looks like);
//   * the MMA warp issues tcgen05.mma with the A operand IN TMEM (128 lanes x K/2 packed columns) against weight
//     tiles streamed by TMA — the A tile never exists in shared memory.
// Why TMEM: the first version scattered the window into three 128B-swizzled smem tiles per unit.  It was correct and
// removed the LSU issue limit, but ran no faster: per unit (N = 64) it moved 24 KB (window, TMA) + ~18 KB (LDS) +
// 48 KB (STS) + 48 KB (UMMA reads A) + 48 KB (weights in + out) = 186 KB through shared memory, 1450 clk at
// 128 B/clk/SM against 1800 measured — shared-memory bandwidth, not the LSU, is the roofline of a small-N implicit
// GEMM.  With A in TMEM the tile is neither written to nor read back from shared memory (~111 KB per unit).
// The plan of a tile (window start per ky, window row per output row and tap) depends only on the rulebook, which all
// SubM convs of a backbone stage share (spconv's `indice_key`): built ONCE per rulebook by k_win_plan
// (pn_conv_window_plan) and streamed into shared memory one tile ahead by a bulk copy.  (First version: a mapper warp
// inside the conv kernel — a single warp needs ~5000 clk per tile for it, 2.5 us of exposed wait per tile; measured.)
//
// Warp roles (608 threads, one persistent CTA per SM): 0-11 builders, 12-15 epilogue (TMEM lane quarter = warp % 4),
// 16 MMA issuer, 17 loader (plan copies, window TMA), 18 weight loader.
#include <cuda.h>

#include <climits>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace {

using namespace pn_tc;

constexpr int BLOCK_M = 128;
constexpr int kWin = 160;                 // staged rows per kernel row (TMA box rows; must be <= 256).  Measured spans:
                                          // median 129, 90 % <= 143, 99 % <= 165, max 191; rows past the window are fetched
                                          // from global memory by the builders
constexpr int kBuilderWarps = 12;         // warp w: TMEM lane quarter w % 4 (rows 32(w%4)..+31), kernel row ky = w / 4
constexpr int kEpilogueWarp0 = kBuilderWarps;          // 12..15 (index % 4 == TMEM lane quarter)
constexpr int kEpilogueThreads = 128;
constexpr int kMmaWarp = kBuilderWarps + 4;            // 16
constexpr int kLoaderWarp = kMmaWarp + 1;              // 17: plan copies + staged windows
constexpr int kWeightWarp = kMmaWarp + 2;              // 18: weight tiles
constexpr int kThreads = (kWeightWarp + 1) * 32;       // 608
// Tile plan (bytes): src[9][128] = window row of the neighbour of output row i under tap t (0xFF absent, 0xFE outside
// the window: fetched from global) | lo[3] + pad (int32)
constexpr int kSrcAbsent = 0xFF, kSrcFar = 0xFE;
constexpr int kPlanLo = 9 * BLOCK_M;                   // byte offset
constexpr int kPlanBytes = kPlanLo + 16;               // 1168
static_assert(kPlanBytes % 16 == 0, "bulk copy granularity");
static_assert(kWin <= kSrcFar, "window offsets must fit the byte map");

// Row partition shared by the plan kernel and the conv kernel: CTA c of `grid` owns rows [begin, end), walked in
// 128-row tiles.  Equal contiguous shares (multiple of 8 rows, at least one tile), as conv_tcgen05.cu.
__host__ __device__ inline int win_share(int rows, int grid) {
  const int s = (((rows + grid - 1) / grid) + 7) & ~7;
  return s < BLOCK_M ? BLOCK_M : s;   // never less than one full tile: an MMA costs the same for 64 rows as for 128
}

template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct WArgs {
  const __nv_bfloat16* in;
  int in_ld;
  const int* nbr;
  const uint8_t* plan;   // tile plans of pn_conv_window_plan: [grid][tiles_per_cta][kPlanBytes]
  int tiles_per_cta;
  int plan_grid;         // row shares the plan was built for (win_geom)
  int n_split;           // CTAs per output-column block: CTA c computes columns [BN * (c % n_split), +BN) of the row
                         // shares c / n_split + k * ceil(plan_grid / n_split), k < n_split
  int n_chunks;          // cin / KU
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int res_ld;
  __nv_bfloat16* out;
  int out_ld;
  int out_coff;
  int relu;
  const int* num_rows;
  int rows_cap;
  int cin;
  int cout;
  int tma_store;
  int tma_res;           // residual rows through TMA boxes (tmap_r is valid)
  unsigned long long* dbg;
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PW_DBG(slot) do { if (P.dbg) P.dbg[blockIdx.x * 16 + slot] = gtime_ns(); } while (0)
#define PW_T0() const long long _w0 = P.dbg ? clock64() : 0
#define PW_ACC(var) do { if (P.dbg) var += clock64() - _w0; } while (0)
#define PW_OUT(slot, var) do { if (P.dbg) P.dbg[blockIdx.x * 16 + slot] = (unsigned long long)(var); } while (0)

// RES: all 9 * n_chunks weight tiles of the layer stay resident in shared memory (loaded once per CTA; NBT = their
// number); otherwise NBT tiles form a ring of NBT / 3 units (3 taps per barrier).  WS = staged-window ring depth,
// AS = TMEM A ring depth in units.
template <int BN, int KU, bool RES, int NBT, int WS, int AS>
struct WSmem {
  static_assert(WS % AS == 0 && AS <= 3, "a window slot must always belong to the same builder group");
  static_assert(BN <= 128, "wider layers are split into 128-column blocks across CTAs (WArgs::n_split)");
  static constexpr int NBU = RES ? 1 : NBT / 3;
  alignas(1024) uint8_t b[NBT][BN * 128];                // weight tiles, one per tap (SWIZZLE_128B, K-major)
  alignas(1024) uint8_t win[WS][kWin * KU * 2];          // staged input windows (KU = 64: SWIZZLE_128B, 32: SWIZZLE_64B)
  alignas(1024) uint8_t stage_out[4 * 2048];             // epilogue boxes for TMA stores (one per epilogue warp)
  alignas(1024) uint8_t stage_res[4][BN >= 32 ? BN / 32 : 1][2048];   // residual boxes (32 rows x 32 cols, SWIZZLE_64B)
  alignas(16) uint8_t plan[3][kPlanBytes];               // tile plans (ring of three, copied one tile ahead)
  alignas(8) uint64_t a_full[AS];
  uint64_t a_empty[AS], win_full[WS], win_empty[WS], b_full[NBU], b_empty[NBU], map_full[3], map_empty[3];
  uint64_t tmem_full[2], tmem_empty[2], res_full[4];
  uint32_t tmem_base;
  alignas(16) float2 ss[BN];   // {scale, shift} per output column, read two columns per LDS.128 (a broadcast LDS.32
                               // costs the shared-memory pipe a full wavefront)
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// registers -> TMEM: lane t of the warp writes 32 (16) consecutive 32-bit columns of TMEM lane (quarter base + t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem descriptor]: the A operand (128 lanes x K/2 packed bf16x2 columns) comes from TMEM
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int tmem_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int BN, int KU, bool RES, int NBT, int WS, int AS>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_win(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_in,
           const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_r, const WArgs P) {
  extern __shared__ uint8_t smem_raw[];
  using S = WSmem<BN, KU, RES, NBT, WS, AS>;
  constexpr int NBU = S::NBU;
  constexpr int kASlots = AS, kWinSlots = WS;
  S& sm = *reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KS = KU / 16;                 // MMA k-steps per tap and chunk
  constexpr int ACOLS = KU / 2;               // TMEM columns of one tap's A operand (bf16x2 per column)
  constexpr int ROWB = KU * 2;                // bytes per staged row
  constexpr int A_COL0 = 2 * BN;              // TMEM: [0, 2BN) double-buffered accumulator, then the A ring
  constexpr int TCOLS = tmem_pow2(A_COL0 + kASlots * 3 * ACOLS);
  static_assert(A_COL0 + kASlots * 3 * ACOLS <= 512, "TMEM budget");
  if (threadIdx.x == 0) PW_DBG(0);
  pdl_launch_dependents();
  // Column split (wide layers with few rows: 61 row tiles of a 256-channel layer leave 87 SMs idle and the busy ones
  // MMA-bound; two CTAs per row share, one per 128-column half, halve that): CTA c owns column block c % n_split.
  const int col_blk = (int)blockIdx.x % P.n_split, share_a = (int)blockIdx.x / P.n_split;
  const int share_b = share_a + (P.plan_grid + P.n_split - 1) / P.n_split;      // second row share of a split CTA
  const int col0 = col_blk * BN, cout_l = min(BN, P.cout - col0);
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kASlots; ++s) {
        mbar_init(&sm.a_full[s], 4);
        mbar_init(&sm.a_empty[s], 1);
      }
      for (int s = 0; s < kWinSlots; ++s) {
        mbar_init(&sm.win_full[s], 1);
        mbar_init(&sm.win_empty[s], 4);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&sm.tmem_full[s], 1);
        mbar_init(&sm.tmem_empty[s], kEpilogueThreads);
      }
      for (int s = 0; s < NBU; ++s) {
        mbar_init(&sm.b_full[s], 1);
        mbar_init(&sm.b_empty[s], 1);
      }
      for (int s = 0; s < 4; ++s) mbar_init(&sm.res_full[s], 1);
      for (int s = 0; s < 3; ++s) {
        mbar_init(&sm.map_full[s], 1);
        mbar_init(&sm.map_empty[s], 4 * kASlots);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(&sm.tmem_base);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  // Programmatic dependent launch: everything above (barriers, TMEM) and the resident weights below are independent of
  // the predecessor, so they overlap its tail; the row count, the plan and every activation are read after the wait.
  if constexpr (RES) {
    if (warp == kWeightWarp && lane == 0) {
      mbar_arrive_expect_tx(&sm.b_full[0], (uint32_t)(9 * P.n_chunks) * BN * 128);
      for (int kc = 0; kc < P.n_chunks; ++kc)
        for (int t = 0; t < 9; ++t)
          tma_load_2d(smem_u32(sm.b[kc * 9 + t]), &tmap_w, t * P.cin + kc * KU, col0, &sm.b_full[0]);
    }
  }
  pdl_wait();
  // the window starts of this CTA's first tile do not depend on the live row count: their load travels together with it
  int4 lo_first = make_int4(0, 0, 0, 0);
  if (warp == kLoaderWarp && lane == 0)
    lo_first = __ldg(reinterpret_cast<const int4*>(P.plan + (size_t)share_a * P.tiles_per_cta * kPlanBytes + kPlanLo));
  const int rows = P.num_rows ? min(*P.num_rows, P.rows_cap) : P.rows_cap;
  // balanced schedule (as conv_tcgen05.cu): equal contiguous row shares, walked in 128-row tiles
  const int share = win_share(rows, P.plan_grid);
  const int beg_a = min(rows, share_a * share), end_a = min(rows, beg_a + share);
  const int tiles_a = (end_a - beg_a + BLOCK_M - 1) / BLOCK_M;
  const bool has_b = P.n_split > 1 && share_b < P.plan_grid;
  const int beg_b = has_b ? min(rows, share_b * share) : rows, end_b = has_b ? min(rows, beg_b + share) : rows;
  const int n_tiles = tiles_a + (end_b - beg_b + BLOCK_M - 1) / BLOCK_M;
  // tile t of this CTA: first row, end of its share, plan record
  auto tile_row0 = [&](int t) { return t < tiles_a ? beg_a + t * BLOCK_M : beg_b + (t - tiles_a) * BLOCK_M; };
  auto tile_end = [&](int t) { return t < tiles_a ? end_a : end_b; };
  auto tile_plan = [&](int t) {
    return P.plan + ((size_t)(t < tiles_a ? share_a : share_b) * P.tiles_per_cta + (t < tiles_a ? t : t - tiles_a)) * kPlanBytes;
  };

  if (threadIdx.x == 0) PW_DBG(1);

  if (warp < kBuilderWarps) {
    // ===================== builders: staged window -> A operands of one unit in TMEM =====================
    // thread = output row i = 32 (warp % 4) + lane of the tile; warp group g = warp / 4 builds whole units (all three
    // taps of a kernel row), AS units in flight at once, one per group — that hides the window / TMEM latencies a single
    // chain per unit exposed (measured: builders and MMA warp spent half their time waiting for each other).  Per tap
    // the thread reads the window row that holds its neighbour with KU/8 16-byte loads (conflict-free on the swizzled
    // window when the rows of neighbouring lanes are consecutive, as they are in raster order) and stores it to its
    // TMEM lane.
    const int quarter = warp & 3, grp = warp >> 2;
    const int i = quarter * 32 + lane;
    const char* in_bytes = reinterpret_cast<const char*>(P.in);
    const uint32_t in_ld_bytes = (uint32_t)P.in_ld * 2u;
    long long w_aempty = 0, w_win = 0, w_map = 0;
    // Group g owns the units u = g (mod AS): exactly the units that use TMEM A slot g, so every a_full / a_empty
    // barrier has ONE producer group and parity waits cannot alias (a group that may run several phases ahead of a
    // barrier it shares with others would pass a parity wait it must block on).  Groups >= AS stay idle (BN = 128 has
    // TMEM columns for two A slots only).
    const int upt = 3 * P.n_chunks;                       // units per tile
    const uint32_t n_units = (uint32_t)(n_tiles * upt);
    int cur_tile = -1, row0 = 0;
    const uint8_t* s_plan = nullptr;
    if (grp < kASlots) {
      for (uint32_t u = (uint32_t)grp; u < n_units; u += kASlots) {
        const int tile = (int)u / upt, r = (int)u - tile * upt, kc = r / 3, ky = r - kc * 3;
        if (tile != cur_tile) {
          if (cur_tile >= 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.map_empty[cur_tile % 3]);     // done with the previous tile's plan
          }
          { PW_T0(); mbar_wait(&sm.map_full[tile % 3], (uint32_t)(tile / 3) & 1u); PW_ACC(w_map); }
          cur_tile = tile;
          row0 = tile_row0(tile);
          s_plan = sm.plan[tile % 3];
        }
        const uint32_t aslot = (uint32_t)grp, aph = (u / kASlots) & 1u;
        const uint32_t wslot = u % kWinSlots, wph = (u / kWinSlots) & 1u;
        uint32_t src3[3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) src3[kx] = s_plan[(ky * 3 + kx) * BLOCK_M + i];
        { PW_T0(); mbar_wait(&sm.win_full[wslot], wph); PW_ACC(w_win); }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t s = src3[kx];
          uint32_t v[32];
          if (s < (uint32_t)kSrcFar) {
            const uint8_t* rowp = sm.win[wslot] + s * ROWB;
            // 16-byte chunk j of window row s sits at chunk j ^ (s & 7) (SWIZZLE_128B) / j ^ ((s >> 1) & 3) (SWIZZLE_64B)
            const uint32_t x = KU == 64 ? (s & 7u) : ((s >> 1) & 3u);
#pragma unroll
            for (int j = 0; j < KU / 8; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(rowp + (((uint32_t)j ^ x) << 4));
              v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
          } else if (s == (uint32_t)kSrcFar) {
            const int srow = __ldg(P.nbr + (long long)(row0 + i) * 9 + ky * 3 + kx);
            const uint4* gp = reinterpret_cast<const uint4*>(in_bytes + (size_t)srow * in_ld_bytes + (size_t)(kc * KU) * 2u);
#pragma unroll
            for (int j = 0; j < KU / 8; ++j) {
              const uint4 q = __ldg(gp + j);
              v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < KU / 2; ++j) v[j] = 0u;
          }
          if (kx == 0) {
            PW_T0(); mbar_wait(&sm.a_empty[aslot], aph ^ 1u); PW_ACC(w_aempty);
            tcgen05_fence_after();
          }
          __syncwarp();     // tcgen05.st is .sync.aligned: the lanes diverged on present / far / absent above
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + A_COL0 + (aslot * 3 + kx) * ACOLS;
          if (KU == 64) tmem_st32(taddr, v); else tmem_st16(taddr, v);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.win_empty[wslot]);      // the window rows went through registers into TMEM stores
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.a_full[aslot]);
      }
      if (cur_tile >= 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.map_empty[cur_tile % 3]);
      }
    }
    if (threadIdx.x == 0) { PW_OUT(8, w_aempty); PW_OUT(9, w_win); PW_OUT(7, w_map); }
  } else if (warp == kLoaderWarp) {
    // ===================== loader: plan copies, staged windows, weight tiles =====================
    if (lane == 0) {
      uint32_t u = 0;
      // plan of `tile` -> ring slot tile % 3 (free once every builder warp has read tile - 3's entries); its window
      // starts come straight from global memory so the first window load does not wait for the copy
      auto prefetch_plan = [&](int tile, int4& lo) {
        const int buf = tile % 3;
        mbar_wait(&sm.map_empty[buf], ((uint32_t)(tile / 3) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sm.map_full[buf], kPlanBytes);
        const uint8_t* plan_g = tile_plan(tile);
        bulk_copy_g2s(smem_u32(sm.plan[buf]), plan_g, kPlanBytes, &sm.map_full[buf]);
        lo = __ldg(reinterpret_cast<const int4*>(plan_g + kPlanLo));
      };
      int4 lo_next = make_int4(0, 0, 0, 0);
      if (n_tiles > 0) {
        int4 dummy;
        prefetch_plan(0, dummy);
        lo_next = lo_first;
      }
      for (int tile = 0; tile < n_tiles; ++tile) {
        const int lo0 = lo_next.x, lo1 = lo_next.y, lo2 = lo_next.z;
        if (tile + 1 < n_tiles) prefetch_plan(tile + 1, lo_next);
        for (int kc = 0; kc < P.n_chunks; ++kc) {
          for (int ky = 0; ky < 3; ++ky, ++u) {
            const uint32_t slot = u % kWinSlots, ph = (u / kWinSlots) & 1u;
            mbar_wait(&sm.win_empty[slot], ph ^ 1u);
            mbar_arrive_expect_tx(&sm.win_full[slot], kWin * ROWB);
            tma_load_2d(smem_u32(sm.win[slot]), &tmap_in, kc * KU, ky == 0 ? lo0 : (ky == 1 ? lo1 : lo2),
                        &sm.win_full[slot]);
          }
        }
      }
    }
  } else if (warp == kWeightWarp) {
    // ===================== weight tiles =====================
    if constexpr (RES) {
      // resident weights were requested before the dependency wait (above); a CTA without tiles still has to see them
      // land before it may exit
      if (lane == 0 && n_tiles == 0) mbar_wait(&sm.b_full[0], 0u);
    }
    if (lane == 0 && n_tiles > 0) {
      if constexpr (!RES) {
        uint32_t u = 0;
        for (int tile = 0; tile < n_tiles; ++tile)
          for (int kc = 0; kc < P.n_chunks; ++kc)
            for (int ky = 0; ky < 3; ++ky, ++u) {
              const uint32_t bs = u % NBU, bph = (u / NBU) & 1u;
              mbar_wait(&sm.b_empty[bs], bph ^ 1u);
              mbar_arrive_expect_tx(&sm.b_full[bs], 3 * BN * 128);
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
                tma_load_2d(smem_u32(sm.b[bs * 3 + kx]), &tmap_w, (ky * 3 + kx) * P.cin + kc * KU, col0, &sm.b_full[bs]);
            }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    const bool issuer = elect_one();
    constexpr uint32_t idesc = make_idesc<BN>();
    const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(sm.b[0]));
    constexpr uint32_t kBStep = (uint32_t)(BN * 128) >> 4;
    uint32_t u = 0;
    bool first = true;
    if constexpr (RES) {
      if (n_tiles > 0) mbar_wait(&sm.b_full[0], 0u);      // resident weights: one wait for the whole kernel
    }
    long long w_afull = 0, w_bfull = 0, w_tempty = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const uint32_t acc = (uint32_t)tile & 1u, acc_ph = ((uint32_t)tile >> 1) & 1u;
      { PW_T0(); mbar_wait(&sm.tmem_empty[acc], acc_ph ^ 1u); PW_ACC(w_tempty); }
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      uint32_t accumulate = 0u;
      for (int kc = 0; kc < P.n_chunks; ++kc) {
        for (int ky = 0; ky < 3; ++ky, ++u) {
          const uint32_t aslot = u % kASlots, aph = (u / kASlots) & 1u;
          const uint32_t bs = RES ? 0u : u % NBU, bph = (u / NBU) & 1u;
          if constexpr (!RES) { PW_T0(); mbar_wait(&sm.b_full[bs], bph); if (!first) PW_ACC(w_bfull); }
          { PW_T0(); mbar_wait(&sm.a_full[aslot], aph); if (!first) PW_ACC(w_afull); }
          tcgen05_fence_after();
          if (first) { if (issuer) PW_DBG(2); first = false; }
          const uint32_t b_tile0 = RES ? (uint32_t)(kc * 9 + ky * 3) : bs * 3u;
          if (issuer) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t a_tmem = tmem_base + A_COL0 + (aslot * 3 + (uint32_t)kx) * ACOLS;
              const uint64_t b_desc = b_desc0 + (uint64_t)((b_tile0 + (uint32_t)kx) * kBStep);
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                umma_bf16_ts(d_tmem, a_tmem + 8 * k, b_desc + 2 * k, idesc, accumulate);
                accumulate = 1u;
              }
            }
            if constexpr (!RES) umma_commit(&sm.b_empty[bs]);
            umma_commit(&sm.a_empty[aslot]);
          }
          accumulate = 1u;
        }
      }
      if (issuer) {
        umma_commit(&sm.tmem_full[acc]);
        PW_DBG(3);
      }
      __syncwarp();
    }
    if (issuer) { PW_OUT(10, w_afull); PW_OUT(11, w_bfull); PW_OUT(12, w_tempty); }
  } else {
    // ===================== epilogue: TMEM -> scale/shift (+residual, ReLU) -> bf16 rows =====================
    const int e = warp - kEpilogueWarp0;
    const int etid = threadIdx.x - kEpilogueWarp0 * 32;
    for (int i = etid; i < BN; i += kEpilogueThreads) {
      sm.ss[i] = make_float2((i < cout_l && P.scale) ? __ldg(P.scale + col0 + i) : 1.f,
                             (i < cout_l && P.shift) ? __ldg(P.shift + col0 + i) : 0.f);
    }
    named_bar_sync(2, kEpilogueThreads);
    uint32_t res_ph = 0u;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const uint32_t acc = (uint32_t)tile & 1u, acc_ph = ((uint32_t)tile >> 1) & 1u;
      const int row = tile_row0(tile) + e * 32 + lane, row_end = tile_end(tile);
      const bool row_ok = row < row_end;
      const bool full_box = BN >= 32 && row - lane + 32 <= row_end;     // the warp's 32 rows are all live rows of this CTA
      const bool box_ok = full_box && P.tma_store != 0;
      // Residual rows of this tile travel as TMA boxes while the tile's MMAs are still running: a thread owns a row, so
      // direct loads touch 32 different lines per warp instruction and sat on the critical path after tmem_full
      // (measured: residual layers 2.5 us slower than their twins, MMA warp waiting on tmem_empty).
      const bool res_box = full_box && P.tma_res != 0 && P.residual != nullptr;
      if (res_box) {
        if (lane == 0) {
          constexpr int NCH = BN >= 32 ? BN / 32 : 1;
          mbar_arrive_expect_tx(&sm.res_full[e], NCH * 2048);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
            tma_load_2d(smem_u32(sm.stage_res[e][c]), &tmap_r, col0 + c * 32, row, &sm.res_full[e]);
        }
        __syncwarp();
      }
      mbar_wait_relaxed(&sm.tmem_full[acc], acc_ph);
      tcgen05_fence_after();
      if (etid == 0) PW_DBG(4);
      if (res_box) { mbar_wait(&sm.res_full[e], res_ph); res_ph ^= 1u; }
      constexpr int CH = BN < 32 ? 16 : 32;
      // the accumulator is handed back to the MMA warp as soon as this thread's LAST chunk of it sits in registers —
      // not after that chunk has been converted and stored (the 64-channel layers had the MMA warp waiting 20-30 % of
      // its time on tmem_empty)
      const int c0_last = cout_l > 0 ? ((min(BN, cout_l) - 1) / CH) * CH : -1;
      if (c0_last < 0) { tcgen05_fence_before(); mbar_arrive(&sm.tmem_empty[acc]); }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        if (c0 >= cout_l) break;   // warp-uniform
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * BN + c0;
        if (CH == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_wait_ld();
        if (c0 == c0_last) { tcgen05_fence_before(); mbar_arrive(&sm.tmem_empty[acc]); }
        if (row_ok) {
          const int nvalid = min(CH, cout_l - c0);
          float f[CH];
#pragma unroll
          for (int j = 0; j < CH; j += 2) {
            const float4 s2 = *reinterpret_cast<const float4*>(&sm.ss[c0 + j]);      // c0, j even: 16-byte aligned
            f[j] = fmaf(__uint_as_float(v[j]), s2.x, s2.y);
            f[j + 1] = fmaf(__uint_as_float(v[j + 1]), s2.z, s2.w);
          }
          __nv_bfloat16* op = P.out + (long long)row * P.out_ld + P.out_coff + col0 + c0;
          if (P.residual) {
            const __nv_bfloat16* rp = P.residual + (long long)row * P.res_ld + col0 + c0;
            if (CH == 32 && res_box) {
              const uint8_t* rb = sm.stage_res[e][c0 / 32];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(rb + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int w2 = 0; w2 < 4; ++w2) {
                  const float2 ff = __bfloat1622float2(h[w2]);
                  f[8 * j + 2 * w2] += ff.x;
                  f[8 * j + 2 * w2 + 1] += ff.y;
                }
              }
            } else if (nvalid == CH && ((reinterpret_cast<uintptr_t>(rp) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int w2 = 0; w2 < 4; ++w2) {
                  const float2 ff = __bfloat1622float2(h[w2]);
                  f[j + 2 * w2] += ff.x;
                  f[j + 2 * w2 + 1] += ff.y;
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
          if (CH == 32 && box_ok && nvalid == CH) {
            uint8_t* stg = sm.stage_out + e * 2048;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous box was read
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
              for (int w2 = 0; w2 < 4; ++w2) h[w2] = __floats2bfloat162_rn(f[8 * j + 2 * w2], f[8 * j + 2 * w2 + 1]);
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = q;   // SWIZZLE_64B
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_o, smem_u32(stg), P.out_coff + col0 + c0, row - lane);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else if (nvalid == CH && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              uint4 q;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
              for (int w2 = 0; w2 < 4; ++w2) h[w2] = __floats2bfloat162_rn(f[j + 2 * w2], f[j + 2 * w2 + 1]);
              *reinterpret_cast<uint4*>(op + j) = q;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (j < nvalid) op[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
      if (etid == 0) PW_DBG(5);
    }
    if (BN >= 32 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
  if (threadIdx.x == 0) { PW_DBG(6); PW_OUT(15, n_tiles); }
}

// ---- tile plans: once per rulebook -------------------------------------------------------------------------------
// One CTA of 128 threads per (conv CTA c, tile k); thread i owns output row row_begin(c) + 128k + i.  Nothing about the
// rulebook is assumed: whatever does not fit the window of its kernel row is flagged and fetched from global memory by
// the conv kernel — for the raster-sorted submanifold tables this library builds that is a fraction of a percent.
__global__ void __launch_bounds__(BLOCK_M)
k_win_plan(const int* __restrict__ nbr, const int* __restrict__ num_rows, int rows_cap, int grid, int tiles_per_cta,
           uint8_t* __restrict__ plan) {
  __shared__ int s_lo[3];
  const int c = blockIdx.x / tiles_per_cta, k = blockIdx.x - c * tiles_per_cta;
  const int i = threadIdx.x;
  const int rows = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const int share = win_share(rows, grid);
  const int row_begin = min(rows, c * share), row_end = min(rows, row_begin + share);
  const int row0 = row_begin + k * BLOCK_M;
  if (row0 >= row_end) return;                 // tile not walked by the conv kernel
  const int row = row0 + i;
  int v[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) v[t] = row < row_end ? __ldg(nbr + (long long)row * 9 + t) : -1;
  if (i < 3) s_lo[i] = INT_MAX;
  __syncthreads();
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int m = INT_MAX;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
      if (v[ky * 3 + kx] >= 0) m = min(m, v[ky * 3 + kx]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((i & 31) == 0 && m != INT_MAX) atomicMin(&s_lo[ky], m);
  }
  __syncthreads();
  uint8_t* out = plan + (size_t)blockIdx.x * kPlanBytes;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int lo = s_lo[t / 3] == INT_MAX ? 0 : s_lo[t / 3];
    int code = kSrcAbsent;
    if (v[t] >= 0) code = (v[t] - lo < kWin) ? v[t] - lo : kSrcFar;
    out[t * BLOCK_M + i] = (uint8_t)code;
  }
  if (i < 4) reinterpret_cast<int*>(out + kPlanLo)[i] = (i < 3 && s_lo[i] != INT_MAX) ? s_lo[i] : 0;
}

struct WinGeom {
  int grid, tiles_per_cta;
};
inline WinGeom win_geom(int rows_cap) {
  const int sms = pn_detail::sm_count();
  const long long shares = PN_DIVUP((long long)rows_cap, (long long)BLOCK_M);
  WinGeom g;
  g.grid = (int)(shares < sms ? (shares < 1 ? 1 : shares) : sms);
  g.tiles_per_cta = PN_DIVUP(win_share(rows_cap, g.grid), BLOCK_M);
  return g;
}

template <int BN, int KU, bool RES, int NBT, int WS, int AS>
int launch(const CUtensorMap& map_w, const CUtensorMap& map_in, const CUtensorMap& map_o, const CUtensorMap& map_r,
           const WArgs& wa, int grid, cudaStream_t stream) {
  constexpr size_t smem = sizeof(WSmem<BN, KU, RES, NBT, WS, AS>) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static pn_detail::PerDeviceOnce once;
  if (once.need())
    PN_CUDA(cudaFuncSetAttribute(k_conv_win<BN, KU, RES, NBT, WS, AS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  static const bool timeline = [] { const char* e = getenv("PN_CONV_TIMELINE"); return e && e[0] == '1'; }();
  static unsigned long long* dbg_buf = nullptr;
  WArgs w = wa;
  if (timeline) {
    if (!dbg_buf) PN_CUDA(cudaMalloc(&dbg_buf, 16 * 1024 * sizeof(unsigned long long)));
    PN_CUDA(cudaMemsetAsync(dbg_buf, 0, 16 * 1024 * sizeof(unsigned long long), stream));
    PN_CUDA(cudaStreamSynchronize(stream));
    w.dbg = dbg_buf;
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
  PN_CUDA(cudaLaunchKernelEx(&cfg, k_conv_win<BN, KU, RES, NBT, WS, AS>, map_w, map_in, map_o, map_r, w));
  PN_CHECK_LAUNCH();
  if (timeline) {
    PN_CUDA(cudaStreamSynchronize(stream));
    static unsigned long long t[16 * 1024];
    PN_CUDA(cudaMemcpy(t, dbg_buf, sizeof(t), cudaMemcpyDeviceToHost));
    unsigned long long t_min = ~0ull, t_max = 0;
    int n = 0;
    double s_setup = 0, s_first = 0, s_mma = 0, s_epi = 0, s_tot = 0, m_mma = 0, s_w[8] = {0};
    for (int c = 0; c < grid && c < 1024; ++c) {
      const unsigned long long* q = t + c * 16;
      if (q[3] == 0) continue;
      ++n;
      if (q[0] < t_min) t_min = q[0];
      if (q[6] > t_max) t_max = q[6];
      const double mm = (double)(q[3] - q[2]);
      s_setup += (double)(q[1] - q[0]); s_first += (double)(q[2] - q[1]); s_mma += mm; s_epi += (double)(q[5] - q[4]);
      s_tot += (double)(q[6] - q[0]);
      if (mm > m_mma) m_mma = mm;
      for (int k = 0; k < 6; ++k) s_w[k] += (double)q[7 + k];
      s_w[6] += (double)q[15];
    }
    if (n > 0) {
      fprintf(stderr, "[conv_win<%d,%d,%s> cin %d cout %d rows_cap %d grid %d busy %d] span %.1f us | setup avg %.1f | first operands "
                      "avg %.1f | mma phase avg %.1f max %.1f | last epilogue avg %.1f | CTA total avg %.1f\n",
              BN, KU, RES ? "resident" : "streamed", wa.cin, wa.cout, wa.rows_cap, grid, n, (t_max - t_min) / 1e3, s_setup / n / 1e3, s_first / n / 1e3,
              s_mma / n / 1e3, m_mma / 1e3, s_epi / n / 1e3, s_tot / n / 1e3);
      fprintf(stderr, "    stalls per CTA (kclk; builder = warp 0): builder map %.1f, a_empty %.1f, win_full %.1f | mma a_full %.1f, b_full %.1f, "
                      "tmem_empty %.1f | tiles %.1f\n",
              s_w[0] / n / 1e3, s_w[1] / n / 1e3, s_w[2] / n / 1e3, s_w[3] / n / 1e3, s_w[4] / n / 1e3, s_w[5] / n / 1e3,
              s_w[6] / n);
    }
  }
  return PN_OK;
}

}  // namespace

namespace pn_detail {

// Returns PN_ERR_UNSUPPORTED when the layer is not one this kernel handles (the caller then uses the gather kernel).
int conv_win(const pn_conv_args* a, cudaStream_t stream) {
  static const bool enabled = [] { const char* e = getenv("PN_CONV_WIN"); return !(e && e[0] == '0'); }();
  if (!enabled || a->nbr_kind != PN_NBR_SUBM_SORTED || a->nbr_plan == nullptr) return PN_ERR_UNSUPPORTED;
  if (a->in_dtype != PN_BF16 || a->out_dtype != PN_BF16 || a->taps != 9 || a->nbr == nullptr || a->deconv_cout != 0 ||
      a->out_hp != 0 || a->in_rows <= 0)
    return PN_ERR_UNSUPPORTED;
  const int ku = a->cin == 32 ? 32 : 64;
  if (a->cin % ku != 0 || a->in_ld % 8 != 0 || (reinterpret_cast<uintptr_t>(a->in) & 15u) != 0) return PN_ERR_UNSUPPORTED;
  if (a->k_pad % 64 != 0 || (reinterpret_cast<uintptr_t>(a->weight) & 15u) != 0) return PN_ERR_UNSUPPORTED;
  if (a->cout > 256 || a->cout % 8 != 0) return PN_ERR_UNSUPPORTED;
  if (a->cout > 128 && ku != 64) return PN_ERR_UNSUPPORTED;
  if ((long long)a->in_rows * a->in_ld * 2 >= (1ll << 32)) return PN_ERR_UNSUPPORTED;
  // 256 output channels: two CTAs per row share, one per 128-column half.  (One 256-wide tile per CTA measured 18-20 us
  // against 13 us on the 7.8k-row stage-4 layers: 61 busy SMs, each MMA-bound.)
  const int bn = a->cout <= 32 ? 32 : a->cout <= 64 ? 64 : 128;
  CUtensorMap map_w, map_in, map_o;
  int rc = pn_tmap::get(a->weight, a->cout, a->k_pad, a->k_pad, 64, bn, CU_TENSOR_MAP_SWIZZLE_128B, &map_w);
  if (rc != PN_OK) return rc;
  rc = pn_tmap::get(a->in, a->in_rows, a->cin, a->in_ld, ku, kWin,
                    ku == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, &map_in);
  if (rc != PN_OK) return rc;
  WArgs w;
  w.in = reinterpret_cast<const __nv_bfloat16*>(a->in);
  w.in_ld = a->in_ld;
  w.nbr = a->nbr;
  const WinGeom geo = win_geom(a->rows_cap);
  if (geo.grid <= 0) return PN_ERR_CUDA;
  w.plan = reinterpret_cast<const uint8_t*>(a->nbr_plan);
  w.tiles_per_cta = geo.tiles_per_cta;
  w.plan_grid = geo.grid;
  w.n_split = PN_DIVUP(a->cout, bn);
  w.n_chunks = a->cin / ku;
  w.scale = a->scale;
  w.shift = a->shift;
  w.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  w.res_ld = a->res_ld;
  w.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  w.out_ld = a->out_ld;
  w.out_coff = a->out_coff;
  w.relu = a->relu;
  w.num_rows = a->num_rows;
  w.rows_cap = a->rows_cap;
  w.cin = a->cin;
  w.cout = a->cout;
  w.dbg = nullptr;
  static const bool tma_store_enabled = [] { const char* e = getenv("PN_CONV_TMA_STORE"); return !(e && e[0] == '0'); }();
  map_o = map_w;
  w.tma_store = 0;
  if (tma_store_enabled && a->cout % 32 == 0 && a->out_coff % 8 == 0 && a->out_ld % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(a->out) & 15u) == 0 &&
      pn_tmap::get(a->out, a->rows_cap, a->out_ld, a->out_ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, &map_o) == PN_OK)
    w.tma_store = 1;
  CUtensorMap map_r = map_w;
  w.tma_res = 0;
  if (tma_store_enabled && a->residual != nullptr && a->cout % 32 == 0 && a->res_ld % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(a->residual) & 15u) == 0 &&
      pn_tmap::get(a->residual, a->rows_cap, a->cout, a->res_ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, &map_r) == PN_OK)
    w.tma_res = 1;
  const int grid = PN_DIVUP(geo.grid, w.n_split) * w.n_split;
  // weights resident in shared memory when the whole layer fits beside the window ring (72 KB)
  const bool res = (long long)9 * w.n_chunks * bn * 128 <= 72 * 1024 && w.n_chunks == 1;
  // WS (window ring) is a multiple of AS (builder groups): a window slot is then always consumed by the same group,
  // which keeps every parity wait within one phase of its barrier.
  if (res) {
    if (bn == 32 && ku == 32) return launch<32, 32, true, 9, 6, 3>(map_w, map_in, map_o, map_r, w, grid, stream);
    if (bn == 32) return launch<32, 64, true, 9, 6, 3>(map_w, map_in, map_o, map_r, w, grid, stream);
    if (bn == 64 && ku == 32) return launch<64, 32, true, 9, 6, 3>(map_w, map_in, map_o, map_r, w, grid, stream);
    if (bn == 64) return launch<64, 64, true, 9, 6, 3>(map_w, map_in, map_o, map_r, w, grid, stream);
  }
  if (bn == 32) return launch<32, 64, false, 6, 6, 3>(map_w, map_in, map_o, map_r, w, grid, stream);
  if (bn == 64) return launch<64, 64, false, 6, 6, 3>(map_w, map_in, map_o, map_r, w, grid, stream);
  if (ku == 32) return launch<128, 32, false, 6, 4, 2>(map_w, map_in, map_o, map_r, w, grid, stream);
  return launch<128, 64, false, 6, 4, 2>(map_w, map_in, map_o, map_r, w, grid, stream);
}

}  // namespace pn_detail

extern "C" {

size_t pn_conv_window_plan_bytes(int rows_cap) {
  if (rows_cap <= 0) return 16;
  const WinGeom g = win_geom(rows_cap);
  return (size_t)g.grid * g.tiles_per_cta * kPlanBytes;
}

int pn_conv_window_plan(const int* nbr, const int* num_rows, int rows_cap, void* plan, size_t plan_bytes,
                        pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(rows_cap >= 0);
  if (rows_cap == 0) return PN_OK;
  PN_REQUIRE(nbr && plan && (reinterpret_cast<uintptr_t>(plan) & 15u) == 0);
  const WinGeom g = win_geom(rows_cap);
  if (g.grid <= 0) return PN_ERR_CUDA;
  if (plan_bytes < (size_t)g.grid * g.tiles_per_cta * kPlanBytes) return PN_ERR_WORKSPACE;
  k_win_plan<<<g.grid * g.tiles_per_cta, BLOCK_M, 0, stream>>>(nbr, num_rows, rows_cap, g.grid, g.tiles_per_cta,
                                                               reinterpret_cast<uint8_t*>(plan));
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
