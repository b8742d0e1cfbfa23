// Dense 3x3 / stride-1 / pad-1 convolution on tcgen05 over a ZERO-PADDED NHWC layout, fully TMA fed.
//
// Replaces the cuDNN Conv2d(3x3)+BatchNorm2d+ReLU stacks of the dense BEV path
// (det3d/models/backbones/PillarResNet.py:110-117 conv5, det3d/models/necks/rpn.py:172-185 blocks,
// det3d/models/bbox_heads/center_head.py:27-33,101-105 shared + first-level head convs).
//
// Layout trick: a (B,H,W,C) map is stored with a one-pixel zero border, flattened to rows
// q = (b*Hp + y)*Wp + x, Hp = H+2, Wp = W+2.  A 3x3 tap (dy,dx) of output q then reads input row
// q + (dy-1)*Wp + (dx-1): a CONSTANT row offset, for every pixel including the image border (the
// border rows are the padding).  So for a tile of 128*MT consecutive q and one 64-channel K chunk, the
// three taps of a kernel row dy are three windows, shifted by one 128-byte row, of ONE contiguous
// segment of 128*MT+2 activation rows — loaded once by TMA (SWIZZLE_128B) and addressed by three UMMA
// descriptors.  Compared with the gather formulation (conv_tcgen05.cu) this reads each activation row
// 3x instead of 9x per tile and needs no rulebook; with MT = 2 the 256-row CTA also reuses every weight
// tile for two MMAs, halving weight bytes per FLOP (the kernel is L2->SM-bandwidth bound).
//
// Warp roles (384 threads, persistent): warp 0 lane 0 = activation TMA producer, warp 1 lane 0 = weight
// TMA producer, warp 2 = TMEM owner + MMA issuer, warps 4-11 = epilogue (folded BN/bias, ReLU, zeroed
// border rows so the output is again a valid padded map — or compact rows for consumers that want
// un-padded NHWC).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace {

constexpr int BLOCK_K = 64;
constexpr int kThreads = 640;            // warps 0-2: producers + MMA, 3: idle, 4-19: epilogue
// 16 epilogue warps: four per TMEM lane quarter, each takes every fourth 32-column chunk.  Measured with per-phase
// clocks (8 warps): a 32x32 chunk cost ~1800 clk of a warp's time, its ~250 instructions issuing at one per ~7 clk
// (two warps per scheduler, dependent chains, TMEM / shared-memory latencies), so every short-K layer ran at the
// epilogue's pace (64 -> 2304 head conv: 94 us, 62 us with the stores compiled out).  Latency is hidden with warps,
// not with software pipelining: 102 registers per thread leave no room for a prefetched second chunk anyway.
constexpr int kEpilogueWarps = 16;
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kColSets = kEpilogueWarps / 4;
constexpr int kTailRows = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();  // protocol bug: trap instead of hanging the box
  }
}
// Long waits (epilogue waiting for a whole tile of MMAs): back off so the spinning warps do not steal
// issue slots from the producer / MMA warps that share their schedulers.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// multicast: the box lands at the same CTA-relative offset of every CTA in `mask`, and each destination
// CTA's barrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("PN_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// arrive on the same-offset barrier of every CTA in `mask` once this CTA's prior MMAs have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// ---- cta_group::2 (one UMMA across the two SMs of a CTA pair) ----
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issuing CTA's prior MMAs retired) on the same-offset barrier of both CTAs of the pair
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
// shared::cluster address of `p` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// TMA load into OUR shared memory whose completion bytes are credited to a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                                uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B descriptor whose start address may sit on ANY 128-byte row of a 1024-byte
// aligned tile (window shifted by dx rows).  Measured on B200 (tests/test_gpu_dense_conv.py probe): the
// 128B swizzle is a function of the absolute shared-memory address, so the shifted window needs NO
// base_offset (bits 49-51 = 0 gives max-abs error 7e-7; setting (addr >> 7) & 7 there gives garbage).
// base_offset_mode != 0 exists only for that probe.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int base_offset_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (base_offset_mode) d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct DArgs {
  int cin;         // multiple of 64
  int in_coff;     // first input channel inside the (rows, in_ld) matrix
  int cout;
  int Hp, Wp;      // padded map size
  int n_pos;       // B*Hp*Wp
  const float* scale;
  const float* shift;
  void* out;
  int out_f32;
  int out_ld;
  int out_coff;
  int out_compact; // 0: padded rows (borders zeroed); 1: compact rows b*H*W + (y-1)*W + (x-1)
  int relu;
  int base_offset_mode;
  // grouped mode (n_groups > 0): N tile g is an independent conv reading channels
  // [in_coff + g*cin, +cin), weight rows [g*BN, +BN), writing group_tab[g] = {out_coff, cout} columns
  int n_groups;
  const int* group_tab;
  // group-major ("planar") intermediate: channels [g*gc, (g+1)*gc) of a logical (rows, C) matrix live in
  // their own contiguous (n_pos, gc) map, map g at row offset g*n_pos of one tall (G*n_pos, gc) matrix.
  // out_group_cols > 0: this conv WRITES that layout (out_ld = gc); in_planar != 0: the grouped conv READS it.
  int out_group_cols;
  int in_planar;
  // development aid (PN_DENSE_TIMELINE=1): CTA 0 records %globaltimer at its pipeline milestones
  unsigned long long* dbg;
  int dbg_mode;   // development only: bit 0 = skip the global stores, bit 1 = skip the TMEM loads
  int tma_store;  // bf16 padded / planar output through shared memory + TMA tensor stores (tmap_o is valid)
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PN_DBG(slot)                                                        \
  do {                                                                      \
    if (P.dbg) P.dbg[blockIdx.x * 16 + slot] = gtime_ns();                   \
  } while (0)

template <int MT, int BN, int SA, int SB, int BROWS = BN>
struct DSmem {
  static constexpr int kSegRows = 128 * MT + kTailRows;
  alignas(1024) uint8_t a[SA][kSegRows * 128];
  alignas(1024) uint8_t b[SB][BROWS * 128];   // BROWS = BN, or BN/2 when a CTA pair shares every weight tile
  alignas(8) uint64_t full_a[SA];
  uint64_t empty_a[SA];
  uint64_t full_b[SB];
  uint64_t empty_b[SB];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  alignas(16) float2 ss[BN];       // {scale, shift} per output column of the current N tile, read two columns per LDS.128:
                                   // a broadcast LDS.32 costs the shared-memory pipe a full wavefront, like 128 useful bytes
  // Epilogue staging for TMA tensor stores: 32 rows x 32 bf16 columns (2 KB, SWIZZLE_64B) per epilogue warp.  A
  // thread owns an output ROW, so its direct 16-byte stores of a warp instruction hit 32 different lines; the LSU
  // charges per line (~2 clk each, like the gathers of conv_tcgen05.cu): 27 clk per (128 rows x column) measured,
  // which made every short-K layer epilogue bound.  Staged, the warp issues four conflict-free STS.128 and one lane
  // hands the 2 KB box to the TMA unit.  Only in the variants whose operand stages leave 16 KB.
  static constexpr bool kTmaStore =
      BN >= 32 && (size_t)SA * kSegRows * 128 + (size_t)SB * BROWS * 128 + 2048 + kEpilogueWarps * 2048 + 2 * BN * 4 + 1024 <= 227 * 1024;
  alignas(1024) uint8_t stage_out[kTmaStore ? kEpilogueWarps * 2048 : 16];
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// CL > 1: thread-block cluster of CL CTAs that work on CL consecutive row tiles of the SAME N tile; every
// weight tile is fetched once per cluster — each CTA loads BN/CL rows and TMA-multicasts them into all CL
// shared memories — so weight bytes per CTA drop CL-fold (the kernel is L2->SM bound and weights are ~75 % of
// its traffic at BN = 256).  A weight stage is recycled only after all CL consumers released it (multicast commit).
// BS ("B stationary", grouped mode only): the CTA works on ONE group (g = blockIdx.x % n_groups) for all its row
// tiles, so the group's nine weight tiles are loaded once into the SB (>= 9) weight slots and never recycled —
// no per-tap weight TMA, barrier wait or commit in the steady state.
// TWO (requires CL == 2): the pair of CTAs runs ONE tcgen05.mma.cta_group::2 per K step — M = 256 (128 rows from
// each CTA's activation segment, accumulated in each CTA's own TMEM), N = BN with each CTA holding half of the
// weight tile.  Every SM then ingests half of the weight bytes.  Rank 0 issues all MMAs; both CTAs' TMA loads
// credit rank 0's full barriers, its commits are multicast to both CTAs' empty / tmem_full barriers, and both
// epilogues release the accumulator on rank 0's tmem_empty barrier.
template <int MT, int BN, int SA, int SB, int CL, bool BS = false, bool TWO = false>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_dense(const __grid_constant__ CUtensorMap tmap_a_main, const __grid_constant__ CUtensorMap tmap_a_tail,
             const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_o, const DArgs P) {
  extern __shared__ uint8_t smem_raw[];
  static_assert(!TWO || (CL == 2 && !BS), "2-SM mode needs a cluster of exactly two CTAs");
  constexpr int BROWS = TWO ? BN / 2 : BN;
  using S = DSmem<MT, BN, SA, SB, BROWS>;
  S& sm = *reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int M_TILE = 128 * MT;
  constexpr int ACC_COLS = MT * BN;                       // fp32 columns per tile
  constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;     // double-buffer the accumulator when TMEM allows
  constexpr int TCOLS = (NACC * ACC_COLS) < 32 ? 32 : NACC * ACC_COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) PN_DBG(0);
  // Programmatic dependent launch: this grid may start during its predecessor's tail.  Weights, scale/shift and
  // the group table are constants, so the weight producer runs ahead unconditionally; the activation producer
  // and the epilogue (reads nothing, but overwrites `out`) wait for the predecessor first.
  pdl_launch_dependents();
  const int n_n_tiles = P.n_groups > 0 ? P.n_groups : (P.cout + BN - 1) / BN;
  const int n_m_tiles = (P.n_pos + M_TILE - 1) / M_TILE;
  // work unit = (N tile, group of CL consecutive row tiles); CTA `rank` of the cluster takes row tile
  // m_group*CL + rank (past the end for the last group: its loads are zero-filled, its stores masked)
  static_assert(!BS || (CL == 1 && SB >= 9), "B-stationary mode: no cluster, one slot per tap");
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  // BS: the loop variable is the row tile itself; CTAs sharing a group interleave over its row tiles
  const int g_fixed = BS ? (int)(blockIdx.x % n_n_tiles) : 0;
  const int ctas_of_g = BS ? ((int)gridDim.x - 1 - g_fixed) / n_n_tiles + 1 : 1;
  const int n_tiles = BS ? n_m_tiles : ((n_m_tiles + CL - 1) / CL) * n_n_tiles;
  const int unit0 = BS ? (int)(blockIdx.x / n_n_tiles) : (int)(blockIdx.x / CL);
  const int unit_step = BS ? ctas_of_g : (int)(gridDim.x / CL);
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  const int n_cc = P.cin / BLOCK_K;
  // K-loop rotation: CTA (cluster) c starts its (channel chunk, kernel row) walk at step c mod (3*n_cc).  With one
  // tile per CTA every CTA would otherwise request the very same weight tile at the very same moment (130 SMs
  // hitting one L2 line set); the accumulation order is free, so the walk is staggered instead.
  const int n_u = 3 * n_cc;
  const int rot = (BS || P.dbg_mode & 4) ? 0 : (int)((blockIdx.x / CL) % n_u);

  if (warp == 2) {
    if (lane == 0) {
      for (int s = 0; s < SA; ++s) { mbar_init(&sm.full_a[s], 1); mbar_init(&sm.empty_a[s], 1); }
      for (int s = 0; s < SB; ++s) { mbar_init(&sm.full_b[s], 1); mbar_init(&sm.empty_b[s], TWO ? 1 : CL); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sm.tmem_full[i], 1);
        mbar_init(&sm.tmem_empty[i], TWO ? 2 * kEpilogueThreads : kEpilogueThreads);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (TWO) tmem_alloc2<TCOLS>(&sm.tmem_base); else tmem_alloc<TCOLS>(&sm.tmem_base);
  }
  tcgen05_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // barriers of every CTA initialised before any remote signal
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  if (threadIdx.x == 0) PN_DBG(1);

  if (warp == 0) {
    // ===================== activation segments (TMA) =====================
    if (lane == 0) {
      pdl_wait();
      uint32_t g = 0;
      for (int tile = unit0; tile < n_tiles; tile += unit_step) {
        const int m_tile = BS ? tile : (tile / n_n_tiles) * CL + crank;
        const int q0 = m_tile * M_TILE;
        const int grp = BS ? g_fixed : (P.n_groups > 0 ? tile % n_n_tiles : 0);
        const int ch0 = P.in_coff + (P.in_planar ? 0 : grp * P.cin);
        // planar input: group g's map starts g*n_pos rows further down; rows that fall outside a map land in the
        // zero border rows of the neighbouring map (or outside the matrix: TMA zero fill) — zeros either way
        const int row_base = P.in_planar ? grp * P.n_pos : 0;
        for (int u = 0; u < n_u; ++u, ++g) {
          {
            int u2 = u + rot;
            if (u2 >= n_u) u2 -= n_u;
            const int cc = u2 / 3, dy = u2 - cc * 3;
            const uint32_t s = g % SA, ph = (g / SA) & 1u;
            mbar_wait(&sm.empty_a[s], ph ^ 1u);
            const int row = row_base + q0 + (dy - 1) * P.Wp - 1;   // may be negative / past the end: TMA zero-fills
            const int ch = ch0 + cc * BLOCK_K;
            if constexpr (TWO) {
              // the leader's barrier counts both CTAs' segments
              if (crank == 0) mbar_arrive_expect_tx(&sm.full_a[s], (uint32_t)(2 * S::kSegRows * 128));
              const uint32_t lb = map_to_cta(&sm.full_a[s], 0);
              tma_load_2d_2sm(smem_u32(sm.a[s]), &tmap_a_main, ch, row, lb);
              tma_load_2d_2sm(smem_u32(sm.a[s]) + M_TILE * 128, &tmap_a_tail, ch, row + M_TILE, lb);
            } else {
              mbar_arrive_expect_tx(&sm.full_a[s], (uint32_t)(S::kSegRows * 128));
              tma_load_2d(smem_u32(sm.a[s]), &tmap_a_main, ch, row, &sm.full_a[s]);
              tma_load_2d(smem_u32(sm.a[s]) + M_TILE * 128, &tmap_a_tail, ch, row + M_TILE, &sm.full_a[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== weight tiles (TMA) =====================
    if (lane == 0 && BS) {
      if (unit0 < n_tiles) {   // a CTA without row tiles must not leave a TMA in flight
        mbar_arrive_expect_tx(&sm.full_b[0], (uint32_t)(9 * BN * 128));
        for (int tap = 0; tap < 9; ++tap)
          tma_load_2d(smem_u32(sm.b[tap]), &tmap_w, tap * P.cin, g_fixed * BN, &sm.full_b[0]);
      }
    } else if (lane == 0) {
      uint32_t g = 0;
      for (int tile = unit0; tile < n_tiles; tile += unit_step) {
        const int n_tile = tile % n_n_tiles;
        for (int u = 0; u < n_u; ++u) {
          int u2 = u + rot;
          if (u2 >= n_u) u2 -= n_u;
          const int cc = u2 / 3, dy = u2 - cc * 3;
          for (int dx = 0; dx < 3; ++dx, ++g) {
            const int tap = dy * 3 + dx;
            const uint32_t s = g % SB, ph = (g / SB) & 1u;
            mbar_wait(&sm.empty_b[s], ph ^ 1u);
            if constexpr (TWO) {
              // this CTA's half of the weight tile, into its own shared memory; bytes credited to the leader
              if (crank == 0) mbar_arrive_expect_tx(&sm.full_b[s], (uint32_t)(BN * 128));
              tma_load_2d_2sm(smem_u32(sm.b[s]), &tmap_w, tap * P.cin + cc * BLOCK_K, n_tile * BN + crank * BROWS,
                              map_to_cta(&sm.full_b[s], 0));
              continue;
            }
            mbar_arrive_expect_tx(&sm.full_b[s], (uint32_t)(BN * 128));
            if (CL == 1) {
              tma_load_2d(smem_u32(sm.b[s]), &tmap_w, tap * P.cin + cc * BLOCK_K, n_tile * BN, &sm.full_b[s]);
            } else {
              constexpr int kSlice = BN / CL;   // rows of the tile this CTA fetches for the whole cluster
              tma_load_2d_mc(smem_u32(sm.b[s]) + crank * kSlice * 128, &tmap_w, tap * P.cin + cc * BLOCK_K,
                             n_tile * BN + crank * kSlice, &sm.full_b[s], kMask);
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop converged (barrier waits, stage counters and descriptors are warp-uniform, so
    // they live in uniform registers); one elected lane issues the tcgen05 instructions.  Entering the loop with a
    // single active lane instead made the compiler wrap every UTCHMMA in an ELECT / broadcast / retry sequence.
    const bool issuer = elect_one();
    if (!TWO || crank == 0) {
      // The issuing thread is a single in-order instruction stream: at BN = 128 an MMA retires every
      // 64 cycles, so descriptor arithmetic per MMA must be a couple of integer adds.  Stage base
      // descriptors are built once; window / K-step offsets are compile-time constants added to the
      // low word (the 14-bit start-address field never carries within a 192 KB tile region).
      // cta_group::2: M = 256 in the instruction descriptor (bits 24-28 hold M >> 4)
      constexpr uint32_t idesc = TWO ? (make_idesc<BN>() + ((uint32_t)(128 >> 4) << 24)) : make_idesc<BN>();
      // Lean issue loop (SASS-checked): the first version spent ~90 instructions per tap on stage select chains,
      // modulo/divide stage counters and timing reads — ~450 clk of one thread's dependent issue against the 256 clk
      // four N=128 MMAs take, which made every small-tile layer issue bound.  Stage descriptors are now
      // base + stage * stride (stages are contiguous arrays), stage/phase counters are incremental.
      const uint64_t a_desc0 = make_desc(smem_u32(sm.a[0]), 0);
      const uint64_t b_desc0 = make_desc(smem_u32(sm.b[0]), 0);
      constexpr uint32_t kAStep = (uint32_t)(S::kSegRows * 128) >> 4;   // descriptor units (16 B) between A stages
      constexpr uint32_t kBStep = (uint32_t)(BROWS * 128) >> 4;
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tcount = 0;
      bool first_tap = true, a_ahead = false, b_ahead = false;
      long long w_fa = 0, w_fb = 0, w_te = 0;          // PN_DENSE_TIMELINE: clocks the MMA warp waited for operands / TMEM
#define DW_T0() const long long _w0 = P.dbg ? clock64() : 0
#define DW_ACC(v) do { if (P.dbg) v += clock64() - _w0; } while (0)
      const bool look_ahead = (P.dbg_mode & 8) == 0;   // PN_DENSE_DBGMODE=8 disables it (A/B measurements)
      if (BS && unit0 < n_tiles) {
        mbar_wait(&sm.full_b[0], 0);
        tcgen05_fence_after();
      }
      for (int tile = unit0; tile < n_tiles; tile += unit_step, ++tcount) {
        const uint32_t acc = NACC == 2 ? (tcount & 1u) : 0u;
        const uint32_t acc_ph = NACC == 2 ? ((tcount >> 1) & 1u) : (tcount & 1u);
        { DW_T0(); mbar_wait(&sm.tmem_empty[acc], acc_ph ^ 1u); DW_ACC(w_te); }
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        int dy = rot % 3;                       // kernel row of this step (the walk is rotated per CTA)
        const bool more_tiles = tile + unit_step < n_tiles;
        for (int u = 0; u < n_u; ++u) {
          if (!a_ahead) { DW_T0(); mbar_wait(&sm.full_a[sa], pha); DW_ACC(w_fa); }
          a_ahead = false;
          const uint64_t a_stage = a_desc0 + (uint64_t)(sa * kAStep);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            uint64_t b_stage;
            if constexpr (BS) {
              b_stage = b_desc0 + (uint64_t)((uint32_t)(dy * 3 + dx) * kBStep);   // resident weight slot of this tap
              if (dx == 0) tcgen05_fence_after();
            } else {
              if (!b_ahead) { DW_T0(); mbar_wait(&sm.full_b[sb], phb); DW_ACC(w_fb); }
              b_ahead = false;
              tcgen05_fence_after();
              b_stage = b_desc0 + (uint64_t)(sb * kBStep);
            }
            if (first_tap) { if (issuer) PN_DBG(2); first_tap = false; }
            const uint32_t first = (u == 0 && dx == 0) ? 0u : 1u;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) {
                if (m == MT - 1 && k == BLOCK_K / 16 - 1) {
                  // Look ahead before the tap's LAST MMA: the barrier poll of the next tap's operands (~90 clk even
                  // when they have landed) then runs while the MMAs issued so far execute, instead of after them with
                  // the tensor pipe drained (measured before: a tap cost its MMAs + ~150 clk).  No deadlock: the
                  // stage waited for was released by a commit issued SA / SB taps ago, never by the pending one.
                  if (look_ahead && !(u == n_u - 1 && dx == 2 && !more_tiles)) {
                    if constexpr (!BS) {
                      const uint32_t sbn = sb + 1 == SB ? 0u : sb + 1, phbn = sb + 1 == SB ? phb ^ 1u : phb;
                      { DW_T0(); mbar_wait(&sm.full_b[sbn], phbn); DW_ACC(w_fb); }
                      b_ahead = true;
                    }
                    if (dx == 2) {
                      const uint32_t san = sa + 1 == SA ? 0u : sa + 1, phan = sa + 1 == SA ? pha ^ 1u : pha;
                      { DW_T0(); mbar_wait(&sm.full_a[san], phan); DW_ACC(w_fa); }
                      a_ahead = true;
                    }
                  }
                }
                // (m*128 + dx) rows * 128 B + k * 32 B, in 16-byte units
                const uint64_t a_desc = a_stage + (uint64_t)((m * 128 + dx) * 8 + k * 2);
                const uint64_t b_desc = b_stage + (uint64_t)(k * 2);
                if (issuer) {
                  if constexpr (TWO) umma_bf16_2(d_tmem + m * BN, a_desc, b_desc, idesc, k == 0 ? first : 1u);
                  else umma_bf16(d_tmem + m * BN, a_desc, b_desc, idesc, k == 0 ? first : 1u);
                }
              }
            }
            if (issuer) {
              if constexpr (TWO) {
                umma_commit2_mc(&sm.empty_b[sb]);
              } else if constexpr (!BS) {
                if (CL == 1) umma_commit(&sm.empty_b[sb]); else umma_commit_mc(&sm.empty_b[sb], kMask);
              }
            }
            if constexpr (!BS) {
              if (++sb == SB) { sb = 0; phb ^= 1u; }
            }
          }
          if (issuer) {
            if constexpr (TWO) umma_commit2_mc(&sm.empty_a[sa]); else umma_commit(&sm.empty_a[sa]);
          }
          if (++sa == SA) { sa = 0; pha ^= 1u; }
          if (++dy == 3) dy = 0;
        }
        if (issuer) {
          if constexpr (TWO) umma_commit2_mc(&sm.tmem_full[acc]); else umma_commit(&sm.tmem_full[acc]);
          PN_DBG(3);
        }
        __syncwarp();
      }
      if (issuer && P.dbg) {
        P.dbg[blockIdx.x * 16 + 8] = w_fa; P.dbg[blockIdx.x * 16 + 9] = w_fb; P.dbg[blockIdx.x * 16 + 10] = w_te;
      }
#undef DW_T0
#undef DW_ACC
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int e = (warp - 4) & 3;          // TMEM lane quarter this warp may access (warp id % 4)
    const int cset = (warp - 4) >> 2;      // which of the kColSets interleaved column-chunk sets
    const int etid = threadIdx.x - 4 * 32;
    const int hw_p = P.Hp * P.Wp;
    uint32_t tcount = 0;
    pdl_wait();
    for (int tile = unit0; tile < n_tiles; tile += unit_step, ++tcount) {
      const int n_tile = BS ? g_fixed : tile % n_n_tiles;
      const int m_tile = BS ? tile : (tile / n_n_tiles) * CL + crank;
      const int n0 = n_tile * BN;          // first weight row / scale index of this N tile
      int cout_t = P.cout, ocol0 = P.out_coff + n0, nbase = n0;
      if (P.n_groups > 0) {                // grouped: own output columns and channel count
        ocol0 = __ldg(P.group_tab + 2 * n_tile);
        cout_t = __ldg(P.group_tab + 2 * n_tile + 1);
        nbase = 0;
      }
      const uint32_t acc = NACC == 2 ? (tcount & 1u) : 0u;
      const uint32_t acc_ph = NACC == 2 ? ((tcount >> 1) & 1u) : (tcount & 1u);
      named_bar_sync(2, kEpilogueThreads);
      for (int i = etid; i < BN; i += kEpilogueThreads) {
        const bool ok = nbase + i < cout_t;
        sm.ss[i] = make_float2((ok && P.scale) ? __ldg(P.scale + n0 + i) : 1.f, (ok && P.shift) ? __ldg(P.shift + n0 + i) : 0.f);
      }
      named_bar_sync(2, kEpilogueThreads);
      mbar_wait_relaxed(&sm.tmem_full[acc], acc_ph);
      tcgen05_fence_after();
      if (etid == 0) PN_DBG(4);
      bool released = false;   // the accumulator goes back to the MMA warp once this thread's last chunk is in registers
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int q0 = m_tile * M_TILE + m * 128 + e * 32;   // first of this warp's 32 rows
        const bool box_ok = q0 + 32 <= P.n_pos;
        const int q = q0 + lane;
        const bool valid = q < P.n_pos;
        const int b = valid ? q / hw_p : 0;
        const int r = q - b * hw_p;
        const int y = r / P.Wp, x = r - y * P.Wp;
        const bool border = (x == 0) || (x == P.Wp - 1) || (y == 0) || (y == P.Hp - 1);
        long long orow = q;
        bool store = valid;
        if (P.out_compact) {
          store = valid && !border;
          orow = ((long long)b * (P.Hp - 2) + (y - 1)) * (P.Wp - 2) + (x - 1);
        }
        constexpr int CH = BN < 32 ? 16 : 32;
        const int n_ch = min(BN, cout_t - nbase + CH - 1) / CH;   // warp-uniform number of live chunks
        const uint32_t tbase = tmem_base + ((uint32_t)(e * 32) << 16) + acc * ACC_COLS + m * BN;
#pragma unroll 1
        for (int ci = cset; ci < n_ch; ci += kColSets) {
          const int c0 = ci * CH;
          uint32_t w[32];
          if (!(P.dbg_mode & 2)) {
            if (CH == 32) tmem_ld32(tbase + c0, w); else tmem_ld16(tbase + c0, w);
            tmem_wait_ld();
          }
          if (m == MT - 1 && ci + kColSets >= n_ch) {
            tcgen05_fence_before();
            if constexpr (TWO) mbar_arrive_cluster(map_to_cta(&sm.tmem_empty[acc], 0));   // the leader's MMA thread waits
            else mbar_arrive(&sm.tmem_empty[acc]);
            released = true;
          }
          // Fast path (every bf16 padded / planar output of the neck and head): the warp's 32 rows exist, the chunk is
          // whole -> folded affine, ReLU and the bf16 pack in one pass (cvt.rn.relu.bf16x2), border rows zeroed on the
          // packed words, four conflict-free STS.128 into the warp's staging box, one TMA tensor store.
          if (S::kTmaStore && CH == 32 && P.tma_store && box_ok && !P.out_f32 && cout_t - (nbase + c0) >= CH &&
              !(P.dbg_mode & 1)) {
            const float4* ss4 = reinterpret_cast<const float4*>(&sm.ss[c0]);   // c0 % 16 == 0: 16-byte aligned
            const bool affine = P.scale != nullptr || P.shift != nullptr;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float t0 = __uint_as_float(w[2 * j]), t1 = __uint_as_float(w[2 * j + 1]);
              if (affine) {
                const float4 s2 = ss4[j];
                t0 = fmaf(t0, s2.x, s2.y);
                t1 = fmaf(t1, s2.z, s2.w);
              }
              if (P.relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(t1), "f"(t0));
              else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(t1), "f"(t0));
            }
            if (border) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = 0u;
            }
            uint8_t* stg = sm.stage_out + (warp - 4) * 2048;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous box was read
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // SWIZZLE_64B: 16-byte chunk index ^ address bits 7-8 = (row >> 1) & 3 for 64-byte rows
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              int col = ocol0 + c0;
              int row0 = q0;
              if (P.out_group_cols > 0) {
                const int g = col / P.out_group_cols;
                col -= g * P.out_group_cols;
                row0 += g * P.n_pos;
              }
              tma_store_2d(&tmap_o, smem_u32(stg), col, row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
          }
          if (store && !(P.dbg_mode & 1)) {
            const int nvalid = min(CH, cout_t - (nbase + c0));
            float f[32];
            const float4* ss4 = reinterpret_cast<const float4*>(&sm.ss[c0]);   // c0 % 16 == 0: 16-byte aligned
            const bool affine = P.scale != nullptr || P.shift != nullptr;
#pragma unroll
            for (int j = 0; j < CH; j += 2) {
              float t0 = __uint_as_float(w[j]), t1 = __uint_as_float(w[j + 1]);
              if (affine) {
                const float4 s2 = ss4[j >> 1];
                t0 = fmaf(t0, s2.x, s2.y);
                t1 = fmaf(t1, s2.z, s2.w);
              }
              if (P.relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); }
              f[j] = (border && !P.out_compact) ? 0.f : t0;
              f[j + 1] = (border && !P.out_compact) ? 0.f : t1;
            }
            long long ooff = orow * P.out_ld + ocol0 + c0;
            if (P.out_group_cols > 0) {   // planar output: a CH-wide chunk never straddles two maps (gc % CH == 0)
              const int col = ocol0 + c0;
              const int g = col / P.out_group_cols;
              ooff = ((long long)g * P.n_pos + orow) * P.out_ld + (col - g * P.out_group_cols);
            }
            if (P.out_f32) {
              float* op = reinterpret_cast<float*>(P.out) + ooff;
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
              if (nvalid == CH && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
                for (int j = 0; j < CH; j += 8) {
                  uint4 qv;
                  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&qv);
#pragma unroll
                  for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(f[j + 2 * u], f[j + 2 * u + 1]);
                  *reinterpret_cast<uint4*>(op + j) = qv;
                }
              } else {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                  if (j < nvalid) op[j] = __float2bfloat16_rn(f[j]);
              }
            }
          }
        }
      }
      if (!released) {         // a warp without a chunk of this tile
        tcgen05_fence_before();
        if constexpr (TWO) mbar_arrive_cluster(map_to_cta(&sm.tmem_empty[acc], 0));
        else mbar_arrive(&sm.tmem_empty[acc]);
      }
      if (etid == 0) PN_DBG(5);
    }
    if (S::kTmaStore && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged boxes are out
  }
  tcgen05_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA may exit while peers still signal its barriers
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (TWO) tmem_dealloc2<TCOLS>(tmem_base); else tmem_dealloc<TCOLS>(tmem_base);
  }
  if (threadIdx.x == 0) PN_DBG(6);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

struct Key {
  const void* ptr;
  long long rows;
  int cols, ld, box_rows;
  bool operator==(const Key& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.rows * 1000003u ^ (size_t)k.cols * 10007u ^
           (size_t)k.ld * 131u ^ (size_t)k.box_rows;
  }
};

// 2-D bf16 tensor map over a row-major (rows, cols) matrix with row stride ld; box = 64 cols x box_rows.
int get_map(const void* base, long long rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  Key key{base, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return PN_OK; }
  }
  EncodeTiledFn enc = encode_fn();
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
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  Key key{base, rows, cols, ld, -32};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return PN_OK; }
  }
  EncodeTiledFn enc = encode_fn();
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

// `units` = work units (see the kernel); the grid is CL x min(units, co-resident clusters).
template <int MT, int BN, int SA, int SB, int CL = 1, bool BS = false, bool TWO = false>
int launch(const CUtensorMap& ma, const CUtensorMap& mt, const CUtensorMap& mw, const CUtensorMap& mo, const DArgs& a,
           long long units, cudaStream_t stream) {
  constexpr size_t smem = sizeof(DSmem<MT, BN, SA, SB, TWO ? BN / 2 : BN>) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static_assert(CL == 1 || (BN / CL) % 8 == 0, "a weight slice must keep the 8-row swizzle period");
  static int max_clusters = 0;
  auto kern = k_conv_dense<MT, BN, SA, SB, CL, BS, TWO>;
  static pn_detail::PerDeviceOnce once;
  if (once.need()) PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (max_clusters == 0) {
    const int sms = pn_detail::sm_count();
    if (CL == 1) {
      max_clusters = sms;
    } else {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(sms / CL * CL);
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      int n = 0;
      PN_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
      if (n <= 0) return PN_ERR_UNSUPPORTED;
      max_clusters = n;
    }
  }
  // balanced persistent grid: every CTA (cluster) gets the same number of work units (260 units on 148 SMs run
  // as 130 x 2, not 148 with a ragged second wave)
  const long long per_cta = PN_DIVUP(units, (long long)max_clusters);
  const int clusters = (int)PN_DIVUP(units, per_cta);
  static const bool timeline = [] { const char* e = getenv("PN_DENSE_TIMELINE"); return e && e[0] == '1'; }();
  static unsigned long long* dbg_buf = nullptr;
  DArgs a_dbg = a;
  { const char* e = getenv("PN_DENSE_DBGMODE"); a_dbg.dbg_mode = e ? atoi(e) : 0; }
  if (timeline) {
    if (!dbg_buf) PN_CUDA(cudaMalloc(&dbg_buf, 16 * 1024 * sizeof(unsigned long long)));   // device memory: no page faults
    PN_CUDA(cudaMemsetAsync(dbg_buf, 0, 16 * 1024 * sizeof(unsigned long long), stream));
    PN_CUDA(cudaStreamSynchronize(stream));
    a_dbg.dbg = dbg_buf;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  if (CL > 1) {
    at[n_at].id = cudaLaunchAttributeClusterDimension;
    at[n_at].val.clusterDim.x = CL; at[n_at].val.clusterDim.y = 1; at[n_at].val.clusterDim.z = 1;
    ++n_at;
  }
  if (pdl_enabled()) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  cfg.attrs = at;
  cfg.numAttrs = n_at;
  PN_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mt, mw, mo, a_dbg));
  PN_CHECK_LAUNCH();
  if (timeline) {
    PN_CUDA(cudaStreamSynchronize(stream));
    static unsigned long long t[16 * 1024];
    PN_CUDA(cudaMemcpy(t, dbg_buf, sizeof(t), cudaMemcpyDeviceToHost));
    const int n = clusters * CL;
    unsigned long long t_min = ~0ull, t_max = 0;
    for (int c = 0; c < n; ++c) {
      if (t[c * 16] < t_min) t_min = t[c * 16];
      if (t[c * 16 + 6] > t_max) t_max = t[c * 16 + 6];
    }
    double s_start = 0, s_first = 0, s_mma = 0, s_epi = 0, s_tot = 0, m_start = 0, m_mma = 0, m_tot = 0, m_epi = 0;
    double s_w[3] = {0, 0, 0};
    int n_issuers = 0;
    for (int c = 0; c < n; ++c) {
      const unsigned long long* q = t + c * 16;
      const double st = (double)(q[0] - t_min), fi = (double)(q[2] - q[1]), mm = (double)(q[3] - q[2]),
                   ep = (double)(q[5] - q[4]), to = (double)(q[6] - q[0]);
      s_start += st; s_first += fi; s_mma += mm; s_epi += ep; s_tot += to;
      if (q[8] + q[9] + q[10] > 0) { ++n_issuers; for (int k = 0; k < 3; ++k) s_w[k] += (double)q[8 + k]; }
      if (st > m_start) m_start = st;
      if (mm > m_mma) m_mma = mm;
      if (to > m_tot) m_tot = to;
      if (ep > m_epi) m_epi = ep;
    }
    fprintf(stderr, "[dense<%d,%d,%d,%d,cl%d> grid %d units %lld] span %.1f us | CTA start skew avg %.1f max %.1f | first "
                    "operands avg %.1f | mma phase avg %.1f max %.1f | last epilogue avg %.1f max %.1f | CTA total avg %.1f "
                    "max %.1f\n",
            MT, BN, SA, SB, CL, n, units, (t_max - t_min) / 1e3, s_start / n / 1e3, m_start / 1e3, s_first / n / 1e3,
            s_mma / n / 1e3, m_mma / 1e3, s_epi / n / 1e3, m_epi / 1e3, s_tot / n / 1e3, m_tot / 1e3);
    if (n_issuers > 0)
      fprintf(stderr, "    MMA warp waits per issuing CTA (kclk): full_a %.1f, full_b %.1f, tmem_empty %.1f\n",
              s_w[0] / n_issuers / 1e3, s_w[1] / n_issuers / 1e3, s_w[2] / n_issuers / 1e3);
  }
  return PN_OK;
}

}  // namespace

namespace {

int dense_run(const void* in, int in_ld, int in_coff, int cin, int n_frames, int H, int W,
              const void* weight, int k_pad, int cout, const float* scale, const float* shift,
              void* out, int out_dtype, int out_ld, int out_coff, int out_compact, int out_group_cols,
              int relu, int tile_hint, cudaStream_t stream) {
  PN_REQUIRE(in && weight && out && n_frames >= 1 && H > 0 && W > 0 && cout >= 1);
  PN_REQUIRE(cin % BLOCK_K == 0 && in_ld % 8 == 0 && in_coff % 8 == 0 && k_pad % BLOCK_K == 0 && k_pad >= 9 * cin);
  PN_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15u) == 0);
  PN_REQUIRE(out_dtype == PN_F32 || out_dtype == PN_BF16);
  // planar output: padded rows only, whole maps, 32-column chunks must not straddle maps
  PN_REQUIRE(out_group_cols == 0 || (out_group_cols % 32 == 0 && !out_compact && out_coff == 0 &&
                                     cout % out_group_cols == 0 && out_ld == out_group_cols));
  const int Hp = H + 2, Wp = W + 2;
  const long long n_pos = (long long)n_frames * Hp * Wp;
  PN_REQUIRE(n_pos < (1ll << 31));
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  // tile shape: minimise waves x per-tap cost, cost = max(load bytes / ~20 B/clk, MMA cycles)
  struct Cand { int mt, bn; };
  const Cand cands[4] = {{2, 256}, {1, 256}, {2, 128}, {1, 128}};
  int best = -1;
  double best_cost = 0;
  for (int i = 0; i < 4; ++i) {
    const int mt = cands[i].mt, bn = cands[i].bn;
    if (bn > 128 && cout <= 128) continue;
    const long long tiles = PN_DIVUP(n_pos, (long long)(128 * mt)) * PN_DIVUP(cout, bn);
    const double waves = (double)PN_DIVUP(tiles, (long long)sms);
    const double bytes = bn * 128.0 + (128.0 * mt + kTailRows) * 128.0 / 3.0;
    const double mma = mt * 4.0 * (bn / 2.0);
    // measured with PN_DENSE_TIMELINE after the issue loop was made lean: a tap costs its MMAs plus ~150 clk of
    // barrier hand-shakes, or its bytes at ~56 B/clk/SM, whichever is larger; the epilogue of a tile costs ~27 clk
    // per (128-row block x column) — a burst of global stores — and is hidden behind the next tile's MMAs only
    // when TMEM holds two accumulators; the last one never is
    const double per_tap = bytes / 56.0 > mma + 150.0 ? bytes / 56.0 : mma + 150.0;
    const double tile_clk = 9.0 * (cin / 64) * per_tap;
    const double epi_clk = mt * bn * 27.0;
    const bool two_acc = 2 * mt * bn <= 512;
    const double cost = waves * tile_clk + (two_acc ? 1.0 : waves) * epi_clk;
    if (best < 0 || cost < best_cost) { best = i; best_cost = cost; }
  }
  if ((tile_hint & 0xF) >= 1 && (tile_hint & 0xF) <= 4) best = (tile_hint & 0xF) - 1;
  const int mt = cands[best].mt, bn = cands[best].bn;
  // cluster size for the weight multicast: tile_hint bits 9/10 force 2 / 4, bit 11 forces 1
  int cl = 1;   // measured on B200: multicast of the weight tiles buys nothing here (the kernel is MMA/latency bound)
  if (tile_hint & 0x200) cl = 2;
  if (tile_hint & 0x400) cl = 4;
  if (tile_hint & 0x800) cl = 1;
  const bool two_sm = (tile_hint & 0x1000) != 0;      // cta_group::2 pair kernel
  if (two_sm) cl = 2;
  const long long m_tiles = PN_DIVUP(n_pos, (long long)(128 * mt));
  if (m_tiles < cl) cl = 1;
  CUtensorMap ma, mtail, mw;
  int rc = get_map(in, n_pos, in_ld, in_ld, 128 * mt, &ma);
  if (rc != PN_OK) return rc;
  rc = get_map(in, n_pos, in_ld, in_ld, kTailRows, &mtail);
  if (rc != PN_OK) return rc;
  rc = get_map(weight, cout, k_pad, k_pad, bn / cl, &mw);   // multicast slice or the pair's half tile: bn/2 rows
  if (rc != PN_OK) return rc;
  DArgs a;
  a.cin = cin; a.in_coff = in_coff; a.cout = cout; a.Hp = Hp; a.Wp = Wp; a.n_pos = (int)n_pos;
  a.scale = scale; a.shift = shift; a.out = out; a.out_f32 = out_dtype == PN_F32; a.out_ld = out_ld;
  a.out_coff = out_coff; a.out_compact = out_compact; a.relu = relu;
  a.base_offset_mode = (tile_hint & 0x100) ? 1 : 0;
  a.n_groups = 0;
  a.group_tab = nullptr;
  a.dbg = nullptr;
  a.out_group_cols = out_group_cols;
  a.in_planar = 0;
  // epilogue through TMA tensor stores: bf16 rows of the padded (or planar) map, whole 32-column chunks
  static const bool tma_store_enabled = [] { const char* e = getenv("PN_DENSE_TMA_STORE"); return !(e && e[0] == '0'); }();
  CUtensorMap mo = mw;
  a.tma_store = 0;
  if (tma_store_enabled && out_dtype == PN_BF16 && !out_compact && cout % 32 == 0 && out_coff % 8 == 0 &&
      out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const long long o_rows = out_group_cols > 0 ? (long long)(cout / out_group_cols) * n_pos : n_pos;
    if (o_rows < (1ll << 31) && get_map_out(out, o_rows, out_ld, out_ld, &mo) == PN_OK) a.tma_store = 1;
  }
  const long long units = PN_DIVUP(m_tiles, (long long)cl) * PN_DIVUP(cout, bn);
  if (two_sm && m_tiles >= 2) {
    // half-size weight stages: the freed shared memory buys deeper pipelines
    if (mt == 2 && bn == 256) return launch<2, 256, 3, 5, 2, false, true>(ma, mtail, mw, mo, a, units, stream);
    if (mt == 1 && bn == 256) return launch<1, 256, 4, 7, 2, false, true>(ma, mtail, mw, mo, a, units, stream);
    if (mt == 2 && bn == 128) return launch<2, 128, 4, 6, 2, false, true>(ma, mtail, mw, mo, a, units, stream);
    return launch<1, 128, 6, 10, 2, false, true>(ma, mtail, mw, mo, a, units, stream);
  }
#define PN_DENSE_LAUNCH(MT_, BN_, SA_, SB_)                                                       \
  do {                                                                                            \
    if (cl == 4) return launch<MT_, BN_, SA_, SB_, 4>(ma, mtail, mw, mo, a, units, stream);           \
    if (cl == 2) return launch<MT_, BN_, SA_, SB_, 2>(ma, mtail, mw, mo, a, units, stream);           \
    return launch<MT_, BN_, SA_, SB_, 1>(ma, mtail, mw, mo, a, units, stream);                        \
  } while (0)
  // stage counts leave 32 KB for the epilogue's staging boxes (one 2 KB box per epilogue warp)
  if (mt == 2 && bn == 256) PN_DENSE_LAUNCH(2, 256, 2, 3);
  if (mt == 1 && bn == 256) PN_DENSE_LAUNCH(1, 256, 3, 4);
  if (mt == 2 && bn == 128) PN_DENSE_LAUNCH(2, 128, 3, 5);
  PN_DENSE_LAUNCH(1, 128, 4, 7);
#undef PN_DENSE_LAUNCH
}

// First-use tile selection.  The cost model in dense_run ranks the four tile shapes from first principles; measured
// (tools/kbench_dense.py) it misses by up to 8 %: e.g. 128x256 tiles win every 180 x 180 layer of the neck although the
// model prefers 256x128.  So the first un-hinted call of a shape that is not inside a stream capture runs every
// candidate (1 warm + 3 timed launches, CUDA events on the caller's stream; the output is simply written again) and
// remembers the fastest; calls during a capture before that use the model.  PN_DENSE_AUTOTUNE=0 keeps the model.
struct TuneKey {
  int dev, cin, cout, n_frames, H, W, out_dtype, out_compact, planar;
  bool operator==(const TuneKey& o) const {
    return dev == o.dev && cin == o.cin && cout == o.cout && n_frames == o.n_frames && H == o.H && W == o.W &&
           out_dtype == o.out_dtype && out_compact == o.out_compact && planar == o.planar;
  }
};
struct TuneKeyHash {
  size_t operator()(const TuneKey& k) const {
    size_t h = (size_t)k.dev;
    for (int v : {k.cin, k.cout, k.n_frames, k.H, k.W, k.out_dtype, k.out_compact, k.planar}) h = h * 1000003u + (size_t)v;
    return h;
  }
};

}  // namespace

extern "C" {

int pn_conv_dense3x3(const void* in, int in_ld, int in_coff, int cin, int n_frames, int H, int W,
                     const void* weight, int k_pad, int cout, const float* scale, const float* shift,
                     void* out, int out_dtype, int out_ld, int out_coff, int out_compact, int out_group_cols,
                     int relu, int tile_hint, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  static const bool autotune = [] { const char* e = getenv("PN_DENSE_AUTOTUNE"); return !(e && e[0] == '0'); }();
  if ((tile_hint & 0xF) == 0 && autotune) {
    static std::mutex mu;
    static std::unordered_map<TuneKey, int, TuneKeyHash> tuned;
    TuneKey key{0, cin, cout, n_frames, H, W, out_dtype, out_compact, out_group_cols > 0 ? 1 : 0};
    if (cudaGetDevice(&key.dev) != cudaSuccess) return PN_ERR_CUDA;
    int choice = 0;
    {
      std::lock_guard<std::mutex> lk(mu);
      auto it = tuned.find(key);
      if (it != tuned.end()) choice = it->second;
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (choice == 0 && cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
      cudaEvent_t e0, e1;
      if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return PN_ERR_CUDA;
      float best_ms = 0.f;
      for (int c = 1; c <= 4; ++c) {
        if (c <= 2 && cout <= 128) continue;           // 256-wide tiles are not offered for narrow layers
        float ms_min = 0.f;
        bool ok = true;
        for (int rep = 0; rep < 4 && ok; ++rep) {
          cudaEventRecord(e0, stream);
          const int rc = dense_run(in, in_ld, in_coff, cin, n_frames, H, W, weight, k_pad, cout, scale, shift, out, out_dtype,
                                   out_ld, out_coff, out_compact, out_group_cols, relu, tile_hint | c, stream);
          cudaEventRecord(e1, stream);
          if (rc != PN_OK || cudaEventSynchronize(e1) != cudaSuccess) { ok = false; break; }
          float ms = 0.f;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep >= 1 && (ms_min == 0.f || ms < ms_min)) ms_min = ms;
        }
        if (ok && ms_min > 0.f && (choice == 0 || ms_min < best_ms)) { choice = c; best_ms = ms_min; }
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      if (choice != 0) {
        std::lock_guard<std::mutex> lk(mu);
        tuned[key] = choice;
      }
    }
    tile_hint |= choice;
  }
  return dense_run(in, in_ld, in_coff, cin, n_frames, H, W, weight, k_pad, cout, scale, shift, out, out_dtype, out_ld,
                   out_coff, out_compact, out_group_cols, relu, tile_hint, stream);
}


// Grouped variant: n_groups independent 3x3 convs (cin channels each, <= 16 outputs each) in one launch —
// the last conv of every CenterHead branch (center_head.py:34-35) on the tensor cores.  weight: bf16
// [n_groups*16][k_pad] (rows g*16+j = output j of group g, zero rows above its cout); scale/shift: f32
// [n_groups*16]; group_tab: device int32 [n_groups][2] = {first output column, cout}.
int pn_conv_dense3x3_grouped(const void* in, int in_ld, int in_coff, int cin, int n_groups, int n_frames, int H,
                             int W, const void* weight, int k_pad, const float* scale, const float* shift,
                             const int* group_tab, void* out, int out_dtype, int out_ld, int out_compact,
                             int relu, int in_planar, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in && weight && out && group_tab && n_groups >= 1 && n_frames >= 1 && H > 0 && W > 0);
  PN_REQUIRE(cin % BLOCK_K == 0 && in_ld % 8 == 0 && in_coff % 8 == 0 && k_pad % BLOCK_K == 0 && k_pad >= 9 * cin);
  PN_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15u) == 0);
  PN_REQUIRE(out_dtype == PN_F32 || out_dtype == PN_BF16);
  const int Hp = H + 2, Wp = W + 2;
  const long long n_pos = (long long)n_frames * Hp * Wp;
  PN_REQUIRE(n_pos < (1ll << 31));
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  constexpr int mt = 2, bn = 16;
  PN_REQUIRE(!in_planar || (in_ld == cin && in_coff == 0 && n_pos * n_groups < (1ll << 31)));
  const long long in_rows = in_planar ? n_pos * n_groups : n_pos;
  CUtensorMap ma, mtail, mw;
  int rc = get_map(in, in_rows, in_ld, in_ld, 128 * mt, &ma);
  if (rc != PN_OK) return rc;
  rc = get_map(in, in_rows, in_ld, in_ld, kTailRows, &mtail);
  if (rc != PN_OK) return rc;
  rc = get_map(weight, (long long)n_groups * bn, k_pad, k_pad, bn, &mw);
  if (rc != PN_OK) return rc;
  DArgs a;
  a.cin = cin; a.in_coff = in_coff; a.cout = bn; a.Hp = Hp; a.Wp = Wp; a.n_pos = (int)n_pos;
  a.scale = scale; a.shift = shift; a.out = out; a.out_f32 = out_dtype == PN_F32; a.out_ld = out_ld;
  a.out_coff = 0; a.out_compact = out_compact; a.relu = relu; a.base_offset_mode = 0;
  a.n_groups = n_groups; a.group_tab = group_tab; a.dbg = nullptr;
  a.out_group_cols = 0; a.in_planar = in_planar; a.tma_store = 0;
  const long long tiles = PN_DIVUP(n_pos, (long long)(128 * mt)) * n_groups;
  if (cin == BLOCK_K && n_groups <= sms) return launch<2, 16, 6, 9, 1, true>(ma, mtail, mw, mw, a, tiles, stream);   // weights stay resident
  return launch<2, 16, 6, 8, 1>(ma, mtail, mw, mw, a, tiles, stream);
}

}  // extern "C"
