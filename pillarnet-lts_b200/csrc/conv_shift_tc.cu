// Last conv of every CenterHead branch (3x3, 64 -> 1..3 channels, det3d/models/bbox_heads/center_head.py:34-35) as ONE
// 1x1 GEMM per branch plus nine shifted sums, on tcgen05.
//
//   out[q, c] = bias[c] + sum_t Y[q + off(t), 3t + c],   Y[q', 3t + c] = sum_k h[q', k] * W[c, k, t],   t = 3dy + dx
//
// on the zero-bordered padded layout of conv_dense_tc.cu (row q = (b*Hp + y)*Wp + x, off(t) = (dy-1)*Wp + (dx-1)): a
// tap is a constant row offset, so the nine taps of a Cout <= 3 conv are 27 output columns of one K = 64 GEMM whose
// rows are then added with nine shifts.
//
// Why (profiles/r2_ncu_full_convs_nusc18.json): the implicit-GEMM form (k_conv_dense<2,16,...> grouped) issues
// 36 N = 16 MMAs per (128 rows, branch), each bound by reading its 4 KB A tile from shared memory (~45-73 clk): 74 us
// for 2.7 GFLOP, as long as the 86 GFLOP conv in front of it.  Here a (128 rows, branch) tile is 4 MMAs of N = 32 and
// the kernel is left with streaming the planar 64-channel intermediate once (n_groups * n_pos * 128 B; 149 MB for the
// nuScenes head) — HBM bound.
//
// Work unit = (branch g, strip of consecutive 128-row tiles): the branch's 4 KB weight tile stays resident, tiles are
// walked in order, Y tiles land in a ring of eight in shared memory (column-major: lanes = consecutive rows, conflict
// free) and tile t is summed once t+2 has arrived (|off| <= Wp + 1 <= 256 rows).  A strip recomputes two halo tiles on
// each side (4 of ~65).
//
// Warp roles (512 threads, one persistent CTA per SM): 0 = TMA producer, 1 = TMEM owner + MMA issuer, 4-7 = drain
// (TMEM -> Y ring), 8-15 = shifted sums + bias -> compact f32 output rows (two groups of four, alternate tiles).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace {

using namespace pn_tc;

constexpr int TM = 128;            // rows (padded positions) per tile
constexpr int KC = 64;             // input channels (one 128-byte swizzle row)
constexpr int NCOL = 32;           // GEMM N: 27 used columns (tap-major, 3 output slots per tap)
constexpr int CP = 3;              // output slots per tap
constexpr int NY = 9 * CP;
constexpr int kStages = 6;         // input tiles in flight (16 KB each)
constexpr int kHalo = 2;           // halo tiles on each side of a strip: covers |off| <= 256 rows
constexpr int kYSlots = 2 * kHalo + 4;   // two sum groups at tiles t, t+1 hold t-2 .. t+3; two more for the drain to run ahead
constexpr int kThreads = 512;
constexpr int kDrainWarp0 = 4, kSumWarp0 = 8;

struct SArgs {
  int n_groups, n_pos, Hp, Wp, H, W;
  int tiles_total, strips, strip_tiles;
  const float* shift;      // [n_groups][4]
  const int* group_tab;    // [n_groups][2] = {first output column, cout}
  float* out;              // compact rows b*H*W + (y-1)*W + (x-1)
  int out_ld;
  unsigned long long* dbg; // PN_SHIFT_TIMELINE=1: per-CTA stall clocks [grid][16]
};

// stall accounting (debug launches only): clocks spent in a wait, summed per role by one lane
#define SW_T0() const long long _w0 = P.dbg ? clock64() : 0
#define SW_ACC(var) do { if (P.dbg) var += clock64() - _w0; } while (0)
#define SW_OUT(slot, var) do { if (P.dbg) P.dbg[blockIdx.x * 16 + slot] = (unsigned long long)(var); } while (0)

struct SSmem {
  alignas(1024) uint8_t h[kStages][TM * 128];
  alignas(1024) uint8_t w[2][NCOL * 128];
  alignas(16) float y[NY][kYSlots * TM];     // Y ring, column-major: tile with ring counter c at positions (c % kYSlots) * 128 + row
  alignas(8) uint64_t h_full[kStages];
  uint64_t h_empty[kStages], w_full[2], w_empty[2], tmem_full[2], tmem_empty[2], y_full[kYSlots], y_empty[kYSlots];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
k_conv_shift(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w, const SArgs P) {
  extern __shared__ uint8_t smem_raw[];
  SSmem& sm = *reinterpret_cast<SSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&sm.h_full[s], 1); mbar_init(&sm.h_empty[s], 1); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], 1);
        mbar_init(&sm.tmem_full[s], 1); mbar_init(&sm.tmem_empty[s], 128);
      }
      for (int s = 0; s < kYSlots; ++s) { mbar_init(&sm.y_full[s], 128); mbar_init(&sm.y_empty[s], 256); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<64>(&sm.tmem_base);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  pdl_wait();                        // the intermediate is the predecessor's output
  const int n_units = P.n_groups * P.strips;

  // every role walks the same units and tiles: unit u -> branch g = u / strips, tiles [t0, t1) plus the halo
  auto unit_range = [&](int u, int& g, int& t0, int& n) {
    g = u / P.strips;
    t0 = (u - g * P.strips) * P.strip_tiles;
    const int t1 = min(P.tiles_total, t0 + P.strip_tiles);
    n = t1 > t0 ? t1 - t0 + 2 * kHalo : 0;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, uc = 0;
      long long w_hempty = 0;
      const long long t_begin = P.dbg ? clock64() : 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        int g, t0, n;
        unit_range(u, g, t0, n);
        if (n == 0) continue;
        const uint32_t wb = uc & 1u;
        mbar_wait(&sm.w_empty[wb], ((uc >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sm.w_full[wb], NCOL * 128);
        tma_load_2d(smem_u32(sm.w[wb]), &tmap_w, 0, g * NCOL, &sm.w_full[wb]);
        ++uc;
        // rows before / after the branch's map belong to the neighbouring branch (or lie outside the matrix: zero
        // fill); the sums never reach them (an interior q + off stays inside its own frame)
        const int row_base = g * P.n_pos + (t0 - kHalo) * TM;
        for (int i = 0; i < n; ++i, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          { SW_T0(); mbar_wait(&sm.h_empty[s], ph ^ 1u); SW_ACC(w_hempty); }
          mbar_arrive_expect_tx(&sm.h_full[s], TM * 128);
          tma_load_2d(smem_u32(sm.h[s]), &tmap_h, 0, row_base + i * TM, &sm.h_full[s]);
        }
      }
      SW_OUT(0, w_hempty); SW_OUT(1, (P.dbg ? clock64() : 0) - t_begin); SW_OUT(2, it);
    }
  } else if (warp == 1) {
    const bool issuer = elect_one();
    // M = 128, N = 32, bf16 x bf16 -> f32, both operands K-major
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    uint32_t it = 0, uc = 0;
    long long w_tempty = 0, w_hfull = 0;
    const long long t_begin = P.dbg ? clock64() : 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      int g, t0, n;
      unit_range(u, g, t0, n);
      if (n == 0) continue;
      const uint32_t wb = uc & 1u;
      mbar_wait(&sm.w_full[wb], (uc >> 1) & 1u);
      const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(sm.w[wb]));
      for (int i = 0; i < n; ++i, ++it) {
        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
        const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
        { SW_T0(); mbar_wait(&sm.tmem_empty[acc], aph ^ 1u); SW_ACC(w_tempty); }
        { SW_T0(); mbar_wait(&sm.h_full[s], ph); SW_ACC(w_hfull); }
        tcgen05_fence_after();
        if (issuer) {
          const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(sm.h[s]));
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            umma_bf16(tmem_base + acc * NCOL, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0 ? 1u : 0u);
          umma_commit(&sm.h_empty[s]);
          umma_commit(&sm.tmem_full[acc]);
        }
        __syncwarp();
      }
      if (issuer) umma_commit(&sm.w_empty[wb]);
      __syncwarp();
      ++uc;
    }
    if (issuer) { SW_OUT(3, w_tempty); SW_OUT(4, w_hfull); SW_OUT(5, (P.dbg ? clock64() : 0) - t_begin); }
  } else if (warp >= kDrainWarp0 && warp < kDrainWarp0 + 4) {
    // ===================== drain: accumulator (row = TMEM lane) -> Y ring, column-major =====================
    const int quarter = warp & 3, r = quarter * 32 + lane;
    uint32_t it = 0;
    long long w_tfull = 0, w_yempty = 0;
    const long long t_begin = P.dbg ? clock64() : 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      int g, t0, n;
      unit_range(u, g, t0, n);
      for (int i = 0; i < n; ++i, ++it) {
        const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
        const uint32_t ys = it % kYSlots, yph = (it / kYSlots) & 1u;
        { SW_T0(); mbar_wait(&sm.tmem_full[acc], aph); SW_ACC(w_tfull); }
        tcgen05_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * NCOL, v);
        tmem_wait_ld();
        tcgen05_fence_before();
        mbar_arrive(&sm.tmem_empty[acc]);
        { SW_T0(); mbar_wait(&sm.y_empty[ys], yph ^ 1u); SW_ACC(w_yempty); }
#pragma unroll
        for (int j = 0; j < NY; ++j) sm.y[j][ys * TM + r] = __uint_as_float(v[j]);
        mbar_arrive(&sm.y_full[ys]);       // release: the sum warps' acquire wait orders these stores before their loads
      }
    }
    if (warp == kDrainWarp0 && lane == 0) { SW_OUT(6, w_tfull); SW_OUT(7, w_yempty); SW_OUT(8, (P.dbg ? clock64() : 0) - t_begin); }
  } else if (warp >= kSumWarp0) {
    // ===================== shifted sums: thread = row of the tile, two warp groups take alternate tiles ==========
    // (one group alone: ~660 clk of dependent index math + 27 LDS per tile on one warp per scheduler; measured with
    // PN_SHIFT_TIMELINE, it paced the whole ring.)  Both groups release every tile: group p is done with tile k once its
    // own sum of tile k+1 or k+2 has finished.
    const int grp = (warp - kSumWarp0) >> 2, r = ((warp - kSumWarp0) & 3) * 32 + lane;
    constexpr int kRing = kYSlots * TM;
    int off[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) off[t] = (t / 3 - 1) * P.Wp + (t % 3 - 1);
    uint32_t it0 = 0;                 // ring counter of the unit's tile 0
    long long w_yfull = 0;
    const long long t_begin = P.dbg ? clock64() : 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      int g, t0, n;
      unit_range(u, g, t0, n);
      if (n == 0) continue;
      const int col = __ldg(P.group_tab + 2 * g), cout = __ldg(P.group_tab + 2 * g + 1);
      float bias[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) bias[c] = (c < cout && P.shift) ? __ldg(P.shift + 4 * g + c) : 0.f;
      // (frame, y, x) of this thread's row in the group's first tile; advanced by two tiles per step
      int q = (t0 + grp) * TM + r;
      int b = q / (P.Hp * P.Wp), y = (q - b * P.Hp * P.Wp) / P.Wp, x = q - (b * P.Hp + y) * P.Wp;
      int rel = 0;                    // next tile (unit-relative) this group has to release
      for (int t = kHalo + grp; t < n - kHalo; t += 2) {
        // tile t + kHalo has arrived (and, in order, every tile before it): tile t has its whole neighbourhood
        const uint32_t cw = it0 + (uint32_t)(t + kHalo);
        { SW_T0(); mbar_wait(&sm.y_full[cw % kYSlots], (cw / kYSlots) & 1u); SW_ACC(w_yfull); }
        if (q < P.n_pos && y >= 1 && y <= P.H && x >= 1 && x <= P.W) {
          float a[CP];
#pragma unroll
          for (int c = 0; c < CP; ++c) a[c] = bias[c];
          const int pos0 = (int)((it0 + (uint32_t)t) % kYSlots) * TM + r;      // ring position of this row
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            int pos = pos0 + off[k];                                            // |off| <= 2 tiles < ring
            pos += pos < 0 ? kRing : 0;
            pos -= pos >= kRing ? kRing : 0;
            const float* yp = &sm.y[k * CP][pos];
#pragma unroll
            for (int c = 0; c < CP; ++c)
              if (c < cout) a[c] += yp[c * kRing];
          }
          float* op = P.out + ((long long)(b * P.H + (y - 1)) * P.W + (x - 1)) * P.out_ld + col;
#pragma unroll
          for (int c = 0; c < CP; ++c)
            if (c < cout) op[c] = a[c];
        }
        q += 2 * TM;
        x += 2 * TM;
        while (x >= P.Wp) { x -= P.Wp; ++y; }
        while (y >= P.Hp) { y -= P.Hp; ++b; }
        // this group's next sum is tile t + 2, which reads tiles >= t: everything below is released
        for (; rel < t; ++rel) mbar_arrive(&sm.y_empty[(it0 + (uint32_t)rel) % kYSlots]);
      }
      // unit end: the remaining tiles, once they have been written (an arrival must land in the phase of ITS use of the slot)
      {
        const uint32_t cw = it0 + (uint32_t)(n - 1);
        SW_T0(); mbar_wait(&sm.y_full[cw % kYSlots], (cw / kYSlots) & 1u); SW_ACC(w_yfull);
      }
      for (; rel < n; ++rel) mbar_arrive(&sm.y_empty[(it0 + (uint32_t)rel) % kYSlots]);
      it0 += (uint32_t)n;
    }
    if (warp == kSumWarp0 && lane == 0) { SW_OUT(9, w_yfull); SW_OUT(10, (P.dbg ? clock64() : 0) - t_begin); }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

}  // namespace

extern "C" {

// in: bf16 planar intermediate [(n_groups * n_pos), 64] (n_pos = n_frames*(H+2)*(W+2), borders zero: what
// pn_conv_dense3x3 writes with out_group_cols = 64); weight: bf16 [n_groups*32][64], row g*32 + 3*tap + c = W_g[c, :, tap]
// (rows 27..31 and slots c >= cout zero); shift: f32 [n_groups*4] biases; group_tab: int32 [n_groups][2] = {first output
// column, cout <= 3}; out: f32 compact rows (n_frames*H*W, out_ld).
int pn_conv_dense3x3_grouped_shift(const void* in, int n_groups, int n_frames, int H, int W, const void* weight,
                                   const float* shift, const int* group_tab, float* out, int out_ld,
                                   pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in && weight && out && group_tab && n_groups >= 1 && n_frames >= 1 && H > 0 && W > 0 && out_ld >= 1);
  PN_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15u) == 0);
  const int Hp = H + 2, Wp = W + 2;
  if (Wp + 1 > kHalo * TM) return PN_ERR_UNSUPPORTED;          // the shifts must stay inside the halo tiles
  const long long n_pos = (long long)n_frames * Hp * Wp;
  PN_REQUIRE(n_pos * n_groups + 4 * TM < (1ll << 31));
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  CUtensorMap mh, mw;
  int rc = pn_tmap::get(in, n_pos * n_groups, KC, KC, KC, TM, CU_TENSOR_MAP_SWIZZLE_128B, &mh);
  if (rc != PN_OK) return rc;
  rc = pn_tmap::get(weight, (long long)n_groups * NCOL, KC, KC, KC, NCOL, CU_TENSOR_MAP_SWIZZLE_128B, &mw);
  if (rc != PN_OK) return rc;
  SArgs a;
  a.n_groups = n_groups; a.n_pos = (int)n_pos; a.Hp = Hp; a.Wp = Wp; a.H = H; a.W = W;
  a.tiles_total = (int)PN_DIVUP(n_pos, (long long)TM);
  // strips per branch: fill the SMs, but keep a strip long against its 4 halo tiles
  int strips = sms / n_groups;
  if (strips < 1) strips = 1;
  const int max_strips = a.tiles_total / 16 > 1 ? a.tiles_total / 16 : 1;
  if (strips > max_strips) strips = max_strips;
  a.strip_tiles = PN_DIVUP(a.tiles_total, strips);
  a.strips = PN_DIVUP(a.tiles_total, a.strip_tiles);
  a.shift = shift; a.group_tab = group_tab; a.out = out; a.out_ld = out_ld;
  a.dbg = nullptr;
  static const bool timeline = [] { const char* e = getenv("PN_SHIFT_TIMELINE"); return e && e[0] == '1'; }();
  static unsigned long long* dbg_buf = nullptr;
  if (timeline) {
    if (!dbg_buf) PN_CUDA(cudaMalloc(&dbg_buf, 16 * 1024 * sizeof(unsigned long long)));
    PN_CUDA(cudaMemsetAsync(dbg_buf, 0, 16 * 1024 * sizeof(unsigned long long), stream));
    a.dbg = dbg_buf;
  }
  const int units = n_groups * a.strips;
  constexpr size_t smem = sizeof(SSmem) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static pn_detail::PerDeviceOnce once;
  if (once.need())
    PN_CUDA(cudaFuncSetAttribute(k_conv_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units < sms ? units : sms);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  PN_CUDA(cudaLaunchKernelEx(&cfg, k_conv_shift, mh, mw, a));
  PN_CHECK_LAUNCH();
  if (timeline) {
    PN_CUDA(cudaStreamSynchronize(stream));
    static unsigned long long t[16 * 1024];
    PN_CUDA(cudaMemcpy(t, dbg_buf, sizeof(t), cudaMemcpyDeviceToHost));
    const int n = (int)cfg.gridDim.x < 1024 ? (int)cfg.gridDim.x : 1024;
    double s[11] = {0};
    for (int c = 0; c < n; ++c)
      for (int k = 0; k < 11; ++k) s[k] += (double)t[c * 16 + k] / n;
    fprintf(stderr, "[conv_shift groups %d strips %d x %d tiles, grid %d] per CTA (kclk): tiles %.1f | producer total %.1f, "
                    "h_empty %.1f | mma total %.1f, tmem_empty %.1f, h_full %.1f | drain total %.1f, tmem_full %.1f, y_empty "
                    "%.1f | sum total %.1f, y_full %.1f\n",
            n_groups, a.strips, a.strip_tiles, n, s[2], s[1] / 1e3, s[0] / 1e3, s[5] / 1e3, s[3] / 1e3, s[4] / 1e3, s[8] / 1e3,
            s[6] / 1e3, s[7] / 1e3, s[10] / 1e3, s[9] / 1e3);
  }
  return PN_OK;
}

}  // extern "C"
