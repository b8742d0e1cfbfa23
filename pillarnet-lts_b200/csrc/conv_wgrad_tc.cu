// Weight gradient of the gather-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[co][t*cin + ci] = sum_o dy[o][co] * x[nbr[o,t]][ci]
//
// This is the backward-weight half of spconv's SubMConv2d / SparseConv2d autograd (external in the
// reference; used by det3d/models/backbones/base.py:38-63, PillarResNet.py:87,95,103 in training,
// SURVEY §8 a25).  The reduction runs over the *rows* (active sites), so both UMMA operands are
// MN-major: a tile of rows x 64 channels, stored exactly as the forward kernel stores its K-major
// activation tile (row r at r*128 B, 16-byte chunks XOR-swizzled with r & 7), is read by the tensor
// core as a 64(MN) x rows(K) operand:  canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
// units with SBO = 1024 B between 8-row groups and LBO = the distance between 64-channel atoms.
//
//   A = dy tile   : M = 128 couts (2 atoms) x K = 64 rows per stage   (cp.async, rows contiguous)
//   B = x gather  : N = BN cins (BN/64 atoms) x K = 64 rows per stage (cp.async gather, zero fill)
//   D = 128 x BN fp32 in TMEM, accumulated over the CTA's whole row range, then reduced into dW with
//       vector fp32 RED (split-K across CTAs).
// One CTA = (row split, group of TPC taps, 128-cout tile, BN-cin tile).  TPC = 3 (a kernel row) for 3x3 layers up to 128
// input channels: the dy tile of a stage is loaded ONCE and multiplied with the three taps' gathered x tiles into three
// TMEM accumulators — one third of the dy traffic and 45 % fewer cp.async per tap than one tap per CTA (measured on the
// 64-channel stage of PillarNet-34 training: 167 us per launch, L2 / LSU bound, ten times its FLOP time).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace pn_tc;

constexpr int BK_ROWS = 64;                 // rows (= MMA K) per pipeline stage
constexpr int ATOM_BYTES = BK_ROWS * 128;   // one 64-channel x 64-row atom
constexpr int kProducerThreads = 512;
constexpr int kProducerWarps = kProducerThreads / 32;
constexpr int kEpilogueWarp0 = kProducerWarps;
constexpr int kMmaWarp = kProducerWarps + 4;
constexpr int kThreads = (kMmaWarp + 1) * 32;

struct WArgs {
  const __nv_bfloat16* x;
  int x_ld;
  const __nv_bfloat16* dy;
  int dy_ld;
  const int* nbr;
  int taps;
  const int* num_rows;
  int rows_cap;
  int cin, cout;
  int n_ci_tiles, n_co_tiles, splits;
  float* dw;
  int dw_ld;
};

// MN-major SWIZZLE_128B descriptor: LBO = bytes between 64-element atoms along M/N, SBO = 1024 B between
// 8-row (K) groups, version 1, layout type 2.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16: D f32, A/B bf16, both MN-major (bits 15, 16), M = 128, N = BN.
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc_mn() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

template <int BN, int STAGES, int TPC>
struct WSmem {
  alignas(1024) uint8_t a[STAGES][2 * ATOM_BYTES];
  alignas(1024) uint8_t b[STAGES][TPC * (BN / 64) * ATOM_BYTES];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int wgrad_tmem_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int BN, int STAGES, int TPC>
__global__ void __launch_bounds__(kThreads, 1)
k_wgrad_tc(const WArgs P) {
  extern __shared__ uint8_t smem_raw[];
  using S = WSmem<BN, STAGES, TPC>;
  S& sm = *reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = P.num_rows ? min(*P.num_rows, P.rows_cap) : P.rows_cap;
  int w = blockIdx.x;
  const int split = w % P.splits; w /= P.splits;
  const int ci0 = (w % P.n_ci_tiles) * BN; w /= P.n_ci_tiles;
  const int co0 = (w % P.n_co_tiles) * 128; w /= P.n_co_tiles;
  const int t0 = w * TPC;                               // first tap of this CTA's group
  int rps = (rows + P.splits - 1) / P.splits;
  rps = (rps + BK_ROWS - 1) / BK_ROWS * BK_ROWS;
  const int r_begin = split * rps;
  const int r_end = min(rows, r_begin + rps);
  if (r_begin >= r_end) return;                       // whole CTA, before any barrier / TMEM use
  const int n_chunks = (r_end - r_begin + BK_ROWS - 1) / BK_ROWS;
  constexpr int TCOLS = wgrad_tmem_cols(TPC * BN);
  static_assert(TPC * BN <= 512, "TMEM budget");

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&sm.full[s], kProducerThreads);
        mbar_init(&sm.empty[s], 1);
      }
      mbar_init(&sm.tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TCOLS>(&sm.tmem_base);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp < kProducerWarps) {
    // ===================== producers: dy rows + gathered x rows =====================
    const int tid = threadIdx.x;
    const int piece = tid & 7, m = tid >> 3;            // row m of the chunk, 16-byte piece of a 128-byte row
    const uint32_t dst_off = (uint32_t)m * 128u + (uint32_t)((piece ^ (m & 7)) << 4);
    const char* x_bytes = reinterpret_cast<const char*>(P.x);
    const char* dy_bytes = reinterpret_cast<const char*>(P.dy);
    auto fetch = [&](int chunk, int (&src)[TPC]) {
      const int o = r_begin + chunk * BK_ROWS + m;
#pragma unroll
      for (int j = 0; j < TPC; ++j) {
        src[j] = -1;
        if (chunk < n_chunks && o < r_end) src[j] = P.nbr ? __ldg(P.nbr + (long long)o * P.taps + t0 + j) : o;
      }
    };
    int src_next[TPC];
    fetch(0, src_next);
    for (int kc = 0; kc < n_chunks; ++kc) {
      const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1u;
      int src[TPC];
      bool any = false;
#pragma unroll
      for (int j = 0; j < TPC; ++j) { src[j] = src_next[j]; any = any || src[j] >= 0; }
      fetch(kc + 1, src_next);
      const int o = r_begin + kc * BK_ROWS + m;
      mbar_wait(&sm.empty[s], ph ^ 1u);
      // A: two 64-cout atoms of row o of dy, once for all taps of the group (zero when none of its pairs exists or the
      // row lies past the range: the x tiles are zero there too)
      const uint32_t a_dst = smem_u32(sm.a[s]) + dst_off;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = co0 + q * 64 + piece * 8;
        const bool ok = any && c < P.cout;
        const char* g = dy_bytes + (ok ? ((size_t)o * P.dy_ld + c) * 2u : 0u);
        cp_async16(a_dst + q * ATOM_BYTES, g, ok ? 16u : 0u);
      }
#pragma unroll
      for (int j = 0; j < TPC; ++j) {
        const uint32_t b_dst = smem_u32(sm.b[s]) + (uint32_t)(j * (BN / 64) * ATOM_BYTES) + dst_off;
#pragma unroll
        for (int q = 0; q < BN / 64; ++q) {
          const int c = ci0 + q * 64 + piece * 8;
          const bool ok = src[j] >= 0 && c < P.cin;
          const char* g = x_bytes + (ok ? ((size_t)src[j] * P.x_ld + c) * 2u : 0u);
          cp_async16(b_dst + q * ATOM_BYTES, g, ok ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive_noinc(&sm.full[s]);
    }
    cp_async_wait_all();
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // whole warp converged, one elected lane issues (see conv_tcgen05.cu)
    {
      const bool issuer = elect_one();
      constexpr uint32_t idesc = make_idesc_mn<BN>();
      const uint64_t a_desc0 = make_mnmajor_sw128_desc(smem_u32(sm.a[0]), ATOM_BYTES);
      const uint64_t b_desc0 = make_mnmajor_sw128_desc(smem_u32(sm.b[0]), ATOM_BYTES);
      constexpr uint32_t kAStep = (uint32_t)(2 * ATOM_BYTES) >> 4, kBTap = (uint32_t)((BN / 64) * ATOM_BYTES) >> 4;
      constexpr uint32_t kBStep = (uint32_t)TPC * kBTap;
      uint32_t s = 0, ph = 0;
      for (int kc = 0; kc < n_chunks; ++kc) {
        mbar_wait(&sm.full[s], ph);
        fence_proxy_async_smem();   // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
        tcgen05_fence_after();
        const uint64_t a_desc = a_desc0 + (uint64_t)(s * kAStep), b_desc = b_desc0 + (uint64_t)(s * kBStep);
        if (issuer) {
#pragma unroll
          for (int j = 0; j < TPC; ++j) {
#pragma unroll
            for (int k = 0; k < BK_ROWS / 16; ++k) {
              // 16 rows (K) further = 2048 bytes = +128 in the (>>4) start-address field
              umma_bf16(tmem_base + (uint32_t)(j * BN), a_desc + 128 * k, b_desc + (uint64_t)(j * kBTap) + 128 * k, idesc,
                        (kc | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&sm.empty[s]);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (issuer) umma_commit(&sm.tmem_full);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: split-K reduction into dW =====================
    const int e = warp - kEpilogueWarp0;
    mbar_wait_relaxed(&sm.tmem_full, 0);
    tcgen05_fence_after();
    const int co = co0 + e * 32 + lane;
#pragma unroll 1
    for (int jc = 0; jc < TPC * (BN / 32); ++jc) {
      const int j = jc / (BN / 32), c0 = (jc - j * (BN / 32)) * 32;
      if (ci0 + c0 >= P.cin) continue;   // warp-uniform
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(j * BN + c0), v);
      tmem_wait_ld();
      if (co < P.cout) {
        float* dst = P.dw + (size_t)co * P.dw_ld + (size_t)(t0 + j) * P.cin + ci0 + c0;
        const int nvalid = min(32, P.cin - (ci0 + c0));
        if (nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

template <int BN, int STAGES, int TPC>
int launch(const WArgs& a, int grid, cudaStream_t stream) {
  constexpr size_t smem = sizeof(WSmem<BN, STAGES, TPC>) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static pn_detail::PerDeviceOnce once;
  if (once.need())
    PN_CUDA(cudaFuncSetAttribute(k_wgrad_tc<BN, STAGES, TPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_wgrad_tc<BN, STAGES, TPC><<<grid, kThreads, smem, stream>>>(a);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // namespace

namespace pn_detail {

int conv_wgrad_tcgen05(const void* x, int x_ld, const void* dy, int dy_ld, const int* nbr, int taps,
                       const int* num_rows, int rows_cap, int cin, int cout, float* dw, int dw_ld,
                       cudaStream_t stream) {
  if (cin % 8 != 0 || cout % 8 != 0 || x_ld % 8 != 0 || dy_ld % 8 != 0) return PN_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0 || (reinterpret_cast<uintptr_t>(dy) & 15u) != 0)
    return PN_ERR_UNSUPPORTED;
  const int bn = cin <= 64 ? 64 : cin <= 128 ? 128 : 256;
  WArgs a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x);
  a.x_ld = x_ld;
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  a.dy_ld = dy_ld;
  a.nbr = nbr;
  a.taps = taps;
  a.num_rows = num_rows;
  a.rows_cap = rows_cap;
  a.cin = cin;
  a.cout = cout;
  a.n_ci_tiles = PN_DIVUP(cin, bn);
  a.n_co_tiles = PN_DIVUP(cout, 128);
  a.dw = dw;
  a.dw_ld = dw_ld;
  const int sms = sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  static const bool group_taps = [] { const char* e = getenv("PN_WGRAD_TAP_GROUP"); return !(e && e[0] == '0'); }();
  const int tpc = (group_taps && taps % 3 == 0 && bn <= 128) ? 3 : 1;
  const int tiles = (taps / tpc) * a.n_ci_tiles * a.n_co_tiles;
  // split the rows so that ~2 CTAs per SM exist, but keep >= 4 stages' worth of rows per CTA
  int splits = PN_DIVUP(2 * sms, tiles);
  const int max_splits = PN_DIVUP(rows_cap, 4 * BK_ROWS);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.splits = splits;
  const int grid = tiles * splits;
  if (tpc == 3) return bn == 64 ? launch<64, 5, 3>(a, grid, stream) : launch<128, 3, 3>(a, grid, stream);
  switch (bn) {
    case 64: return launch<64, 6, 1>(a, grid, stream);
    case 128: return launch<128, 5, 1>(a, grid, stream);
    default: return launch<256, 4, 1>(a, grid, stream);
  }
}

}  // namespace pn_detail
