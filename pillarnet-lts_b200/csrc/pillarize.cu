// Dynamic pillarization + fused PFN/scatter-max for sm_100a.
//
// Reference path being replaced (PillarNet-LTS):
//   det3d/models/readers/dynamic_pillar_encoder.py:29-50   cell coords, range mask, compaction
//   det3d/ops/pillar_ops/pillar_utils.py:22-57              pillar set / order / point->pillar / offsets
//   det3d/ops/pillar_ops/src/pillar_ops_gpu.cu:13-78        index + indices kernels
//   det3d/ops/pillar_ops/src/group_ops_gpu.cu:8-33          gathers
//   det3d/ops/pillar_ops/pillar_modules.py:26-33,71-72      Linear+BN1d+ReLU, scatter_max
//   det3d/ops/pillar_ops/src/scatter_ops_gpu.cu:13-45       scatter max / argmax / grad
//
// Design: occupancy is a 1-bit/cell mask (L2 resident: 259 KB per 1440^2 frame) ranked by a
// popcount scan, so nothing B*H*W-sized and wider than a bit is ever touched; points are staged
// through shared memory with float4 loads; the PFN never materialises the (L,7)/(L,32) matrices.
#include "common.cuh"
#include "mask_scan.cuh"

namespace {

constexpr int kPtThreads = 256;
constexpr int kMaxPointDim = 8;

// Coalesced stage of `count` consecutive points (point_dim floats each) starting at point `first`
// into shared memory.  The block's first float index is a multiple of 4 when kPtThreads*dim is,
// which holds for kPtThreads = 256, so the bulk goes as float4.
__device__ __forceinline__ void stage_points(const float* __restrict__ pts, long long first_float,
                                             int n_floats, float* __restrict__ smem) {
  const float* src = pts + first_float;
  const int n4 = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) ? (n_floats >> 2) : 0;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  float4* dst4 = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) dst4[i] = __ldg(src4 + i);
  for (int i = (n4 << 2) + threadIdx.x; i < n_floats; i += blockDim.x) smem[i] = __ldg(src + i);
}

__device__ __forceinline__ int frame_of(const int* __restrict__ off, int n_frames, int p) {
  // frame b owns [off[b], off[b+1]); offsets are non-decreasing, empty frames allowed.
  int lo = 0, hi = n_frames;  // invariant: off[lo] <= p < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// cell coordinate exactly as the reference's CUDA expression
//   ((points[:,0] - pc_range[0]) / pillar_size).floor().int()
// evaluates: fp32 subtract, multiply by the fp32 reciprocal (ATen scalar-division fast path), floor,
// saturating convert.  __fsub_rn/__fmul_rn forbid FMA contraction.
__device__ __forceinline__ int cell_coord(float v, float v0, float inv) {
  return (int)floorf(__fmul_rn(__fsub_rn(v, v0), inv));
}

constexpr int kMarkPerThread = 4;   // points per thread (strided by the block size: coalesced)

// One pass over the points: cell id per point + occupancy bits.  Instruction-lean on purpose (the first
// version staged points through shared memory and binary-searched the frame per point: 245 instructions per
// point, issue bound at 1.7 TB/s): x/y are read straight from global (the 20-byte records of a warp are one
// contiguous 640-byte run, so both loads hit the same L1 lines), the frame is resolved once per block when the
// block does not straddle a frame boundary, and each thread handles kMarkPerThread points.
__global__ void __launch_bounds__(kPtThreads)
k_mark(const float* __restrict__ pts, int dim, const int* __restrict__ frame_off, int n_points,
       int n_frames, int H, int W, float x0, float y0, float inv, uint32_t* __restrict__ words,
       int* __restrict__ point_cell, int* __restrict__ scan_state, int n_state) {
  __shared__ int s_frame[2];
  pn_detail::zero_scan_state(scan_state, n_state);   // for the single-pass scan launched next (mask_scan.cu)
  n_points = min(n_points, __ldg(frame_off + n_frames));  // live count on the device, capacity on the host
  const int first = blockIdx.x * (kPtThreads * kMarkPerThread);
  if (first >= n_points) return;
  const int last = min(first + kPtThreads * kMarkPerThread, n_points) - 1;
  if (threadIdx.x < 2) s_frame[threadIdx.x] = frame_of(frame_off, n_frames, threadIdx.x == 0 ? first : last);
  __syncthreads();
  const int b_first = s_frame[0];
  const bool uniform = b_first == s_frame[1];
  const int lane = threadIdx.x & 31;
  // Three phases, each with all of the thread's memory operations in flight together: at one frame (a few blocks
  // per SM, cold L2) the kernel is a chain of dependent round trips, and the first version walked its points one
  // after the other — load, mark, load, mark: eight round trips instead of three (ncu: 2/3 of the stall samples
  // on the two dependent loads per point; 13.8 us for 5 MB).
  float x[kMarkPerThread], y[kMarkPerThread];
#pragma unroll
  for (int j = 0; j < kMarkPerThread; ++j) {
    const int p = first + j * kPtThreads + threadIdx.x;
    x[j] = y[j] = 0.f;
    if (p <= last) {
      const float* q = pts + (long long)p * dim;
      x[j] = __ldg(q);
      y[j] = __ldg(q + 1);
    }
  }
  // warp-aggregated marking: lanes that hit the same 32-cell word (scan-order neighbours usually do) combine their
  // bits and ONE lane issues the atomicOr — and only if the word still lacks a bit (plain L2 read first; a stale
  // read merely costs a redundant atomic).
  int lead_word[kMarkPerThread];
  unsigned lead_bits[kMarkPerThread];
#pragma unroll
  for (int j = 0; j < kMarkPerThread; ++j) {
    const int p = first + j * kPtThreads + threadIdx.x;
    const bool live = p <= last;     // warp-uniform except in the last warp of the grid
    int cell = -1;
    if (live) {
      const int cx = cell_coord(x[j], x0, inv);
      const int cy = cell_coord(y[j], y0, inv);
      if (cx >= 0 && cx < W && cy >= 0 && cy < H) {
        const int b = uniform ? b_first : frame_of(frame_off, n_frames, p);
        cell = (b * H + cy) * W + cx;
      }
      point_cell[p] = cell;
    }
    const int word = cell >= 0 ? (cell >> 5) : -1;
    const unsigned peers = __match_any_sync(__activemask(), word);
    lead_word[j] = -1;
    lead_bits[j] = 0u;
    if (word >= 0) {
      const unsigned bits = __reduce_or_sync(peers, 1u << (cell & 31));
      if (lane == __ffs(peers) - 1) { lead_word[j] = word; lead_bits[j] = bits; }
    }
  }
  unsigned seen[kMarkPerThread];
#pragma unroll
  for (int j = 0; j < kMarkPerThread; ++j) seen[j] = lead_word[j] >= 0 ? __ldcg(words + lead_word[j]) : 0xffffffffu;
#pragma unroll
  for (int j = 0; j < kMarkPerThread; ++j)
    if (lead_word[j] >= 0 && (seen[j] & lead_bits[j]) != lead_bits[j]) atomicOr(words + lead_word[j], lead_bits[j]);
}

__global__ void __launch_bounds__(kPtThreads)
k_rank(const uint32_t* __restrict__ words, const int* __restrict__ prefix, int n_points,
       const int* __restrict__ n_live, int* __restrict__ point_pillar) {
  const int p = blockIdx.x * kPtThreads + threadIdx.x;
  if (p >= min(n_points, __ldg(n_live))) return;
  const int cell = point_pillar[p];
  if (cell >= 0) point_pillar[p] = pn_rank_of(words, prefix, cell);
}

template <int C>
__global__ void __launch_bounds__(256)
k_zero_rows(float* __restrict__ out, int* __restrict__ arg, const int* __restrict__ num_rows,
            int m_cap) {
  const int n = min(*num_rows, m_cap);
  const long long total4 = (long long)n * C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    reinterpret_cast<float4*>(out)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arg) reinterpret_cast<int4*>(arg)[i] = make_int4(-1, -1, -1, -1);
  }
}

// PFN + scatter-max.  Phase 1: thread = point computes all C channels (7 FMA per channel, weights read
// as constant-bank operands of the FFMA, so no load instructions) and parks them in shared memory.
// Phase 2: each warp walks its 32 points with lane = channel, merges runs of consecutive points of
// the same pillar in registers (scan-order locality) and issues ONE coalesced 128-byte RED.MAX per run
// and channel group.  The earlier lane=channel-only version spent ~40 warp instructions per point (all
// lanes redundantly decoding the same point) and was issue bound at 0.4-0.8 TB/s.
// Post-ReLU values are >= 0, so max on the raw int bits is exact and order-independent: the result is
// deterministic, unlike the reference's CAS loop (atomics.cuh:74-86).
template <int C>
struct PfnParams {
  float w[C][kMaxPointDim + 2];
  float scale[C];
  float shift[C];
};

template <int C>
__global__ void __launch_bounds__(kPtThreads)
k_pfn_scatter_max(const __grid_constant__ PfnParams<C> W, const float* __restrict__ pts, int dim, int n_points,
                  const int* __restrict__ n_live, const int* __restrict__ point_pillar, float x0, float y0,
                  float inv, float ps, float xoff, float yoff, float* __restrict__ out, int m_cap) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_pts = s_dyn;                                   // kPtThreads * dim
  float* s_h = s_dyn + ((kPtThreads * dim + 3) & ~3);     // kPtThreads * (C + 1)
  __shared__ int s_rank[kPtThreads];
  constexpr int R = C / 32;
  if (n_live) n_points = min(n_points, __ldg(n_live));
  const int first = blockIdx.x * kPtThreads;
  const int count = min(kPtThreads, n_points - first);
  if (count <= 0) return;
  stage_points(pts, (long long)first * dim, count * dim, s_pts);
  const int t = threadIdx.x;
  int rank = -1;
  if (t < count) {
    rank = __ldg(point_pillar + first + t);
    if (rank >= m_cap) rank = -1;
  }
  s_rank[t] = rank;
  __syncthreads();
  if (rank >= 0) {
    const float* p = s_pts + t * dim;
    float f[kMaxPointDim + 2];
    const float x = p[0], y = p[1];
    // pillar centre: int->float, *pillar_size, +offset as three separately rounded steps
    // (pillar_utils.py:51-52), then the offset features (pillar_utils.py:54).
    f[0] = __fsub_rn(x, __fadd_rn(__fmul_rn((float)cell_coord(x, x0, inv), ps), xoff));
    f[1] = __fsub_rn(y, __fadd_rn(__fmul_rn((float)cell_coord(y, y0, inv), ps), yoff));
#pragma unroll
    for (int k = 0; k < kMaxPointDim; ++k) f[k + 2] = k < dim ? p[k] : 0.f;
    float* hrow = s_h + t * (C + 1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float z = W.w[c][0] * f[0];
#pragma unroll
      for (int k = 1; k < kMaxPointDim + 2; ++k) z = fmaf(W.w[c][k], f[k], z);   // zero weights beyond dim+2
      hrow[c] = fmaxf(fmaf(z, W.scale[c], W.shift[c]), 0.f);
    }
  }
  __syncthreads();
  const int lane = t & 31, warp = t >> 5;
  const int j0 = warp * 32, j1 = min(j0 + 32, count);
  int cur = -1;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  for (int j = j0; j <= j1; ++j) {
    const int rk = j < j1 ? s_rank[j] : -2;   // sentinel flushes the last run
    if (rk == -1) continue;
    if (rk != cur) {
      if (cur >= 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (acc[r] > 0.f)
            atomicMax(reinterpret_cast<int*>(out) + (long long)cur * C + lane + 32 * r, __float_as_int(acc[r]));
          acc[r] = 0.f;
        }
      }
      cur = rk;
    }
    if (rk >= 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaxf(acc[r], s_h[j * (C + 1) + lane + 32 * r]);
    }
  }
}

// bf16-output variant (the tensor-core backbone consumes bf16 rows): max commutes with the monotone
// fp32->bf16 rounding, so scattering the rounded values gives bit-identical results to rounding the fp32
// max, and sm_100a has REDG.MAX.BF16x8: one lane-atomic covers 8 channels (16 bytes).  A pillar row of 32
// channels is 4 lane-atomics instead of 32, which takes the kernel off the per-SM RED issue limit
// (measured: ~1 lane-atomic per cycle per SM bounded the f32 version at 0.75 TB/s).
__device__ __forceinline__ void red_max_bf16x8(void* addr, uint4 v) {
  asm volatile("red.global.v4.bf16x2.max.noftz [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

template <int C>
__global__ void __launch_bounds__(kPtThreads)
k_pfn_scatter_max_bf16(const __grid_constant__ PfnParams<C> W, const float* __restrict__ pts, int dim, int n_points,
                       const int* __restrict__ n_live, const int* __restrict__ point_pillar, float x0, float y0,
                       float inv, float ps, float xoff, float yoff, __nv_bfloat16* __restrict__ out, int m_cap) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_pts = s_dyn;                                                       // kPtThreads * dim
  constexpr int RW = C / 2 + 4;                                               // u32 per point row (+pad: conflict-free v4)
  uint32_t* s_h = reinterpret_cast<uint32_t*>(s_dyn + ((kPtThreads * dim + 3) & ~3));  // kPtThreads * RW
  __shared__ int s_rank[kPtThreads];
  if (n_live) n_points = min(n_points, __ldg(n_live));
  const int first = blockIdx.x * kPtThreads;
  const int count = min(kPtThreads, n_points - first);
  if (count <= 0) return;
  stage_points(pts, (long long)first * dim, count * dim, s_pts);
  const int t = threadIdx.x;
  int rank = -1;
  if (t < count) {
    rank = __ldg(point_pillar + first + t);
    if (rank >= m_cap) rank = -1;
  }
  s_rank[t] = rank;
  __syncthreads();
  if (rank >= 0) {
    const float* p = s_pts + t * dim;
    float f[kMaxPointDim + 2];
    const float x = p[0], y = p[1];
    f[0] = __fsub_rn(x, __fadd_rn(__fmul_rn((float)cell_coord(x, x0, inv), ps), xoff));
    f[1] = __fsub_rn(y, __fadd_rn(__fmul_rn((float)cell_coord(y, y0, inv), ps), yoff));
#pragma unroll
    for (int k = 0; k < kMaxPointDim; ++k) f[k + 2] = k < dim ? p[k] : 0.f;
    uint32_t* hrow = s_h + t * RW;
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      float h[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float z = W.w[c + e][0] * f[0];
#pragma unroll
        for (int k = 1; k < kMaxPointDim + 2; ++k) z = fmaf(W.w[c + e][k], f[k], z);
        h[e] = fmaxf(fmaf(z, W.scale[c + e], W.shift[c + e]), 0.f);
      }
      const __nv_bfloat162 pk = __floats2bfloat162_rn(h[0], h[1]);
      hrow[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
    }
  }
  __syncthreads();
  // lane = (point j8, quad q): one RED instruction flushes 8 points x 4 quads x 8 channels
  constexpr int QUADS = C / 8;               // lane-atomics per point
  constexpr int PPI = 32 / QUADS;            // points per RED instruction
  const int lane = t & 31, warp = t >> 5;
  const int q = lane % QUADS, jl = lane / QUADS;
  for (int j = warp * 32 + jl; j < min(warp * 32 + 32, count); j += PPI) {
    const int rk = s_rank[j];
    if (rk < 0) continue;
    // (folding runs of same-pillar points before the atomic was measured slower: the serial follower
    //  loop costs more than the 16-byte vector atomics it saves)
    const uint4 v = *reinterpret_cast<const uint4*>(s_h + j * RW + q * 4);
    if ((v.x | v.y | v.z | v.w) != 0u) red_max_bf16x8(out + (long long)rk * C + q * 8, v);
  }
}

template <int C>
__global__ void __launch_bounds__(256)
k_zero_rows_bf16(__nv_bfloat16* __restrict__ out, const int* __restrict__ num_rows, int m_cap) {
  const int n = min(*num_rows, m_cap);
  const long long total8 = (long long)n * C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(out)[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Training only: deterministic argmax = lowest point id attaining the max (second pass, like the
// reference's scatter_arg_max_kernel but exact-compare and tie-broken).
template <int C>
__global__ void __launch_bounds__(kPtThreads)
k_pfn_argmax(const __grid_constant__ PfnParams<C> W, const float* __restrict__ pts, int dim, int n_points,
             const int* __restrict__ n_live, const int* __restrict__ point_pillar, float x0, float y0, float inv,
             float ps, float xoff, float yoff, const float* __restrict__ out, int* __restrict__ arg, int m_cap) {
  const int p = blockIdx.x * (kPtThreads / 32) + (threadIdx.x >> 5);
  if (p >= n_points || (n_live && p >= __ldg(n_live))) return;
  const int rank = __ldg(point_pillar + p);
  if (rank < 0 || rank >= m_cap) return;
  const int lane = threadIdx.x & 31;
  const float* q = pts + (long long)p * dim;
  const float x = __ldg(q), y = __ldg(q + 1);
  const float f0 = __fsub_rn(x, __fadd_rn(__fmul_rn((float)cell_coord(x, x0, inv), ps), xoff));
  const float f1 = __fsub_rn(y, __fadd_rn(__fmul_rn((float)cell_coord(y, y0, inv), ps), yoff));
  float f[kMaxPointDim + 2];
  f[0] = f0;
  f[1] = f1;
#pragma unroll
  for (int k = 0; k < kMaxPointDim; ++k) f[k + 2] = k < dim ? __ldg(q + k) : 0.f;
  for (int c = lane; c < C; c += 32) {
    // same operation order as k_pfn_scatter_max, so the recomputed value is bit-identical to the stored max
    float z = W.w[c][0] * f[0];
#pragma unroll
    for (int k = 1; k < kMaxPointDim + 2; ++k) z = fmaf(W.w[c][k], f[k], z);
    const float h = fmaxf(fmaf(z, W.scale[c], W.shift[c]), 0.f);
    if (h == out[(long long)rank * C + c]) {
      // arg holds -1 (0xFFFFFFFF) initially: unsigned min keeps the lowest flat index
      atomicMin(reinterpret_cast<unsigned*>(arg) + (long long)rank * C + c, (unsigned)(p * C + c));
    }
  }
}

template <int C>
__global__ void __launch_bounds__(256)
k_bf16_rows(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
            const int* __restrict__ num_rows, int m_cap) {
  const int n = min(*num_rows, m_cap);
  const long long total2 = (long long)n * C / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total2;
       i += (long long)gridDim.x * blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(in)[i];
    reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(v.x, v.y);
  }
}

__global__ void __launch_bounds__(256)
k_scatter_max_grad(const float* __restrict__ grad_out, const int* __restrict__ arg,
                   const int* __restrict__ num_rows, int m_cap, int C,
                   float* __restrict__ grad_src) {
  const int n = num_rows ? min(*num_rows, m_cap) : m_cap;
  const long long total = (long long)n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int a = arg[i];
    if (a >= 0) grad_src[a] = grad_out[i];
  }
}

}  // namespace

extern "C" {

size_t pn_pillarize_scratch_bytes(int n_frames, int H, int W) {
  return pn_detail::scan_scratch_bytes(pn_detail::n_words((long long)n_frames * H * W));
}

int pn_pillarize(const float* points, int point_dim, const int* frame_offsets, int n_points,
                 int n_frames, int H, int W, float x0, float y0, float inv_pillar,
                 uint32_t* occ_words, int* word_prefix, int* pillar_coords, int m_cap,
                 int* point_pillar, int* num_pillars, void* scratch, size_t scratch_bytes,
                 pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(point_dim >= 2 && point_dim <= kMaxPointDim);
  PN_REQUIRE(n_frames >= 1 && H > 0 && W > 0 && n_points >= 0 && m_cap >= 0);
  PN_REQUIRE((long long)n_frames * H * W < (1ll << 31));
  PN_REQUIRE(occ_words && word_prefix && num_pillars && scratch);
  PN_REQUIRE(n_points == 0 || (points && frame_offsets && point_pillar));
  const long long nw = pn_detail::n_words((long long)n_frames * H * W);
  if (scratch_bytes < pn_detail::scan_scratch_bytes(nw)) return PN_ERR_WORKSPACE;
  PN_CUDA(cudaMemsetAsync(occ_words, 0, nw * sizeof(uint32_t), stream));
  const int blocks = PN_DIVUP(n_points, kPtThreads);
  if (n_points > 0) {
    k_mark<<<PN_DIVUP(n_points, kPtThreads * kMarkPerThread), kPtThreads, 0, stream>>>(
        points, point_dim, frame_offsets, n_points, n_frames, H, W, x0, y0, inv_pillar, occ_words,
        point_pillar, reinterpret_cast<int*>(scratch), pn_detail::scan_state_words(nw));
    PN_CHECK_LAUNCH();
  }
  int rc = pn_detail::mask_scan_emit(occ_words, word_prefix, nw, H * W, W, pillar_coords, m_cap,
                                     num_pillars, scratch, scratch_bytes, stream, /*state_is_zero=*/n_points > 0);
  if (rc != PN_OK) return rc;
  if (n_points > 0) {
    k_rank<<<blocks, kPtThreads, 0, stream>>>(occ_words, word_prefix, n_points, frame_offsets + n_frames,
                                              point_pillar);
    PN_CHECK_LAUNCH();
  }
  return PN_OK;
}

int pn_pfn_scatter_max(const float* points, int point_dim, int n_points, const int* n_points_live,
                       const int* point_pillar, const int* num_pillars, int m_cap, float x0, float y0,
                       float inv_pillar, float pillar_size, float x_offset, float y_offset,
                       const float* weight, const float* scale, const float* shift, int c_out,
                       float* out_f32, void* out_bf16, int* arg, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(point_dim >= 2 && point_dim <= kMaxPointDim);
  PN_REQUIRE(c_out == 32 || c_out == 64);
  PN_REQUIRE(num_pillars && weight && scale && shift && (out_f32 || out_bf16) && m_cap >= 0 && n_points >= 0);
  PN_REQUIRE(out_f32 || !arg);
  if (m_cap == 0) return PN_OK;
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  const int zero_blocks = sms * 8;
  const int blocks = PN_DIVUP(n_points, kPtThreads);
  auto smem_pfn = [&](int C) { return (size_t)(((kPtThreads * point_dim + 3) & ~3) + kPtThreads * (C + 1)) * sizeof(float); };
  // weight/scale/shift are HOST pointers: they travel as kernel parameters (constant bank), so the FFMAs
  // read them as constant operands instead of issuing loads.
  static_assert(sizeof(PfnParams<64>) <= 3600, "kernel parameter space");
  PfnParams<64> hp64;
  PfnParams<32>& hp32 = *reinterpret_cast<PfnParams<32>*>(&hp64);  // filled below for the C that is used
  const int fdim = point_dim + 2;
  auto fill = [&](auto& hp, int C) {
    for (int c = 0; c < C; ++c) {
      for (int k = 0; k < kMaxPointDim + 2; ++k) hp.w[c][k] = k < fdim ? weight[c * fdim + k] : 0.f;
      hp.scale[c] = scale[c];
      hp.shift[c] = shift[c];
    }
  };
  if (c_out == 32) fill(hp32, 32); else fill(hp64, 64);
  const void* wparams = c_out == 32 ? (const void*)&hp32 : (const void*)&hp64;
  if (!out_f32) {
    // bf16-only fast path: vector bf16 RED straight into the bf16 rows
    auto smem_bf = [&](int C) { return (size_t)(((kPtThreads * point_dim + 3) & ~3) + kPtThreads * (C / 2 + 4)) * 4; };
#define PN_PFN_BF16(C)                                                                                \
    do {                                                                                              \
      static pn_detail::PerDeviceOnce ok##C;                                                          \
      if (ok##C.need()) {                                                                             \
        PN_CUDA(cudaFuncSetAttribute(k_pfn_scatter_max_bf16<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem_bf(C) + 4096));                                        \
      }                                                                                               \
      k_zero_rows_bf16<C><<<zero_blocks, 256, 0, stream>>>((__nv_bfloat16*)out_bf16, num_pillars, m_cap); \
      PN_CHECK_LAUNCH();                                                                              \
      if (n_points > 0) {                                                                             \
        k_pfn_scatter_max_bf16<C><<<blocks, kPtThreads, smem_bf(C), stream>>>(                        \
            *reinterpret_cast<const PfnParams<C>*>(wparams), points, point_dim, n_points, n_points_live, \
            point_pillar, x0, y0, inv_pillar, pillar_size, x_offset, y_offset, (__nv_bfloat16*)out_bf16, m_cap); \
        PN_CHECK_LAUNCH();                                                                            \
      }                                                                                               \
    } while (0)
    if (c_out == 32) PN_PFN_BF16(32); else PN_PFN_BF16(64);
#undef PN_PFN_BF16
    return PN_OK;
  }
#define PN_PFN_LAUNCH(C)                                                                           \
  do {                                                                                             \
    static pn_detail::PerDeviceOnce smem_ok##C;                                                    \
    if (smem_ok##C.need()) {                                                                       \
      PN_CUDA(cudaFuncSetAttribute(k_pfn_scatter_max<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,\
                                   (int)smem_pfn(C) + 4096));                                      \
    }                                                                                             \
    k_zero_rows<C><<<zero_blocks, 256, 0, stream>>>(out_f32, arg, num_pillars, m_cap);             \
    PN_CHECK_LAUNCH();                                                                             \
    if (n_points > 0) {                                                                            \
      k_pfn_scatter_max<C><<<blocks, kPtThreads, smem_pfn(C), stream>>>(                                \
          *reinterpret_cast<const PfnParams<C>*>(wparams), points, point_dim, n_points, n_points_live,     \
          point_pillar, x0, y0, inv_pillar, pillar_size, x_offset, y_offset, out_f32, m_cap);          \
      PN_CHECK_LAUNCH();                                                                           \
      if (arg) {                                                                                   \
        k_pfn_argmax<C><<<PN_DIVUP(n_points, kPtThreads / 32), kPtThreads, 0, stream>>>(           \
            *reinterpret_cast<const PfnParams<C>*>(wparams), points, point_dim, n_points, n_points_live,   \
            point_pillar, x0, y0, inv_pillar, pillar_size, x_offset, y_offset, out_f32, arg, m_cap);     \
        PN_CHECK_LAUNCH();                                                                         \
      }                                                                                            \
    }                                                                                              \
    if (out_bf16) {                                                                                \
      k_bf16_rows<C><<<zero_blocks, 256, 0, stream>>>(out_f32, (__nv_bfloat16*)out_bf16,           \
                                                      num_pillars, m_cap);                         \
      PN_CHECK_LAUNCH();                                                                           \
    }                                                                                              \
  } while (0)
  if (c_out == 32) PN_PFN_LAUNCH(32); else PN_PFN_LAUNCH(64);
#undef PN_PFN_LAUNCH
  return PN_OK;
}

int pn_scatter_max_grad(const float* grad_out, const int* arg, const int* num_pillars, int m_cap,
                        int c_out, float* grad_src, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(grad_out && arg && grad_src && c_out > 0);   // num_pillars may be NULL: all m_cap rows (arg < 0 = none)
  if (m_cap == 0) return PN_OK;
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  k_scatter_max_grad<<<sms * 8, 256, 0, stream>>>(grad_out, arg, num_pillars, m_cap, c_out, grad_src);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
