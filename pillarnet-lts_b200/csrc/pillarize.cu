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

__global__ void __launch_bounds__(kPtThreads)
k_mark(const float* __restrict__ pts, int dim, const int* __restrict__ frame_off, int n_points,
       int n_frames, int H, int W, float x0, float y0, float inv, uint32_t* __restrict__ words,
       int* __restrict__ point_cell) {
  extern __shared__ __align__(16) float s_pts[];
  n_points = min(n_points, __ldg(frame_off + n_frames));  // live count on the device, capacity on the host
  const int first = blockIdx.x * kPtThreads;
  const int count = min(kPtThreads, n_points - first);
  if (count <= 0) return;
  stage_points(pts, (long long)first * dim, count * dim, s_pts);
  __syncthreads();
  if ((int)threadIdx.x >= count) return;
  const int p = first + threadIdx.x;
  const float x = s_pts[threadIdx.x * dim + 0];
  const float y = s_pts[threadIdx.x * dim + 1];
  const int cx = cell_coord(x, x0, inv);
  const int cy = cell_coord(y, y0, inv);
  int cell = -1;
  if (cx >= 0 && cx < W && cy >= 0 && cy < H) {
    const int b = frame_of(frame_off, n_frames, p);
    cell = (b * H + cy) * W + cx;
    atomicOr(words + (cell >> 5), 1u << (cell & 31));
  }
  point_cell[p] = cell;
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

// One warp walks 32 staged points; lane = output channel (C/32 channels per lane).  All lanes see
// the same point, so range checks are warp-uniform and the 128-byte row of out[] is hit by one
// coalesced RED.MAX per point.  Post-ReLU values are >= 0, so max on the raw int bits is exact and
// order-independent: the result is deterministic, unlike the reference's CAS loop.
template <int C>
__global__ void __launch_bounds__(kPtThreads)
k_pfn_scatter_max(const float* __restrict__ pts, int dim, int n_points, const int* __restrict__ n_live,
                  const int* __restrict__ point_pillar, float x0, float y0,
                  float inv, float ps, float xoff, float yoff, const float* __restrict__ weight,
                  const float* __restrict__ scale, const float* __restrict__ shift,
                  float* __restrict__ out, int m_cap) {
  extern __shared__ __align__(16) float s_pts[];
  __shared__ int s_rank[kPtThreads];
  constexpr int R = C / 32;
  if (n_live) n_points = min(n_points, __ldg(n_live));
  const int first = blockIdx.x * kPtThreads;
  const int count = min(kPtThreads, n_points - first);
  if (count <= 0) return;
  stage_points(pts, (long long)first * dim, count * dim, s_pts);
  if ((int)threadIdx.x < count) s_rank[threadIdx.x] = __ldg(point_pillar + first + threadIdx.x);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fdim = dim + 2;
  float w[R][kMaxPointDim + 2], sc[R], sh[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
#pragma unroll
    for (int k = 0; k < kMaxPointDim + 2; ++k) w[r][k] = k < fdim ? __ldg(weight + c * fdim + k) : 0.f;
    sc[r] = __ldg(scale + c);
    sh[r] = __ldg(shift + c);
  }
  __syncthreads();
  const int j0 = warp * 32;
  const int j1 = min(j0 + 32, count);
  for (int j = j0; j < j1; ++j) {
    const int rank = s_rank[j];
    if (rank < 0 || rank >= m_cap) continue;
    const float* p = s_pts + j * dim;
    const float x = p[0], y = p[1];
    // pillar centre: int->float, *pillar_size, +offset as three separately rounded steps
    // (pillar_utils.py:51-52), then the offset features (pillar_utils.py:54).
    const float cxf = (float)cell_coord(x, x0, inv);
    const float cyf = (float)cell_coord(y, y0, inv);
    const float f0 = __fsub_rn(x, __fadd_rn(__fmul_rn(cxf, ps), xoff));
    const float f1 = __fsub_rn(y, __fadd_rn(__fmul_rn(cyf, ps), yoff));
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float z = w[r][0] * f0;
      z = fmaf(w[r][1], f1, z);
#pragma unroll
      for (int k = 0; k < kMaxPointDim; ++k)
        if (k < dim) z = fmaf(w[r][k + 2], p[k], z);
      const float h = fmaf(z, sc[r], sh[r]);
      if (h > 0.f)
        atomicMax(reinterpret_cast<int*>(out) + (long long)rank * C + lane + 32 * r,
                  __float_as_int(h));
    }
  }
}

// Training only: deterministic argmax = lowest point id attaining the max (second pass, like the
// reference's scatter_arg_max_kernel but exact-compare and tie-broken).
template <int C>
__global__ void __launch_bounds__(kPtThreads)
k_pfn_argmax(const float* __restrict__ pts, int dim, int n_points, const int* __restrict__ n_live,
             const int* __restrict__ point_pillar, float x0, float y0, float inv, float ps,
             float xoff, float yoff, const float* __restrict__ weight,
             const float* __restrict__ scale, const float* __restrict__ shift,
             const float* __restrict__ out, int* __restrict__ arg, int m_cap) {
  const int p = blockIdx.x * (kPtThreads / 32) + (threadIdx.x >> 5);
  if (p >= n_points || (n_live && p >= __ldg(n_live))) return;
  const int rank = __ldg(point_pillar + p);
  if (rank < 0 || rank >= m_cap) return;
  const int lane = threadIdx.x & 31;
  const float* q = pts + (long long)p * dim;
  const float x = __ldg(q), y = __ldg(q + 1);
  const float f0 = __fsub_rn(x, __fadd_rn(__fmul_rn((float)cell_coord(x, x0, inv), ps), xoff));
  const float f1 = __fsub_rn(y, __fadd_rn(__fmul_rn((float)cell_coord(y, y0, inv), ps), yoff));
  const int fdim = dim + 2;
  for (int c = lane; c < C; c += 32) {
    const float* wr = weight + c * fdim;
    float z = __ldg(wr) * f0;
    z = fmaf(__ldg(wr + 1), f1, z);
    for (int k = 0; k < dim; ++k) z = fmaf(__ldg(wr + k + 2), __ldg(q + k), z);
    float h = fmaf(z, __ldg(scale + c), __ldg(shift + c));
    h = fmaxf(h, 0.f);
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
  const int n = min(*num_rows, m_cap);
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
  PN_CUDA(cudaMemsetAsync(occ_words, 0, nw * sizeof(uint32_t), stream));
  const int blocks = PN_DIVUP(n_points, kPtThreads);
  if (n_points > 0) {
    k_mark<<<blocks, kPtThreads, kPtThreads * point_dim * sizeof(float), stream>>>(
        points, point_dim, frame_offsets, n_points, n_frames, H, W, x0, y0, inv_pillar, occ_words,
        point_pillar);
    PN_CHECK_LAUNCH();
  }
  int rc = pn_detail::mask_scan_emit(occ_words, word_prefix, nw, H * W, W, pillar_coords, m_cap,
                                     num_pillars, scratch, scratch_bytes, stream);
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
  PN_REQUIRE(num_pillars && weight && scale && shift && out_f32 && m_cap >= 0 && n_points >= 0);
  if (m_cap == 0) return PN_OK;
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  const int zero_blocks = sms * 8;
  const int blocks = PN_DIVUP(n_points, kPtThreads);
  const size_t smem = kPtThreads * point_dim * sizeof(float);
#define PN_PFN_LAUNCH(C)                                                                           \
  do {                                                                                             \
    k_zero_rows<C><<<zero_blocks, 256, 0, stream>>>(out_f32, arg, num_pillars, m_cap);             \
    PN_CHECK_LAUNCH();                                                                             \
    if (n_points > 0) {                                                                            \
      k_pfn_scatter_max<C><<<blocks, kPtThreads, smem, stream>>>(                                  \
          points, point_dim, n_points, n_points_live, point_pillar, x0, y0, inv_pillar,       \
          pillar_size, x_offset, y_offset, weight, scale, shift, out_f32, m_cap);                               \
      PN_CHECK_LAUNCH();                                                                           \
      if (arg) {                                                                                   \
        k_pfn_argmax<C><<<PN_DIVUP(n_points, kPtThreads / 32), kPtThreads, 0, stream>>>(           \
            points, point_dim, n_points, n_points_live, point_pillar, x0, y0, inv_pillar,         \
            pillar_size, x_offset, y_offset, weight, scale, shift, out_f32, arg, m_cap);                                  \
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
  PN_REQUIRE(grad_out && arg && num_pillars && grad_src && c_out > 0);
  if (m_cap == 0) return PN_OK;
  const int sms = pn_detail::sm_count();
  if (sms <= 0) return PN_ERR_CUDA;
  k_scatter_max_grad<<<sms * 8, 256, 0, stream>>>(grad_out, arg, num_pillars, m_cap, c_out, grad_src);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
