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

// Each active input marks the (up to 4) outputs whose 3x3/s2/p1 window contains it:
// input y is tap ky of output oy iff 2*oy - 1 + ky == y.
__global__ void __launch_bounds__(256)
k_down_mark(const int* __restrict__ coords, const int* __restrict__ num_rows, int m_cap, int Ho,
            int Wo, uint32_t* __restrict__ out_words) {
  const int n = min(*num_rows, m_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = __ldg(coords + 3 * i), y = __ldg(coords + 3 * i + 1), x = __ldg(coords + 3 * i + 2);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        const int cell = (b * Ho + oy) * Wo + ox;
        atomicOr(out_words + (cell >> 5), 1u << (cell & 31));
      }
    }
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
  const long long cap = (long long)(sms > 0 ? sms : 148) * 16;
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
  PN_CUDA(cudaMemsetAsync(out_words, 0, nw * sizeof(uint32_t), stream));
  if (in_m_cap > 0) {
    k_down_mark<<<grid_for(in_m_cap, 256), 256, 0, stream>>>(in_coords, in_num_rows, in_m_cap, Ho,
                                                            Wo, out_words);
    PN_CHECK_LAUNCH();
  }
  int rc = pn_detail::mask_scan_emit(out_words, out_prefix, nw, Ho * Wo, Wo, out_coords, out_m_cap,
                                     out_num_rows, scratch, scratch_bytes, stream);
  if (rc != PN_OK) return rc;
  if (out_m_cap > 0) {
    k_down_nbr<<<grid_for((long long)out_m_cap * 9, 256), 256, 0, stream>>>(
        in_words, in_prefix, H_in, W_in, out_coords, out_num_rows, out_m_cap, nbr);
    PN_CHECK_LAUNCH();
  }
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
