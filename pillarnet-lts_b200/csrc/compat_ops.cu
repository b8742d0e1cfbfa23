// Operator-level drop-ins for the reference's `pillar_cuda` extension (det3d/ops/pillar_ops/src/pillar_api.cpp:10-21):
// the individual ops det3d's own Python (pillar_utils.py:34-50, group_utils.py:6-37) calls one by one.  The product
// path does not use them — pn_pillarize / pn_pfn_scatter_max fuse the whole sequence — they exist so that the
// reference's unmodified PillarQueryAndGroup / PillarMaxPooling Python can run against this library
// (pillarnet_lts_b200.compat.pillar_cuda), which gives the parity tests a second, reference-driven channel.
//
// All are HBM-bound streaming kernels: one pass, coalesced 16-byte accesses where the layout allows it, grid-stride
// loops over a grid sized from the SM count.
#include "common.cuh"

namespace {

inline int grid_for(long long work, int threads) {
  const int sms = pn_detail::sm_count();
  long long g = PN_DIVUP(work, (long long)threads);
  const long long cap = (long long)(sms > 0 ? sms : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// pillar_ops_gpu.cu:13-39: frame of a point from the per-frame counts, cell id = (b*H + y)*W + x for in-range cells,
// occupancy byte set.  pts_xy is read as one 8-byte load per point.
__global__ void __launch_bounds__(256)
k_compat_point_pillar_index(const int2* __restrict__ pts_xy, const int* __restrict__ batch_cnt, int n, int B, int H,
                            int W, unsigned char* __restrict__ mask, int* __restrict__ index) {
  extern __shared__ int s_end[];   // inclusive prefix of the per-frame counts
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) { acc += batch_cnt[b]; s_end[b] = acc; }
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 c = pts_xy[i];
    if (c.x < 0 || c.x >= W || c.y < 0 || c.y >= H) continue;   // index keeps the caller's initial value (-1)
    int b = 0;
    while (b < B - 1 && i >= s_end[b]) ++b;    // the reference puts every point past the last count into frame B-1
    const int cell = (b * H + c.y) * W + c.x;
    mask[cell] = 1;
    index[i] = cell;
  }
}

// pillar_ops_gpu.cu:60-78: position (B,H,W) holds the rank of an occupied cell or a negative value
__global__ void __launch_bounds__(256)
k_compat_pillar_indices(const int* __restrict__ position, long long cells, int H, int W, int* __restrict__ out) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells;
       c += (long long)gridDim.x * blockDim.x) {
    const int r = position[c];
    if (r < 0) continue;
    const int x = (int)(c % W);
    const long long t = c / W;
    out[3 * (long long)r + 0] = (int)(t / H);
    out[3 * (long long)r + 1] = (int)(t % H);
    out[3 * (long long)r + 2] = x;
  }
}

// group_ops_gpu.cu:8-17
__global__ void __launch_bounds__(256)
k_compat_gather_indice(const int* __restrict__ index, const int* __restrict__ indices, int L, int* __restrict__ outs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) outs[i] = indices[index[i]];
}

// group_ops_gpu.cu:20-32: outs[i,:] = features[index[i],:]; thread = (row, channel) so a row is read coalesced
__global__ void __launch_bounds__(256)
k_compat_gather_feature(const int* __restrict__ index, const float* __restrict__ features, long long total, int C,
                        float* __restrict__ outs) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long i = g / C;
    const int c = (int)(g - i * C);
    outs[g] = features[(long long)index[i] * C + c];
  }
}

// group_ops_gpu.cu:35-48: grad_features[index[i],:] += grad_outs[i,:]
__global__ void __launch_bounds__(256)
k_compat_gather_feature_grad(const int* __restrict__ index, const float* __restrict__ grad_outs, long long total, int C,
                             float* __restrict__ grad_features) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long i = g / C;
    const int c = (int)(g - i * C);
    atomicAdd(grad_features + (long long)index[i] * C + c, grad_outs[g]);
  }
}

}  // namespace

extern "C" {

int pn_compat_point_pillar_index(const int* pts_xy, const int* pts_batch_cnt, int n_points, int n_frames, int H, int W,
                                 unsigned char* pillars_mask, int* point_pillar_index, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n_points >= 0 && n_frames >= 1 && n_frames <= 4096 && H > 0 && W > 0);
  if (n_points == 0) return PN_OK;
  PN_REQUIRE(pts_xy && pts_batch_cnt && pillars_mask && point_pillar_index);
  k_compat_point_pillar_index<<<grid_for(n_points, 256), 256, n_frames * sizeof(int), stream>>>(
      (const int2*)pts_xy, pts_batch_cnt, n_points, n_frames, H, W, pillars_mask, point_pillar_index);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_compat_pillar_indices(const int* pillars_position, int n_frames, int H, int W, int* pillar_indices,
                             pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n_frames >= 0 && H > 0 && W > 0);
  const long long cells = (long long)n_frames * H * W;
  if (cells == 0) return PN_OK;
  PN_REQUIRE(pillars_position && pillar_indices);
  k_compat_pillar_indices<<<grid_for(cells, 256), 256, 0, stream>>>(pillars_position, cells, H, W, pillar_indices);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_compat_gather_indice(const int* index, const int* indices, int n, int* outs, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n >= 0);
  if (n == 0) return PN_OK;
  PN_REQUIRE(index && indices && outs);
  k_compat_gather_indice<<<grid_for(n, 256), 256, 0, stream>>>(index, indices, n, outs);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_compat_gather_feature(const int* index, const float* features, int n, int c, float* outs, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n >= 0 && c > 0);
  if (n == 0) return PN_OK;
  PN_REQUIRE(index && features && outs);
  k_compat_gather_feature<<<grid_for((long long)n * c, 256), 256, 0, stream>>>(index, features, (long long)n * c, c, outs);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_compat_gather_feature_grad(const int* index, const float* grad_outs, int n, int c, float* grad_features,
                                  pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n >= 0 && c > 0);
  if (n == 0) return PN_OK;
  PN_REQUIRE(index && grad_outs && grad_features);
  k_compat_gather_feature_grad<<<grid_for((long long)n * c, 256), 256, 0, stream>>>(index, grad_outs, (long long)n * c, c,
                                                                                  grad_features);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
