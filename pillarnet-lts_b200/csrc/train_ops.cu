// Training-path operators: the pieces of the reader / sparse-backbone backward that the
// reference gets from pillar_cuda and from spconv's autograd.
//
//   pn_point_features      det3d/ops/pillar_ops/pillar_utils.py:51-56 (+ gather_feature, group_ops_gpu.cu:20-33)
//   pn_scatter_max         det3d/ops/pillar_ops/src/scatter_ops_gpu.cu:13-36 (scatter_max_wrapper)
//   pn_rulebook_transpose  input-stationary view of a rulebook: the indice pairs spconv's dgrad walks
//   pn_conv_wgrad          spconv's weight-gradient implicit GEMM, fp32 SIMT form (the tcgen05 form is in
//                          conv_wgrad_tc.cu); dW[co][t*cin+ci] = sum_o dy[o][co] * x[nbr[o,t]][ci]
#include "common.cuh"

namespace pn_detail {
int conv_wgrad_tcgen05(const void* x, int x_ld, const void* dy, int dy_ld, const int* nbr, int taps,
                       const int* num_rows, int rows_cap, int cin, int cout, float* dw, int dw_ld,
                       cudaStream_t stream);  // conv_wgrad_tc.cu
}

namespace {

constexpr int kMaxDim = 8;

__device__ __forceinline__ int cell_of(float v, float v0, float inv) {
  return (int)floorf(__fmul_rn(__fsub_rn(v, v0), inv));
}

__global__ void __launch_bounds__(256)
k_point_features(const float* __restrict__ pts, int dim, int n, float x0, float y0, float inv, float ps,
                 float xoff, float yoff, float* __restrict__ out) {
  const int od = dim + 2;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const float* q = pts + (long long)p * dim;
    const float x = __ldg(q), y = __ldg(q + 1);
    // centre = (float)cell * pillar_size + offset, two roundings (pillar_utils.py:52-53)
    const float cx = __fadd_rn(__fmul_rn((float)cell_of(x, x0, inv), ps), xoff);
    const float cy = __fadd_rn(__fmul_rn((float)cell_of(y, y0, inv), ps), yoff);
    float* o = out + (long long)p * od;
    o[0] = __fsub_rn(x, cx);
    o[1] = __fsub_rn(y, cy);
    for (int k = 0; k < dim; ++k) o[2 + k] = __ldg(q + k);
  }
}

// out is zero-initialised and only values > 0 can raise it, so the max runs on the int bit patterns.
__global__ void __launch_bounds__(256)
k_scatter_max(const float* __restrict__ src, const int* __restrict__ index, long long total, int C, int M,
              float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / C), c = (int)(i - (long long)p * C);
    const int m = __ldg(index + p);
    if (m < 0 || m >= M) continue;
    const float v = src[i];
    if (v > 0.f) atomicMax(reinterpret_cast<int*>(out) + (long long)m * C + c, __float_as_int(v));
  }
}

// arg = lowest flat index i = p*C + c whose value equals the stored max (the reference accepts any point
// within 1e-5 and lets the last writer win, scatter_ops_gpu.cu:25-36).
__global__ void __launch_bounds__(256)
k_scatter_argmax(const float* __restrict__ src, const int* __restrict__ index, long long total, int C, int M,
                 const float* __restrict__ out, int* __restrict__ arg) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / C), c = (int)(i - (long long)p * C);
    const int m = __ldg(index + p);
    if (m < 0 || m >= M) continue;
    const long long o = (long long)m * C + c;
    if (src[i] == out[o]) atomicMin(reinterpret_cast<unsigned*>(arg) + o, (unsigned)i);
  }
}

__global__ void __launch_bounds__(256)
k_rulebook_transpose(const int* __restrict__ nbr, const int* __restrict__ num_out, int out_cap, int taps,
                     int in_cap, int* __restrict__ nbr_t) {
  const int n = num_out ? min(*num_out, out_cap) : out_cap;
  const long long total = (long long)n * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int src = __ldg(nbr + i);
    if (src < 0 || src >= in_cap) continue;
    const int o = (int)(i / taps), t = (int)(i - (long long)o * taps);
    nbr_t[(long long)src * taps + t] = o;   // (input, tap) pairs are unique: at most one output per tap
  }
}

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

constexpr int WT = 64, WR = 32;

// One CTA: (row split, tap, 64 couts, 64 cins); fp32 accumulation, one atomicAdd per element at the end.
template <typename T>
__global__ void __launch_bounds__(256)
k_wgrad_simt(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld,
             const int* __restrict__ nbr, int taps, const int* __restrict__ num_rows, int rows_cap, int cin,
             int cout, int rows_per_split, int n_ci_tiles, int n_co_tiles, float* __restrict__ dw, int dw_ld) {
  __shared__ __align__(16) float sY[WR][WT];
  __shared__ __align__(16) float sX[WR][WT];
  __shared__ int sN[WR];
  const int rows = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const int r_begin = blockIdx.x * rows_per_split;
  const int r_end = min(rows, r_begin + rows_per_split);
  if (r_begin >= r_end) return;
  int w = blockIdx.y;
  const int ci0 = (w % n_ci_tiles) * WT; w /= n_ci_tiles;
  const int co0 = (w % n_co_tiles) * WT; w /= n_co_tiles;
  const int t = w;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += WR) {
    __syncthreads();
    if (threadIdx.x < WR) {
      const int r = r0 + threadIdx.x;
      int src = -1;
      if (r < r_end) src = nbr ? __ldg(nbr + (long long)r * taps + t) : r;
      sN[threadIdx.x] = src;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < WR * WT; i += 256) {
      const int m = i / WT, c = i % WT;
      const int src = sN[m];
      float vy = 0.f, vx = 0.f;
      if (src >= 0) {
        if (co0 + c < cout) vy = ldf<T>(dy + (long long)(r0 + m) * dy_ld + co0 + c);
        if (ci0 + c < cin) vx = ldf<T>(x + (long long)src * x_ld + ci0 + c);
      }
      sY[m][c] = vy;
      sX[m][c] = vx;
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < WR; ++m) {
      const float4 a = *reinterpret_cast<const float4*>(&sY[m][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sX[m][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci >= cin) continue;
      atomicAdd(dw + (long long)co * dw_ld + (long long)t * cin + ci, acc[i][j]);
    }
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

int pn_point_features(const float* points, int point_dim, int n_points, float x0, float y0, float inv_pillar,
                      float pillar_size, float x_offset, float y_offset, float* out, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(point_dim >= 2 && point_dim <= kMaxDim && n_points >= 0);
  if (n_points == 0) return PN_OK;
  PN_REQUIRE(points && out);
  k_point_features<<<grid_for(n_points, 256), 256, 0, stream>>>(points, point_dim, n_points, x0, y0, inv_pillar,
                                                               pillar_size, x_offset, y_offset, out);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_scatter_max(const float* src, const int* index, int n_points, int n_pillars, int c, float* out, int* arg,
                   pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n_points >= 0 && n_pillars >= 0 && c > 0);
  if (n_pillars == 0) return PN_OK;
  PN_REQUIRE(out);
  PN_CUDA(cudaMemsetAsync(out, 0, (size_t)n_pillars * c * sizeof(float), stream));
  if (arg) PN_CUDA(cudaMemsetAsync(arg, 0xFF, (size_t)n_pillars * c * sizeof(int), stream));
  if (n_points == 0) return PN_OK;
  PN_REQUIRE(src && index);
  const long long total = (long long)n_points * c;
  PN_REQUIRE(total < 0x7FFFFFFFll);   // arg holds flat indices in int32, as the reference's does
  k_scatter_max<<<grid_for(total, 256), 256, 0, stream>>>(src, index, total, c, n_pillars, out);
  PN_CHECK_LAUNCH();
  if (arg) {
    k_scatter_argmax<<<grid_for(total, 256), 256, 0, stream>>>(src, index, total, c, n_pillars, out, arg);
    PN_CHECK_LAUNCH();
  }
  return PN_OK;
}

int pn_rulebook_transpose(const int* nbr, const int* num_out, int out_cap, int taps, int in_cap, int* nbr_t,
                          pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(taps >= 1 && out_cap >= 0 && in_cap >= 0);
  if (in_cap == 0) return PN_OK;
  PN_REQUIRE(nbr_t);
  PN_CUDA(cudaMemsetAsync(nbr_t, 0xFF, (size_t)in_cap * taps * sizeof(int), stream));
  if (out_cap == 0) return PN_OK;
  PN_REQUIRE(nbr);
  k_rulebook_transpose<<<grid_for((long long)out_cap * taps, 256), 256, 0, stream>>>(nbr, num_out, out_cap, taps,
                                                                                    in_cap, nbr_t);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_conv_wgrad(const void* x, int x_dtype, int x_ld, const void* dy, int dy_dtype, int dy_ld, const int* nbr,
                  int taps, const int* num_rows, int rows_cap, int cin, int cout, float* dw, int dw_ld, int impl,
                  pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(x && dy && dw && taps >= 1 && cin > 0 && cout > 0 && rows_cap >= 0);
  PN_REQUIRE(x_dtype == dy_dtype && (x_dtype == PN_F32 || x_dtype == PN_BF16));
  PN_REQUIRE(dw_ld >= taps * cin && x_ld >= cin && dy_ld >= cout);
  PN_CUDA(cudaMemset2DAsync(dw, (size_t)dw_ld * sizeof(float), 0, (size_t)taps * cin * sizeof(float), cout,
                            stream));
  if (rows_cap == 0) return PN_OK;
  if (impl == PN_IMPL_TCGEN05) {
    PN_REQUIRE(x_dtype == PN_BF16);
    return pn_detail::conv_wgrad_tcgen05(x, x_ld, dy, dy_ld, nbr, taps, num_rows, rows_cap, cin, cout, dw, dw_ld,
                                         stream);
  }
  PN_REQUIRE(impl == PN_IMPL_SIMT);
  const int n_ci = PN_DIVUP(cin, WT), n_co = PN_DIVUP(cout, WT);
  const int tiles = taps * n_ci * n_co;
  const int sms = pn_detail::sm_count() > 0 ? pn_detail::sm_count() : 148;
  int splits = PN_DIVUP(sms * 4, tiles);
  const int max_splits = PN_DIVUP(rows_cap, 4 * WR);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rps = PN_DIVUP(rows_cap, splits);
  rps = PN_DIVUP(rps, WR) * WR;
  splits = PN_DIVUP(rows_cap, rps);
  dim3 grid(splits, tiles);
  if (x_dtype == PN_F32)
    k_wgrad_simt<float><<<grid, 256, 0, stream>>>((const float*)x, x_ld, (const float*)dy, dy_ld, nbr, taps, num_rows,
                                                  rows_cap, cin, cout, rps, n_ci, n_co, dw, dw_ld);
  else
    k_wgrad_simt<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)dy,
                                                          dy_ld, nbr, taps, num_rows, rows_cap, cin, cout, rps, n_ci,
                                                          n_co, dw, dw_ld);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
