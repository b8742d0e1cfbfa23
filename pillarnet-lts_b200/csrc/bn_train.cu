// Batch-statistics BatchNorm1d over the ACTIVE rows of a sparse feature matrix, fused with residual add and ReLU,
// forward and backward — the training-mode form of det3d/models/backbones/base.py:155-213
// (SparseSequential(conv, BN1d(eps 1e-3, momentum 0.01)[, SparseReLU]); `out = relu(bn(conv(x)) + identity)`).
//
// The reference runs nn.BatchNorm1d + add + ReLU as ~4 PyTorch kernels per conv on an exactly sized (M, C) matrix, which
// needs the live row count on the host.  Here the row count stays a DEVICE scalar (`num_rows`) next to a host capacity,
// so a whole training step is sync-free and CUDA-graph capturable; rows >= *num_rows are neither read nor written.
//
//   pn_bn_stats        sums[0:C] = sum_r x[r,c], sums[C:2C] = sum_r x[r,c]^2           (rows < n; fp32 atomics)
//   pn_bn_finalize     mean, rstd, scale = gamma*rstd, shift = beta - mean*scale; running stats updated like
//                      torch (momentum, unbiased variance)
//   pn_bn_apply        y = act(x*scale + shift + residual)
//   pn_bn_bwd_stats    g = dy * (y > 0 if relu); sums[0:C] = sum g, sums[C:2C] = sum g * xhat
//   pn_bn_bwd_apply    dx = gamma*rstd*(g - sum_g/n - xhat*sum_gx/n); dres = g
// All are HBM-bound streaming kernels (coalesced along the channel dimension, grid-stride, grid sized from the SM count).
#include "common.cuh"

namespace {

inline int grid_for(long long work, int threads, int per_sm = 8) {
  const int sms = pn_detail::sm_count();
  long long g = PN_DIVUP(work, (long long)threads);
  const long long cap = (long long)(sms > 0 ? sms : 148) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T>
__device__ __forceinline__ float ld(const T* p);
template <>
__device__ __forceinline__ float ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st(T* p, float v);
template <>
__device__ __forceinline__ void st<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// eight consecutive channels per thread: one 16-byte access for bf16 rows, two for fp32 rows
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}

// thread = (row lane, channel): block of 256 threads covers 256 / C rows per pass (C <= 256, power of two here)
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
k_bn_stats(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld,
           const float* __restrict__ mean, const float* __restrict__ rstd, int relu, const int* __restrict__ num_rows,
           int rows_cap, int C, float* __restrict__ sums) {
  __shared__ float s_a[256], s_b[256];
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const int rpb = 256 / C;                       // rows per block pass
  const int c = threadIdx.x % C, rl = threadIdx.x / C;
  float a = 0.f, b = 0.f;
  float m = 0.f, rs = 0.f;
  if (BWD) { m = mean[c]; rs = rstd[c]; }
  if (rl < rpb) {
    for (long long r = (long long)blockIdx.x * rpb + rl; r < n; r += (long long)gridDim.x * rpb) {
      if (BWD) {
        float g = ld(dy + r * dy_ld + c);
        if (relu && !(ld(y + r * y_ld + c) > 0.f)) g = 0.f;
        const float xh = (ld(x + r * x_ld + c) - m) * rs;
        a += g;
        b += g * xh;
      } else {
        const float v = ld(x + r * x_ld + c);
        a += v;
        b += v * v;
      }
    }
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x < C) {
    for (int k = 1; k < rpb; ++k) { a += s_a[threadIdx.x + k * C]; b += s_b[threadIdx.x + k * C]; }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, b);
  }
}

__global__ void __launch_bounds__(256)
k_bn_finalize(const float* __restrict__ sums, const int* __restrict__ num_rows, int rows_cap, int C,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
              float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean,
              float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
    const float m = sums[c] * inv_n;
    const float var = fmaxf(sums[C + c] * inv_n - m * m, 0.f);      // biased (what normalises the batch)
    const float rs = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    mean[c] = m;
    rstd[c] = rs;
    scale[c] = g * rs;
    shift[c] = b - m * g * rs;
    if (running_mean && n > 0) {
      // torch: running = (1 - momentum) * running + momentum * batch, variance unbiased (n / (n - 1))
      const float unbiased = n > 1 ? var * ((float)n / (float)(n - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bn_apply(const T* __restrict__ x, int x_ld, const float* __restrict__ scale, const float* __restrict__ shift,
           const T* __restrict__ res, int res_ld, int relu, const int* __restrict__ num_rows, int rows_cap, int C,
           T* __restrict__ out, int out_ld) {
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const long long total = (long long)n * C;            // rows of the capacity past the live count are never read
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    float v = fmaf(ld(x + r * x_ld + c), scale[c], shift[c]);
    if (res) v += ld(res + r * res_ld + c);
    if (relu) v = fmaxf(v, 0.f);
    st(out + r * out_ld + c, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ x, int x_ld,
               const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
               const float* __restrict__ sums, int relu, const int* __restrict__ num_rows, int rows_cap, int C,
               T* __restrict__ dx, int dx_ld, T* __restrict__ dres, int dres_ld) {
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const long long total = (long long)n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    float g = ld(dy + r * dy_ld + c);
    if (relu && !(ld(y + r * y_ld + c) > 0.f)) g = 0.f;
    const float rs = rstd[c];
    const float xh = (ld(x + r * x_ld + c) - mean[c]) * rs;
    const float gm = gamma ? gamma[c] : 1.f;
    st(dx + r * dx_ld + c, gm * rs * (g - sums[c] * inv_n - xh * sums[C + c] * inv_n));
    if (dres) st(dres + r * dres_ld + c, g);
  }
}

// rows of a dense NHWC map at the sites of a rank table: out[r, :] = dense[b, y, x, :]
template <typename T>
__global__ void __launch_bounds__(256)
k_dense_to_sparse(const T* __restrict__ dense, int dense_ld, const int* __restrict__ coords,
                  const int* __restrict__ num_rows, int rows_cap, int H, int W, int C, T* __restrict__ out, int out_ld) {
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const long long total = (long long)n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const int b = coords[3 * r], y = coords[3 * r + 1], x = coords[3 * r + 2];
    out[r * out_ld + c] = dense[(((long long)b * H + y) * W + x) * dense_ld + c];
  }
}

// ---- 8-channel-per-thread variants (C % 8 == 0, 16-byte aligned rows): what the backbone uses ----------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_bn_apply8(const T* __restrict__ x, int x_ld, const float* __restrict__ scale, const float* __restrict__ shift,
            const T* __restrict__ res, int res_ld, int relu, const int* __restrict__ num_rows, int rows_cap, int C,
            T* __restrict__ out, int out_ld) {
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const int c8 = C >> 3;
  const long long total = (long long)n * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c8;
    const int c = (int)(i - r * c8) << 3;
    float v[8], sc[8], sh[8];
    ld8(x + r * x_ld + c, v);
    ld8(scale + c, sc);
    ld8(shift + c, sh);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
    if (res) {
      float q[8];
      ld8(res + r * res_ld + c, q);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += q[k];
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    st8(out + r * out_ld + c, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply8(const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ x, int x_ld,
                const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                const float* __restrict__ sums, int relu, const int* __restrict__ num_rows, int rows_cap, int C,
                T* __restrict__ dx, int dx_ld, T* __restrict__ dres, int dres_ld) {
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const int c8 = C >> 3;
  const long long total = (long long)n * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c8;
    const int c = (int)(i - r * c8) << 3;
    float g[8], xv[8], m[8], rs[8], s0[8], s1[8], gm[8], o[8];
    ld8(dy + r * dy_ld + c, g);
    if (relu) {
      float yv[8];
      ld8(y + r * y_ld + c, yv);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = yv[k] > 0.f ? g[k] : 0.f;
    }
    ld8(x + r * x_ld + c, xv);
    ld8(mean + c, m);
    ld8(rstd + c, rs);
    ld8(sums + c, s0);
    ld8(sums + C + c, s1);
    if (gamma) ld8(gamma + c, gm);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (xv[k] - m[k]) * rs[k];
      o[k] = (gamma ? gm[k] : 1.f) * rs[k] * (g[k] - s0[k] * inv_n - xh * s1[k] * inv_n);
    }
    st8(dx + r * dx_ld + c, o);
    if (dres) st8(dres + r * dres_ld + c, g);
  }
}

// statistics: thread = (row lane, 8-channel group); 256 threads cover 256 / (C/8) rows per pass
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
k_bn_stats8(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld,
            const float* __restrict__ mean, const float* __restrict__ rstd, int relu, const int* __restrict__ num_rows,
            int rows_cap, int C, float* __restrict__ sums) {
  __shared__ float s_a[256 * 8], s_b[256 * 8];
  const int n = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const int c8 = C >> 3;                       // <= 32
  const int rpb = 256 / c8;
  const int cg = threadIdx.x % c8, rl = threadIdx.x / c8;
  const int c = cg << 3;
  float a[8], b[8], m[8], rs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = 0.f; b[k] = 0.f; m[k] = 0.f; rs[k] = 0.f; }
  if (BWD) { ld8(mean + c, m); ld8(rstd + c, rs); }
  for (long long r = (long long)blockIdx.x * rpb + rl; r < n; r += (long long)gridDim.x * rpb) {
    float xv[8];
    ld8(x + r * x_ld + c, xv);
    if (BWD) {
      float g[8];
      ld8(dy + r * dy_ld + c, g);
      if (relu) {
        float yv[8];
        ld8(y + r * y_ld + c, yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = yv[k] > 0.f ? g[k] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] += g[k]; b[k] += g[k] * (xv[k] - m[k]) * rs[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] += xv[k]; b[k] += xv[k] * xv[k]; }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { s_a[k * 256 + threadIdx.x] = a[k]; s_b[k * 256 + threadIdx.x] = b[k]; }
  __syncthreads();
  // thread t < C reduces channel t over the row lanes
  if (threadIdx.x < C) {
    const int g = threadIdx.x >> 3, k = threadIdx.x & 7;
    float sa = 0.f, sb = 0.f;
    for (int l = 0; l < rpb; ++l) { sa += s_a[k * 256 + l * c8 + g]; sb += s_b[k * 256 + l * c8 + g]; }
    atomicAdd(sums + threadIdx.x, sa);
    atomicAdd(sums + C + threadIdx.x, sb);
  }
}

inline bool vec8_ok(const void* p, int ld, int c, int bytes) {
  return c % 8 == 0 && (ld * bytes) % 16 == 0 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

bool pow2_le256(int c) { return c >= 1 && c <= 256 && (c & (c - 1)) == 0; }

}  // namespace

extern "C" {

int pn_bn_stats(const void* x, int dtype, int x_ld, const int* num_rows, int rows_cap, int c, float* sums,
                pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(sums && rows_cap >= 0 && pow2_le256(c));
  PN_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)c * sizeof(float), stream));
  if (rows_cap == 0) return PN_OK;
  PN_REQUIRE(x);
  const int g = grid_for((long long)rows_cap * c, 256, 4);
  const int es = dtype == PN_F32 ? 4 : 2;
  if (c >= 8 && vec8_ok(x, x_ld, c, es)) {
    const int g8 = grid_for((long long)rows_cap * (c / 8), 256, 4);
    if (dtype == PN_F32)
      k_bn_stats8<float, false><<<g8, 256, 0, stream>>>((const float*)x, x_ld, nullptr, 0, nullptr, 0, nullptr, nullptr, 0,
                                                       num_rows, rows_cap, c, sums);
    else
      k_bn_stats8<__nv_bfloat16, false><<<g8, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, nullptr, 0, nullptr, 0,
                                                               nullptr, nullptr, 0, num_rows, rows_cap, c, sums);
    PN_CHECK_LAUNCH();
    return PN_OK;
  }
  if (dtype == PN_F32)
    k_bn_stats<float, false><<<g, 256, 0, stream>>>((const float*)x, x_ld, nullptr, 0, nullptr, 0, nullptr, nullptr, 0,
                                                    num_rows, rows_cap, c, sums);
  else
    k_bn_stats<__nv_bfloat16, false><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, nullptr, 0, nullptr, 0, nullptr,
                                                            nullptr, 0, num_rows, rows_cap, c, sums);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_bn_finalize(const float* sums, const int* num_rows, int rows_cap, int c, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                   float* scale, float* shift, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(sums && mean && rstd && scale && shift && c >= 1);
  k_bn_finalize<<<1, 256, 0, stream>>>(sums, num_rows, rows_cap, c, gamma, beta, eps, momentum, running_mean, running_var,
                                       mean, rstd, scale, shift);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_bn_apply(const void* x, int dtype, int x_ld, const float* scale, const float* shift, const void* residual,
                int res_ld, int relu, const int* num_rows, int rows_cap, int c, void* out, int out_ld,
                pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(rows_cap >= 0 && c >= 1);
  if (rows_cap == 0) return PN_OK;
  PN_REQUIRE(x && scale && shift && out);
  const int g = grid_for((long long)rows_cap * c, 256);
  const int es = dtype == PN_F32 ? 4 : 2;
  if (vec8_ok(x, x_ld, c, es) && vec8_ok(out, out_ld, c, es) && (!residual || vec8_ok(residual, res_ld, c, es)) &&
      vec8_ok(scale, 8, 8, 4) && vec8_ok(shift, 8, 8, 4)) {
    const int g8 = grid_for((long long)rows_cap * (c / 8), 256);
    if (dtype == PN_F32)
      k_bn_apply8<float><<<g8, 256, 0, stream>>>((const float*)x, x_ld, scale, shift, (const float*)residual, res_ld, relu,
                                                 num_rows, rows_cap, c, (float*)out, out_ld);
    else
      k_bn_apply8<__nv_bfloat16><<<g8, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, scale, shift,
                                                         (const __nv_bfloat16*)residual, res_ld, relu, num_rows, rows_cap,
                                                         c, (__nv_bfloat16*)out, out_ld);
    PN_CHECK_LAUNCH();
    return PN_OK;
  }
  if (dtype == PN_F32)
    k_bn_apply<float><<<g, 256, 0, stream>>>((const float*)x, x_ld, scale, shift, (const float*)residual, res_ld, relu,
                                             num_rows, rows_cap, c, (float*)out, out_ld);
  else
    k_bn_apply<__nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, scale, shift,
                                                     (const __nv_bfloat16*)residual, res_ld, relu, num_rows, rows_cap, c,
                                                     (__nv_bfloat16*)out, out_ld);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_bn_bwd_stats(const void* dy, int dtype, int dy_ld, const void* y, int y_ld, const void* x, int x_ld,
                    const float* mean, const float* rstd, int relu, const int* num_rows, int rows_cap, int c,
                    float* sums, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(sums && rows_cap >= 0 && pow2_le256(c));
  PN_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)c * sizeof(float), stream));
  if (rows_cap == 0) return PN_OK;
  PN_REQUIRE(dy && x && mean && rstd && (y || !relu));
  const int g = grid_for((long long)rows_cap * c, 256, 4);
  const int es = dtype == PN_F32 ? 4 : 2;
  if (c >= 8 && vec8_ok(x, x_ld, c, es) && vec8_ok(dy, dy_ld, c, es) && (!relu || vec8_ok(y, y_ld, c, es)) &&
      vec8_ok(mean, 8, 8, 4) && vec8_ok(rstd, 8, 8, 4)) {
    const int g8 = grid_for((long long)rows_cap * (c / 8), 256, 4);
    if (dtype == PN_F32)
      k_bn_stats8<float, true><<<g8, 256, 0, stream>>>((const float*)x, x_ld, (const float*)dy, dy_ld, (const float*)y, y_ld,
                                                      mean, rstd, relu, num_rows, rows_cap, c, sums);
    else
      k_bn_stats8<__nv_bfloat16, true><<<g8, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)dy,
                                                              dy_ld, (const __nv_bfloat16*)y, y_ld, mean, rstd, relu,
                                                              num_rows, rows_cap, c, sums);
    PN_CHECK_LAUNCH();
    return PN_OK;
  }
  if (dtype == PN_F32)
    k_bn_stats<float, true><<<g, 256, 0, stream>>>((const float*)x, x_ld, (const float*)dy, dy_ld, (const float*)y, y_ld,
                                                   mean, rstd, relu, num_rows, rows_cap, c, sums);
  else
    k_bn_stats<__nv_bfloat16, true><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)dy, dy_ld,
                                                           (const __nv_bfloat16*)y, y_ld, mean, rstd, relu, num_rows,
                                                           rows_cap, c, sums);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_bn_bwd_apply(const void* dy, int dtype, int dy_ld, const void* y, int y_ld, const void* x, int x_ld,
                    const float* mean, const float* rstd, const float* gamma, const float* sums, int relu,
                    const int* num_rows, int rows_cap, int c, void* dx, int dx_ld, void* dres, int dres_ld,
                    pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(rows_cap >= 0 && c >= 1);
  if (rows_cap == 0) return PN_OK;
  PN_REQUIRE(dy && x && mean && rstd && sums && dx && (y || !relu));
  const int g = grid_for((long long)rows_cap * c, 256);
  const int es = dtype == PN_F32 ? 4 : 2;
  if (vec8_ok(x, x_ld, c, es) && vec8_ok(dy, dy_ld, c, es) && (!relu || vec8_ok(y, y_ld, c, es)) &&
      vec8_ok(dx, dx_ld, c, es) && (!dres || vec8_ok(dres, dres_ld, c, es)) && vec8_ok(mean, 8, 8, 4) &&
      vec8_ok(rstd, 8, 8, 4) && vec8_ok(sums, 8, 8, 4) && (!gamma || vec8_ok(gamma, 8, 8, 4))) {
    const int g8 = grid_for((long long)rows_cap * (c / 8), 256);
    if (dtype == PN_F32)
      k_bn_bwd_apply8<float><<<g8, 256, 0, stream>>>((const float*)dy, dy_ld, (const float*)y, y_ld, (const float*)x, x_ld,
                                                     mean, rstd, gamma, sums, relu, num_rows, rows_cap, c, (float*)dx,
                                                     dx_ld, (float*)dres, dres_ld);
    else
      k_bn_bwd_apply8<__nv_bfloat16><<<g8, 256, 0, stream>>>(
          (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)y, y_ld, (const __nv_bfloat16*)x, x_ld, mean, rstd, gamma,
          sums, relu, num_rows, rows_cap, c, (__nv_bfloat16*)dx, dx_ld, (__nv_bfloat16*)dres, dres_ld);
    PN_CHECK_LAUNCH();
    return PN_OK;
  }
  if (dtype == PN_F32)
    k_bn_bwd_apply<float><<<g, 256, 0, stream>>>((const float*)dy, dy_ld, (const float*)y, y_ld, (const float*)x, x_ld,
                                                 mean, rstd, gamma, sums, relu, num_rows, rows_cap, c, (float*)dx, dx_ld,
                                                 (float*)dres, dres_ld);
  else
    k_bn_bwd_apply<__nv_bfloat16><<<g, 256, 0, stream>>>(
        (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)y, y_ld, (const __nv_bfloat16*)x, x_ld, mean, rstd, gamma,
        sums, relu, num_rows, rows_cap, c, (__nv_bfloat16*)dx, dx_ld, (__nv_bfloat16*)dres, dres_ld);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_dense_to_sparse(const void* dense, int dtype, int dense_ld, const int* coords, const int* num_rows, int rows_cap,
                       int H, int W, int c, void* out, int out_ld, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(rows_cap >= 0 && c >= 1 && H > 0 && W > 0);
  if (rows_cap == 0) return PN_OK;
  PN_REQUIRE(dense && coords && out);
  const int g = grid_for((long long)rows_cap * c, 256);
  if (dtype == PN_F32)
    k_dense_to_sparse<float><<<g, 256, 0, stream>>>((const float*)dense, dense_ld, coords, num_rows, rows_cap, H, W, c,
                                                    (float*)out, out_ld);
  else
    k_dense_to_sparse<__nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)dense, dense_ld, coords, num_rows,
                                                            rows_cap, H, W, c, (__nv_bfloat16*)out, out_ld);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
