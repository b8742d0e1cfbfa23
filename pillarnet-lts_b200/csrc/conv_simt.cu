// Gather-GEMM convolution, fp32-FMA path (PN_IMPL_SIMT), plus layout helpers.
//
// This is the precise (true fp32 accumulate, no tensor cores) implementation of pn_conv_gather used as
// the validation mode for the tcgen05 path and for the 1e-3 fp32 tolerance contract.  It restates
//   spconv SubMConv2d / SparseConv2d  (used at det3d/models/backbones/base.py:38-63)
//   nn.Conv2d 3x3 / nn.ConvTranspose2d 2x2 s2 (det3d/models/necks/rpn.py:147-207, center_head.py:27-35)
// as out[o,:] = sum_t W_t . in[nbr[o,t],:], with the per-channel affine (BN eval + bias), residual
// add and ReLU of base.py:155-213 fused into the epilogue.
#include "common.cuh"

namespace pn_detail {
int conv_tcgen05(const pn_conv_args* a, cudaStream_t stream);  // conv_tcgen05.cu
}

namespace {

constexpr int TM = 32, TN = 64, TK = 16, THREADS = 128;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(THREADS)
k_conv_simt(const TIn* __restrict__ in, int in_ld, const int* __restrict__ nbr, int taps,
            const TIn* __restrict__ weight, int k_pad, const float* __restrict__ scale,
            const float* __restrict__ shift, const TOut* __restrict__ residual, int res_ld,
            TOut* __restrict__ out, int out_ld, int out_coff, int relu,
            const int* __restrict__ num_rows, int rows_cap, int cin, int cout, int out_hp, int out_wp) {
  __shared__ float sA[TK][TM + 1];
  __shared__ float sW[TK][TN + 1];
  __shared__ int sNbr[TM];
  const int rows = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const int row0 = blockIdx.x * TM;
  if (row0 >= rows) return;
  const int n0 = blockIdx.y * TN;
  const int tx = threadIdx.x & 15;  // 16 column groups x 4 couts
  const int ty = threadIdx.x >> 4;  // 8 row groups x 4 rows
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < taps; ++t) {
    __syncthreads();
    if (threadIdx.x < TM) {
      const int r = row0 + threadIdx.x;
      int src = -1;
      if (r < rows) src = nbr ? __ldg(nbr + (long long)r * taps + t) : r;
      sNbr[threadIdx.x] = src;
    }
    __syncthreads();
    for (int c0 = 0; c0 < cin; c0 += TK) {
      // A tile: TM rows x TK channels
      for (int i = threadIdx.x; i < TM * TK; i += THREADS) {
        const int m = i / TK, k = i % TK;
        const int src = sNbr[m];
        float v = 0.f;
        if (src >= 0 && c0 + k < cin) v = to_f32<TIn>(in[(long long)src * in_ld + c0 + k]);
        sA[k][m] = v;
      }
      // W tile: TN couts x TK
      for (int i = threadIdx.x; i < TN * TK; i += THREADS) {
        const int n = i / TK, k = i % TK;
        float v = 0.f;
        if (n0 + n < cout && c0 + k < cin)
          v = to_f32<TIn>(weight[(long long)(n0 + n) * k_pad + (long long)t * cin + c0 + k]);
        sW[k][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = sW[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= rows) continue;
    bool border = false;
    if (out_wp > 0) {
      const int q = r % (out_hp * out_wp);
      const int y = q / out_wp, x = q - y * out_wp;
      border = x == 0 || x == out_wp - 1 || y == 0 || y == out_hp - 1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= cout) continue;
      float v = acc[i][j];
      const float sc = scale ? __ldg(scale + n) : 1.f;
      const float sh = shift ? __ldg(shift + n) : 0.f;
      v = fmaf(v, sc, sh);
      if (residual) v += to_f32<TOut>(residual[(long long)r * res_ld + n]);
      if (relu) v = fmaxf(v, 0.f);
      if (border) v = 0.f;
      out[(long long)r * out_ld + out_coff + n] = from_f32<TOut>(v);
    }
  }
}

template <typename TIn, typename TOut>
int launch_simt(const pn_conv_args* a, cudaStream_t stream) {
  dim3 grid(PN_DIVUP(a->rows_cap, TM), PN_DIVUP(a->cout, TN));
  if (grid.x == 0) return PN_OK;
  k_conv_simt<TIn, TOut><<<grid, THREADS, 0, stream>>>(
      (const TIn*)a->in, a->in_ld, a->nbr, a->taps, (const TIn*)a->weight, a->k_pad, a->scale,
      a->shift, (const TOut*)a->residual, a->res_ld, (TOut*)a->out, a->out_ld, a->out_coff, a->relu,
      a->num_rows, a->rows_cap, a->cin, a->cout, a->out_hp, a->out_wp);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

__global__ void __launch_bounds__(256)
k_pack_weight(const float* __restrict__ w, int cout, int k, int k_pad, __nv_bfloat16* __restrict__ o) {
  const long long total = (long long)cout * k_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / k_pad), kk = (int)(i - (long long)n * k_pad);
    o[i] = __float2bfloat16_rn(kk < k ? w[(long long)n * k + kk] : 0.f);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
k_cast_rows(const TI* __restrict__ in, int in_ld, TO* __restrict__ out, int out_ld, int cols,
            const int* __restrict__ num_rows, int rows_cap) {
  const int rows = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[r * out_ld + c] = from_f32<TO>(to_f32<TI>(in[r * in_ld + c]));
  }
}

// f32 row -> [hi | lo | hi] bf16 with hi = bf16(x), lo = bf16(x - hi): the input side of the split-bf16 tensor-core
// mode (x * w ~= hi_x * hi_w + lo_x * hi_w + hi_x * lo_w against weights laid out [hi | hi | lo] per tap; the dropped
// lo * lo term and the rounding of the lo parts are ~2^-17 relative).
__global__ void __launch_bounds__(256)
k_split_bf16x3(const float* __restrict__ in, int in_ld, __nv_bfloat16* __restrict__ out, int cols,
               const int* __restrict__ num_rows, int rows_cap) {
  const int rows = num_rows ? min(*num_rows, rows_cap) : rows_cap;
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float x = in[r * in_ld + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * 3ll * cols + c;
    o[0] = hi;
    o[cols] = lo;
    o[2 * cols] = hi;
  }
}

// Dense-driven densify: one thread per (pixel, 8-byte chunk); absent pixels get zeros, so no memset
// and every output byte is written exactly once, coalesced along channels.  pad = 1: the output rows
// are the zero-padded (H+2, W+2) map (borders written as zeros).
template <typename T>
__global__ void __launch_bounds__(256)
k_sparse_to_dense(const T* __restrict__ feat, int feat_ld, const uint32_t* __restrict__ words,
                  const int* __restrict__ prefix, int n_frames, int H, int W, int pad, int C,
                  T* __restrict__ out, int out_ld, int out_coff) {
  constexpr int V = 8 / sizeof(T);  // elements per 8-byte chunk
  const int chunks = C / V;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long total = (long long)n_frames * Hp * Wp * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / chunks;
    const int ch = (int)(i - pos * chunks);
    const int x = (int)(pos % Wp) - pad, y = (int)((pos / Wp) % Hp) - pad;
    const int b = (int)(pos / ((long long)Wp * Hp));
    int r = -1;
    if (x >= 0 && x < W && y >= 0 && y < H) r = pn_rank_of(words, prefix, (b * H + y) * W + x);
    uint2 v = make_uint2(0u, 0u);
    if (r >= 0) v = *reinterpret_cast<const uint2*>(feat + (long long)r * feat_ld + ch * V);
    *reinterpret_cast<uint2*>(out + pos * out_ld + out_coff + ch * V) = v;
  }
}

// 16-byte variant: a group of C*sizeof(T)/16 consecutive lanes copies one position, so the rank lookup and the
// position arithmetic are shared by the group's coalesced 16-byte loads and stores (the 8-byte kernel above
// repeats them per chunk: 15.2 -> 8.7 us for the 180x180x256 map of the nuScenes frame).
template <typename T>
__global__ void __launch_bounds__(256)
k_sparse_to_dense16(const T* __restrict__ feat, int feat_ld, const uint32_t* __restrict__ words,
                    const int* __restrict__ prefix, int n_frames, int H, int W, int pad, int C,
                    T* __restrict__ out, int out_ld, int out_coff) {
  constexpr int V = 16 / sizeof(T);
  const int chunks = C / V;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long total = (long long)n_frames * Hp * Wp * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / chunks;
    const int ch = (int)(i - pos * chunks);
    const int q = (int)(pos % ((long long)Wp * Hp));
    const int b = (int)(pos / ((long long)Wp * Hp));
    const int yy = q / Wp, x = q - yy * Wp - pad, y = yy - pad;
    int r = -1;
    if (x >= 0 && x < W && y >= 0 && y < H) r = pn_rank_of(words, prefix, (b * H + y) * W + x);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r >= 0) v = __ldg(reinterpret_cast<const uint4*>(feat + (long long)r * feat_ld + ch * V));
    *reinterpret_cast<uint4*>(out + pos * out_ld + out_coff + ch * V) = v;
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

int pn_conv_gather(const pn_conv_args* a, int impl, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(a && a->in && a->weight && a->out);
  PN_REQUIRE(a->taps >= 1 && a->cin >= 1 && a->cout >= 1 && a->rows_cap >= 0);
  PN_REQUIRE(a->nbr != nullptr || a->taps == 1);
  PN_REQUIRE(a->k_pad >= a->taps * a->cin);
  PN_REQUIRE(a->in_dtype == PN_F32 || a->in_dtype == PN_BF16);
  PN_REQUIRE(a->out_dtype == PN_F32 || a->out_dtype == PN_BF16);
  if (a->rows_cap == 0) return PN_OK;
  if (impl == PN_IMPL_TCGEN05) return pn_detail::conv_tcgen05(a, stream);
  if (impl != PN_IMPL_SIMT) return PN_ERR_INVALID_ARG;
  if (a->deconv_cout != 0) return PN_ERR_UNSUPPORTED;   // the GEMM form of the transposed conv is tensor-core only
  if (a->in_dtype == PN_F32 && a->out_dtype == PN_F32) return launch_simt<float, float>(a, stream);
  if (a->in_dtype == PN_BF16 && a->out_dtype == PN_BF16)
    return launch_simt<__nv_bfloat16, __nv_bfloat16>(a, stream);
  if (a->in_dtype == PN_BF16 && a->out_dtype == PN_F32)
    return launch_simt<__nv_bfloat16, float>(a, stream);
  return launch_simt<float, __nv_bfloat16>(a, stream);
}

size_t pn_sizeof_conv_args(void) { return sizeof(pn_conv_args); }
size_t pn_sizeof_task_args(void) { return sizeof(pn_task_args); }

int pn_conv_pack_weight_bf16(const float* w_f32, int cout, int k, int k_pad, void* w_bf16,
                             pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(w_f32 && w_bf16 && cout > 0 && k > 0 && k_pad >= k);
  k_pack_weight<<<grid_for((long long)cout * k_pad, 256), 256, 0, stream>>>(
      w_f32, cout, k, k_pad, (__nv_bfloat16*)w_bf16);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_cast_f32_to_bf16(const float* in, int in_ld, void* out, int out_ld, int cols,
                        const int* num_rows, int rows_cap, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in && out && cols > 0 && rows_cap >= 0);
  if (rows_cap == 0) return PN_OK;
  k_cast_rows<float, __nv_bfloat16><<<grid_for((long long)rows_cap * cols, 256), 256, 0, stream>>>(
      in, in_ld, (__nv_bfloat16*)out, out_ld, cols, num_rows, rows_cap);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_cast_bf16_to_f32(const void* in, int in_ld, float* out, int out_ld, int cols,
                        const int* num_rows, int rows_cap, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in && out && cols > 0 && rows_cap >= 0);
  if (rows_cap == 0) return PN_OK;
  k_cast_rows<__nv_bfloat16, float><<<grid_for((long long)rows_cap * cols, 256), 256, 0, stream>>>(
      (const __nv_bfloat16*)in, in_ld, out, out_ld, cols, num_rows, rows_cap);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_split_bf16x3(const float* in, int in_ld, void* out, int cols, const int* num_rows, int rows_cap,
                    pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in && out && cols > 0 && rows_cap >= 0);
  if (rows_cap == 0) return PN_OK;
  k_split_bf16x3<<<grid_for((long long)rows_cap * cols, 256), 256, 0, stream>>>(in, in_ld, (__nv_bfloat16*)out, cols,
                                                                                 num_rows, rows_cap);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_sparse_to_dense(const void* feat, int dtype, int feat_ld, const uint32_t* occ_words,
                       const int* word_prefix, int n_frames, int H, int W, int C, void* out,
                       int out_ld, int out_coff, int out_padded, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(feat && occ_words && word_prefix && out && n_frames >= 1 && H > 0 && W > 0 && C > 0);
  const int pad = out_padded ? 1 : 0;
  const long long n_cells = (long long)n_frames * (H + 2 * pad) * (W + 2 * pad);
  if (dtype == PN_F32) {
    PN_REQUIRE(C % 2 == 0 && feat_ld % 2 == 0 && out_ld % 2 == 0 && out_coff % 2 == 0);
    k_sparse_to_dense<float><<<grid_for(n_cells * (C / 2), 256), 256, 0, stream>>>(
        (const float*)feat, feat_ld, occ_words, word_prefix, n_frames, H, W, pad, C, (float*)out, out_ld, out_coff);
  } else if (dtype == PN_BF16) {
    PN_REQUIRE(C % 4 == 0 && feat_ld % 4 == 0 && out_ld % 4 == 0 && out_coff % 4 == 0);
    if (C % 8 == 0 && feat_ld % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(feat) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
      k_sparse_to_dense16<__nv_bfloat16><<<grid_for(n_cells * (C / 8), 256), 256, 0, stream>>>(
          (const __nv_bfloat16*)feat, feat_ld, occ_words, word_prefix, n_frames, H, W, pad, C, (__nv_bfloat16*)out,
          out_ld, out_coff);
      PN_CHECK_LAUNCH();
      return PN_OK;
    }
    k_sparse_to_dense<__nv_bfloat16><<<grid_for(n_cells * (C / 4), 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)feat, feat_ld, occ_words, word_prefix, n_frames, H, W, pad, C, (__nv_bfloat16*)out,
        out_ld, out_coff);
  } else {
    return PN_ERR_INVALID_ARG;
  }
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
