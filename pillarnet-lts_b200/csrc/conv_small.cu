// Dense 3x3 (pad 1, stride 1) convolution with very few output channels, grouped: the last layer of
// every CenterHead branch (nn.Conv2d(64, c, 3, padding=1) with c in {1,2,3}, bbox_heads/center_head.py:34-35),
// all branches of all tasks in ONE launch.
//
// Why not the tensor-core gather-GEMM: with N <= 3 the MMA is idle and a gather formulation re-reads
// each 64-channel input pixel 9 times from L2 (measured 26 us per branch, 36 branches per frame).
// Here a CTA stages a 32x16 pixel tile plus halo in shared memory once (bf16, transposed to
// channel-pair-major so that threads reading neighbouring pixels hit distinct banks), every thread
// produces 4 horizontally adjacent pixels so both the staged inputs and the broadcast weights are
// reused from registers, and accumulation is fp32.
#include "common.cuh"

namespace {

constexpr int TW = 32, TH = 16;           // output tile
constexpr int HW_ = TW + 2, HH_ = TH + 2; // halo tile 34 x 18
constexpr int RS = 37;                    // padded halo row stride (== 1 mod 4: conflict-free, see below)
constexpr int CHUNK = 32;                 // input channels staged per pass
constexpr int PAIRS = CHUNK / 2;
constexpr int THREADS = 128;              // 16 rows x 8 column quads

struct GroupDesc {
  int in_coff;   // first input channel of the group
  int cout;      // 1..4
  int w_off;     // float offset of [cout][9][cin] weights in the packed buffer
  int s_off;     // float offset of [cout] shift (bias) values
  int out_coff;  // first output column
};

// weights of one (tap, channel pair) are WV floats: [j][2] for j < COUT, zero padded to a multiple of 4 so
// that they are fetched with broadcast LDS.128 (COUT <= 2: one, COUT 3..4: two) instead of 2*COUT LDS.32
template <int COUT> struct WVec { static constexpr int value = COUT <= 2 ? 4 : 8; };

template <int COUT>
__device__ __forceinline__ void compute_chunk(const uint32_t* __restrict__ s_in, const float* __restrict__ s_w,
                                              int ty, int tx4, float (&acc)[4][COUT]) {
  constexpr int WV = WVec<COUT>::value;
  // s_in[pair][row*RS + col] (bf16x2), s_w[tap][pair][WV]
#pragma unroll 2
  for (int p = 0; p < PAIRS; ++p) {
    const uint32_t* base = s_in + p * (HH_ * RS) + ty * RS + tx4;
    float2 v[3][6];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const uint32_t u = base[dy * RS + c];
        v[dy][c] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
      }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float w[WV];
        const float4* wp = reinterpret_cast<const float4*>(s_w + ((dy * 3 + dx) * PAIRS + p) * WV);
#pragma unroll
        for (int i = 0; i < WV / 4; ++i) {
          const float4 t = wp[i];
          w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][j] = fmaf(v[dy][q + dx].x, w[2 * j], acc[q][j]);
            acc[q][j] = fmaf(v[dy][q + dx].y, w[2 * j + 1], acc[q][j]);
          }
        }
      }
  }
}

template <int COUT>
__device__ __forceinline__ void run_group(const __nv_bfloat16* __restrict__ in, int in_ld, int cin, int H, int W, int pad,
                                          const GroupDesc& g, const float* __restrict__ wbuf,
                                          float* __restrict__ out, int out_ld, uint32_t* s_in, float* s_w,
                                          int b, int y0, int x0) {
  const int tid = threadIdx.x;
  const int ty = tid >> 3, tx4 = (tid & 7) * 4;
  float acc[4][COUT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[q][j] = 0.f;
  for (int c0 = 0; c0 < cin; c0 += CHUNK) {
    __syncthreads();
    // stage halo: half-warp per pixel, lane -> channel pair (64 contiguous bytes per pixel), written
    // transposed.  4-byte cp.async with zero-fill keeps the WHOLE halo in flight at once; staging
    // through registers (8 loads per thread at a time) left the kernel L2-latency bound (measured).
    const int half = tid >> 4, pr = tid & 15;
    constexpr int kPix = HH_ * HW_, kStep = THREADS / 16;
    const uint32_t s_in_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_in));
    {
      // incremental (hy,hx): kStep = 8 pixels per iteration, no div/mod in the loop
      int hy = 0, hx = half;
      const int Wp = W + 2 * pad;
      const __nv_bfloat16* in_g = in + g.in_coff + c0;
      const long long frame_base = (long long)b * (H + 2 * pad) * Wp;
      for (int px = half; px < kPix; px += kStep) {
        const int gy = y0 + hy - 1, gx = x0 + hx - 1;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const __nv_bfloat16* p = in_g + (ok ? (frame_base + (long long)(gy + pad) * Wp + gx + pad) * in_ld : 0);
        const uint32_t dst = s_in_addr + 4u * (uint32_t)(pr * (HH_ * RS) + hy * RS + hx);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst),
                     "l"(reinterpret_cast<const uint32_t*>(p) + pr), "r"(ok ? 4u : 0u)
                     : "memory");
        hx += kStep;
        if (hx >= HW_) { hx -= HW_; ++hy; }
      }
    }
    // stage weights of this channel chunk: s_w[tap][pair][WV] = W[j][tap][c0 + 2*pair + {0,1}], zero padded
    constexpr int WV = WVec<COUT>::value;
    for (int i = tid; i < 9 * PAIRS * WV; i += THREADS) {
      const int e = i % WV, pp = (i / WV) % PAIRS, tap = i / (WV * PAIRS);
      const int j = e >> 1;
      s_w[i] = j < COUT ? __ldg(wbuf + g.w_off + (j * 9 + tap) * cin + c0 + 2 * pp + (e & 1)) : 0.f;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    compute_chunk<COUT>(s_in, s_w, ty, tx4, acc);
  }
  const int gy = y0 + ty;
  if (gy < H) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gx = x0 + tx4 + q;
      if (gx < W) {
        float* o = out + ((long long)(b * H + gy) * W + gx) * out_ld + g.out_coff;
#pragma unroll
        for (int j = 0; j < COUT; ++j) o[j] = acc[q][j] + __ldg(wbuf + g.s_off + j);
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS)
k_conv3x3_small(const __nv_bfloat16* __restrict__ in, int in_ld, int cin, int H, int W, int pad, int tiles_x,
                int tiles_y, const GroupDesc* __restrict__ groups, const float* __restrict__ wbuf,
                float* __restrict__ out, int out_ld) {
  __shared__ uint32_t s_in[PAIRS * HH_ * RS];
  __shared__ __align__(16) float s_w[9 * PAIRS * 8];
  const GroupDesc g = groups[blockIdx.y];
  const int t = blockIdx.x;
  const int b = t / (tiles_x * tiles_y);
  const int r = t - b * tiles_x * tiles_y;
  const int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW;
  switch (g.cout) {
    case 1: run_group<1>(in, in_ld, cin, H, W, pad, g, wbuf, out, out_ld, s_in, s_w, b, y0, x0); break;
    case 2: run_group<2>(in, in_ld, cin, H, W, pad, g, wbuf, out, out_ld, s_in, s_w, b, y0, x0); break;
    case 3: run_group<3>(in, in_ld, cin, H, W, pad, g, wbuf, out, out_ld, s_in, s_w, b, y0, x0); break;
    default: run_group<4>(in, in_ld, cin, H, W, pad, g, wbuf, out, out_ld, s_in, s_w, b, y0, x0); break;
  }
}

}  // namespace

extern "C" {

// groups: device array of n_groups x 5 int32 {in_coff, cout(1..4), w_off, s_off, out_coff};
// wbuf: packed f32 weights ([cout][9][cin] per group) and shifts.  in: bf16 NHWC rows, out: f32 rows.
int pn_conv3x3_small_cout(const void* in, int in_ld, int cin, int n_frames, int H, int W, int in_padded,
                          const int* groups, int n_groups, const float* wbuf, float* out, int out_ld,
                          pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(in && groups && wbuf && out && n_groups >= 1 && n_groups <= 65535);
  PN_REQUIRE(cin % CHUNK == 0 && in_ld % 2 == 0 && n_frames >= 1 && H > 0 && W > 0);
  PN_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3u) == 0);
  const int tiles_x = PN_DIVUP(W, TW), tiles_y = PN_DIVUP(H, TH);
  dim3 grid(n_frames * tiles_x * tiles_y, n_groups);
  k_conv3x3_small<<<grid, THREADS, 0, stream>>>((const __nv_bfloat16*)in, in_ld, cin, H, W, in_padded ? 1 : 0, tiles_x, tiles_y,
                                                reinterpret_cast<const GroupDesc*>(groups), wbuf, out, out_ld);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
