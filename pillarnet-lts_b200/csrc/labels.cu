// CenterPoint training-target assignment on the GPU (SURVEY §8 f rank 1): replaces the CPU AssignLabel stage of
// the reference's data pipeline (det3d/datasets/pipelines/preprocess.py:248-317) and the Gaussian helpers it calls
// (det3d/core/utils/center_utils.py:16-64: gaussian_radius, gaussian2D, draw_umich_gaussian).
//
// One warp per (frame, object slot).  The heat-map is a running maximum of Gaussian patches, so objects can be
// drawn concurrently with an atomic max on the float bit patterns (all values are in [0, 1]).
// Arithmetic follows the reference as NumPy >= 2 evaluates it (NEP 50: fp32 scalars stay fp32 when combined with
// Python scalars); the Gaussian itself is evaluated in double and rounded to fp32, as np.maximum(out=float32) does.
#include "common.cuh"

namespace {

// center_utils.py:16-38, every operation rounded to fp32 in the reference's order
__device__ float gaussian_radius_f32(float height, float width, float min_overlap) {
  const float b1 = __fadd_rn(height, width);
  const float c1 = __fdiv_rn(__fmul_rn(__fmul_rn(width, height), 1.0f - min_overlap), 1.0f + min_overlap);
  const float sq1 = __fsqrt_rn(__fsub_rn(__fmul_rn(b1, b1), __fmul_rn(4.0f, c1)));
  const float r1 = __fdiv_rn(__fadd_rn(b1, sq1), 2.0f);
  const float b2 = __fmul_rn(2.0f, __fadd_rn(height, width));
  const float c2 = __fmul_rn(__fmul_rn(1.0f - min_overlap, width), height);
  const float sq2 = __fsqrt_rn(__fsub_rn(__fmul_rn(b2, b2), __fmul_rn(16.0f, c2)));
  const float r2 = __fdiv_rn(__fadd_rn(b2, sq2), 2.0f);
  const float a3 = __fmul_rn(4.0f, min_overlap);
  const float b3 = __fmul_rn(__fmul_rn(-2.0f, min_overlap), __fadd_rn(height, width));
  const float c3 = __fmul_rn(__fmul_rn(min_overlap - 1.0f, width), height);
  const float sq3 = __fsqrt_rn(__fsub_rn(__fmul_rn(b3, b3), __fmul_rn(__fmul_rn(4.0f, a3), c3)));
  const float r3 = __fdiv_rn(__fadd_rn(b3, sq3), 2.0f);
  return fminf(r1, fminf(r2, r3));
}

struct LabelParams {
  int n_frames, max_objs, num_cls, H, W, box_dim;
  float x0, y0, cell, min_overlap;
  int min_radius[8];
  int n_min_radius;
};

__global__ void __launch_bounds__(128)
k_assign_labels(const __grid_constant__ LabelParams P, const float* __restrict__ gt_boxes,
                const int* __restrict__ gt_cls, float* __restrict__ hm, long long* __restrict__ ind,
                unsigned char* __restrict__ mask, long long* __restrict__ cat, float* __restrict__ anno_box,
                float* __restrict__ gt_box) {
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (frame, object)
  if (slot >= P.n_frames * P.max_objs) return;
  const int lane = threadIdx.x & 31;
  const int frame = slot / P.max_objs;
  const int cls1 = gt_cls[slot];                   // 1-based class id within the task, 0 = empty slot
  if (cls1 <= 0 || cls1 > P.num_cls) return;
  const int cls_id = cls1 - 1;
  const float* b = gt_boxes + (long long)slot * P.box_dim;
  const float w = __fdiv_rn(b[3], P.cell), l = __fdiv_rn(b[4], P.cell);
  if (!(w > 0.f && l > 0.f)) return;
  int radius = (int)gaussian_radius_f32(l, w, P.min_overlap);                      // int(): truncation
  radius = max(P.n_min_radius > 1 ? P.min_radius[cls_id] : P.min_radius[0], radius);
  const float ctx = __fdiv_rn(__fsub_rn(b[0], P.x0), P.cell), cty = __fdiv_rn(__fsub_rn(b[1], P.y0), P.cell);
  const int x = (int)ctx, y = (int)cty;                                            // astype(int32): truncation
  if (!(x >= 0 && x < P.W && y >= 0 && y < P.H)) return;
  if (lane == 0) {
    cat[slot] = cls_id;
    ind[slot] = (long long)y * P.W + x;
    mask[slot] = 1;
    float* g = gt_box + (long long)slot * 7;
    g[0] = b[0]; g[1] = b[1]; g[2] = b[2]; g[3] = b[3]; g[4] = b[4]; g[5] = b[5]; g[6] = b[P.box_dim - 1];
    float* a = anno_box + (long long)slot * 10;
    a[0] = __fsub_rn(ctx, (float)x);
    a[1] = __fsub_rn(cty, (float)y);
    a[2] = b[2];
    a[3] = logf(b[3]); a[4] = logf(b[4]); a[5] = logf(b[5]);
    a[6] = P.box_dim >= 9 ? b[6] : 0.f;
    a[7] = P.box_dim >= 9 ? b[7] : 0.f;
    const float rot = b[P.box_dim - 1];
    a[8] = sinf(rot);
    a[9] = cosf(rot);
  }
  // draw_umich_gaussian: sigma = (2r+1)/6, patch clipped to the map
  const int left = min(x, radius), right = min(P.W - x, radius + 1);
  const int top = min(y, radius), bottom = min(P.H - y, radius + 1);
  const int pw = left + right, ph = top + bottom;
  if (pw <= 0 || ph <= 0) return;
  const double sigma = (double)(2 * radius + 1) / 6.0;
  const double inv = 1.0 / (2.0 * sigma * sigma);
  float* hm_f = hm + (long long)frame * P.H * P.W * P.num_cls;
  for (int i = lane; i < pw * ph; i += 32) {
    const int py = i / pw, px = i - py * pw;
    const double dx = (double)(px - left), dy = (double)(py - top);
    const double e = exp(-(dx * dx + dy * dy) * inv);
    const float v = (float)e;                                  // cast on the fp32 write of np.maximum(out=...)
    if (e < 2.220446049250313e-16) continue;                   // gaussian2D: h[h < eps * h.max()] = 0
    const long long o = ((long long)(y - top + py) * P.W + (x - left + px)) * P.num_cls + cls_id;
    atomicMax(reinterpret_cast<int*>(hm_f) + o, __float_as_int(v));
  }
}

}  // namespace

extern "C" {

int pn_assign_labels(const float* gt_boxes, int box_dim, const int* gt_cls, int n_frames, int max_objs, int num_cls,
                     int H, int W, float x0, float y0, float cell, float gaussian_overlap, const int* min_radius,
                     int n_min_radius, float* hm, long long* ind, unsigned char* mask, long long* cat,
                     float* anno_box, float* gt_box, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(n_frames >= 1 && max_objs >= 1 && num_cls >= 1 && num_cls <= 8 && H > 0 && W > 0);
  PN_REQUIRE(box_dim == 7 || box_dim == 9);
  PN_REQUIRE(gt_boxes && gt_cls && hm && ind && mask && cat && anno_box && gt_box && min_radius);
  PN_REQUIRE(n_min_radius == 1 || (n_min_radius >= num_cls && n_min_radius <= 8));
  const long long slots = (long long)n_frames * max_objs;
  PN_CUDA(cudaMemsetAsync(hm, 0, sizeof(float) * (size_t)n_frames * H * W * num_cls, stream));
  PN_CUDA(cudaMemsetAsync(ind, 0, sizeof(long long) * slots, stream));
  PN_CUDA(cudaMemsetAsync(mask, 0, slots, stream));
  PN_CUDA(cudaMemsetAsync(cat, 0, sizeof(long long) * slots, stream));
  PN_CUDA(cudaMemsetAsync(anno_box, 0, sizeof(float) * slots * 10, stream));
  PN_CUDA(cudaMemsetAsync(gt_box, 0, sizeof(float) * slots * 7, stream));
  LabelParams P;
  P.n_frames = n_frames; P.max_objs = max_objs; P.num_cls = num_cls; P.H = H; P.W = W; P.box_dim = box_dim;
  P.x0 = x0; P.y0 = y0; P.cell = cell; P.min_overlap = gaussian_overlap;
  P.n_min_radius = n_min_radius;
  for (int i = 0; i < 8; ++i) P.min_radius[i] = i < n_min_radius ? min_radius[i] : 0;
  k_assign_labels<<<(unsigned)PN_DIVUP(slots, 4ll), 128, 0, stream>>>(P, gt_boxes, gt_cls, hm, ind, mask, cat, anno_box,
                                                                   gt_box);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
