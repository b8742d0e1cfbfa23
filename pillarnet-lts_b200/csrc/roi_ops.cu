// Second stage (Pillar R-CNN), inference path: RoI grid points + bilinear interpolation of the fused BEV map, and the
// box refinement of the RoI head.
//
// Replaces det3d/models/second_stage/bev_interpolation.py:85-123 (`get_pooling_points` = center_to_grid_box2d,
// det3d/core/bbox/box_torch_ops.py:220-251 with rotation_2d :159-172; `interpolate_from_bev_features` =
// bilinear_interpolate_torch, det3d/core/utils/center_utils.py:91-120 — one indexed gather per corner and frame, four
// (N, C) temporaries) and det3d/models/roi_heads/roi_head_template.py:189-219 (`generate_predicted_boxes`) +
// det3d/models/detectors/pillar_rcnn.py:141-170 (`post_process`: score fusion and validity mask).
#include "common.cuh"

namespace {

__device__ __forceinline__ float ld_feat(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_feat(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// One warp per (roi, grid point); lanes stride over the channels, so each of the four corner reads is one coalesced row
// segment.  Arithmetic follows the reference operation by operation in fp32:
//   grid offset  g = (i + 0.5) / G * dim - dim / 2       (i = x index = point / G, j = y index = point % G)
//   rotation     x' = gx cos + gy sin, y' = -gx sin + gy cos, then + centre
//   map coords   u = (x' - x0) / cell, v = (y' - y0) / cell, corners floor / floor + 1 clamped to the map,
//   weights from the CLAMPED corners (as the reference: a point outside the map gets weights that do not sum to 1).
template <typename T>
__global__ void __launch_bounds__(256)
k_roi_grid_bilinear(const float* __restrict__ rois, int roi_ld, int ry_col, int n_rois, int rois_per_frame, int G,
                    const T* __restrict__ feat, int feat_ld, int feat_coff, int H, int W, int pad, int C,
                    float x0, float y0, float cell, float* __restrict__ points_out, T* __restrict__ out) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int P = G * G;
  if (warp >= n_rois * P) return;
  const int r = warp / P, p = warp - r * P;
  const int i = p / G, j = p - i * G;
  const float* roi = rois + (long long)r * roi_ld;
  const float cx = roi[0], cy = roi[1], dx = roi[3], dy = roi[4], ang = roi[ry_col];
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn((float)i + 0.5f, (float)G), dx), __fdiv_rn(dx, 2.f));
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn((float)j + 0.5f, (float)G), dy), __fdiv_rn(dy, 2.f));
  const float s = sinf(ang), c = cosf(ang);
  const float px = __fadd_rn(__fadd_rn(__fmul_rn(gx, c), __fmul_rn(gy, s)), cx);
  const float py = __fadd_rn(__fadd_rn(__fmul_rn(gx, -s), __fmul_rn(gy, c)), cy);
  if (points_out && lane == 0) {
    points_out[2ll * warp] = px;
    points_out[2ll * warp + 1] = py;
  }
  const float u = __fdiv_rn(__fsub_rn(px, x0), cell), v = __fdiv_rn(__fsub_rn(py, y0), cell);
  // floor -> int64 in the reference; clamp before the int conversion so far-away RoIs cannot overflow
  const float fu = floorf(u), fv = floorf(v);
  const int xa = (int)fminf(fmaxf(fu, -2.f), (float)W + 1.f), ya = (int)fminf(fmaxf(fv, -2.f), (float)H + 1.f);
  const int x0i = min(max(xa, 0), W - 1), x1i = min(max(xa + 1, 0), W - 1);
  const int y0i = min(max(ya, 0), H - 1), y1i = min(max(ya + 1, 0), H - 1);
  const float wx1 = __fsub_rn((float)x1i, u), wx0 = __fsub_rn(u, (float)x0i);
  const float wy1 = __fsub_rn((float)y1i, v), wy0 = __fsub_rn(v, (float)y0i);
  const float wa = __fmul_rn(wx1, wy1), wb = __fmul_rn(wx1, wy0), wc = __fmul_rn(wx0, wy1), wd = __fmul_rn(wx0, wy0);
  const int b = r / rois_per_frame;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long base = (long long)b * Hp * Wp;
  const T* fa = feat + (base + (long long)(y0i + pad) * Wp + x0i + pad) * feat_ld + feat_coff;   // Ia = im[y0, x0]
  const T* fb = feat + (base + (long long)(y1i + pad) * Wp + x0i + pad) * feat_ld + feat_coff;   // Ib = im[y1, x0]
  const T* fc = feat + (base + (long long)(y0i + pad) * Wp + x1i + pad) * feat_ld + feat_coff;   // Ic = im[y0, x1]
  const T* fd = feat + (base + (long long)(y1i + pad) * Wp + x1i + pad) * feat_ld + feat_coff;   // Id = im[y1, x1]
  T* o = out + (long long)warp * C;
  for (int ch = lane; ch < C; ch += 32) {
    // ((Ia*wa + Ib*wb) + Ic*wc) + Id*wd, each product rounded: the reference's four temporaries and three adds
    float acc = __fmul_rn(ld_feat(fa + ch), wa);
    acc = __fadd_rn(acc, __fmul_rn(ld_feat(fb + ch), wb));
    acc = __fadd_rn(acc, __fmul_rn(ld_feat(fc + ch), wc));
    acc = __fadd_rn(acc, __fmul_rn(ld_feat(fd + ch), wd));
    if constexpr (sizeof(T) == 4) o[ch] = acc; else o[ch] = __float2bfloat16_rn(acc);
  }
}

// rois (n, roi_ld) [x, y, z, dx, dy, dz, ry (+ extras)], reg (n, code) residuals, cls (n) logits.
// boxes = rotate_z(reg + [0, 0, 0, dx, dy, dz, ry, ...], ry)[:3] + centre; scores = sqrt(sigmoid(cls) * roi_score);
// valid = label != 0 and every refined dim > 0.
__global__ void __launch_bounds__(256)
k_roi_refine(const float* __restrict__ rois, int roi_ld, const float* __restrict__ reg, int code,
             const float* __restrict__ cls, const float* __restrict__ roi_scores, const long long* __restrict__ roi_labels,
             int n, float* __restrict__ boxes, float* __restrict__ scores, uint8_t* __restrict__ valid) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* roi = rois + (long long)r * roi_ld;
  const float* d = reg + (long long)r * code;
  float* o = boxes + (long long)r * code;
  const float ry = roi[6];
  const float s = sinf(ry), c = cosf(ry);
  const float x = d[0], y = d[1], z = d[2];               // local_rois[:, 0:3] = 0
  // points[:, :, 0:3] @ [[c, -s, 0], [s, c, 0], [0, 0, 1]]
  o[0] = __fadd_rn(__fadd_rn(__fmul_rn(x, c), __fmul_rn(y, s)), roi[0]);
  o[1] = __fadd_rn(__fadd_rn(__fmul_rn(x, -s), __fmul_rn(y, c)), roi[1]);
  o[2] = __fadd_rn(z, roi[2]);
  bool ok = roi_labels ? roi_labels[r] != 0 : true;
  for (int k = 3; k < code; ++k) {
    o[k] = __fadd_rn(d[k], roi[k]);
    if (k < 6) ok = ok && o[k] > 0.f;
  }
  const float sg = 1.f / (1.f + expf(-cls[r]));
  scores[r] = sqrtf(__fmul_rn(sg, roi_scores[r]));
  valid[r] = ok ? 1 : 0;
}

inline int grid_for(long long work, int threads) {
  const long long g = (work + threads - 1) / threads;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" {

int pn_roi_grid_bilinear(const float* rois, int roi_ld, int ry_col, int n_rois, int rois_per_frame, int grid_size,
                         const void* feat, int feat_dtype, int feat_ld, int feat_coff, int n_frames, int H, int W,
                         int feat_padded, int C, float x0, float y0, float cell, float* points_out, void* out,
                         pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(rois && feat && out && n_rois >= 0 && rois_per_frame >= 1 && grid_size >= 1 && roi_ld >= 5);
  PN_REQUIRE(ry_col >= 5 && ry_col < roi_ld && n_frames >= 1 && H > 0 && W > 0 && C > 0 && cell > 0.f);
  PN_REQUIRE(feat_dtype == PN_F32 || feat_dtype == PN_BF16);
  PN_REQUIRE((long long)n_rois <= (long long)n_frames * rois_per_frame);
  if (n_rois == 0) return PN_OK;
  const long long warps = (long long)n_rois * grid_size * grid_size;
  PN_REQUIRE(warps * 32 < (1ll << 40));
  const int pad = feat_padded ? 1 : 0;
  if (feat_dtype == PN_F32)
    k_roi_grid_bilinear<float><<<grid_for(warps * 32, 256), 256, 0, stream>>>(
        rois, roi_ld, ry_col, n_rois, rois_per_frame, grid_size, (const float*)feat, feat_ld, feat_coff, H, W, pad, C,
        x0, y0, cell, points_out, (float*)out);
  else
    k_roi_grid_bilinear<__nv_bfloat16><<<grid_for(warps * 32, 256), 256, 0, stream>>>(
        rois, roi_ld, ry_col, n_rois, rois_per_frame, grid_size, (const __nv_bfloat16*)feat, feat_ld, feat_coff, H, W,
        pad, C, x0, y0, cell, points_out, (__nv_bfloat16*)out);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

int pn_roi_refine(const float* rois, int roi_ld, const float* reg, int code_size, const float* cls,
                  const float* roi_scores, const long long* roi_labels, int n_rois, float* boxes, float* scores,
                  unsigned char* valid, pn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PN_REQUIRE(rois && reg && cls && roi_scores && boxes && scores && valid && n_rois >= 0);
  PN_REQUIRE(code_size >= 7 && roi_ld >= code_size);
  if (n_rois == 0) return PN_OK;
  k_roi_refine<<<grid_for(n_rois, 256), 256, 0, stream>>>(rois, roi_ld, reg, code_size, cls, roi_scores, roi_labels, n_rois,
                                                          boxes, scores, valid);
  PN_CHECK_LAUNCH();
  return PN_OK;
}

}  // extern "C"
