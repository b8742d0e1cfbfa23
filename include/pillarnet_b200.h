/*
 * libpillarnet_b200 — C ABI of the B200-native (sm_100a) PillarNet point->BEV hot path.
 *
 * Drop-in boundary for the reference's native operators (all citations relative to the
 * PillarNet-LTS tree):
 *   det3d/ops/pillar_ops/src/pillar_api.cpp:10-21     (pybind module `pillar_cuda`, 7 wrappers)
 *   det3d/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-19   (pybind module `iou3d_nms_cuda`)
 *   spconv.pytorch SubMConv2d / SparseConv2d          (external dependency, not vendored)
 *   torch.nn.Conv2d / ConvTranspose2d (cuDNN)         (dense BEV neck / head)
 *
 * Conventions
 *   - plain C: POD arguments only (device pointers, ints, floats, a cudaStream_t passed as void*);
 *   - the caller owns all memory (outputs and scratch); sizes come from the pn_*_bytes() helpers;
 *   - every call is asynchronous on `stream`, never synchronises the host, keeps no global state
 *     besides a cached device-property lookup, and is CUDA-graph capturable;
 *   - returns PN_OK (0) or a PN_ERR_* code; pn_last_error() gives a message. Never exit()s
 *     (the reference does: pillar_ops_gpu.cu:53-57, iou3d_nms.cpp:14-25).
 *   - row counts that only the device knows (number of pillars, number of active sites) are passed
 *     as `const int*` device scalars next to a host-side capacity; kernels are sized by the capacity
 *     and exit early past the device count.
 */
#ifndef PILLARNET_B200_H_
#define PILLARNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_OK 0
#define PN_ERR_INVALID_ARG 1
#define PN_ERR_CUDA 2
#define PN_ERR_WORKSPACE 3
#define PN_ERR_UNSUPPORTED 4

typedef void* pn_stream_t; /* cudaStream_t */

/* element types for conv activations / weights */
#define PN_F32 0
#define PN_BF16 1

int pn_abi_version(void);
const char* pn_last_error(void);
/* multiProcessorCount of the current device (148 on B200); <0 on error. */
int pn_device_sm_count(void);
/* number of kernels this library has launched in this process (memsets not counted). */
long long pn_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (1) Dynamic pillarization.
 * Replaces DynamicPFE.forward's coordinate/mask code (models/readers/dynamic_pillar_encoder.py:33-47),
 * pillar_cuda.create_point_pillar_index_stack_wrapper (pillar_ops.cpp:15-36, pillar_ops_gpu.cu:13-39),
 * the torch.cumsum + .item() rank pass (pillar_utils.py:43-45),
 * pillar_cuda.create_pillar_indices_wrapper (pillar_ops.cpp:39-55, pillar_ops_gpu.cu:60-78) and
 * pillar_cuda.gather_indice_wrapper (group_ops_gpu.cu:8-17).
 *
 *   points        (n_points, point_dim) f32, frames concatenated in order; n_points is a host
 *                 CAPACITY, the live count is frame_offsets[n_frames] on the device (so a captured
 *                 CUDA graph serves frames of any size up to the capacity)
 *   frame_offsets (n_frames+1) i32 device; frame b owns rows [off[b], off[b+1])
 *   cell coords   cx = (int)floorf((x - x0) * inv_pillar) with inv_pillar = 1.0f/(float)pillar_size:
 *                 this is what the reference's CUDA expression evaluates (torch scalar division).
 *   occ_words     (ceil(B*H*W/32)) u32, written: bit c set <=> cell c = b*H*W + cy*W + cx occupied
 *   word_prefix   (same length) i32, written: exclusive popcount prefix of occ_words
 *   pillar_coords (m_cap,3) i32 [b, cy, cx], ascending cell order  == reference pillar_indices
 *   point_pillar  (n_points) i32: pillar rank of each point, -1 for out-of-range points.
 *                 point_pillar[point_pillar >= 0] == reference point_pillar_indices (order kept).
 *   num_pillars   (1) i32 device
 * scratch: pn_pillarize_scratch_bytes(n_frames,H,W).
 */
size_t pn_mask_words(int n_frames, int H, int W);
size_t pn_pillarize_scratch_bytes(int n_frames, int H, int W);
int pn_pillarize(const float* points, int point_dim, const int* frame_offsets, int n_points,
                 int n_frames, int H, int W, float x0, float y0, float inv_pillar,
                 uint32_t* occ_words, int* word_prefix, int* pillar_coords, int m_cap,
                 int* point_pillar, int* num_pillars, void* scratch, size_t scratch_bytes,
                 pn_stream_t stream);

/* Multi-sweep accumulation in front of pn_pillarize (SURVEY §8 f rank 2; det3d/datasets/pipelines/loading.py:
 * 37-61,118-141): sweep 0 (the key frame) is kept whole; every other sweep drops the points with |x| < min_distance
 * and |y| < min_distance (sensor frame), is transformed by its 3x4 float64 matrix (rounded once to fp32) and gets
 * its time lag as the extra last column.  Order preserving.
 *   raw (n_raw, in_dim) f32 device, sweeps concatenated; the first n_feat columns are kept (x,y,z,features...)
 *   sweep_offsets (n_sweeps+1) HOST ints, [0] = 0, [n_sweeps] = n_raw; n_sweeps <= 16
 *   transforms (n_sweeps,12) HOST doubles row-major 3x4, a row of NaNs = no transform; time_lag (n_sweeps) HOST
 *   out (out_cap, n_feat+1) f32; rows are written from *out_base (device scalar, NULL = 0) on;
 *   n_total (1) i32 device = *out_base + points kept (clamped to out_cap): chain frames by passing it as the next
 *   call's out_base; the vector of bases/totals is pn_pillarize's frame_offsets. */
size_t pn_merge_sweeps_scratch_bytes(int n_raw);
int pn_merge_sweeps(const float* raw, int in_dim, int n_feat, const int* sweep_offsets, int n_sweeps,
                    const double* transforms, const float* time_lag, float min_distance, const int* out_base,
                    float* out, int out_cap, int* n_total, void* scratch, size_t scratch_bytes, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (2) PFN (Linear -> BatchNorm1d(eval, folded to scale/shift) -> ReLU) fused with scatter-max.
 * Replaces PillarQueryAndGroup's centre/offset features (pillar_utils.py:51-56, gather_feature
 * group_ops_gpu.cu:20-33), PillarMaxPooling.shared_mlps (pillar_modules.py:26-33,71) and
 * pillar_cuda.scatter_max_wrapper (scatter_ops.cpp:7-24, scatter_ops_gpu.cu:13-36).
 *
 *   per point: cx recomputed from x as in pn_pillarize (equals pillar_indices[:,2] of its pillar);
 *              ctr = (float)cx*pillar + offset   (mul, add rounded separately, as the reference);
 *              f = [x-ctr_x, y-ctr_y, p[0..point_dim)]            (2+point_dim values)
 *              h[c] = max(0, dot(weight[c,:], f) * scale[c] + shift[c])
 *   out[m,c] = max over points of pillar m (>= 0 by construction; reference zero-inits out)
 *   weight (c_out, 2+point_dim) f32 row-major (== nn.Linear.weight); scale/shift (c_out) f32 —
 *   these three are HOST pointers (tiny; they are passed to the kernel as launch parameters).
 *   out_f32 (m_cap, c_out) f32 ; out_bf16 optional (may be NULL) same shape, bf16 copy for the
 *   tensor-core backbone.  out_f32 may be NULL when out_bf16 is given: bf16-only fast path (vector
 *   bf16 atomics; bit-identical to rounding the fp32 result because fp32->bf16 rounding is monotone).  arg (m_cap, c_out) i32 optional (NULL for inference): index of a
 *   point attaining the max (lowest point id; the reference's is racy, scatter_ops_gpu.cu:33-35).
 *   c_out must be 32 or 64.  n_points_live: optional device scalar (e.g. frame_offsets + n_frames)
 *   bounding the live points below the host capacity n_points; NULL = all n_points.
 */
int pn_pfn_scatter_max(const float* points, int point_dim, int n_points, const int* n_points_live,
                       const int* point_pillar, const int* num_pillars, int m_cap, float x0, float y0, float inv_pillar,
                       float pillar_size, float x_offset, float y_offset, const float* weight,
                       const float* scale, const float* shift, int c_out, float* out_f32,
                       void* out_bf16, int* arg, pn_stream_t stream);

/* Backward of scatter-max (scatter_ops_gpu.cu:38-45): grad_src[arg[m,c]] = grad_out[m,c]. */
int pn_scatter_max_grad(const float* grad_out, const int* arg, const int* num_pillars, int m_cap,
                        int c_out, float* grad_src, pn_stream_t stream);

/* Training-path reader ops.
 * pn_point_features: out (n_points, 2+point_dim) f32 = [x-ctr_x, y-ctr_y, p[0..point_dim)] with the pillar
 *   centre recomputed from the point's own cell (pillar_utils.py:51-56 without the gather).
 * pn_scatter_max: drop-in for pillar_cuda.scatter_max_wrapper (scatter_ops.cpp:7-24): out (n_pillars,c) f32 =
 *   max(0, max over points with index==m of src) — written in full (zero-init inside), arg (n_pillars,c) i32 =
 *   LOWEST flat index p*c+ch attaining the max, -1 if none (the reference: any point within 1e-5, racy).
 *   index values outside [0,n_pillars) are ignored.  n_points*c must fit int32. */
int pn_point_features(const float* points, int point_dim, int n_points, float x0, float y0, float inv_pillar,
                      float pillar_size, float x_offset, float y_offset, float* out, pn_stream_t stream);
int pn_scatter_max(const float* src, const int* index, int n_points, int n_pillars, int c, float* out, int* arg,
                   pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Rulebooks (replace spconv's indice-pair generation).
 * Output-stationary neighbour tables: nbr[o*9 + ky*3+kx] = input row or -1.
 *  - submanifold 3x3 (SubMConv2d, backbones/base.py:38-52): input site (y+ky-1, x+kx-1), same frame;
 *  - strided 3x3 s2 p1 (SparseConv2d, PillarResNet.py:87,95,103): Ho = (H+2-3)/2+1, output site
 *    active iff any input at (2oy-1+ky, 2ox-1+kx); output rows in ascending b*Ho*Wo+oy*Wo+ox order.
 */
int pn_rulebook_subm3x3(const uint32_t* occ_words, const int* word_prefix, const int* coords,
                        const int* num_rows, int m_cap, int H, int W, int* nbr, pn_stream_t stream);

size_t pn_rulebook_down_scratch_bytes(int n_frames, int H_out, int W_out);
int pn_rulebook_down3x3s2(const uint32_t* in_words, const int* in_prefix, const int* in_coords,
                          const int* in_num_rows, int in_m_cap, int n_frames, int H_in, int W_in,
                          uint32_t* out_words, int* out_prefix, int* out_coords, int* out_num_rows,
                          int out_m_cap, int* nbr, void* scratch, size_t scratch_bytes,
                          pn_stream_t stream);

/* All strided levels of the backbone at once (PillarResNet.py:87,95,103 applied in sequence, plus the submanifold
 * table of each new level, backbones/base.py:38-52 with indice_key res2..res4).  Same results as n_levels calls of
 * pn_rulebook_down3x3s2 + pn_rulebook_subm3x3, in n_levels + 2 launches instead of 5 per level: the occupancy
 * masks are chained first (each depends only on the previous mask), then one launch scans every level and one
 * launch writes every level's two neighbour tables.  Level l (0-based) has H_l = (H_{l-1}+2-3)/2+1.
 * `levels` is a host array; every pointer in it is device memory owned by the caller. */
typedef struct pn_rulebook_level {
  uint32_t* words;      /* out: occupancy words, pn_mask_words(n_frames, H_l, W_l) */
  int* prefix;          /* out: exclusive popcount prefix per word */
  int* coords;          /* out: (m_cap,3) [b,y,x] */
  int* num_rows;        /* out: device scalar, active rows of this level */
  int m_cap;
  int* nbr_down;        /* out: (m_cap,9) rows of the previous level (the strided conv's rulebook) */
  int* nbr_subm;        /* out: (m_cap,9) rows of this level (its submanifold rulebook) */
} pn_rulebook_level;
#define PN_MAX_RULEBOOK_LEVELS 4
size_t pn_rulebook_pyramid_scratch_bytes(int n_frames, int H0, int W0, int n_levels);
int pn_rulebook_pyramid3x3s2(const uint32_t* words0, const int* prefix0, int n_frames, int H0, int W0,
                             int n_levels, const pn_rulebook_level* levels, void* scratch,
                             size_t scratch_bytes, pn_stream_t stream);

/* Static gather tables for the dense BEV convs (NHWC rows = b*H*W + y*W + x), computed once per shape:
 *   mode 0: 3x3 stride s pad 1 (Conv2d / ZeroPad2d+valid conv), taps 9
 *   mode 1: ConvTranspose2d(k=2,s=2): taps 4, exactly one valid tap (dy*2+dx) per output pixel.
 * pad_flags bit 0: input rows index a zero-padded (H+2,W+2) map; bit 1: output rows do (border rows get
 * all -1 taps; the conv then zeroes them, see pn_conv_args.out_hp). */
int pn_dense_nbr_table(int mode, int n_frames, int H_in, int W_in, int stride, int pad_flags, int* nbr,
                       pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (4) Gather-GEMM convolution (sparse SubM / strided sparse / dense 3x3 / transposed 2x2).
 *   out[o, coff + n] = act( (sum_t sum_c in[nbr[o,t], c] * weight[n, t*cin + c]) * scale[n] + shift[n]
 *                           + residual[o, n] )
 * weight layout [cout][taps*cin] == spconv 2.x (Cout,kH,kW,Cin) (torchie/trainer/checkpoint.py:78-87).
 * impl PN_IMPL_SIMT  : fp32 FMA reference path (any dtype combination below);
 * impl PN_IMPL_TCGEN05: bf16 operands, fp32 accumulation in TMEM; weight must be K-padded to a
 *                       multiple of 64 (k_pad) — see pn_conv_pack_weight_bf16.
 */
#define PN_IMPL_SIMT 0
#define PN_IMPL_TCGEN05 1

typedef struct pn_conv_args {
  const void* in;        /* (rows_in, in_ld) channels-last */
  int in_dtype;          /* PN_F32 | PN_BF16 */
  int in_ld;
  const int* nbr;        /* (rows_cap, taps) or NULL => identity with taps == 1 */
  int taps;
  const void* weight;    /* [cout][k_pad] of in_dtype */
  int k_pad;             /* >= taps*cin, row stride of weight in elements */
  const float* scale;    /* (cout) or NULL */
  const float* shift;    /* (cout) or NULL */
  const void* residual;  /* (rows, res_ld) of out_dtype or NULL */
  int res_ld;
  void* out;             /* (rows, out_ld) */
  int out_dtype;
  int out_ld;
  int out_coff;
  int relu;
  const int* num_rows;   /* device row count or NULL => rows_cap */
  int rows_cap;
  int cin;
  int cout;
  int rows_hint;         /* expected live rows (tile-shape heuristic only); 0 => rows_cap */
  int out_hp, out_wp;    /* != 0: output rows are a zero-padded (B,out_hp,out_wp) map; border rows are written as 0 */
  int in_rows;           /* allocated rows of `in` (enables the TMA gather4 activation path); 0 = unknown */
  /* ConvTranspose2d(k=2,s=2) as ONE GEMM (necks/rpn.py:150-154,184-189): deconv_cout != 0 selects it.  `in` rows
   * index the zero-padded (B, deconv_hp_in, deconv_wp_in) input map, taps == 1, nbr == NULL, cout == 4*deconv_cout
   * with weight row (dy*2+dx)*deconv_cout + o, scale/shift repeated per tap; the result of input pixel (py,px)
   * and tap (dy,dx) goes to row (2py-1+dy, 2px-1+dx) of the padded (B,out_hp,out_wp) output map, columns
   * [out_coff, out_coff+deconv_cout).  A quarter of the MMAs and gathers of the 4-tap gather formulation (each
   * output pixel has exactly one valid tap).  PN_IMPL_TCGEN05 only. */
  int deconv_cout, deconv_hp_in, deconv_wp_in;
  /* What the caller knows about `nbr` (a hint: results never depend on it).  PN_NBR_SUBM_SORTED: the 3x3
   * output-stationary table (tap = ky*3+kx) of a site set whose rows are in raster order, as pn_rulebook_subm3x3 /
   * pn_rulebook_pyramid3x3s2 produce for the output of pn_pillarize / a strided level — then the neighbours of 128
   * consecutive outputs under one kernel row are a short contiguous run of input rows, and PN_IMPL_TCGEN05 stages that
   * run once per tile by TMA instead of gathering it once per tap (conv_win_tc.cu). */
  int nbr_kind;
  const void* nbr_plan;  /* pn_conv_window_plan(nbr) or NULL (then the gather kernel runs) */
} pn_conv_args;
#define PN_NBR_ANY 0
#define PN_NBR_SUBM_SORTED 1

/* Tile plans of a 3x3 rulebook for the window-staged kernel: per 128-row output tile and kernel row the start of the
 * input-row window, the destination of every staged row under each kx, and bit masks of absent / out-of-window
 * neighbours.  A function of (nbr, num_rows, rows_cap) only — build it once per rulebook (spconv's `indice_key`) and
 * pass it to every conv that uses that rulebook with the same rows_cap / num_rows.  plan: 16-byte aligned device
 * buffer of pn_conv_window_plan_bytes(rows_cap) bytes. */
size_t pn_conv_window_plan_bytes(int rows_cap);
int pn_conv_window_plan(const int* nbr, const int* num_rows, int rows_cap, void* plan, size_t plan_bytes,
                        pn_stream_t stream);

int pn_conv_gather(const pn_conv_args* args, int impl, pn_stream_t stream);
/* sizeof() of the two argument structs, so a foreign binding can verify its layout. */
size_t pn_sizeof_conv_args(void);
size_t pn_sizeof_task_args(void);

/* Grouped dense 3x3 (pad 1) conv with 1..4 output channels per group, all groups in one launch:
 * the last conv of every CenterHead branch (center_head.py:34-35).  in: bf16 NHWC rows
 * (n_frames*H*W, in_ld); group g reads channels [in_coff, in_coff+cin) and writes f32 columns
 * [out_coff, out_coff+cout) of out (rows, out_ld).  groups: device (n_groups,5) int32
 * {in_coff, cout, w_off, s_off, out_coff}; wbuf: packed f32 [cout][9][cin] weights + [cout] biases.
 * cin % 32 == 0.  fp32 accumulation.  in_padded != 0: `in` rows index the zero-padded (H+2,W+2) map. */
int pn_conv3x3_small_cout(const void* in, int in_ld, int cin, int n_frames, int H, int W, int in_padded,
                          const int* groups, int n_groups, const float* wbuf, float* out, int out_ld,
                          pn_stream_t stream);

/* Dense 3x3 stride-1 pad-1 conv + scale/shift + ReLU on a ZERO-PADDED NHWC layout (tensor cores, TMA).
 * in : bf16 rows (n_frames*(H+2)*(W+2), in_ld), border rows are zeros; channels [in_coff, in_coff+cin)
 * weight: bf16 [cout][k_pad] (tap-major, k_pad >= 9*cin, multiple of 64) as pn_conv_pack_weight_bf16 makes
 * out: padded rows again (borders written as zeros; out_compact = 0) or compact rows
 *      b*H*W + y*W + x (out_compact = 1); bf16 or f32; columns [out_coff, out_coff+cout).
 * out_group_cols = gc > 0: planar output for a following grouped conv — output channels [g*gc, (g+1)*gc) go to
 *      their own contiguous padded map; the maps are stacked as one (cout/gc * n_rows, gc) matrix (out_ld = gc,
 *      out_coff = 0, out_compact = 0).  Each group's activations are then contiguous in memory.
 * cin % 64 == 0.  tile_hint: 0 = automatic tile shape (1..4 force one; testing only). */
int pn_conv_dense3x3(const void* in, int in_ld, int in_coff, int cin, int n_frames, int H, int W,
                     const void* weight, int k_pad, int cout, const float* scale, const float* shift,
                     void* out, int out_dtype, int out_ld, int out_coff, int out_compact, int out_group_cols,
                     int relu, int tile_hint, pn_stream_t stream);

/* Grouped form of pn_conv_dense3x3: n_groups independent 3x3 convs in one launch (the last conv of every
 * CenterHead branch, center_head.py:34-35).  Group g reads channels [in_coff + g*cin, +cin) of the padded
 * input, uses weight rows [16g, 16g+16) of the bf16 [n_groups*16][k_pad] matrix (rows >= its cout are
 * zero) and scale/shift entries [16g, 16g+16), and writes group_tab[g] = {first output column, cout}
 * (device int32 [n_groups][2]) columns of `out`.  cin % 64 == 0, cout <= 16.
 * in_planar != 0: `in` is the planar layout pn_conv_dense3x3 writes with out_group_cols = cin (in_ld = cin,
 * in_coff = 0): group g reads its own contiguous map. */
int pn_conv_dense3x3_grouped(const void* in, int in_ld, int in_coff, int cin, int n_groups, int n_frames, int H,
                             int W, const void* weight, int k_pad, const float* scale, const float* shift,
                             const int* group_tab, void* out, int out_dtype, int out_ld, int out_compact,
                             int relu, int in_planar, pn_stream_t stream);

/* The same grouped last conv (center_head.py:34-35: Conv2d(64, cout <= 3, 3, padding=1, bias=True) of every branch) as
 * ONE 1x1 GEMM per branch plus nine shifted sums: out[q, c] = bias[c] + sum_tap Y[q + off(tap), 3 tap + c] with
 * Y = h_g * W'_g (N = 27 of 32).  Reads the planar intermediate once (HBM bound) where the implicit-GEMM form is bound
 * by N = 16 MMAs.  bf16 tensor-core mode only; the fp32 twin stays pn_conv3x3_small_cout.
 *   in: bf16 [(n_groups * n_pos), 64], n_pos = n_frames*(H+2)*(W+2), borders zero (pn_conv_dense3x3 with
 *       out_group_cols = 64); weight: bf16 [n_groups*32][64], row g*32 + 3*tap + c = W_g[c, :, tap], unused rows zero;
 *   shift: f32 [n_groups*4] biases (NULL: none); group_tab: int32 [n_groups][2] = {first output column, cout <= 3};
 *   out: f32 compact rows (n_frames*H*W, out_ld).
 * PN_ERR_UNSUPPORTED when W + 3 > 256 (a shift must stay within two 128-row tiles). */
int pn_conv_dense3x3_grouped_shift(const void* in, int n_groups, int n_frames, int H, int W, const void* weight,
                                   const float* shift, const int* group_tab, float* out, int out_ld,
                                   pn_stream_t stream);

/* ---- Pillar R-CNN second stage, inference path ------------------------------------------------
 * Rulebook of spconv's SparseConv2d(k = s, stride = s, padding = 0) (lateral layers of
 * det3d/models/second_stage/bev_interpolation.py:66-72): output grid (H_in / s, W_in / s), cell active iff any input of
 * its s x s block is; nbr (out_m_cap, s*s) int32, tap ky*s + kx reads input (s*oy + ky, s*ox + kx) or -1.
 * Same mask / prefix / coords / device-count conventions as pn_rulebook_down3x3s2; scratch =
 * pn_rulebook_down_scratch_bytes(n_frames, H_in / s, W_in / s). */
int pn_rulebook_block(const uint32_t* in_words, const int* in_prefix, int n_frames, int H_in, int W_in, int s,
                      uint32_t* out_words, int* out_prefix, int* out_coords, int* out_num_rows, int out_m_cap,
                      int* nbr, void* scratch, size_t scratch_bytes, pn_stream_t stream);

/* RoI grid points + bilinear interpolation of an NHWC map (bev_interpolation.py:85-123, box_torch_ops.py:159-251,
 * center_utils.py:91-120) in one launch.  rois (n_rois, roi_ld) f32 [x, y, z, dx, dy, dz, ...] with the yaw at column
 * ry_col; RoI r belongs to frame r / rois_per_frame.  feat: rows of a (n_frames, H, W) map (feat_padded: the zero-bordered
 * (H+2, W+2) layout), channels [feat_coff, +C) of rows with stride feat_ld, f32 or bf16; cell = bev_stride * pillar_size.
 * out (n_rois, grid_size^2, C) in feat's dtype; points_out (n_rois, grid_size^2, 2) f32 or NULL.  Point order = the
 * reference's (x index major). */
int pn_roi_grid_bilinear(const float* rois, int roi_ld, int ry_col, int n_rois, int rois_per_frame, int grid_size,
                         const void* feat, int feat_dtype, int feat_ld, int feat_coff, int n_frames, int H, int W,
                         int feat_padded, int C, float x0, float y0, float cell, float* points_out, void* out,
                         pn_stream_t stream);

/* RoI head output -> refined boxes (roi_head_template.py:189-219) and fused scores / validity
 * (detectors/pillar_rcnn.py:141-170): boxes (n, code_size) = rotate_z(reg + [0,0,0, roi dims, yaw, ...], yaw) + centre,
 * scores = sqrt(sigmoid(cls) * roi_score), valid = (label != 0) & (dims > 0).  roi_labels: int64 or NULL. */
int pn_roi_refine(const float* rois, int roi_ld, const float* reg, int code_size, const float* cls,
                  const float* roi_scores, const long long* roi_labels, int n_rois, float* boxes, float* scores,
                  unsigned char* valid, pn_stream_t stream);

/* Training targets on the GPU (SURVEY §8 f rank 1): AssignLabel of the reference's data pipeline
 * (det3d/datasets/pipelines/preprocess.py:248-317) for ONE task: per object the Gaussian radius
 * (center_utils.py:16-38), the heat-map patch (draw_umich_gaussian, :48-64) and the regression targets.
 *   gt_boxes (n_frames, max_objs, box_dim) f32, box_dim 9 [x,y,z,w,l,h,vx,vy,rot] or 7 [x,y,z,w,l,h,rot];
 *   gt_cls   (n_frames, max_objs) i32: class id WITHIN the task, 1-based, 0 = empty slot (slot k = object k);
 *   H, W = task grid (BEV grid // stride); cell = pillar_size*stride as fp32; min_radius: HOST ints, 1 or per class.
 * Outputs (all written in full): hm (n_frames,H,W,num_cls) f32, ind/cat (n_frames,max_objs) i64,
 * mask (n_frames,max_objs) u8, anno_box (n_frames,max_objs,10) f32 [dx,dy,z,log w,log l,log h,vx,vy,sin,cos],
 * gt_box (n_frames,max_objs,7) f32. */
int pn_assign_labels(const float* gt_boxes, int box_dim, const int* gt_cls, int n_frames, int max_objs, int num_cls,
                     int H, int W, float x0, float y0, float cell, float gaussian_overlap, const int* min_radius,
                     int n_min_radius, float* hm, long long* ind, unsigned char* mask, long long* cat,
                     float* anno_box, float* gt_box, pn_stream_t stream);

/* Backward of pn_conv_gather (spconv's autograd for SubMConv2d / SparseConv2d, external in the reference).
 * pn_rulebook_transpose: nbr_t[i*taps + t] = o  <=>  nbr[o*taps + t] = i  (else -1): the input-stationary
 *   table; the data gradient is then a forward gather conv  dx = pn_conv_gather(dy, nbr_t, W^T)  with
 *   W^T[ci][t*cout + co] = W[co][t*cin + ci].  num_out: device count of live output rows (NULL = out_cap).
 * pn_conv_wgrad: dw[co*dw_ld + t*cin + ci] = sum_o dy[o][co] * x[nbr[o,t]][ci]   (f32, overwritten).
 *   nbr NULL = identity (1x1 / linear).  impl PN_IMPL_TCGEN05: bf16 operands, MN-major UMMA over the rows,
 *   fp32 accumulation in TMEM, split-K fp32 RED (summation order not fixed); PN_IMPL_SIMT: f32 or bf16. */
int pn_rulebook_transpose(const int* nbr, const int* num_out, int out_cap, int taps, int in_cap, int* nbr_t,
                          pn_stream_t stream);
int pn_conv_wgrad(const void* x, int x_dtype, int x_ld, const void* dy, int dy_dtype, int dy_ld, const int* nbr,
                  int taps, const int* num_rows, int rows_cap, int cin, int cout, float* dw, int dw_ld, int impl,
                  pn_stream_t stream);

/* f32 -> bf16 weight packing with zero padding of K to k_pad (multiple of 64). */
/* ---- training: batch-statistics BatchNorm over the live rows of a sparse feature matrix (bn_train.cu) -------------
 * Replaces nn.BatchNorm1d(+ residual add + ReLU) in train mode on `.features` (backbones/base.py:155-213) with the row
 * count kept on the device (num_rows; NULL => rows_cap).  dtype = PN_F32 | PN_BF16 for x / y / dy / dx / residual.
 * c must be a power of two <= 256 for the two *_stats entry points. */
/* sums (2c f32, zeroed here): sum x, sum x^2 over rows < *num_rows */
int pn_bn_stats(const void* x, int dtype, int x_ld, const int* num_rows, int rows_cap, int c, float* sums,
                pn_stream_t stream);
/* mean, rstd = 1/sqrt(var_biased + eps), scale = gamma*rstd, shift = beta - mean*scale (all (c) f32); when
 * running_mean/var are given they are updated in place as torch does (momentum, unbiased variance) */
int pn_bn_finalize(const float* sums, const int* num_rows, int rows_cap, int c, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                   float* scale, float* shift, pn_stream_t stream);
/* out[r,:] = act(x[r,:]*scale + shift + residual[r,:]) for r < *num_rows (the other rows of the capacity are
 * neither read nor written: every consumer on the path masks by the same count) */
int pn_bn_apply(const void* x, int dtype, int x_ld, const float* scale, const float* shift, const void* residual,
                int res_ld, int relu, const int* num_rows, int rows_cap, int c, void* out, int out_ld,
                pn_stream_t stream);
/* g = dy * (y > 0 if relu); sums (2c f32, zeroed here): sum g, sum g * xhat  with xhat = (x - mean) * rstd
 * (these are d beta and d gamma) */
int pn_bn_bwd_stats(const void* dy, int dtype, int dy_ld, const void* y, int y_ld, const void* x, int x_ld,
                    const float* mean, const float* rstd, int relu, const int* num_rows, int rows_cap, int c,
                    float* sums, pn_stream_t stream);
/* dx = gamma*rstd*(g - sums[0:c]/n - xhat*sums[c:2c]/n); dres (may be NULL) = g; rows < *num_rows only */
int pn_bn_bwd_apply(const void* dy, int dtype, int dy_ld, const void* y, int y_ld, const void* x, int x_ld,
                    const float* mean, const float* rstd, const float* gamma, const float* sums, int relu,
                    const int* num_rows, int rows_cap, int c, void* dx, int dx_ld, void* dres, int dres_ld,
                    pn_stream_t stream);
/* backward of pn_sparse_to_dense / SparseConvTensor.dense(): out[r,:] = dense[b,y,x,:] at coords[r] = [b,y,x] of a
 * compact (B,H,W,dense_ld) NHWC map; rows < *num_rows only */
int pn_dense_to_sparse(const void* dense, int dtype, int dense_ld, const int* coords, const int* num_rows, int rows_cap,
                       int H, int W, int c, void* out, int out_ld, pn_stream_t stream);

int pn_conv_pack_weight_bf16(const float* w_f32, int cout, int k, int k_pad, void* w_bf16,
                             pn_stream_t stream);
/* rows x cols cast (first *num_rows rows when num_rows != NULL). */
int pn_cast_f32_to_bf16(const float* in, int in_ld, void* out, int out_ld, int cols,
                        const int* num_rows, int rows_cap, pn_stream_t stream);
int pn_cast_bf16_to_f32(const void* in, int in_ld, float* out, int out_ld, int cols,
                        const int* num_rows, int rows_cap, pn_stream_t stream);

/* Split-bf16 ("bf16x3") input for a tensor-core conv at fp32-grade accuracy: out (rows, 3*cols) bf16 =
 * [hi | lo | hi] of in[:, :cols] (f32, row stride in_ld), hi = bf16(x), lo = bf16(x - hi).  With weights laid out
 * [hi_w | hi_w | lo_w] per tap (cin' = 3*cin) pn_conv_gather (PN_IMPL_TCGEN05, f32 output) then accumulates
 * hi*hi + lo*hi + hi*lo in fp32: ~2^-17 relative per product, against 2^-9 of the plain bf16 mode.  Rows at or above
 * *num_rows (when given) are not written. */
int pn_split_bf16x3(const float* in, int in_ld, void* out, int cols, const int* num_rows, int rows_cap,
                    pn_stream_t stream);

/* SparseConvTensor.dense() (spconv) in NHWC: out[(b*H+y)*W+x, coff..coff+C) = feat[rank] or 0.
 * out_padded != 0: rows index the zero-padded map (b*(H+2)+y+1)*(W+2)+x+1, borders written as 0. */
int pn_sparse_to_dense(const void* feat, int dtype, int feat_ld, const uint32_t* occ_words,
                       const int* word_prefix, int n_frames, int H, int W, int C, void* out,
                       int out_ld, int out_coff, int out_padded, pn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5) CenterHead decode + NMS.
 * Replaces CenterHead.predict / post_processing (bbox_heads/center_head.py:216-413),
 * rotate_nms_pcdet / rotate_class_nms_pcdet (core/bbox/box_torch_ops.py:296-359),
 * iou3d_nms_cuda.nms_gpu (ops/iou3d_nms/src/iou3d_nms.cpp:113-159 + iou3d_nms_kernel.cu:104-324)
 * and circle_nms (core/utils/circle_nms_jit.py:4-28).
 *
 * A "segment" is one independent NMS problem: (frame, task) for use_rotate_nms / circular_nms,
 * (frame, task, class) for use_multi_class_nms.  Segment id = frame*segs_per_frame + seg_in_frame.
 */
typedef struct pn_task_args {
  const float* maps;   /* (B*H*W, ld) f32 NHWC: all head outputs of this task packed per pixel */
  int ld;
  int off_reg, off_height, off_dim, off_rot, off_vel, off_iou, off_hm; /* channel offsets, -1 = absent */
  int num_cls;
  int H, W;
  int stride;          /* task stride (tasks[i].stride) */
  int seg_base;        /* first segment (within a frame) owned by this task */
  int per_class;       /* 1 => one segment per class (use_multi_class_nms) */
  int activated;       /* 1 => maps already hold sigmoid(hm), exp(clamp(dim)), clamped iou (pn_double_flip_merge) */
} pn_task_args;

/* Double-flip test-time augmentation (center_head.py:233-248,274-304,319-323): frames 4b..4b+3 of
 * task->maps are the original / y-flipped / x-flipped / xy-flipped views of output frame b.  Each view is
 * un-flipped, activated (sigmoid hm, exp(clamp(dim)), clamped iou), sign-corrected (reg, rot, vel) and the
 * four are averaged in torch.mean's order.  out: (n_frames_out*H*W, out_ld) f32 with the same column
 * offsets; decode it with a copy of `task` that has maps = out, ld = out_ld, activated = 1.
 * n_cols = number of packed columns of the task. */
int pn_double_flip_merge(const pn_task_args* task, int n_frames_out, int n_cols, float* out, int out_ld,
                         pn_stream_t stream);

/* Stage A: per pixel sigmoid/max/threshold/range test; appends (score,pixel) keys to its segment.
 * All `n_tasks` (<= 8) tasks of the head go in one launch.
 *   rectifier: host (n_tasks, 8) floats, per task per class (NULL = all 0)
 *   cand_keys (n_frames*segs_per_frame, cand_cap) u64 ; cand_count (n_frames*segs_per_frame) i32 (zeroed by caller)
 *   key = (bits(rect_score) << 32) | (0xFFFFFFFF - pixel)   => descending key order == score desc, pixel asc
 */
int pn_decode_candidates(const pn_task_args* tasks, int n_tasks, int n_frames, int segs_per_frame,
                         float score_thr, const float* center_range6 /*host, may be NULL*/,
                         float pillar_size, float x0, float y0, const float* rectifier /*host*/,
                         unsigned long long* cand_keys, int cand_cap, int* cand_count,
                         pn_stream_t stream);

/* Stage B: per segment top-`pre_max` selection + sort, decode of the surviving boxes (one CTA per
 * segment, every segment of every task and frame in one launch).
 *   seg_pre_max (segs_per_frame) host ints.
 *   sorted_boxes (n_segs, pre_cap, 12) f32: [x,y,z,w,l,h,vx,vy,rot, score, rect_score, label]
 *   sorted_count (n_segs) i32
 */
int pn_select_topk(const pn_task_args* tasks, int n_tasks, int n_frames, int segs_per_frame,
                   const int* seg_pre_max /*host*/, float pillar_size, float x0, float y0,
                   const float* rectifier /*host (n_tasks,8)*/,
                   const unsigned long long* cand_keys, int cand_cap, const int* cand_count,
                   float* sorted_boxes, int pre_cap, int* sorted_count, pn_stream_t stream);

/* Stage C+D: suppression matrix (upper triangle) + greedy sweep on device.
 *   mode 0: rotated BEV IoU > thr (iou3d_nms_kernel.cu:104-234 arithmetic incl. MARGIN/EPS)
 *   mode 1: circle: dist^2 <= thr (circle_nms_jit.py:23-26)
 *   seg_thr / seg_post_max: (segs_per_frame) host arrays.
 *   mask scratch: pn_nms_scratch_bytes(n_segs, pre_cap)
 *   keep_idx (n_segs, post_cap) i32: positions in sorted order; keep_count (n_segs) i32
 *   det_out  (n_segs, post_cap, 11) f32: [box9, score, label] of kept boxes (vx,vy = 0 when absent)
 */
size_t pn_nms_scratch_bytes(int n_segs, int pre_cap);
int pn_nms(int mode, int n_frames, int segs_per_frame, const float* seg_thr /*host*/,
           const int* seg_post_max /*host*/, const int* seg_use_rectified /*host or NULL*/,
           const float* sorted_boxes, int pre_cap, const int* sorted_count, void* scratch,
           size_t scratch_bytes, int* keep_idx, int post_cap, int* keep_count, float* det_out,
           pn_stream_t stream);

/* Standalone drop-ins for iou3d_nms_cuda (boxes (n,7) f32 [x,y,z,dx,dy,dz,heading], device). */
int pn_boxes_iou_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* iou,
                     pn_stream_t stream);
/* drop-in for iou3d_nms_cuda.boxes_aligned_overlap_bev_gpu (iou3d_nms.cpp, kernel iou3d_nms_kernel.cu:251-262):
 * overlap[i] = BEV intersection area of boxes_a[i] and boxes_b[i] (training: IouLoss target). */
int pn_boxes_aligned_overlap_bev(const float* boxes_a, const float* boxes_b, int n, float* overlap,
                                 pn_stream_t stream);
/* keep (n) i32 device, num_keep (1) i32 device; boxes must already be score-sorted (as nms_gpu). */
int pn_nms_rotated(const float* boxes, int n, float thr, void* scratch, size_t scratch_bytes,
                   int* keep, int* num_keep, pn_stream_t stream);
/* drop-in for iou3d_nms_cuda.nms_normal_gpu (iou3d_nms.cpp:162-207, kernel iou3d_nms_kernel.cu:325-380): the same
 * greedy sweep over the axis-aligned IoU of [x, y, ., dx, dy] (heading ignored).  Same scratch as pn_nms_rotated. */
int pn_nms_normal(const float* boxes, int n, float thr, void* scratch, size_t scratch_bytes,
                  int* keep, int* num_keep, pn_stream_t stream);
/* drop-in for iou3d_nms_cuda.boxes_overlap_bev_gpu (iou3d_nms.cpp:66-88, kernel iou3d_nms_kernel.cu:236-249):
 * overlap (na, nb) f32 = BEV intersection AREA of every pair (boxes_iou3d_gpu builds the 3-D IoU from it). */
int pn_boxes_overlap_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* overlap,
                         pn_stream_t stream);

/* ---- operator-level drop-ins for the reference's `pillar_cuda` pybind module --------------------------------------
 * (det3d/ops/pillar_ops/src/pillar_api.cpp:10-21; headers pillar_ops_gpu.h:7-16, group_ops_gpu.h:6-12).  Not used by
 * the fused product path (pn_pillarize / pn_pfn_scatter_max); they let det3d's own pillar_utils.py / group_utils.py
 * run unmodified on this library through pillarnet_lts_b200.compat.pillar_cuda.
 * scatter_max_wrapper / scatter_max_grad_wrapper map to pn_scatter_max / pn_scatter_max_grad above. */
/* create_point_pillar_index_stack_wrapper (pillar_ops.cpp:15-36, pillar_ops_gpu.cu:13-39): pts_xy (n,2) i32 [x,y] cell
 * coordinates, pts_batch_cnt (B) i32 points per frame; sets pillars_mask[(b*H+y)*W+x] = 1 (bool/u8, caller-zeroed)
 * and point_pillar_index[i] = that cell id for in-range points (others keep the caller's initial value, -1). */
int pn_compat_point_pillar_index(const int* pts_xy, const int* pts_batch_cnt, int n_points, int n_frames, int H, int W,
                                 unsigned char* pillars_mask, int* point_pillar_index, pn_stream_t stream);
/* create_pillar_indices_wrapper (pillar_ops.cpp:39-55, pillar_ops_gpu.cu:60-78): pillars_position (B,H,W) i32 = rank of
 * an occupied cell or < 0; writes pillar_indices[rank] = [b, y, x]. */
int pn_compat_pillar_indices(const int* pillars_position, int n_frames, int H, int W, int* pillar_indices,
                             pn_stream_t stream);
/* gather_indice_wrapper (group_ops_gpu.cu:8-17): outs[i] = indices[index[i]]. */
int pn_compat_gather_indice(const int* index, const int* indices, int n, int* outs, pn_stream_t stream);
/* gather_feature_wrapper / gather_feature_grad_wrapper (group_ops_gpu.cu:20-48): outs[i,:] = features[index[i],:];
 * grad_features[index[i],:] += grad_outs[i,:] (grad_features caller-zeroed). */
int pn_compat_gather_feature(const int* index, const float* features, int n, int c, float* outs, pn_stream_t stream);
int pn_compat_gather_feature_grad(const int* index, const float* grad_outs, int n, int c, float* grad_features,
                                  pn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PILLARNET_B200_H_ */
