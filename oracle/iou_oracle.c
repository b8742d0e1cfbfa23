/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle, never part of the product path.
 *
 * Plain-C restatement of the reference's rotated BEV IoU and greedy NMS, used by tests/ and by
 * bench.py's cpu_baseline leg to check the CUDA kernels.
 *
 * Follows (PillarNet-LTS):
 *   det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:36-61    cross / check_rect_cross / check_in_box2d
 *   det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:63-92    intersection
 *   det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:104-225  box_overlap
 *   det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:227-234  iou_bev
 *   det3d/ops/iou3d_nms/src/iou3d_cpu.cpp:128-273        (the reference's own CPU twin, same shape)
 *   det3d/ops/iou3d_nms/src/iou3d_nms.cpp:139-156        greedy sweep over the suppression matrix
 *   det3d/core/utils/circle_nms_jit.py:4-28              circle_nms
 *
 * Pinned against the reference's compiled CPU twin (oracle/_ref, boxes_iou_bev_cpu) in
 * tests/test_oracle_pins.py and against committed golden vectors in tests/golden/.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_EPS 1e-8f
#define ORACLE_MARGIN 1e-2f

typedef struct { float x, y; } pt;

static float cross2(pt a, pt b) { return a.x * b.y - a.y * b.x; }

static float cross3(pt p1, pt p2, pt p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

static int rect_cross(pt p1, pt p2, pt q1, pt q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

static int in_box(const float* box, pt p) {
  float cx = box[0], cy = box[1];
  float c = cosf(-box[6]), s = sinf(-box[6]);
  float rx = (p.x - cx) * c + (p.y - cy) * (-s);
  float ry = (p.x - cx) * s + (p.y - cy) * c;
  return fabsf(rx) < box[3] / 2 + ORACLE_MARGIN && fabsf(ry) < box[4] / 2 + ORACLE_MARGIN;
}

static int seg_intersection(pt p1, pt p0, pt q1, pt q0, pt* ans) {
  if (!rect_cross(p0, p1, q0, q1)) return 0;
  float s1 = cross3(q0, p1, p0);
  float s2 = cross3(p1, q1, p0);
  float s3 = cross3(p0, q1, q0);
  float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > ORACLE_EPS) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    float D = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / D;
    ans->y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

static void rotate(pt ctr, float c, float s, pt* p) {
  float nx = (p->x - ctr.x) * c + (p->y - ctr.y) * (-s) + ctr.x;
  float ny = (p->x - ctr.x) * s + (p->y - ctr.y) * c + ctr.y;
  p->x = nx;
  p->y = ny;
}

float oracle_box_overlap(const float* a, const float* b) {
  float ahx = a[3] / 2, bhx = b[3] / 2, ahy = a[4] / 2, bhy = b[4] / 2;
  pt ca = {a[0], a[1]}, cb = {b[0], b[1]};
  pt A[5] = {{a[0] - ahx, a[1] - ahy}, {a[0] + ahx, a[1] - ahy}, {a[0] + ahx, a[1] + ahy}, {a[0] - ahx, a[1] + ahy}};
  pt B[5] = {{b[0] - bhx, b[1] - bhy}, {b[0] + bhx, b[1] - bhy}, {b[0] + bhx, b[1] + bhy}, {b[0] - bhx, b[1] + bhy}};
  float ac = cosf(a[6]), as = sinf(a[6]), bc = cosf(b[6]), bs = sinf(b[6]);
  for (int k = 0; k < 4; ++k) {
    rotate(ca, ac, as, &A[k]);
    rotate(cb, bc, bs, &B[k]);
  }
  A[4] = A[0];
  B[4] = B[0];
  pt poly[24];
  pt ctr = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      pt x;
      if (seg_intersection(A[i + 1], A[i], B[j + 1], B[j], &x)) {
        poly[cnt] = x;
        ctr.x = ctr.x + x.x;
        ctr.y = ctr.y + x.y;
        ++cnt;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, B[k])) { ctr.x = ctr.x + B[k].x; ctr.y = ctr.y + B[k].y; poly[cnt++] = B[k]; }
    if (in_box(b, A[k])) { ctr.x = ctr.x + A[k].x; ctr.y = ctr.y + A[k].y; poly[cnt++] = A[k]; }
  }
  ctr.x /= cnt;
  ctr.y /= cnt;
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i) {
      float ai = atan2f(poly[i].y - ctr.y, poly[i].x - ctr.x);
      float bi = atan2f(poly[i + 1].y - ctr.y, poly[i + 1].x - ctr.x);
      if (ai > bi) { pt t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t; }
    }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    pt u = {poly[k].x - poly[0].x, poly[k].y - poly[0].y};
    pt v = {poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
    area += cross2(u, v);
  }
  return fabsf(area) / 2.0f;
}

float oracle_iou_bev(const float* a, const float* b) {
  float sa = a[3] * a[4], sb = b[3] * b[4];
  float ov = oracle_box_overlap(a, b);
  return ov / fmaxf(sa + sb - ov, ORACLE_EPS);
}

/* pairwise IoU, (na,nb) row-major */
void oracle_boxes_iou_bev(const float* A, int na, const float* B, int nb, float* out) {
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) out[(long)i * nb + j] = oracle_iou_bev(A + 7 * i, B + 7 * j);
}

/* greedy NMS over score-sorted boxes: box i suppresses later j iff iou(i,j) > thr. returns #kept */
int oracle_nms_rotated(const float* boxes, int n, float thr, int64_t* keep) {
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[kept++] = i;
    for (int j = i + 1; j < n; ++j)
      if (!removed[j] && oracle_iou_bev(boxes + 7 * i, boxes + 7 * j) > thr) removed[j] = 1;
  }
  free(removed);
  return kept;
}

/* circle NMS on score-sorted centres: suppress later j iff dist^2 <= thresh (float32 difference,
 * float64 square/sum, as numba evaluates float32**int). returns #kept */
int oracle_nms_circle(const float* xy, int n, double thresh, int64_t* keep) {
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[kept++] = i;
    for (int j = i + 1; j < n; ++j) {
      if (removed[j]) continue;
      double dx = (double)(float)(xy[2 * i] - xy[2 * j]);
      double dy = (double)(float)(xy[2 * i + 1] - xy[2 * j + 1]);
      if (dx * dx + dy * dy <= thresh) removed[j] = 1;
    }
  }
  free(removed);
  return kept;
}
