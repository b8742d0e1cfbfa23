// Rotated BEV IoU with the arithmetic contract of the reference's iou3d_nms extension
// (det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:36-234): the same fp32 formula trees (so nvcc makes
// the same FMA contraction choices), the same predicates (strict sign tests, |s5-s1| > 1e-8 switch,
// MARGIN = 1e-2 corner test, swap-if-greater angular bubble sort) and precise sinf/cosf/atan2f.
// It is *not* an exact polygon clipper; an exact clipper flips ~1 in 3000 overlap decisions
//  which breaks bit-exact keep lists.
//
// Restructured for the GPU: everything that depends on one box only (rotated corners, the inverse
// rotation used by the corner test, margins, area) is computed once per box into BoxGeom instead of
// once per pair, and polygon vertex angles are computed once per vertex instead of O(cnt^2) times in
// the sort comparator.  Each value is produced by the same expression as in the reference, so the
// bits agree.
#pragma once
#include <cuda_runtime.h>

namespace pn_iou {

struct BoxGeom {
  float cx, cy;        // centre
  float px[4], py[4];  // corners after rotation by heading (order: (x1,y1),(x2,y1),(x2,y2),(x1,y2))
  float ic, is;        // cos(-heading), sin(-heading)
  float mx, my;        // dx/2 + MARGIN, dy/2 + MARGIN
  float area;          // dx*dy
};

__device__ __forceinline__ void make_geom(float x, float y, float dx, float dy, float heading,
                                          BoxGeom& g) {
  const float kMargin = 1e-2f;
  const float hx = dx / 2, hy = dy / 2;
  const float x1 = x - hx, y1 = y - hy, x2 = x + hx, y2 = y + hy;
  const float c = cosf(heading), s = sinf(heading);
  const float qx[4] = {x1, x2, x2, x1};
  const float qy[4] = {y1, y1, y2, y2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g.px[k] = (qx[k] - x) * c + (qy[k] - y) * (-s) + x;
    g.py[k] = (qx[k] - x) * s + (qy[k] - y) * c + y;
  }
  g.cx = x;
  g.cy = y;
  g.ic = cosf(-heading);
  g.is = sinf(-heading);
  g.mx = dx / 2 + kMargin;
  g.my = dy / 2 + kMargin;
  g.area = dx * dy;
}

// z-component of (p1-p0) x (p2-p0)
__device__ __forceinline__ float cross3(float p1x, float p1y, float p2x, float p2y, float p0x,
                                        float p0y) {
  return (p1x - p0x) * (p2y - p0y) - (p2x - p0x) * (p1y - p0y);
}

// segment p0->p1 against q0->q1; writes the crossing point, returns 1 when they properly cross
__device__ __forceinline__ int seg_cross(float p1x, float p1y, float p0x, float p0y, float q1x,
                                         float q1y, float q0x, float q0y, float& ox, float& oy) {
  const float kEps = 1e-8f;
  const int boxes_touch = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                          fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
  if (!boxes_touch) return 0;
  const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
  const float s2 = cross3(p1x, p1y, q1x, q1y, p0x, p0y);
  const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
  const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  const float s5 = cross3(q1x, q1y, p1x, p1y, p0x, p0y);
  if (fabsf(s5 - s1) > kEps) {
    ox = (s5 * q0x - s1 * q1x) / (s5 - s1);
    oy = (s5 * q0y - s1 * q1y) / (s5 - s1);
  } else {
    const float a0 = p0y - p1y, b0 = p1x - p0x, c0 = p0x * p1y - p1x * p0y;
    const float a1 = q0y - q1y, b1 = q1x - q0x, c1 = q0x * q1y - q1x * q0y;
    const float D = a0 * b1 - a1 * b0;
    ox = (b0 * c1 - b1 * c0) / D;
    oy = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

__device__ __forceinline__ int corner_inside(const BoxGeom& g, float px, float py) {
  const float rx = (px - g.cx) * g.ic + (py - g.cy) * (-g.is);
  const float ry = (px - g.cx) * g.is + (py - g.cy) * g.ic;
  return fabsf(rx) < g.mx && fabsf(ry) < g.my;
}

// overlap area of a (first/row box) and b (second/column box)
__device__ __forceinline__ float overlap_area(const BoxGeom& a, const BoxGeom& b) {
  float vx[16], vy[16];
  int cnt = 0;
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int i1 = (i + 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int j1 = (j + 1) & 3;
      float ox, oy;
      if (seg_cross(a.px[i1], a.py[i1], a.px[i], a.py[i], b.px[j1], b.py[j1], b.px[j], b.py[j], ox,
                    oy)) {
        if (cnt < 16) { vx[cnt] = ox; vy[cnt] = oy; }
        sx = sx + ox;
        sy = sy + oy;
        ++cnt;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (corner_inside(a, b.px[k], b.py[k])) {
      sx = sx + b.px[k];
      sy = sy + b.py[k];
      if (cnt < 16) { vx[cnt] = b.px[k]; vy[cnt] = b.py[k]; }
      ++cnt;
    }
    if (corner_inside(b, a.px[k], a.py[k])) {
      sx = sx + a.px[k];
      sy = sy + a.py[k];
      if (cnt < 16) { vx[cnt] = a.px[k]; vy[cnt] = a.py[k]; }
      ++cnt;
    }
  }
  if (cnt > 16) cnt = 16;  // unreachable for convex quadrilaterals (<= 8 crossings + 8 corners)
  const float mx = sx / cnt, my = sy / cnt;
  float ang[16];
  for (int k = 0; k < cnt; ++k) ang[k] = atan2f(vy[k] - my, vx[k] - mx);
  for (int j = 0; j < cnt - 1; ++j) {
    for (int i = 0; i < cnt - j - 1; ++i) {
      if (ang[i] > ang[i + 1]) {
        float t;
        t = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = t;
        t = vx[i]; vx[i] = vx[i + 1]; vx[i + 1] = t;
        t = vy[i]; vy[i] = vy[i + 1]; vy[i + 1] = t;
      }
    }
  }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ax = vx[k] - vx[0], ay = vy[k] - vy[0];
    const float bx = vx[k + 1] - vx[0], by = vy[k + 1] - vy[0];
    area += ax * by - ay * bx;
  }
  return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float iou_bev(const BoxGeom& a, const BoxGeom& b) {
  const float kEps = 1e-8f;
  const float ov = overlap_area(a, b);
  return ov / fmaxf(a.area + b.area - ov, kEps);
}

// Conservative disjointness test: true only when the reference arithmetic provably yields
// cnt == 0 => overlap 0 => IoU 0 (never > thr for thr >= 0).  Circumscribed circles plus 0.1 m slack
// (>> MARGIN*sqrt(2) and any fp32 rounding at |coord| < 1e3).
__device__ __forceinline__ bool surely_disjoint(const BoxGeom& a, const BoxGeom& b) {
  const float ra = sqrtf(a.mx * a.mx + a.my * a.my);
  const float rb = sqrtf(b.mx * b.mx + b.my * b.my);
  const float dx = a.cx - b.cx, dy = a.cy - b.cy;
  const float R = ra + rb + 0.1f;
  return dx * dx + dy * dy > R * R * 1.001f;
}

// Sharper conservative test on the same contract: separating axis over the four box axes with every half extent
// taken as half size + MARGIN (g.mx, g.my) and 0.1 m of slack.  Boxes this far apart have no crossing edges and no
// corner inside the other's margin box, so the reference arithmetic yields cnt == 0 => overlap 0.  The circle test
// above lets every pair within ~(ra + rb) through to the exact evaluation — for 4 x 2 m boxes that is 5x the area a
// box really covers, and the exact evaluations were what k_nms_mask spent its time on (28 us at 6 x 1000 boxes).
// NaN / inf geometry fails every comparison => "not disjoint" => exact path.
__device__ __forceinline__ bool surely_disjoint_sat(const BoxGeom& a, const BoxGeom& b) {
  const float kSlack = 0.1f;
  const float dx = b.cx - a.cx, dy = b.cy - a.cy;
  // axes: u = (ic, -is), v = (is, ic) (ic = cos(heading), is = -sin(heading))
  const float c = fabsf(a.ic * b.ic + a.is * b.is);        // |cos(ha - hb)|
  const float s = fabsf(a.ic * b.is - a.is * b.ic);        // |sin(ha - hb)|
  const float da_u = fabsf(dx * a.ic - dy * a.is), da_v = fabsf(dx * a.is + dy * a.ic);
  const float db_u = fabsf(dx * b.ic - dy * b.is), db_v = fabsf(dx * b.is + dy * b.ic);
  const float k = 1.001f;
  return da_u > (a.mx + b.mx * c + b.my * s + kSlack) * k || da_v > (a.my + b.mx * s + b.my * c + kSlack) * k ||
         db_u > (b.mx + a.mx * c + a.my * s + kSlack) * k || db_v > (b.my + a.mx * s + a.my * c + kSlack) * k;
}

}  // namespace pn_iou
