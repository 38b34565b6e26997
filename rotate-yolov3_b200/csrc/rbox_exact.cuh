// Pinned-arithmetic rotated-rectangle overlap, bit-compatible with the reference NMS kernel.
//
// What this mirrors: the device functions of the reference
//   utils/nms/src/rotate_polygon_nms_kernel.cu:22-260 (trangle_area, area, reorder_pts, inter2line,
//   in_rect, inter_pts, convert_region, inter, devRotateIoU) AS COMPILED by nvcc with its default
//   -fmad=true for rotate_nms_kernel (:262-308).  The reference source was compiled for sm_100a
//   (oracle/build_ref.sh) and its PTX + SASS read; every floating-point operation below is written
//   with a round-to-nearest intrinsic (__fmaf_rn/__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/
//   __fsqrt_rn), which neither NVVM nor ptxas may contract or re-associate, in exactly the shape
//   the reference build has:
//     * corners   x = cx + fma(c, lx, -(s*ly)),  y = cy + fma(s, lx, c*ly)        (:196-229)
//     * dot       a.b = fma(ax, bx, ay*by)                                         (:134-160)
//     * cross     p x q = fma(px, qy, -(py*qx)), then * 0.5 (the `/ 2.0`)          (:22-24)
//     * area_cdb = (area_abc + area_cda) - area_abd                                (:119)
//     * t = area_cda / (area_abd - area_abc); pt = fma(b - a, t, a)                (:124-129)
//     * key = vx/|v| (or -2 - vx/|v| when vy/|v| < 0), |v| = sqrt(fma(vx,vx,vy*vy)) (:52-66)
//     * union = fma(w2, h2, w1*h1) - inter  (row area hoisted out of the column loop, column
//       area fused -- this is the kernel's form, not the textbook one)             (:251-258)
//   Not a copy: the evaluation is restructured (per-box corner/edge invariants precomputed once
//   by a prep kernel instead of 2 sincos per pair; the 32 corner-difference vectors shared by
//   the 48 triangle areas), but each value is produced by the same operation on the same
//   operands, so results are bitwise those of the reference kernel.
#pragma once
#include <cuda_runtime.h>

namespace ryolo {

// Per-box invariants, computed once per (score-sorted) box.  24 floats = 96 B.
struct BoxGeom {
  float px[4], py[4];  // corners P0..P3 in the reference's storage order (convert_region :226-227)
  float abx, aby;      // P1 - P0
  float adx, ady;      // P3 - P0
  float abab, adad;    // |ab|^2, |ad|^2 as fma(x,x,y*y)
  float w, h, area;    // area = w*h (single rounding), used when this box is the ROW box
  // ---- conservative pre-filter data (NOT part of the pinned arithmetic) ----
  float cx, cy, rad;   // padded bounding-circle radius; +inf when the box is ill-conditioned
  float ux, uy;        // (cos, sin) of theta
  float hw, hh;        // padded half extents; +inf when ill-conditioned
};
static_assert(sizeof(BoxGeom) == 24 * 4, "BoxGeom layout");

constexpr int kGeomFields = 24;

// convert_region (:196-229) with the contraction pattern of the reference build.
__device__ __forceinline__ void exact_corners(float cx, float cy, float w, float h, float angle,
                                              float* px, float* py, float* c_out, float* s_out) {
  const float c = cosf(angle);  // same libdevice code as the reference build (no fast-math)
  const float s = sinf(angle);
  const float lx[4] = {__fmul_rn(w, -0.5f), __fmul_rn(w, 0.5f), __fmul_rn(w, 0.5f), __fmul_rn(w, -0.5f)};
  const float ly[4] = {__fmul_rn(h, -0.5f), __fmul_rn(h, -0.5f), __fmul_rn(h, 0.5f), __fmul_rn(h, 0.5f)};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float x = __fadd_rn(cx, __fmaf_rn(c, lx[i], -__fmul_rn(s, ly[i])));
    const float y = __fadd_rn(cy, __fmaf_rn(s, lx[i], __fmul_rn(c, ly[i])));
    px[3 - i] = x;  // pts[7 - 2i - 1]
    py[3 - i] = y;  // pts[7 - 2i]
  }
  *c_out = c;
  *s_out = s;
}

// in_rect (:134-160): point (x,y) against a rect given by P0, ab, ad, |ab|^2, |ad|^2.
__device__ __forceinline__ bool exact_in_rect(float x, float y, float p0x, float p0y, float abx, float aby,
                                              float adx, float ady, float abab, float adad) {
  const float apx = __fsub_rn(x, p0x);
  const float apy = __fsub_rn(y, p0y);
  const float abap = __fmaf_rn(abx, apx, __fmul_rn(aby, apy));
  const float adap = __fmaf_rn(adx, apx, __fmul_rn(ady, apy));
  return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}

// trangle_area numerator form: ((a-c) x (b-c)) / 2 given the two difference vectors.
__device__ __forceinline__ float exact_half_cross(float px, float py, float qx, float qy) {
  return __fmul_rn(__fmaf_rn(px, qy, -__fmul_rn(py, qx)), 0.5f);
}

// Scratch view: one thread's candidate points/keys, strided by the block size so that
// consecutive threads hit consecutive banks.  slot s of thread t lives at base[s * stride + t].
struct PtScratch {
  float* x;
  float* y;
  float* k;
  int stride;
  __device__ __forceinline__ float& X(int i) { return x[i * stride]; }
  __device__ __forceinline__ float& Y(int i) { return y[i * stride]; }
  __device__ __forceinline__ float& K(int i) { return k[i * stride]; }
};

constexpr int kMaxPts = 8;  // the reference buffer int_pts[16] holds 8 points (:235); more is UB there

// Intersection area of rect1 (row box) and rect2 (column box): inter() (:231-249).
// g1/g2 are read through the accessor F(field, which) so that callers can keep them in shared memory.
template <typename G1, typename G2>
__device__ __forceinline__ float exact_inter_area(const G1& g1, const G2& g2, PtScratch sc) {
  float p1x[4], p1y[4], p2x[4], p2y[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    p1x[i] = g1.px(i);
    p1y[i] = g1.py(i);
    p2x[i] = g2.px(i);
    p2y[i] = g2.py(i);
  }
  int n = 0;
  // inter_pts (:162-194): corners first, interleaved box1/box2
  {
    const float a1x = g1.abx(), a1y = g1.aby(), d1x = g1.adx(), d1y = g1.ady(), ab1 = g1.abab(), ad1 = g1.adad();
    const float a2x = g2.abx(), a2y = g2.aby(), d2x = g2.adx(), d2y = g2.ady(), ab2 = g2.abab(), ad2 = g2.adad();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (exact_in_rect(p1x[i], p1y[i], p2x[0], p2y[0], a2x, a2y, d2x, d2y, ab2, ad2)) {
        if (n < kMaxPts) { sc.X(n) = p1x[i]; sc.Y(n) = p1y[i]; }
        n++;
      }
      if (exact_in_rect(p2x[i], p2y[i], p1x[0], p1y[0], a1x, a1y, d1x, d1y, ab1, ad1)) {
        if (n < kMaxPts) { sc.X(n) = p2x[i]; sc.Y(n) = p2y[i]; }
        n++;
      }
    }
  }
  // inter2line (:90-132) for the 16 edge pairs.  D[i][j] = P1[i] - P2[j].
  {
    float dx[4][4], dy[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        dx[i][j] = __fsub_rn(p1x[i], p2x[j]);
        dy[i][j] = __fsub_rn(p1y[i], p2y[j]);
      }
    // abc[i][j] = trangle_area(P1[i], P1[i+1], P2[j]);  area_abd(i,j) == abc[i][j+1]
    float abc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int i1 = (i + 1) & 3;
        abc[i][j] = exact_half_cross(dx[i][j], dy[i][j], dx[i1][j], dy[i1][j]);
      }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int i1 = (i + 1) & 3;
      const float ex = __fsub_rn(p1x[i1], p1x[i]);  // b - a
      const float ey = __fsub_rn(p1y[i1], p1y[i]);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int j1 = (j + 1) & 3;
        const float area_abc = abc[i][j];
        const float area_abd = abc[i][j1];
        if (__fmul_rn(area_abc, area_abd) >= 0.f) continue;
        // trangle_area(c, d, a) = ((c-a) x (d-a))/2 == ((a-c) x (a-d))/2 bitwise (both factors negated)
        const float area_cda = exact_half_cross(dx[i][j], dy[i][j], dx[i][j1], dy[i][j1]);
        const float area_cdb = __fsub_rn(__fadd_rn(area_abc, area_cda), area_abd);
        if (__fmul_rn(area_cda, area_cdb) >= 0.f) continue;
        const float t = __fdiv_rn(area_cda, __fsub_rn(area_abd, area_abc));
        if (n < kMaxPts) {
          sc.X(n) = __fmaf_rn(ex, t, p1x[i]);
          sc.Y(n) = __fmaf_rn(ey, t, p1y[i]);
        }
        n++;
      }
    }
  }
  if (n > kMaxPts) n = kMaxPts;
  if (n < 3) return 0.f;  // area() loop body never runs (:28)

  // reorder_pts (:35-89)
  float sx = 0.f, sy = 0.f;
  for (int i = 0; i < n; i++) {
    sx = __fadd_rn(sx, sc.X(i));
    sy = __fadd_rn(sy, sc.Y(i));
  }
  const float fn = (float)n;
  const float ctrx = __fdiv_rn(sx, fn);
  const float ctry = __fdiv_rn(sy, fn);
  for (int i = 0; i < n; i++) {
    const float vx = __fsub_rn(sc.X(i), ctrx);
    const float vy = __fsub_rn(sc.Y(i), ctry);
    const float d = __fsqrt_rn(__fmaf_rn(vx, vx, __fmul_rn(vy, vy)));
    float kx = __fdiv_rn(vx, d);
    const float ky = __fdiv_rn(vy, d);
    if (ky < 0.f) kx = __fsub_rn(-2.f, kx);
    sc.K(i) = kx;
  }
  for (int i = 1; i < n; ++i) {
    if (sc.K(i - 1) > sc.K(i)) {
      const float temp = sc.K(i), tx = sc.X(i), ty = sc.Y(i);
      int j = i;
      while (j > 0 && sc.K(j - 1) > temp) {
        sc.K(j) = sc.K(j - 1);
        sc.X(j) = sc.X(j - 1);
        sc.Y(j) = sc.Y(j - 1);
        j--;
      }
      sc.K(j) = temp;
      sc.X(j) = tx;
      sc.Y(j) = ty;
    }
  }
  // area (:26-33): fan around point 0
  float ar = 0.f;
  const float ax = sc.X(0), ay = sc.Y(0);
  float bx = sc.X(1), by = sc.Y(1);
  for (int i = 0; i < n - 2; i++) {
    const float cx = sc.X(i + 2), cy = sc.Y(i + 2);
    const float tri = exact_half_cross(__fsub_rn(ax, cx), __fsub_rn(ay, cy), __fsub_rn(bx, cx), __fsub_rn(by, cy));
    ar = __fadd_rn(ar, fabsf(tri));
    bx = cx;
    by = cy;
  }
  return ar;
}

// devRotateIoU (:251-260) in the form rotate_nms_kernel compiles to: row area hoisted, column area fused.
__device__ __forceinline__ float exact_iou_from_inter(float area_row, float w_col, float h_col, float inter) {
  const float uni = __fsub_rn(__fmaf_rn(w_col, h_col, area_row), inter);
  return __fdiv_rn(inter, uni);
}

}  // namespace ryolo
