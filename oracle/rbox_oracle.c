/*
 * oracle/rbox_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference algorithms on the rotated-box hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (libryolo.so and the Python package) never does.
 *
 * Parity pinning (see tests/test_oracle.py):
 *   - orc_ref_iou is checked BIT-FOR-BIT against the reference's own device code compiled as host C++
 *     (oracle/_ref/libref_rnms_host.so, built by oracle/build_ref.sh from /root/reference) on random and
 *     adversarial pairs, and against the committed golden vectors generated from that build
 *     (tests/golden/rnms_ref_golden.npz, script tests/golden/make_golden.py);
 *   - the only fixture the reference itself ships for this path, the 4 boxes of
 *     utils/nms/nms_wrapper_test.py:35-38, is checked against the analytic answer keep=[0,3];
 *   - orc_skew_iou restates skew_bbox_iou, whose arithmetic lives in shapely/GEOS (not vendored, no pinned
 *     version in the reference, absent from this image): PARITY UNPINNED at the shapely boundary.  It is
 *     pinned by us with analytic cases and cross-checked against cv2.rotatedRectangleIntersection; the
 *     corner convention is pinned against the reference's own get_rotated_coors run here (golden fixture).
 *
 * All citations are relative to the reference repository root.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_FMA
/* shape of the reference kernel as nvcc -fmad=true compiles it for sm_100a (read from its PTX/SASS) */
#define DOT2(ax, bx, ay, by) fmaf((ax), (bx), (ay) * (by))
#define CROSS2(px, qy, py, qx) fmaf((px), (qy), -((py) * (qx)))
#define LERP(e, t, a) fmaf((e), (t), (a))
#define SUFFIX(name) name##_fma
#else
/* the source as written, one rounding per operation (gcc -ffp-contract=off) */
#define DOT2(ax, bx, ay, by) ((ax) * (bx) + (ay) * (by))
#define CROSS2(px, qy, py, qx) ((px) * (qy) - (py) * (qx))
#define LERP(e, t, a) ((a) + (t) * (e))
#define SUFFIX(name) name
#endif

/* utils/nms/src/rotate_polygon_nms_kernel.cu:22-24 */
static float trangle_area_(const float* a, const float* b, const float* c) {
  return (float)(CROSS2(a[0] - c[0], b[1] - c[1], a[1] - c[1], b[0] - c[0]) / 2.0);
}

/* :26-33 */
static float poly_area_(const float* pts, int n) {
  float area = 0.0f;
  for (int i = 0; i < n - 2; i++) area += fabsf(trangle_area_(pts, pts + 2 * i + 2, pts + 2 * i + 4));
  return area;
}

/* :35-89 centroid, pseudo-angle key, insertion sort */
static void reorder_pts_(float* pts, int n) {
  if (n <= 0) return;
  float center[2] = {0.0f, 0.0f};
  for (int i = 0; i < n; i++) {
    center[0] += pts[2 * i];
    center[1] += pts[2 * i + 1];
  }
  center[0] /= n;
  center[1] /= n;
  float vs[16];
  for (int i = 0; i < n; i++) {
    float v0 = pts[2 * i] - center[0];
    float v1 = pts[2 * i + 1] - center[1];
    float d = sqrtf(DOT2(v0, v0, v1, v1));
    v0 = v0 / d;
    v1 = v1 / d;
    if (v1 < 0) v0 = -2 - v0;
    vs[i] = v0;
  }
  for (int i = 1; i < n; ++i) {
    if (vs[i - 1] > vs[i]) {
      float temp = vs[i], tx = pts[2 * i], ty = pts[2 * i + 1];
      int j = i;
      while (j > 0 && vs[j - 1] > temp) {
        vs[j] = vs[j - 1];
        pts[j * 2] = pts[j * 2 - 2];
        pts[j * 2 + 1] = pts[j * 2 - 1];
        j--;
      }
      vs[j] = temp;
      pts[j * 2] = tx;
      pts[j * 2 + 1] = ty;
    }
  }
}

/* :90-132 proper segment crossing via signed triangle areas */
static int inter2line_(const float* p1, const float* p2, int i, int j, float* out) {
  const float* a = p1 + 2 * i;
  const float* b = p1 + 2 * ((i + 1) % 4);
  const float* c = p2 + 2 * j;
  const float* d = p2 + 2 * ((j + 1) % 4);
  float area_abc = trangle_area_(a, b, c);
  float area_abd = trangle_area_(a, b, d);
  if (area_abc * area_abd >= 0) return 0;
  float area_cda = trangle_area_(c, d, a);
  float area_cdb = area_cda + area_abc - area_abd;
  if (area_cda * area_cdb >= 0) return 0;
  float t = area_cda / (area_abd - area_abc);
  out[0] = LERP(b[0] - a[0], t, a[0]);
  out[1] = LERP(b[1] - a[1], t, a[1]);
  return 1;
}

/* :134-160 */
static int in_rect_(float x, float y, const float* pts) {
  float ab0 = pts[2] - pts[0], ab1 = pts[3] - pts[1];
  float ad0 = pts[6] - pts[0], ad1 = pts[7] - pts[1];
  float ap0 = x - pts[0], ap1 = y - pts[1];
  float abab = DOT2(ab0, ab0, ab1, ab1);
  float abap = DOT2(ab0, ap0, ab1, ap1);
  float adad = DOT2(ad0, ad0, ad1, ad1);
  float adap = DOT2(ad0, ap0, ad1, ap1);
  return abab >= abap && abap >= 0 && adad >= adap && adap >= 0;
}

int SUFFIX(orc_last_npts) = 0; /* candidate count of the last inter_pts_ call, BEFORE the clamp (diagnostics) */

/* :162-194 ; the reference buffer holds 8 points and is not bounds-checked -- we stop storing at 8 */
static int inter_pts_(const float* p1, const float* p2, float* out) {
  int n = 0;
  for (int i = 0; i < 4; i++) {
    if (in_rect_(p1[2 * i], p1[2 * i + 1], p2)) {
      if (n < 8) { out[2 * n] = p1[2 * i]; out[2 * n + 1] = p1[2 * i + 1]; }
      n++;
    }
    if (in_rect_(p2[2 * i], p2[2 * i + 1], p1)) {
      if (n < 8) { out[2 * n] = p2[2 * i]; out[2 * n + 1] = p2[2 * i + 1]; }
      n++;
    }
  }
  float tmp[2];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (inter2line_(p1, p2, i, j, tmp)) {
        if (n < 8) { out[2 * n] = tmp[0]; out[2 * n + 1] = tmp[1]; }
        n++;
      }
  SUFFIX(orc_last_npts) = n;
  return n > 8 ? 8 : n;
}

/* :196-229 corners stored reversed */
static void convert_region_(float* pts, const float* region) {
  float angle = region[4];
  float a_cos = cosf(angle), a_sin = sinf(angle);
  float ctr_x = region[0], ctr_y = region[1], w = region[2], h = region[3];
  float px[4] = {-w / 2, w / 2, w / 2, -w / 2};
  float py[4] = {-h / 2, -h / 2, h / 2, h / 2};
  for (int i = 0; i < 4; i++) {
#ifdef ORC_FMA
    pts[7 - 2 * i - 1] = ctr_x + fmaf(a_cos, px[i], -(a_sin * py[i]));
    pts[7 - 2 * i] = ctr_y + fmaf(a_sin, px[i], a_cos * py[i]);
#else
    pts[7 - 2 * i - 1] = a_cos * px[i] - a_sin * py[i] + ctr_x;
    pts[7 - 2 * i] = a_sin * px[i] + a_cos * py[i] + ctr_y;
#endif
  }
}

/* :231-249 */
float SUFFIX(orc_ref_inter)(const float* r1, const float* r2) {
  float p1[8], p2[8], ip[16];
  convert_region_(p1, r1);
  convert_region_(p2, r2);
  int n = inter_pts_(p1, p2, ip);
  reorder_pts_(ip, n);
  return poly_area_(ip, n);
}

/* :251-260 ; r1 = row (higher score) box, r2 = column box */
float SUFFIX(orc_ref_iou)(const float* r1, const float* r2) {
  float area_inter = SUFFIX(orc_ref_inter)(r1, r2);
#ifdef ORC_FMA
  /* rotate_nms_kernel as compiled: row area hoisted, column area fused into the sum */
  float area1 = r1[2] * r1[3];
  return area_inter / (fmaf(r2[2], r2[3], area1) - area_inter);
#else
  float area1 = r1[2] * r1[3];
  float area2 = r2[2] * r2[3];
  return area_inter / (area1 + area2 - area_inter);
#endif
}

void SUFFIX(orc_ref_iou_paired)(const float* a, const float* b, int n, int stride, float* out) {
  for (int i = 0; i < n; i++) out[i] = SUFFIX(orc_ref_iou)(a + (size_t)i * stride, b + (size_t)i * stride);
}

/* Upper-triangle mask as consumed by the reference host scan (:262-308 with :371-374): boxes are the
 * score-sorted [n,6]; mask [n, ceil(n/64)] (fully written; lower-triangle words are zero). */
void SUFFIX(orc_mask)(const float* boxes, int n, float thr, unsigned long long* mask) {
  const int cb = (n + 63) / 64;
  memset(mask, 0, sizeof(unsigned long long) * (size_t)n * cb);
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++)
      if (SUFFIX(orc_ref_iou)(boxes + (size_t)i * 6, boxes + (size_t)j * 6) > thr)
        mask[(size_t)i * cb + j / 64] |= 1ULL << (j % 64);
}

#ifndef ORC_FMA
/* greedy scan :358-376 ; returns count, keep = sorted-order positions */
int orc_scan(const unsigned long long* mask, int n, int64_t* keep) {
  const int cb = (n + 63) / 64;
  unsigned long long* remv = (unsigned long long*)calloc((size_t)cb, sizeof(unsigned long long));
  int k = 0;
  for (int i = 0; i < n; i++) {
    int nb = i / 64, ib = i % 64;
    if (!(remv[nb] & (1ULL << ib))) {
      keep[k++] = i;
      const unsigned long long* p = mask + (size_t)i * cb;
      for (int j = nb; j < cb; j++) remv[j] |= p[j];
    }
  }
  free(remv);
  return k;
}

typedef struct { float score; int64_t idx; } orc_si;
static int cmp_desc_(const void* x, const void* y) {
  const orc_si* a = (const orc_si*)x; const orc_si* b = (const orc_si*)y;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx); /* stable: original order on ties */
}
static int cmp_i64_(const void* x, const void* y) {
  int64_t a = *(const int64_t*)x, b = *(const int64_t*)y;
  return (a > b) - (a < b);
}

/* stable descending score order (reference :326-328 uses an unstable torch sort; tests are tie-free) */
void orc_sort_order(const float* dets, int n, int64_t* order) {
  orc_si* s = (orc_si*)malloc(sizeof(orc_si) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) { s[i].score = dets[(size_t)i * 6 + 5]; s[i].idx = i; }
  qsort(s, (size_t)n, sizeof(orc_si), cmp_desc_);
  for (int i = 0; i < n; i++) order[i] = s[i].idx;
  free(s);
}

/* whole r_nms :323-384.  variant 0 = source arithmetic, 1 = nvcc-contracted arithmetic.
 * keep_out receives ascending ORIGINAL indices; returns K. */
void orc_mask_fma(const float* boxes, int n, float thr, unsigned long long* mask);
int orc_rnms(const float* dets, int n, float thr, int64_t* keep_out, int variant) {
  if (n <= 0) return 0;
  const int cb = (n + 63) / 64;
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  float* sorted = (float*)malloc(sizeof(float) * 6 * (size_t)n);
  unsigned long long* mask = (unsigned long long*)malloc(sizeof(unsigned long long) * (size_t)n * cb);
  int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  orc_sort_order(dets, n, order);
  for (int i = 0; i < n; i++) memcpy(sorted + (size_t)i * 6, dets + (size_t)order[i] * 6, 6 * sizeof(float));
  if (variant) orc_mask_fma(sorted, n, thr, mask); else orc_mask(sorted, n, thr, mask);
  int k = orc_scan(mask, n, keep);
  for (int i = 0; i < k; i++) keep_out[i] = order[keep[i]];
  qsort(keep_out, (size_t)k, sizeof(int64_t), cmp_i64_);
  free(order); free(sorted); free(mask); free(keep);
  return k;
}

/* ------------------------------------------------------------------------------------------------
 * skew_bbox_iou restatement in float64 (utils/utils.py:290-320 -> get_rotated_coors :702-725 ->
 * skewiou :663-699).  shapely semantics: convex hulls, intersection area, zero/invalid guards.
 * ------------------------------------------------------------------------------------------------ */
/* :702-725 -- corner order (xmin,ymin),(xmin,ymax),(xmax,ymax),(xmax,ymin), rotated about the centre
 * by cv2.getRotationMatrix2D(angle=-a*180/pi): x' = cos a (x-cx) - sin a (y-cy) + cx, y' = sin a (x-cx) + cos a (y-cy) + cy */
void orc_rotated_coors(const double* box, double* out8) {
  double cx = box[0], cy = box[1], w = box[2], h = box[3], a = box[4];
  double xmin = cx - w * 0.5, xmax = cx + w * 0.5, ymin = cy - h * 0.5, ymax = cy + h * 0.5;
  double tx[4] = {xmin, xmin, xmax, xmax}, ty[4] = {ymin, ymax, ymax, ymin};
  double c = cos(a), s = sin(a);
  for (int i = 0; i < 4; i++) {
    out8[2 * i] = c * (tx[i] - cx) - s * (ty[i] - cy) + cx;
    out8[2 * i + 1] = s * (tx[i] - cx) + c * (ty[i] - cy) + cy;
  }
}

static double shoelace_(const double* p, int n) {
  double a = 0.0;
  for (int i = 0; i < n; i++) {
    int j = (i + 1) % n;
    a += p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
  }
  return 0.5 * a;
}

/* clip convex polygon subj against the half-plane left of edge (ax,ay)->(bx,by) (CCW clip polygon) */
static int clip_edge_(const double* subj, int n, double ax, double ay, double bx, double by, double* out) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const double* P = subj + 2 * i;
    const double* Q = subj + 2 * ((i + 1) % n);
    double dp = (bx - ax) * (P[1] - ay) - (by - ay) * (P[0] - ax);
    double dq = (bx - ax) * (Q[1] - ay) - (by - ay) * (Q[0] - ax);
    if (dp >= 0) { out[2 * m] = P[0]; out[2 * m + 1] = P[1]; m++; }
    if ((dp >= 0) != (dq >= 0)) {
      double t = dp / (dp - dq);
      out[2 * m] = P[0] + t * (Q[0] - P[0]);
      out[2 * m + 1] = P[1] + t * (Q[1] - P[1]);
      m++;
    }
  }
  return m;
}

/* mode 0 'iou', 1 'giou' (= inter / axis-aligned envelope area, :682-685) */
double orc_skew_iou(const double* box1, const double* box2, int mode) {
  double a[8], b[8];
  orc_rotated_coors(box1, a);
  orc_rotated_coors(box2, b);
  for (int i = 0; i < 8; i++) if (!isfinite(a[i]) || !isfinite(b[i])) return 0.0; /* invalid polygon :669-671 */
  double area_a = shoelace_(a, 4), area_b = shoelace_(b, 4);
  if (area_a < 0) { /* make CCW */
    for (int i = 0; i < 2; i++) { double t0 = a[2 * i], t1 = a[2 * i + 1]; a[2 * i] = a[2 * (3 - i)]; a[2 * i + 1] = a[2 * (3 - i) + 1]; a[2 * (3 - i)] = t0; a[2 * (3 - i) + 1] = t1; }
    area_a = -area_a;
  }
  if (area_b < 0) {
    for (int i = 0; i < 2; i++) { double t0 = b[2 * i], t1 = b[2 * i + 1]; b[2 * i] = b[2 * (3 - i)]; b[2 * i + 1] = b[2 * (3 - i) + 1]; b[2 * (3 - i)] = t0; b[2 * (3 - i) + 1] = t1; }
    area_b = -area_b;
  }
  if (area_a == 0.0 || area_b == 0.0) return 0.0; /* :672-673 */
  double buf1[32], buf2[32];
  memcpy(buf1, a, sizeof(a));
  int n = 4;
  for (int e = 0; e < 4 && n > 0; e++) {
    n = clip_edge_(buf1, n, b[2 * e], b[2 * e + 1], b[2 * ((e + 1) % 4)], b[2 * ((e + 1) % 4) + 1], buf2);
    memcpy(buf1, buf2, sizeof(double) * 2 * (size_t)n);
  }
  double inter = n >= 3 ? fabs(shoelace_(buf1, n)) : 0.0;
  double uni;
  if (mode == 1) {
    double x0 = a[0], x1 = a[0], y0 = a[1], y1 = a[1];
    for (int i = 0; i < 4; i++) {
      if (a[2 * i] < x0) x0 = a[2 * i]; if (a[2 * i] > x1) x1 = a[2 * i];
      if (b[2 * i] < x0) x0 = b[2 * i]; if (b[2 * i] > x1) x1 = b[2 * i];
      if (a[2 * i + 1] < y0) y0 = a[2 * i + 1]; if (a[2 * i + 1] > y1) y1 = a[2 * i + 1];
      if (b[2 * i + 1] < y0) y0 = b[2 * i + 1]; if (b[2 * i + 1] > y1) y1 = b[2 * i + 1];
    }
    uni = (x1 - x0) * (y1 - y0);
  } else {
    uni = area_a + area_b - inter;
  }
  if (uni == 0.0) return 0.0; /* :693-694 */
  return inter / uni;
}

/* fp32 boxes in (stride floats per row), float64 arithmetic inside, like the reference (fp32 tensors ->
 * python floats -> GEOS doubles -> FloatTensor) */
void orc_skew_iou_paired(const float* a, const float* b, int n, int sa, int sb, int mode, float* out) {
  for (int i = 0; i < n; i++) {
    double x[5], y[5];
    for (int k = 0; k < 5; k++) { x[k] = a[(size_t)i * sa + k]; y[k] = b[(size_t)i * sb + k]; }
    out[i] = (float)orc_skew_iou(x, y, mode);
  }
}
void orc_skew_iou_pairwise(const float* a, int n, int sa, const float* b, int m, int sb, int mode, float* out) {
  for (int i = 0; i < n; i++) {
    double x[5];
    for (int k = 0; k < 5; k++) x[k] = a[(size_t)i * sa + k];
    for (int j = 0; j < m; j++) {
      double y[5];
      for (int k = 0; k < 5; k++) y[k] = b[(size_t)j * sb + k];
      out[(size_t)i * m + j] = (float)orc_skew_iou(x, y, mode);
    }
  }
}
#endif /* !ORC_FMA */
