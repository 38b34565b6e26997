// Rotated-box IoU for sm_100a with skew_bbox_iou semantics (reference: utils/utils.py:290-320,
// skewiou :663-699, get_rotated_coors :702-725).  The reference computes one pair at a time in Python
// through cv2 + shapely (float64 GEOS); here one launch produces the whole N x M matrix (or the N paired
// values).  Geometry: box2 is expressed in box1's local frame (translation by the centre difference and
// rotation by theta2 - theta1, which keeps the fp32 cancellation error at the scale of the boxes, not of
// the canvas), clipped against box1's four axis-aligned half-planes (Sutherland-Hodgman, <= 8 vertices)
// and measured with the shoelace formula.  This is NOT the fragile corner/edge-crossing enumeration of the
// reference NMS kernel: skew_bbox_iou's oracle is shapely, for which identical boxes have IoU 1.
//
// Pairwise kernel layout: CTA tile = 32 rows x 128 columns, staged in shared memory and written back with
// 128-bit coalesced stores; a conservative separating-axis filter (no margin subtleties here: rejecting is
// only allowed when the polygons are strictly disjoint, where the exact answer is 0) sends ~10% of the
// pairs to the clipper through a shared-memory queue so that the heavy path runs with full warps.
#include "common.cuh"

namespace ryolo {

constexpr int RT = 32;    // tile rows
constexpr int CT = 128;   // tile columns
constexpr int RIOU_THREADS = 256;

struct RBox {  // per-box derived data
  float cx, cy, hw, hh, c, s, th, area, rad;
  bool ok;
};

__device__ __forceinline__ RBox make_rbox(const float* __restrict__ p) {
  RBox b;
  const float cx = p[0], cy = p[1], w = p[2], h = p[3], th = p[4];
  b.cx = cx; b.cy = cy; b.th = th;
  b.hw = 0.5f * fabsf(w);
  b.hh = 0.5f * fabsf(h);
  sincosf(th, &b.s, &b.c);
  b.area = fabsf(w * h);  // Polygon(...).convex_hull.area of the 4 corners (utils/utils.py:667-668)
  b.rad = sqrtf(b.hw * b.hw + b.hh * b.hh);
  b.ok = isfinite(cx) && isfinite(cy) && isfinite(w) && isfinite(h) && isfinite(th);
  return b;
}

// strictly-disjoint test (tiny relative slack so that touching boxes still go to the clipper)
__device__ __forceinline__ bool riou_disjoint(const RBox& a, const RBox& b) {
  const float dx = b.cx - a.cx, dy = b.cy - a.cy;
  const float R = (a.rad + b.rad) * 1.0001f;
  if (dx * dx + dy * dy > R * R) return true;
  const float k = 1.0001f;
  const float C = fabsf(a.c * b.c + a.s * b.s);
  const float S = fabsf(a.c * b.s - a.s * b.c);
  if (fabsf(dx * a.c + dy * a.s) > k * (a.hw + b.hw * C + b.hh * S)) return true;
  if (fabsf(dy * a.c - dx * a.s) > k * (a.hh + b.hw * S + b.hh * C)) return true;
  if (fabsf(dx * b.c + dy * b.s) > k * (b.hw + a.hw * C + a.hh * S)) return true;
  if (fabsf(dy * b.c - dx * b.s) > k * (b.hh + a.hw * S + a.hh * C)) return true;
  return false;
}

// clip polygon (px,py,n) against  sgn * coord(axis) <= lim ; returns new count (<= n + 1)
__device__ __forceinline__ int clip_halfplane(const float* px, const float* py, int n, float* qx, float* qy, int axis,
                                              float sgn, float lim) {
  int m = 0;
  if (n == 0) return 0;
  float ax = px[n - 1], ay = py[n - 1];
  float da = sgn * (axis == 0 ? ax : ay) - lim;  // <= 0 inside
  for (int i = 0; i < n; i++) {
    const float bx = px[i], by = py[i];
    const float db = sgn * (axis == 0 ? bx : by) - lim;
    if ((da <= 0.f) != (db <= 0.f)) {
      const float t = da / (da - db);
      float ix = ax + t * (bx - ax), iy = ay + t * (by - ay);
      if (axis == 0) ix = sgn * lim; else iy = sgn * lim;  // snap onto the clip line
      qx[m] = ix; qy[m] = iy; m++;
    }
    if (db <= 0.f) { qx[m] = bx; qy[m] = by; m++; }
    ax = bx; ay = by; da = db;
  }
  return m;
}

// intersection area of a (axis-aligned in its own frame) and b
__device__ __forceinline__ float clip_inter_area(const RBox& a, const RBox& b) {
  // b's centre and axes in a's frame
  const float dx = b.cx - a.cx, dy = b.cy - a.cy;
  const float rx = dx * a.c + dy * a.s;
  const float ry = dy * a.c - dx * a.s;
  float sd, cd;
  sincosf(b.th - a.th, &sd, &cd);
  const float ux = cd * b.hw, uy = sd * b.hw;    // half-width vector
  const float vx = -sd * b.hh, vy = cd * b.hh;   // half-height vector
  float px[8], py[8], qx[8], qy[8];
  px[0] = rx - ux - vx; py[0] = ry - uy - vy;
  px[1] = rx + ux - vx; py[1] = ry + uy - vy;
  px[2] = rx + ux + vx; py[2] = ry + uy + vy;
  px[3] = rx - ux + vx; py[3] = ry - uy + vy;
  int n = 4;
  n = clip_halfplane(px, py, n, qx, qy, 0, 1.f, a.hw);
  n = clip_halfplane(qx, qy, n, px, py, 0, -1.f, a.hw);
  n = clip_halfplane(px, py, n, qx, qy, 1, 1.f, a.hh);
  n = clip_halfplane(qx, qy, n, px, py, 1, -1.f, a.hh);
  if (n < 3) return 0.f;
  float acc = 0.f;
  const float ox = px[0], oy = py[0];  // fan about vertex 0: differences keep the terms small
  for (int i = 1; i + 1 < n; i++)
    acc += (px[i] - ox) * (py[i + 1] - oy) - (py[i] - oy) * (px[i + 1] - ox);
  return 0.5f * fabsf(acc);
}

// axis-aligned envelope of the 8 corners (mode 'giou', utils/utils.py:682-685)
__device__ __forceinline__ float envelope_area(const RBox& a, const RBox& b) {
  const float ax = fabsf(a.c) * a.hw + fabsf(a.s) * a.hh, ay = fabsf(a.s) * a.hw + fabsf(a.c) * a.hh;
  const float bx = fabsf(b.c) * b.hw + fabsf(b.s) * b.hh, by = fabsf(b.s) * b.hw + fabsf(b.c) * b.hh;
  const float x0 = fminf(a.cx - ax, b.cx - bx), x1 = fmaxf(a.cx + ax, b.cx + bx);
  const float y0 = fminf(a.cy - ay, b.cy - by), y1 = fmaxf(a.cy + ay, b.cy + by);
  return (x1 - x0) * (y1 - y0);
}

__device__ __forceinline__ float riou_value(const RBox& a, const RBox& b, int mode, bool skip_filter) {
  if (!a.ok || !b.ok) return 0.f;               // invalid polygon -> 0 (utils/utils.py:669-671)
  if (a.area == 0.f || b.area == 0.f) return 0.f;  // :672-673
  if (!skip_filter && riou_disjoint(a, b)) return 0.f;
  float inter = clip_inter_area(a, b);
  inter = fminf(inter, fminf(a.area, b.area));
  const float uni = mode == RYOLO_IOU_MODE_GIOU ? envelope_area(a, b) : a.area + b.area - inter;
  if (uni == 0.f) return 0.f;                   // :693-694
  return inter / uni;
}

__global__ void __launch_bounds__(256) riou_paired_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int n, int sa, int sb, int mode, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RBox ba = make_rbox(a + (size_t)i * sa);
  const RBox bb = make_rbox(b + (size_t)i * sb);
  out[i] = riou_value(ba, bb, mode, false);
}

struct RiouSmem {
  RBox row[RT];
  RBox col[CT];
  float out[RT * CT];
  unsigned short queue[RT * CT];
  int cnt;
};

__global__ void __launch_bounds__(RIOU_THREADS) riou_pairwise_kernel(const float* __restrict__ a, int n, int sa,
                                                                     const float* __restrict__ b, int m, int sb,
                                                                     int mode, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RiouSmem& sm = *reinterpret_cast<RiouSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int r0 = blockIdx.y * RT, c0 = blockIdx.x * CT;

  if (tid < RT) {
    const int i = r0 + tid;
    if (i < n) sm.row[tid] = make_rbox(a + (size_t)i * sa);
    else { RBox z = {}; z.ok = false; sm.row[tid] = z; }
  } else if (tid >= 64 && tid < 64 + CT) {
    const int j = c0 + (tid - 64);
    if (j < m) sm.col[tid - 64] = make_rbox(b + (size_t)j * sb);
    else { RBox z = {}; z.ok = false; sm.col[tid - 64] = z; }
  }
  if (tid == 0) sm.cnt = 0;
  __syncthreads();

  // ---- phase A: one column per thread (registers), 16 rows each (shared-memory broadcast) ----
  {
    const int c = tid & (CT - 1);
    const int rg = tid >> 7;  // 0..1
    const RBox cb = sm.col[c];
    const bool col_live = cb.ok && cb.area != 0.f;
#pragma unroll 4
    for (int k = 0; k < RT / 2; k++) {
      const int r = rg * (RT / 2) + k;
      const RBox& rbx = sm.row[r];
      bool cand = col_live && rbx.ok && rbx.area != 0.f;
      if (cand) cand = !riou_disjoint(rbx, cb);
      if (!cand) sm.out[r * CT + c] = 0.f;
      const unsigned mk = __ballot_sync(0xffffffffu, cand);
      if (mk) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sm.cnt, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) sm.queue[base + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)((r << 7) | c);
      }
    }
  }
  __syncthreads();
  // ---- phase B: clip the survivors ----
  {
    const int cnt = sm.cnt;
    for (int q = tid; q < cnt; q += RIOU_THREADS) {
      const int e = sm.queue[q];
      const int r = e >> 7, c = e & (CT - 1);
      sm.out[r * CT + c] = riou_value(sm.row[r], sm.col[c], mode, true);
    }
  }
  __syncthreads();
  // ---- write back ----
  const int rows = min(RT, n - r0), cols = min(CT, m - c0);
  if ((m & 3) == 0 && cols == CT && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    for (int e = tid; e < rows * (CT / 4); e += RIOU_THREADS) {
      const int r = e / (CT / 4), c4 = e % (CT / 4);
      const float4 v = *reinterpret_cast<const float4*>(&sm.out[r * CT + c4 * 4]);
      __stcs(reinterpret_cast<float4*>(out + (size_t)(r0 + r) * m + c0 + c4 * 4), v);  // streaming store
    }
  } else {
    for (int e = tid; e < rows * CT; e += RIOU_THREADS) {
      const int r = e / CT, c = e % CT;
      if (c < cols) out[(size_t)(r0 + r) * m + c0 + c] = sm.out[r * CT + c];
    }
  }
}

}  // namespace ryolo

using namespace ryolo;

extern "C" int ryolo_riou_paired(const float* a, const float* b, int n, int stride_a, int stride_b, int mode, float* out,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(n >= 0 && stride_a >= 5 && stride_b >= 5);
  RYOLO_ARG_CHECK(mode == RYOLO_IOU_MODE_IOU || mode == RYOLO_IOU_MODE_GIOU);
  if (n == 0) return RYOLO_OK;
  RYOLO_ARG_CHECK(a && b && out);
  riou_paired_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, n, stride_a, stride_b, mode, out);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_riou_pairwise(const float* a, int n, int stride_a, const float* b, int m, int stride_b, int mode,
                                   float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(n >= 0 && m >= 0 && stride_a >= 5 && stride_b >= 5);
  RYOLO_ARG_CHECK(mode == RYOLO_IOU_MODE_IOU || mode == RYOLO_IOU_MODE_GIOU);
  if (n == 0 || m == 0) return RYOLO_OK;
  RYOLO_ARG_CHECK(a && b && out);
  static bool attr_set = false;
  if (!attr_set) {
    RYOLO_CUDA_TRY(cudaFuncSetAttribute(riou_pairwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(RiouSmem)));
    attr_set = true;
  }
  dim3 grid((m + CT - 1) / CT, (n + RT - 1) / RT);
  RYOLO_ARG_CHECK(grid.y <= 65535);
  riou_pairwise_kernel<<<grid, RIOU_THREADS, sizeof(RiouSmem), stream>>>(a, n, stride_a, b, m, stride_b, mode, out);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
