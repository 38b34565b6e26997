// Rotated-box IoU for sm_100a with skew_bbox_iou semantics (reference: utils/utils.py:290-320,
// skewiou :663-699, get_rotated_coors :702-725).  The reference computes one pair at a time in Python
// through cv2 + shapely (float64 GEOS); here one launch produces the whole N x M matrix (or the N paired values).
//
// Geometry.  Box b is expressed in box a's local frame (centre difference rotated by -theta_a, relative rotation
// theta_b - theta_a from the per-box sin/cos), which keeps fp32 cancellation at box scale instead of canvas scale.
// The intersection area is the area enclosed by the image of b's boundary under the Euclidean PROJECTION onto a
// (for an axis-aligned box: the coordinate-wise clamp).  For z in int(a) the winding number of that image curve
// about z equals the winding number of the boundary of b (the segment from w to clamp(w) never enters int(a)), and it
// is 0 outside a, so
//        area(a ∩ b) = (1/2) * closed-integral over clamp(boundary b) of (x dy - y dx) = sum over edges of
//                      sum over the linear pieces between clamp breakpoints of (X_i + X_{i+1}) (Y_{i+1} - Y_i) / 2.
// Each edge has at most 4 breakpoints (x = ±hw, y = ±hh); they are sorted with a 5-comparator network, extra or
// coincident breakpoints are harmless (they only subdivide a linear piece), so the routine is branch-free, needs no
// vertex arrays, and is a continuous function of its inputs (identical boxes -> 1, touching boxes -> 0, no
// topological decisions as in edge-crossing enumeration).  ~300 fp32 instructions per pair, all lanes converged.
//
// Pairwise kernel: CTA tile = 32 rows x 128 columns, 256 threads.  Stage A1: bounding-circle test for every pair
// (10 instructions, one LDS.128 broadcast); survivors (~28 %) are compacted once per thread into queue 1.
// Stage A2: separating-axis test on queue 1 with converged warps -> queue 2 (~12 % of all pairs).  Stage B: clamp
// integral on queue 2.  The output tile lives in shared memory (pre-zeroed) and is written back with 128-bit
// streaming stores, so HBM sees exactly one coalesced write of the matrix.
#include <stdlib.h>

#include "common.cuh"

namespace ryolo {

constexpr int RT = 32;    // tile rows
constexpr int CT = 128;   // tile columns
constexpr int RIOU_THREADS = 256;

struct RBox {  // per-box derived data, two float4
  float cx, cy, rad, area;  // rad < 0 marks a box whose IoU with anything is 0 (invalid or zero area)
  float c, s, hw, hh;
};

__device__ __forceinline__ RBox make_rbox(const float* __restrict__ p) {
  RBox b;
  const float cx = p[0], cy = p[1], w = p[2], h = p[3], th = p[4];
  b.cx = cx;
  b.cy = cy;
  b.hw = 0.5f * fabsf(w);
  b.hh = 0.5f * fabsf(h);
  sincosf(th, &b.s, &b.c);
  b.area = fabsf(w * h);  // Polygon(...).convex_hull.area of the 4 corners (utils/utils.py:667-668)
  const bool live = isfinite(cx) && isfinite(cy) && isfinite(w) && isfinite(h) && isfinite(th) &&
                    b.area != 0.f && isfinite(b.area);  // invalid polygon / zero area -> 0 (utils/utils.py:669-673)
  b.rad = live ? sqrtf(b.hw * b.hw + b.hh * b.hh) : -1.f;
  return b;
}

// a's frame: |x| <= hw, |y| <= hh.  Returns area(a ∩ b).
__device__ __forceinline__ float clamp_integral_area(const RBox& a, const RBox& b) {
  const float dx = b.cx - a.cx, dy = b.cy - a.cy;
  const float rx = dx * a.c + dy * a.s;
  const float ry = dy * a.c - dx * a.s;
  const float cd = a.c * b.c + a.s * b.s;   // cos(theta_b - theta_a)
  const float sd = a.c * b.s - a.s * b.c;   // sin(theta_b - theta_a)
  const float ux = cd * b.hw, uy = sd * b.hw;
  const float vx = -sd * b.hh, vy = cd * b.hh;
  const float hw = a.hw, hh = a.hh;
  // corners, counter-clockwise (u x v = hw_b * hh_b > 0)
  const float px[4] = {rx - ux - vx, rx + ux - vx, rx + ux + vx, rx - ux + vx};
  const float py[4] = {ry - uy - vy, ry + uy - vy, ry + uy + vy, ry - uy + vy};
  float cxp[4], cyp[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    cxp[i] = fminf(fmaxf(px[i], -hw), hw);
    cyp[i] = fminf(fmaxf(py[i], -hh), hh);
  }
  float acc = 0.f;
  // edge vectors are +2u, +2v, -2u, -2v: four reciprocals serve the eight x/y breakpoint slopes
  const float ru_x = __fdividef(0.5f, ux), ru_y = __fdividef(0.5f, uy), rv_x = __fdividef(0.5f, vx), rv_y = __fdividef(0.5f, vy);
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int f = (e + 1) & 3;
    const float ax = px[e], ay = py[e];
    const float ex = px[f] - ax, ey = py[f] - ay;
    const float rex = e == 0 ? ru_x : (e == 1 ? rv_x : (e == 2 ? -ru_x : -rv_x));
    const float rey = e == 0 ? ru_y : (e == 1 ? rv_y : (e == 2 ? -ru_y : -rv_y));
    // breakpoints (clamped to [0,1]; NaN from 0*inf is dropped by fmaxf/fminf)
    float t0 = fminf(fmaxf((-hw - ax) * rex, 0.f), 1.f);
    float t1 = fminf(fmaxf((hw - ax) * rex, 0.f), 1.f);
    float t2 = fminf(fmaxf((-hh - ay) * rey, 0.f), 1.f);
    float t3 = fminf(fmaxf((hh - ay) * rey, 0.f), 1.f);
    float lo, hi;
    lo = fminf(t0, t1); hi = fmaxf(t0, t1); t0 = lo; t1 = hi;
    lo = fminf(t2, t3); hi = fmaxf(t2, t3); t2 = lo; t3 = hi;
    lo = fminf(t0, t2); hi = fmaxf(t0, t2); t0 = lo; t2 = hi;
    lo = fminf(t1, t3); hi = fmaxf(t1, t3); t1 = lo; t3 = hi;
    lo = fminf(t1, t2); hi = fmaxf(t1, t2); t1 = lo; t2 = hi;
    float X0 = cxp[e], Y0 = cyp[e];
    const float ts[4] = {t0, t1, t2, t3};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float X1 = fminf(fmaxf(fmaf(ts[k], ex, ax), -hw), hw);
      const float Y1 = fminf(fmaxf(fmaf(ts[k], ey, ay), -hh), hh);
      acc = fmaf(X0 + X1, Y1 - Y0, acc);
      X0 = X1;
      Y0 = Y1;
    }
    acc = fmaf(X0 + cxp[f], cyp[f] - Y0, acc);
  }
  return fmaxf(0.5f * acc, 0.f);
}

// axis-aligned envelope of the 8 corners (mode 'giou', utils/utils.py:682-685)
__device__ __forceinline__ float envelope_area(const RBox& a, const RBox& b) {
  const float ax = fabsf(a.c) * a.hw + fabsf(a.s) * a.hh, ay = fabsf(a.s) * a.hw + fabsf(a.c) * a.hh;
  const float bx = fabsf(b.c) * b.hw + fabsf(b.s) * b.hh, by = fabsf(b.s) * b.hw + fabsf(b.c) * b.hh;
  const float x0 = fminf(a.cx - ax, b.cx - bx), x1 = fmaxf(a.cx + ax, b.cx + bx);
  const float y0 = fminf(a.cy - ay, b.cy - by), y1 = fmaxf(a.cy + ay, b.cy + by);
  return (x1 - x0) * (y1 - y0);
}

// strictly-separated test on the 4 box axes (tiny relative slack: touching boxes go on to the integral)
__device__ __forceinline__ bool sat_disjoint(const RBox& a, const RBox& b) {
  const float dx = b.cx - a.cx, dy = b.cy - a.cy;
  const float k = 1.0001f;
  const float C = fabsf(a.c * b.c + a.s * b.s);
  const float S = fabsf(a.c * b.s - a.s * b.c);
  if (fabsf(dx * a.c + dy * a.s) > k * (a.hw + b.hw * C + b.hh * S)) return true;
  if (fabsf(dy * a.c - dx * a.s) > k * (a.hh + b.hw * S + b.hh * C)) return true;
  if (fabsf(dx * b.c + dy * b.s) > k * (b.hw + a.hw * C + a.hh * S)) return true;
  if (fabsf(dy * b.c - dx * b.s) > k * (b.hh + a.hw * S + a.hh * C)) return true;
  return false;
}

__device__ __forceinline__ float iou_from_inter(const RBox& a, const RBox& b, float inter, int mode) {
  inter = fminf(inter, fminf(a.area, b.area));
  const float uni = mode == RYOLO_IOU_MODE_GIOU ? envelope_area(a, b) : a.area + b.area - inter;
  return uni == 0.f ? 0.f : inter / uni;   // union == 0 -> 0 (utils/utils.py:693-694)
}

__global__ void __launch_bounds__(256) riou_paired_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int n, int sa, int sb, int mode, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RBox ba = make_rbox(a + (size_t)i * sa);
  const RBox bb = make_rbox(b + (size_t)i * sb);
  float r = 0.f;
  if (ba.rad >= 0.f && bb.rad >= 0.f) {
    const float dx = bb.cx - ba.cx, dy = bb.cy - ba.cy, R = (ba.rad + bb.rad) * 1.0001f;
    if (!(dx * dx + dy * dy > R * R) && !sat_disjoint(ba, bb)) r = iou_from_inter(ba, bb, clamp_integral_area(ba, bb), mode);
  }
  out[i] = r;
}

struct RiouSmem {
  float4 row0[RT], row1[RT];   // (cx, cy, rad, area), (c, s, hw, hh)
  float4 col0[CT], col1[CT];
  float out[RT * CT];
  unsigned short q1[RT * CT];
  unsigned short q2[RT * CT];
  int cnt1, cnt2;
};

__device__ __forceinline__ RBox load_rbox(const float4* p0, const float4* p1, int i) {
  const float4 u = p0[i], v = p1[i];
  RBox b;
  b.cx = u.x; b.cy = u.y; b.rad = u.z; b.area = u.w;
  b.c = v.x; b.s = v.y; b.hw = v.z; b.hh = v.w;
  return b;
}

template <int STORE>   // 0: streaming (st.global.cs), 1: plain, 2: write-through (st.global.wt)
__global__ void __launch_bounds__(RIOU_THREADS, 4) riou_pairwise_kernel(const float* __restrict__ a, int n, int sa,
                                                                        const float* __restrict__ b, int m, int sb,
                                                                        int mode, float* __restrict__ out) {
  __shared__ RiouSmem sm;
  const int tid = threadIdx.x, lane = tid & 31;
  const int r0 = blockIdx.y * RT, c0 = blockIdx.x * CT;

  if (tid < RT) {
    RBox bx;
    if (r0 + tid < n) bx = make_rbox(a + (size_t)(r0 + tid) * sa);
    else { bx = RBox{}; bx.rad = -1.f; }
    sm.row0[tid] = make_float4(bx.cx, bx.cy, bx.rad, bx.area);
    sm.row1[tid] = make_float4(bx.c, bx.s, bx.hw, bx.hh);
  } else if (tid >= 64 && tid < 64 + CT) {
    const int j = tid - 64;
    RBox bx;
    if (c0 + j < m) bx = make_rbox(b + (size_t)(c0 + j) * sb);
    else { bx = RBox{}; bx.rad = -1.f; }
    sm.col0[j] = make_float4(bx.cx, bx.cy, bx.rad, bx.area);
    sm.col1[j] = make_float4(bx.c, bx.s, bx.hw, bx.hh);
  }
  if (tid == 0) { sm.cnt1 = 0; sm.cnt2 = 0; }
  {  // pre-zero the output tile: 4096 floats = 1024 float4
    float4* o4 = reinterpret_cast<float4*>(sm.out);
#pragma unroll
    for (int e = 0; e < (RT * CT / 4) / RIOU_THREADS; e++) o4[tid + e * RIOU_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  // ---- stage A1: bounding circles; thread = one column x 16 rows ----
  {
    const int c = tid & (CT - 1);
    const int rg = tid >> 7;
    const float4 cb = sm.col0[c];
    unsigned pass = 0;
    if (cb.z >= 0.f) {
#pragma unroll
      for (int k = 0; k < RT / 2; k++) {
        const float4 rb = sm.row0[rg * (RT / 2) + k];   // broadcast
        const float dx = cb.x - rb.x, dy = cb.y - rb.y;
        const float R = (cb.z + rb.z) * 1.0001f;
        const bool ok = rb.z >= 0.f && !(fmaf(dx, dx, dy * dy) > R * R);
        pass |= (ok ? 1u : 0u) << k;
      }
    }
    // one compaction per thread: warp exclusive scan of the counts + one atomic per warp
    const int cnt = __popc(pass);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    int base = 0;
    if (lane == 31) base = atomicAdd(&sm.cnt1, incl);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
    while (pass) {
      const int k = __ffs(pass) - 1;
      pass &= pass - 1;
      sm.q1[base++] = (unsigned short)(((rg * (RT / 2) + k) << 7) | c);
    }
  }
  __syncthreads();
  // ---- stage A2: separating axes on queue 1 (converged warps) ----
  {
    const int cnt1 = sm.cnt1;
    for (int q0 = 0; q0 < cnt1; q0 += RIOU_THREADS) {
      const int q = q0 + tid;
      bool keep = false;
      int e = 0;
      if (q < cnt1) {
        e = sm.q1[q];
        const RBox ra = load_rbox(sm.row0, sm.row1, e >> 7), cb = load_rbox(sm.col0, sm.col1, e & (CT - 1));
        keep = !sat_disjoint(ra, cb);
      }
      const unsigned mk = __ballot_sync(0xffffffffu, keep);
      if (mk) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sm.cnt2, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) sm.q2[base + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)e;
      }
    }
  }
  __syncthreads();
  // ---- stage B: clamp integral on queue 2 ----
  {
    const int cnt2 = sm.cnt2;
    for (int q = tid; q < cnt2; q += RIOU_THREADS) {
      const int e = sm.q2[q];
      const int r = e >> 7, c = e & (CT - 1);
      const RBox ra = load_rbox(sm.row0, sm.row1, r), cb = load_rbox(sm.col0, sm.col1, c);
      sm.out[r * CT + c] = iou_from_inter(ra, cb, clamp_integral_area(ra, cb), mode);
    }
  }
  __syncthreads();
  // ---- write back ----
  const int rows = min(RT, n - r0), cols = min(CT, m - c0);
  if ((m & 3) == 0 && cols == CT && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    for (int e = tid; e < rows * (CT / 4); e += RIOU_THREADS) {
      const int r = e / (CT / 4), c4 = e % (CT / 4);
      const float4 v = *reinterpret_cast<const float4*>(&sm.out[r * CT + c4 * 4]);
      float4* dst = reinterpret_cast<float4*>(out + (size_t)(r0 + r) * m + c0 + c4 * 4);
      if (STORE == 0) __stcs(dst, v);        // streaming: the matrix is written once and not re-read by this kernel
      else if (STORE == 1) *dst = v;
      else __stwt(dst, v);
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
  dim3 grid((m + CT - 1) / CT, (n + RT - 1) / RT);
  RYOLO_ARG_CHECK(grid.y <= 65535);
  static int store_mode = -1;
  if (store_mode < 0) {   // measurement knob (profiles/): RYOLO_RIOU_STORE=0|1|2, default streaming stores
    const char* e = getenv("RYOLO_RIOU_STORE");
    store_mode = e ? atoi(e) : 0;
    if (store_mode < 0 || store_mode > 2) store_mode = 0;
  }
  if (store_mode == 1) riou_pairwise_kernel<1><<<grid, RIOU_THREADS, 0, stream>>>(a, n, stride_a, b, m, stride_b, mode, out);
  else if (store_mode == 2) riou_pairwise_kernel<2><<<grid, RIOU_THREADS, 0, stream>>>(a, n, stride_a, b, m, stride_b, mode, out);
  else riou_pairwise_kernel<0><<<grid, RIOU_THREADS, 0, stream>>>(a, n, stride_a, b, m, stride_b, mode, out);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
