// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- never linked into the product.
//
// Host-C++ build of the reference's *own* rotated-IoU device code.  The device
// functions of /root/reference/utils/nms/src/rotate_polygon_nms_kernel.cu
// (:19-260, trangle_area .. devRotateIoU) are self-contained CUDA C; build_ref.sh
// extracts that line range at build time into the git-ignored oracle/_ref/ and this
// wrapper compiles it as plain C++ (`__device__` defined away).  The driver below
// restates the host side of nms_cuda() (same file :323-384) without ATen/THC.
//
// Nothing from /root/reference is committed; the .so is rebuilt from the sources
// where they lie.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define __device__
#define __host__
#include REF_EXTRACT_HOST   // -> oracle/_ref/_ref_device_host.inc (generated, deleted after build)

static_assert(sizeof(decltype(cos(1.0f))) == 4, "cos(float) must resolve to the float overload like CUDA");
static_assert(sizeof(decltype(sqrt(1.0f))) == 4, "sqrt(float) must resolve to the float overload like CUDA");

extern "C" {

// devRotateIoU(region1, region2) -- reference :251-260
float ref_host_iou(const float* r1, const float* r2) { return devRotateIoU(r1, r2); }

// intersection area only -- reference inter() :231-249
float ref_host_inter(const float* r1, const float* r2) { return inter(r1, r2); }

// paired: out[i] = devRotateIoU(a + i*stride, b + i*stride)
void ref_host_iou_paired(const float* a, const float* b, int n, int stride, float* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int i = 0; i < n; i++) out[i] = devRotateIoU(a + (size_t)i * stride, b + (size_t)i * stride);
}

// all pairs: out[i*m + j] = devRotateIoU(a_i, b_j)
void ref_host_iou_pairwise(const float* a, int n, const float* b, int m, int stride, float* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) out[(size_t)i * m + j] = devRotateIoU(a + (size_t)i * stride, b + (size_t)j * stride);
}

// Upper-triangle suppression mask, exactly the words the reference's host scan reads
// (rotate_nms_kernel :262-308 restricted to col_block >= row_block; diagonal tile bits only
// for col > row).  boxes = score-sorted [n,6]; mask = [n, col_blocks] uint64 (zero-filled here).
void ref_host_mask(const float* boxes, int n, float thr, unsigned long long* mask, int nthreads) {
  const int TPB = 64;
  const int col_blocks = DIVUP(n, TPB);
  memset(mask, 0, sizeof(unsigned long long) * (size_t)n * col_blocks);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 8)
  for (int i = 0; i < n; i++) {
    const int rb = i / TPB;
    for (int cb = rb; cb < col_blocks; cb++) {
      const int col_size = std::min(n - cb * TPB, TPB);
      int start = (cb == rb) ? (i % TPB) + 1 : 0;
      unsigned long long t = 0;
      for (int k = start; k < col_size; k++)
        if (devRotateIoU(boxes + (size_t)i * 6, boxes + (size_t)(cb * TPB + k) * 6) > thr) t |= 1ULL << k;
      mask[(size_t)i * col_blocks + cb] = t;
    }
  }
}

// Greedy scan of a mask -- reference :358-376.  Returns num_to_keep; keep = sorted-order indices.
int ref_host_scan(const unsigned long long* mask, int n, int64_t* keep) {
  const int TPB = 64;
  const int col_blocks = DIVUP(n, TPB);
  std::vector<unsigned long long> remv(col_blocks, 0ULL);
  int num_to_keep = 0;
  for (int i = 0; i < n; i++) {
    int nblock = i / TPB, inblock = i % TPB;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep[num_to_keep++] = i;
      const unsigned long long* p = mask + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
    }
  }
  return num_to_keep;
}

// Whole nms_cuda() :323-384 on the host.  dets [n,6] (x,y,w,h,theta,score); returns K and writes the
// ascending-sorted ORIGINAL indices of the kept boxes.  Score sort is stable-descending (the
// reference's torch sort is unstable; tests use tie-free scores).
int ref_host_rnms(const float* dets, int n, float thr, int64_t* keep_out, int nthreads, double* t_mask_s, double* t_scan_s) {
  if (n <= 0) return 0;
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return dets[x * 6 + 5] > dets[y * 6 + 5]; });
  std::vector<float> sorted((size_t)n * 6);
  for (int i = 0; i < n; i++) memcpy(&sorted[(size_t)i * 6], dets + order[i] * 6, 6 * sizeof(float));
  const int col_blocks = DIVUP(n, 64);
  std::vector<unsigned long long> mask((size_t)n * col_blocks);
  double t0 = 0, t1 = 0, t2 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  ref_host_mask(sorted.data(), n, thr, mask.data(), nthreads);
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  std::vector<int64_t> keep(n);
  int k = ref_host_scan(mask.data(), n, keep.data());
#ifdef _OPENMP
  t2 = omp_get_wtime();
#endif
  if (t_mask_s) *t_mask_s = t1 - t0;
  if (t_scan_s) *t_scan_s = t2 - t1;
  for (int i = 0; i < k; i++) keep_out[i] = order[keep[i]];
  std::sort(keep_out, keep_out + k);
  return k;
}

int ref_host_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
