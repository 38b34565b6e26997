// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- never linked into the product.
//
// nvcc build (DEFAULT flags, i.e. -fmad=true, exactly like the reference's setup.py which passes
// no nvcc flags: /root/reference/utils/nms/setup.py:4-12) of the reference's own device code and
// rotate_nms_kernel (/root/reference/utils/nms/src/rotate_polygon_nms_kernel.cu:19-308, extracted
// at build time by build_ref.sh, never committed).  The host function below restates
// nms_cuda() (:323-384) with the plain CUDA runtime instead of ATen/THC (which no longer exist in
// torch 2.x, so the reference extension itself cannot be built).  This is the AUTHORITATIVE index
// oracle for r_nms bit-parity on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <numeric>
#include <vector>

#include REF_EXTRACT_CUDA   // -> oracle/_ref/_ref_device_cuda.inc (generated, deleted after build)

__global__ void ref_iou_paired_kernel(const float* a, const float* b, int n, int stride, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = devRotateIoU(a + (size_t)i * stride, b + (size_t)i * stride);
}

extern "C" {

// launch exactly as :344-347 (grid (cb,cb), 64 threads, legacy default stream). Pointers are DEVICE.
int ref_cuda_mask(const float* boxes_sorted_dev, int n, float thr, unsigned long long* mask_dev) {
  const int cb = DIVUP(n, threadsPerBlock);
  dim3 blocks(cb, cb), threads(threadsPerBlock);
  rotate_nms_kernel<<<blocks, threads>>>(n, thr, boxes_sorted_dev, mask_dev);
  return (int)cudaDeviceSynchronize();
}

// device pointers; paired IoU with the reference devRotateIoU
int ref_cuda_iou_paired(const float* a_dev, const float* b_dev, int n, int stride, float* out_dev) {
  ref_iou_paired_kernel<<<(n + 127) / 128, 128>>>(a_dev, b_dev, n, stride, out_dev);
  return (int)cudaDeviceSynchronize();
}

// Whole nms_cuda(): HOST pointers in/out. Returns K (<0 on CUDA error).
int ref_cuda_rnms(const float* dets, int n, float thr, int64_t* keep_out) {
  if (n <= 0) return 0;
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return dets[x * 6 + 5] > dets[y * 6 + 5]; });
  std::vector<float> sorted((size_t)n * 6);
  for (int i = 0; i < n; i++) memcpy(&sorted[(size_t)i * 6], dets + order[i] * 6, 6 * sizeof(float));
  const int col_blocks = DIVUP(n, threadsPerBlock);
  float* boxes_dev = nullptr;
  unsigned long long* mask_dev = nullptr;
  if (cudaMalloc(&boxes_dev, sizeof(float) * 6 * n) != cudaSuccess) return -1;
  if (cudaMalloc(&mask_dev, sizeof(unsigned long long) * (size_t)n * col_blocks) != cudaSuccess) return -2;
  cudaMemcpy(boxes_dev, sorted.data(), sizeof(float) * 6 * n, cudaMemcpyHostToDevice);
  if (ref_cuda_mask(boxes_dev, n, thr, mask_dev) != 0) return -3;
  std::vector<unsigned long long> mask_host((size_t)n * col_blocks);
  cudaMemcpy(mask_host.data(), mask_dev, sizeof(unsigned long long) * (size_t)n * col_blocks, cudaMemcpyDeviceToHost);
  cudaFree(boxes_dev);
  cudaFree(mask_dev);
  std::vector<unsigned long long> remv(col_blocks, 0ULL);
  int num_to_keep = 0;
  std::vector<int64_t> keep(n);
  for (int i = 0; i < n; i++) {
    int nblock = i / threadsPerBlock, inblock = i % threadsPerBlock;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep[num_to_keep++] = i;
      unsigned long long* p = &mask_host[0] + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
    }
  }
  for (int i = 0; i < num_to_keep; i++) keep_out[i] = order[keep[i]];
  std::sort(keep_out, keep_out + num_to_keep);
  return num_to_keep;
}
}
