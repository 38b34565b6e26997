// Greedy prediction <-> ground-truth matching of the mAP evaluation on the device (reference test.py:134-151: a Python
// loop over the predictions of an image, one skew_bbox_iou call -- i.e. one GPU round trip -- per prediction).
// Input: the rotated-IoU matrix of the image (ryolo_riou_pairwise, [P, T]), prediction / target classes; output:
// correct[P] in {0, 1}.  The assignment is sequential in the confidence order by definition (a target can be claimed
// once), but each step is only an argmax over the T targets of the prediction's class: one warp walks the predictions,
// lanes stride over the targets, (value, index) warp reduction with the FIRST maximum winning like torch.max.
#include "common.cuh"

namespace ryolo {

__global__ void __launch_bounds__(32) match_kernel(const float* __restrict__ iou, int p, int t, int iou_stride,
                                                   const float* __restrict__ pcls, int pcls_stride,
                                                   const float* __restrict__ tcls, float thr,
                                                   unsigned char* __restrict__ claimed /*[t] scratch*/,
                                                   unsigned char* __restrict__ correct) {
  const int lane = threadIdx.x;
  for (int j = lane; j < t; j += 32) claimed[j] = 0;
  for (int i = lane; i < p; i += 32) correct[i] = 0;
  __syncwarp();
  int n_claimed = 0;
  for (int i = 0; i < p && n_claimed < t; i++) {          // "break if all targets already located" (test.py:137-139)
    const float c = pcls[(size_t)i * pcls_stride];
    float best = -1.f;                                     // IoU >= 0; -1 = no target of this class
    int bi = 0x7fffffff;
    for (int j = lane; j < t; j += 32) {
      if (tcls[j] == c) {
        const float v = iou[(size_t)i * iou_stride + j];
        if (v > best) { best = v; bi = j; }               // strict: keeps the first maximum of this lane's stride
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    // "continue if predicted class not among image classes" = no target of this class: best stays -1
    if (best > thr && bi != 0x7fffffff && !claimed[bi]) {
      if (lane == 0) {
        claimed[bi] = 1;
        correct[i] = 1;
      }
      n_claimed++;
    }
    __syncwarp();
  }
}

}  // namespace ryolo

using namespace ryolo;

extern "C" int ryolo_match_detections(const float* iou, int p, int t, int iou_stride, const float* pcls, int pcls_stride,
                                      const float* tcls, float iou_thres, unsigned char* claimed_scratch,
                                      unsigned char* correct, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(p >= 0 && t >= 0 && correct != nullptr);
  if (p == 0) return RYOLO_OK;
  if (t == 0) {
    RYOLO_CUDA_TRY(cudaMemsetAsync(correct, 0, (size_t)p, stream));
    return RYOLO_OK;
  }
  RYOLO_ARG_CHECK(iou && pcls && tcls && claimed_scratch && iou_stride >= t && pcls_stride >= 1);
  match_kernel<<<1, 32, 0, stream>>>(iou, p, t, iou_stride, pcls, pcls_stride, tcls, iou_thres, claimed_scratch, correct);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
