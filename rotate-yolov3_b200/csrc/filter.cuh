// Row predicates of the non_max_suppression candidate filter (reference utils/nms/nms.py:34-40), shared by the per-image
// filter (misc.cu) and the batched candidate selection (select.cu).
#pragma once

namespace ryolo {

// class_conf, class_pred = pred[:, 6:].max(1)  (first maximum on ties; NaN propagates like torch.max)
__device__ __forceinline__ void class_max(const float* __restrict__ row, int nc, float* best_out, int* idx_out,
                                          bool* finite_out) {
  float best = row[6];
  int bi = 0;
  bool fin = isfinite(best);
  for (int k = 1; k < nc; k++) {
    const float v = row[6 + k];
    fin = fin && isfinite(v);
    if (v > best || (v != v && best == best)) { best = v; bi = k; }
  }
  *best_out = best; *idx_out = bi; *finite_out = fin;
}

// (pred[:, 5] > conf_thres) & (pred[:, 2:4] > min_wh).all(1) & torch.isfinite(pred).all(1)   (nms.py:40)
__device__ __forceinline__ bool keep_row(const float* __restrict__ row, float conf, bool cls_finite, float conf_thres,
                                         float min_wh) {
  bool fin = cls_finite && isfinite(conf);
#pragma unroll
  for (int k = 0; k < 5; k++) fin = fin && isfinite(row[k]);
  return conf > conf_thres && row[2] > min_wh && row[3] > min_wh && fin;
}

}  // namespace ryolo
