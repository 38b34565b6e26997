// Multi-tensor forms of the per-layer weight pack / weight-gradient unpack kernels.  One training step packs 149 weight
// operands and unpacks 75 weight gradients; as separate launches that is ~2 ms of launch-bound GPU time per step whatever
// the batch size -- a fixed cost that caps strong scaling (at 8 images per rank the whole step is ~10 ms).  Here one
// launch walks a job table (device memory, built once per plan): thread -> global element index -> job by binary search
// over the prefix sums -> the same element function as the single-tensor kernels (pack.cuh).
#include <cuda_bf16.h>

#include "common.cuh"
#include "pack.cuh"

namespace ryolo {

__device__ __forceinline__ int find_job(const long long* __restrict__ prefix, int njobs, long long i) {
  int lo = 0, hi = njobs;          // prefix[j] <= i < prefix[j + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) pack_multi_kernel(const ryolo_pack_job* __restrict__ jobs, int njobs,
                                                         const long long* __restrict__ prefix, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = find_job(prefix, njobs, i);
    const ryolo_pack_job jb = jobs[j];
    const long long e = i - prefix[j];
    const int ci = (int)(e % jb.cin_pad);
    const int co = (int)((e / jb.cin_pad) % jb.cout_pad);
    const int tap = (int)(e / ((long long)jb.cin_pad * jb.cout_pad));
    float v = 0.f;
    if (ci < jb.cin && co < jb.cout) v = pack_value(jb.weight, jb.cout, jb.cin, jb.ks, jb.mode, tap, co, ci);
    reinterpret_cast<__nv_bfloat16*>(jb.packed)[e] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(256) unpack_multi_kernel(const ryolo_unpack_job* __restrict__ jobs, int njobs,
                                                           const long long* __restrict__ prefix, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = find_job(prefix, njobs, i);
    const ryolo_unpack_job jb = jobs[j];
    const long long e = i - prefix[j];
    const int ks = jb.ks;
    const int kw = (int)(e % ks), kh = (int)((e / ks) % ks);
    const int c = (int)((e / ((long long)ks * ks)) % jb.cin);
    const int co = (int)(e / ((long long)ks * ks * jb.cin));
    float v;
    if (jb.mode == 0) {
      v = jb.dw[((size_t)(kh * ks + kw) * jb.cout_pad + co) * jb.cin_pad + c];
    } else {
      const int qy = kh == 0 ? 0 : 1, py = kh == 1 ? 0 : 1, qx = kw == 0 ? 0 : 1, px = kw == 1 ? 0 : 1;
      v = jb.dw[((size_t)(qy * 2 + qx) * jb.cout_pad + co) * jb.cin_pad + (py * 2 + px) * jb.cin + c];
    }
    jb.grad[e] = v;
  }
}

}  // namespace ryolo

using namespace ryolo;

extern "C" int ryolo_conv_pack_weights_multi(const ryolo_pack_job* jobs_dev, int njobs, const long long* prefix_dev,
                                             long long total, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(jobs_dev && prefix_dev && njobs > 0 && total > 0);
  const long long want = (total + 255) / 256;
  const int blocks = (int)(want > 148 * 32 ? 148 * 32 : want);
  pack_multi_kernel<<<blocks, 256, 0, stream>>>(jobs_dev, njobs, prefix_dev, total);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_conv_unpack_wgrad_multi(const ryolo_unpack_job* jobs_dev, int njobs, const long long* prefix_dev,
                                             long long total, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(jobs_dev && prefix_dev && njobs > 0 && total > 0);
  const long long want = (total + 255) / 256;
  const int blocks = (int)(want > 148 * 32 ? 148 * 32 : want);
  unpack_multi_kernel<<<blocks, 256, 0, stream>>>(jobs_dev, njobs, prefix_dev, total);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
