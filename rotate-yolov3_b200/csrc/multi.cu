// Multi-tensor forms of the per-layer weight pack / weight-gradient unpack kernels.  One training step packs 149 weight
// operands and unpacks 75 weight gradients; as separate launches that is ~2 ms of launch-bound GPU time per step whatever
// the batch size -- a fixed cost that caps strong scaling (at 8 images per rank the whole step is ~10 ms).  Here one
// launch walks a job table (device memory, built once per plan): grid.y = job, the block strides over that job's elements
// with the same element function as the single-tensor kernels (pack.cuh).
#include <cuda_bf16.h>

#include "common.cuh"
#include "pack.cuh"

namespace ryolo {

// grid = (chunks, jobs): a block belongs to ONE job (parameters uniform, no per-element search) and strides over its elements
__global__ void __launch_bounds__(256) pack_multi_kernel(const ryolo_pack_job* __restrict__ jobs, int njobs,
                                                         const long long* __restrict__ prefix, long long total) {
  const ryolo_pack_job jb = jobs[blockIdx.y];
  const long long n = prefix[blockIdx.y + 1] - prefix[blockIdx.y];
  const int plane = jb.cin_pad * jb.cout_pad;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(jb.packed);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(e / plane);
    const int r = (int)(e - (long long)tap * plane);
    const int co = r / jb.cin_pad, ci = r - co * jb.cin_pad;
    float v = 0.f;
    if (ci < jb.cin && co < jb.cout) v = pack_value(jb.weight, jb.cout, jb.cin, jb.ks, jb.mode, tap, co, ci);
    out[e] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(256) unpack_multi_kernel(const ryolo_unpack_job* __restrict__ jobs, int njobs,
                                                           const long long* __restrict__ prefix, long long total) {
  const ryolo_unpack_job jb = jobs[blockIdx.y];
  const long long n = prefix[blockIdx.y + 1] - prefix[blockIdx.y];
  const int ks = jb.ks, kk = ks * ks;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e % kk);
    const int kw = t % ks, kh = t / ks;
    const long long q = e / kk;
    const int c = (int)(q % jb.cin);
    const int co = (int)(q / jb.cin);
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
  RYOLO_ARG_CHECK(njobs <= 65535);
  pack_multi_kernel<<<dim3(64, njobs), 256, 0, stream>>>(jobs_dev, njobs, prefix_dev, total);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_conv_unpack_wgrad_multi(const ryolo_unpack_job* jobs_dev, int njobs, const long long* prefix_dev,
                                             long long total, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(jobs_dev && prefix_dev && njobs > 0 && total > 0);
  RYOLO_ARG_CHECK(njobs <= 65535);
  unpack_multi_kernel<<<dim3(64, njobs), 256, 0, stream>>>(jobs_dev, njobs, prefix_dev, total);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
