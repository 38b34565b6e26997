// Multi-tensor forms of the per-layer weight pack / weight-gradient unpack kernels.  One training step packs 149 weight
// operands and unpacks 75 weight gradients; as separate launches that is ~2 ms of launch-bound GPU time per step whatever
// the batch size -- a fixed cost that caps strong scaling (at 8 images per rank the whole step is ~10 ms).  Here one
// launch walks a job table (device memory, built once per plan): grid.y = job, the block strides over that job's elements
// with the same element function as the single-tensor kernels (pack.cuh).
#include <cuda_bf16.h>

#include "common.cuh"
#include "pack.cuh"

namespace ryolo {

// grid = (chunks, jobs): a block belongs to ONE job (parameters uniform, no per-element search).  A thread owns one
// (co, ci) PAIR and walks its taps: the k*k weights of a pair are contiguous in nn.Conv2d.weight (36-byte runs instead of nine
// 4-byte reads a sector each), and each tap plane of the packed operand is written coalesced along ci.
__global__ void __launch_bounds__(256) pack_multi_kernel(const ryolo_pack_job* __restrict__ jobs, int njobs,
                                                         const long long* __restrict__ prefix, long long total) {
  const ryolo_pack_job jb = jobs[blockIdx.y];
  const int plane = jb.cin_pad * jb.cout_pad;
  const int taps = jb.ks * jb.ks;            // ks = taps per side of the PACKED operand (<= 3)
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(jb.packed);
  for (int pair = blockIdx.x * blockDim.x + threadIdx.x; pair < plane; pair += gridDim.x * blockDim.x) {
    const int co = pair / jb.cin_pad, ci = pair - co * jb.cin_pad;
    const bool live = ci < jb.cin && co < jb.cout;
    float v[9];
#pragma unroll
    for (int tap = 0; tap < 9; tap++)
      v[tap] = (live && tap < taps) ? pack_value(jb.weight, jb.cout, jb.cin, jb.ks, jb.mode, tap, co, ci) : 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; tap++)
      if (tap < taps) out[(size_t)tap * plane + pair] = __float2bfloat16_rn(v[tap]);
  }
}

// thread = one (co, c) pair of the parameter gradient: k*k reads, one per tap plane (coalesced along c), k*k contiguous writes
__global__ void __launch_bounds__(256) unpack_multi_kernel(const ryolo_unpack_job* __restrict__ jobs, int njobs,
                                                           const long long* __restrict__ prefix, long long total) {
  const ryolo_unpack_job jb = jobs[blockIdx.y];
  const int ks = jb.ks, kk = ks * ks;        // ks = taps per side of the PARAMETER (1 or 3)
  const int pairs = jb.cout * jb.cin;
  for (int pair = blockIdx.x * blockDim.x + threadIdx.x; pair < pairs; pair += gridDim.x * blockDim.x) {
    const int co = pair / jb.cin, c = pair - co * jb.cin;
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      if (t >= kk) continue;
      const int kw = t % ks, kh = t / ks;
      if (jb.mode == 0) {
        v[t] = jb.dw[((size_t)(kh * ks + kw) * jb.cout_pad + co) * jb.cin_pad + c];
      } else {
        const int qy = kh == 0 ? 0 : 1, py = kh == 1 ? 0 : 1, qx = kw == 0 ? 0 : 1, px = kw == 1 ? 0 : 1;
        v[t] = jb.dw[((size_t)(qy * 2 + qx) * jb.cout_pad + co) * jb.cin_pad + (py * 2 + px) * jb.cin + c];
      }
    }
#pragma unroll
    for (int t = 0; t < 9; t++)
      if (t < kk) jb.grad[(size_t)pair * kk + t] = v[t];
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
