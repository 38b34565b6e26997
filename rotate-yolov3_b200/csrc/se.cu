// Squeeze-and-excitation block of the reference's *_se.cfg graphs (model/models.py:16-31 SELayer; built for `[se]` blocks at
// models.py:88-90; cfg/ICDAR/yolov3_608_se.cfg, cfg/HRSC+/yolov3_512_se.cfg): per image and channel
//     s = sigmoid(W2 . relu(W1 . mean_{h,w} x)),   x <- x * s
// on a padded-NHWC bf16 activation, in place (the scaled tensor is what every later block reads).  Eval path:
// three small memory-bound kernels -- channel means (fp32 sums of the bf16 tensor), the two bias-free linear layers of
// one image in one CTA, and the channel-wise rescale.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ryolo {

__device__ __forceinline__ size_t se_pad_off(int b, int y, int x, int h, int w, int cs) {
  return (((size_t)b * (h + 2) + y + 1) * (w + 2) + x + 1) * cs;
}

// sums[b, c] += sum over a span of interior rows; grid (row spans, batch); thread = 8-channel group x pixel stride
__global__ void __launch_bounds__(256) se_pool_kernel(const __nv_bfloat16* __restrict__ x, int cs, int h, int w, int c,
                                                      int rows_per_block, float* __restrict__ sums) {
  const int b = blockIdx.y;
  const int cgs = c >> 3;
  const int cg = threadIdx.x % cgs, p0 = threadIdx.x / cgs, pstep = 256 / cgs;
  if (p0 >= pstep) return;
  const int y0 = blockIdx.x * rows_per_block, y1 = min(h, y0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; e++) acc[e] = 0.f;
  for (int y = y0; y < y1; y++) {
    const __nv_bfloat16* row = x + se_pad_off(b, y, 0, h, w, cs) + cg * 8;
    for (int xx = p0; xx < w; xx += pstep) {
      const uint4 v = *reinterpret_cast<const uint4*>(row + (size_t)xx * cs);
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float2 f = __bfloat1622float2(hp[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; e++) atomicAdd(&sums[(size_t)b * c + cg * 8 + e], acc[e]);
}

// one CTA per image: hidden = relu(W1 [cr, c] . mean), s = sigmoid(W2 [c, cr] . hidden)
__global__ void __launch_bounds__(256) se_fc_kernel(const float* __restrict__ sums, float inv_hw, const float* __restrict__ w1,
                                                    const float* __restrict__ w2, int c, int cr, float* __restrict__ scale) {
  extern __shared__ float se_sm[];     // [c] means, [cr] hidden
  float* mean = se_sm;
  float* hid = se_sm + c;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < c; i += 256) mean[i] = sums[(size_t)b * c + i] * inv_hw;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < cr; j += 8) {
    float a = 0.f;
    for (int i = lane; i < c; i += 32) a = fmaf(w1[(size_t)j * c + i], mean[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) hid[j] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += 256) {
    float a = 0.f;
    for (int j = 0; j < cr; j++) a = fmaf(w2[(size_t)i * cr + j], hid[j], a);
    scale[(size_t)b * c + i] = 1.f / (1.f + expf(-a));
  }
}

__global__ void __launch_bounds__(256) se_scale_kernel(__nv_bfloat16* __restrict__ x, int cs, int batch, int h, int w, int c,
                                                       const float* __restrict__ scale) {
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cgs = c >> 3;
  const size_t npix = (size_t)batch * h * w;
  if (item >= npix * cgs) return;
  const int cg = (int)(item % cgs);
  const size_t pix = item / cgs;
  const int xx = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((size_t)w * h));
  __nv_bfloat16* p = x + se_pad_off(b, y, xx, h, w, cs) + cg * 8;
  uint4 v = *reinterpret_cast<const uint4*>(p);
  __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&v);
  const float* s = scale + (size_t)b * c + cg * 8;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float2 f = __bfloat1622float2(hp[e]);
    hp[e] = __floats2bfloat162_rn(f.x * s[2 * e], f.y * s[2 * e + 1]);
  }
  *reinterpret_cast<uint4*>(p) = v;
}

}  // namespace ryolo

using namespace ryolo;

extern "C" int ryolo_se_block(void* x, int x_cstride, int batch, int h, int w, int c, const float* w1, const float* w2,
                              int reduced, float* sums_scratch, float* scale_scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(x && w1 && w2 && sums_scratch && scale_scratch && batch > 0 && h > 0 && w > 0);
  RYOLO_ARG_CHECK(c > 0 && c % 8 == 0 && c <= 2048 && 256 % (c >> 3) == 0 && reduced > 0 && reduced <= c && batch <= 65535);
  RYOLO_ARG_CHECK(x_cstride >= c && x_cstride % 8 == 0);
  RYOLO_CUDA_TRY(cudaMemsetAsync(sums_scratch, 0, sizeof(float) * (size_t)batch * c, stream));
  const int rpb = h > 64 ? (h + 63) / 64 : 1;
  dim3 gp((h + rpb - 1) / rpb, batch);
  se_pool_kernel<<<gp, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_cstride, h, w, c, rpb, sums_scratch);
  RYOLO_LAUNCH_CHECK();
  se_fc_kernel<<<batch, 256, (size_t)(c + reduced) * sizeof(float), stream>>>(sums_scratch, 1.0f / ((float)h * w), w1, w2, c,
                                                                             reduced, scale_scratch);
  RYOLO_LAUNCH_CHECK();
  const size_t items = (size_t)batch * h * w * (c >> 3);
  se_scale_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x), x_cstride, batch, h, w, c,
                                                                      scale_scratch);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
