// Training-mode BatchNorm2d + PReLU (+ shortcut add, + nearest x2 upsample) around the tcgen05 convolutions, forward
// and backward, on padded-NHWC bf16 activations.  Reference semantics: nn.BatchNorm2d(momentum=0.1, eps=1e-5) with
// batch statistics -> nn.PReLU(num_parameters=1) (model/models.py:62-65), shortcut add (models.py:281-282), nn.Upsample
// nearest x2 (models.py:93-94), differentiated by autograd (train.py:278-282).
//
// Memory-bound passes, 16-byte (8-channel) accesses, channels innermost:
//   bn_stats          z -> per-channel sum, sum of squares                      (1 read)
//   bn_act_fwd        y = prelu(z*scale + shift) [+ res], optional 2x2 replicated store      (1 read [+1], 1 write)
//   bn_act_bwd_reduce dy, z -> per-channel sum(du), sum(du * zhat), scalar d(slope)            (2 reads)
//   bn_act_bwd_apply  dz = scale*(du - mean(du) - zhat*mean(du*zhat)) written over z; dres (+)= dy   (2 reads, 1-2 writes)
// plus the small layout kernels of the training graph (zero insertion for stride-2 adjoints, NCHW fp32 -> padded
// NHWC bf16 for the head gradients, im2col of the 3-channel image so that the first layer runs on the same GEMMs).
#include <cuda_bf16.h>

#include "common.cuh"

namespace ryolo {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return *reinterpret_cast<uint4*>(h);
}

struct Geo {
  int batch, h, w, c;   // interior size, real channels (multiple of 8)
};

// thread -> (interior pixel, 8-channel group)
__device__ __forceinline__ bool decode_item(const Geo& g, size_t item, int& b, int& y, int& x, int& cg) {
  const int cgs = g.c >> 3;
  const size_t npix = (size_t)g.batch * g.h * g.w;
  if (item >= npix * cgs) return false;
  cg = (int)(item % cgs);
  const size_t pix = item / cgs;
  x = (int)(pix % g.w);
  y = (int)((pix / g.w) % g.h);
  b = (int)(pix / ((size_t)g.w * g.h));
  return true;
}
__device__ __forceinline__ size_t pad_off(int b, int y, int x, int h, int w, int cs) {  // element offset of padded pixel
  return (((size_t)b * (h + 2) + y + 1) * (w + 2) + x + 1) * cs;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ z, int zcs, Geo g,
                                                       float* __restrict__ sums /*[2*c]*/, int items_per_block) {
  extern __shared__ float s_acc[];  // [2 * c]
  for (int i = threadIdx.x; i < 2 * g.c; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int cgs = g.c >> 3;
  // a thread keeps its channel group for the whole block: items advance by a multiple of cgs
  const int stride = (blockDim.x / cgs) * cgs;
  const size_t begin = (size_t)blockIdx.x * items_per_block;
  const size_t end = begin + items_per_block;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; e++) s1[e] = s2[e] = 0.f;
  int cg_mine = -1;
  if ((int)threadIdx.x < stride) {
    for (size_t item = begin + threadIdx.x; item < end; item += stride) {
      int b, y, x, cg;
      if (!decode_item(g, item, b, y, x, cg)) break;
      cg_mine = cg;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(z + pad_off(b, y, x, g.h, g.w, zcs) + cg * 8), f);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        s1[e] += f[e];
        s2[e] = fmaf(f[e], f[e], s2[e]);
      }
    }
  }
  if (cg_mine >= 0) {
#pragma unroll
    for (int e = 0; e < 8; e++) {
      atomicAdd(&s_acc[cg_mine * 8 + e], s1[e]);
      atomicAdd(&s_acc[g.c + cg_mine * 8 + e], s2[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.c; i += blockDim.x) atomicAdd(&sums[i], s_acc[i]);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ z, int zcs, Geo g,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         float slope, int has_act, const __nv_bfloat16* __restrict__ res,
                                                         int rcs, __nv_bfloat16* __restrict__ y, int ycs, int up) {
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b, yy, xx, cg;
  if (!decode_item(g, item, b, yy, xx, cg)) return;
  float f[8];
  unpack8(*reinterpret_cast<const uint4*>(z + pad_off(b, yy, xx, g.h, g.w, zcs) + cg * 8), f);
  const float4 sc0 = *reinterpret_cast<const float4*>(scale + cg * 8), sc1 = *reinterpret_cast<const float4*>(scale + cg * 8 + 4);
  const float4 sh0 = *reinterpret_cast<const float4*>(shift + cg * 8), sh1 = *reinterpret_cast<const float4*>(shift + cg * 8 + 4);
  const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
  const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
  float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (res) unpack8(*reinterpret_cast<const uint4*>(res + pad_off(b, yy, xx, g.h, g.w, rcs) + cg * 8), r);
#pragma unroll
  for (int e = 0; e < 8; e++) {
    float u = fmaf(f[e], sc[e], sh[e]);
    if (has_act) u = u > 0.f ? u : slope * u;
    f[e] = u + r[e];
  }
  const uint4 o = pack8(f);
  if (!up) {
    *reinterpret_cast<uint4*>(y + pad_off(b, yy, xx, g.h, g.w, ycs) + cg * 8) = o;
  } else {
#pragma unroll
    for (int ry = 0; ry < 2; ry++)
#pragma unroll
      for (int rx = 0; rx < 2; rx++)
        *reinterpret_cast<uint4*>(y + pad_off(b, 2 * yy + ry, 2 * xx + rx, 2 * g.h, 2 * g.w, ycs) + cg * 8) = o;
  }
}

// du for one 8-channel group; dy possibly at 2x resolution (sum of the 2x2 block = adjoint of nearest upsample)
__device__ __forceinline__ void load_dy(const __nv_bfloat16* __restrict__ dy, int dcs, const Geo& g, int b, int y, int x,
                                        int cg, int up, float* d) {
  if (!up) {
    unpack8(*reinterpret_cast<const uint4*>(dy + pad_off(b, y, x, g.h, g.w, dcs) + cg * 8), d);
  } else {
#pragma unroll
    for (int e = 0; e < 8; e++) d[e] = 0.f;
#pragma unroll
    for (int ry = 0; ry < 2; ry++)
#pragma unroll
      for (int rx = 0; rx < 2; rx++) {
        float t[8];
        unpack8(*reinterpret_cast<const uint4*>(dy + pad_off(b, 2 * y + ry, 2 * x + rx, 2 * g.h, 2 * g.w, dcs) + cg * 8), t);
#pragma unroll
        for (int e = 0; e < 8; e++) d[e] += t[e];
      }
  }
}

__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, int dcs, int up,
                                                                const __nv_bfloat16* __restrict__ z, int zcs, Geo g,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, float slope, int has_act,
                                                                float* __restrict__ sums /*[2*c + 1]*/,
                                                                int items_per_block) {
  extern __shared__ float s_acc[];  // [2 * c + 1]
  for (int i = threadIdx.x; i < 2 * g.c + 1; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int cgs = g.c >> 3;
  const int stride = (blockDim.x / cgs) * cgs;
  const size_t begin = (size_t)blockIdx.x * items_per_block;
  const size_t end = begin + items_per_block;
  float a1[8], a2[8], asl = 0.f;
#pragma unroll
  for (int e = 0; e < 8; e++) a1[e] = a2[e] = 0.f;
  int cg_mine = -1;
  if ((int)threadIdx.x < stride) {
    for (size_t item = begin + threadIdx.x; item < end; item += stride) {
      int b, y, x, cg;
      if (!decode_item(g, item, b, y, x, cg)) break;
      cg_mine = cg;
      float f[8], d[8];
      unpack8(*reinterpret_cast<const uint4*>(z + pad_off(b, y, x, g.h, g.w, zcs) + cg * 8), f);
      load_dy(dy, dcs, g, b, y, x, cg, up, d);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int ch = cg * 8 + e;
        const float u = fmaf(f[e], scale[ch], shift[ch]);
        float du = d[e];
        if (has_act && !(u > 0.f)) {
          asl = fmaf(d[e], u, asl);
          du *= slope;
        }
        const float zh = (f[e] - mean[ch]) * invstd[ch];
        a1[e] += du;
        a2[e] = fmaf(du, zh, a2[e]);
      }
    }
  }
  if (cg_mine >= 0) {
#pragma unroll
    for (int e = 0; e < 8; e++) {
      atomicAdd(&s_acc[cg_mine * 8 + e], a1[e]);
      atomicAdd(&s_acc[g.c + cg_mine * 8 + e], a2[e]);
    }
    atomicAdd(&s_acc[2 * g.c], asl);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.c + 1; i += blockDim.x) atomicAdd(&sums[i], s_acc[i]);
}

__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dcs, int up,
                                                               __nv_bfloat16* __restrict__ z /*in: z, out: dz*/, int zcs,
                                                               Geo g, const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, float slope, int has_act,
                                                               int has_bn, const float* __restrict__ sums, float inv_n,
                                                               __nv_bfloat16* __restrict__ gres, int gcs, int gres_acc) {
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b, y, x, cg;
  if (!decode_item(g, item, b, y, x, cg)) return;
  float f[8], d[8];
  __nv_bfloat16* zp = z + pad_off(b, y, x, g.h, g.w, zcs) + cg * 8;
  unpack8(*reinterpret_cast<const uint4*>(zp), f);
  load_dy(dy, dcs, g, b, y, x, cg, up, d);
  if (gres) {  // shortcut branch: d(residual) (+)= dy   (never combined with upsample)
    __nv_bfloat16* gp = gres + pad_off(b, y, x, g.h, g.w, gcs) + cg * 8;
    float o[8];
    if (gres_acc) {
      unpack8(*reinterpret_cast<const uint4*>(gp), o);
#pragma unroll
      for (int e = 0; e < 8; e++) o[e] += d[e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) o[e] = d[e];
    }
    *reinterpret_cast<uint4*>(gp) = pack8(o);
  }
  float out[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int ch = cg * 8 + e;
    const float sc = scale[ch];
    const float u = fmaf(f[e], sc, shift[ch]);
    float du = d[e];
    if (has_act && !(u > 0.f)) du *= slope;
    if (has_bn) {
      const float zh = (f[e] - mean[ch]) * invstd[ch];
      out[e] = sc * (du - sums[ch] * inv_n - zh * sums[g.c + ch] * inv_n);
    } else {
      out[e] = du;
    }
  }
  *reinterpret_cast<uint4*>(zp) = pack8(out);
}

// ---------------------------------------------------------------------------------------------------------------
// adjoint of "keep the even pixels" (stride-2 conv): small grid gradient placed on the even interior pixels of the
// (pre-zeroed, never otherwise written) input-resolution buffer
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const __nv_bfloat16* __restrict__ src, int scs, Geo g /*small*/,
                                                            __nv_bfloat16* __restrict__ dst, int dcs, int ih, int iw) {
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b, y, x, cg;
  if (!decode_item(g, item, b, y, x, cg)) return;
  const uint4 v = *reinterpret_cast<const uint4*>(src + pad_off(b, y, x, g.h, g.w, scs) + cg * 8);
  *reinterpret_cast<uint4*>(dst + pad_off(b, 2 * y, 2 * x, ih, iw, dcs) + cg * 8) = v;
}

// fp32 NCHW [B, C, H, W] -> bf16 padded NHWC (interior, channels [0, C)); one thread per (pixel, channel)
__global__ void __launch_bounds__(256) nchw_to_padded_kernel(const float* __restrict__ src, int batch, int c, int h, int w,
                                                             __nv_bfloat16* __restrict__ dst, int dcs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)batch * c * h * w;
  if (i >= total) return;
  const int x = (int)(i % w);
  const int y = (int)((i / w) % h);
  const int ch = (int)((i / ((size_t)w * h)) % c);
  const int b = (int)(i / ((size_t)w * h * c));
  dst[pad_off(b, y, x, h, w, dcs) + ch] = __float2bfloat16_rn(src[i]);
}

// im2col of the 3-channel image for the first 3x3/stride-1/pad-1 conv: column index = c*9 + kh*3 + kw (the flattening
// of nn.Conv2d.weight[co]), 27 real + zero padding to 64 channels, bf16 padded NHWC
__global__ void __launch_bounds__(256) im2col_first_kernel(const float* __restrict__ img, int batch, int h, int w,
                                                           __nv_bfloat16* __restrict__ dst /*cs = 64*/) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)batch * h * w) return;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const int b = (int)(pix / ((size_t)w * h));
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; k++) v[k] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
      for (int dx = 0; dx < 3; dx++) {
        const int yy = y + dy - 1, xx = x + dx - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v[c * 9 + dy * 3 + dx] = __ldg(img + (((size_t)b * 3 + c) * h + yy) * w + xx);
      }
  __nv_bfloat16* o = dst + pad_off(b, y, x, h, w, 64);
#pragma unroll
  for (int q = 0; q < 4; q++) *reinterpret_cast<uint4*>(o + q * 8) = pack8(v + q * 8);
#pragma unroll
  for (int q = 4; q < 8; q++) *reinterpret_cast<uint4*>(o + q * 8) = make_uint4(0, 0, 0, 0);
}

static inline Geo mk_geo(int batch, int h, int w, int c) {
  Geo g;
  g.batch = batch; g.h = h; g.w = w; g.c = c;
  return g;
}
static inline size_t n_items(const Geo& g) { return (size_t)g.batch * g.h * g.w * (g.c >> 3); }

}  // namespace ryolo

using namespace ryolo;

#define GEO_CHECK() RYOLO_ARG_CHECK(batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && c <= 2048)

extern "C" int ryolo_bn_stats(const void* z, int z_cstride, int batch, int h, int w, int c, float* sums, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(z && sums);
  GEO_CHECK();
  const Geo g = mk_geo(batch, h, w, c);
  RYOLO_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, stream));
  const size_t items = n_items(g);
  const int cgs = c >> 3;
  int per_block = ((256 / cgs) * cgs) * 64;
  unsigned blocks = (unsigned)((items + per_block - 1) / per_block);
  bn_stats_kernel<<<blocks, 256, 2 * c * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(z), z_cstride, g, sums,
                                                                   per_block);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_bn_act_fwd(const void* z, int z_cstride, int batch, int h, int w, int c, const float* scale,
                                const float* shift, float slope, int has_act, const void* residual, int res_cstride,
                                void* y, int y_cstride, int upsample2x, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(z && scale && shift && y);
  GEO_CHECK();
  RYOLO_ARG_CHECK(!(residual && upsample2x));
  const Geo g = mk_geo(batch, h, w, c);
  const size_t items = n_items(g);
  bn_act_fwd_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(z), z_cstride, g, scale, shift, slope, has_act,
      static_cast<const __nv_bfloat16*>(residual), res_cstride, static_cast<__nv_bfloat16*>(y), y_cstride, upsample2x);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_bn_act_bwd(const void* dy, int dy_cstride, int upsample2x, void* z_dz, int z_cstride, int batch, int h,
                                int w, int c, const float* scale, const float* shift, const float* mean,
                                const float* invstd, float slope, int has_act, int has_bn, float* sums, void* gres,
                                int gres_cstride, int gres_accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(dy && z_dz && scale && shift && mean && invstd && sums);
  GEO_CHECK();
  RYOLO_ARG_CHECK(!(gres && upsample2x));
  const Geo g = mk_geo(batch, h, w, c);
  const size_t items = n_items(g);
  RYOLO_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(float) * (2 * c + 1), stream));
  const int cgs = c >> 3;
  const int per_block = ((256 / cgs) * cgs) * 64;
  bn_act_bwd_reduce_kernel<<<(unsigned)((items + per_block - 1) / per_block), 256, (2 * c + 1) * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(dy), dy_cstride, upsample2x, static_cast<const __nv_bfloat16*>(z_dz), z_cstride, g,
      scale, shift, mean, invstd, slope, has_act, sums, per_block);
  RYOLO_LAUNCH_CHECK();
  const float inv_n = 1.0f / ((float)batch * h * w);
  bn_act_bwd_apply_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dy), dy_cstride, upsample2x, static_cast<__nv_bfloat16*>(z_dz), z_cstride, g, scale,
      shift, mean, invstd, slope, has_act, has_bn, sums, inv_n, static_cast<__nv_bfloat16*>(gres), gres_cstride,
      gres_accumulate);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_zero_insert2x(const void* src, int src_cstride, int batch, int h, int w, int c, void* dst,
                                   int dst_cstride, int dst_h, int dst_w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(src && dst);
  GEO_CHECK();
  RYOLO_ARG_CHECK(dst_h >= 2 * h - 1 && dst_w >= 2 * w - 1);
  const Geo g = mk_geo(batch, h, w, c);
  const size_t items = n_items(g);
  zero_insert2x_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src),
                                                                           src_cstride, g, static_cast<__nv_bfloat16*>(dst),
                                                                           dst_cstride, dst_h, dst_w);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_nchw_to_padded(const float* src, int batch, int c, int h, int w, void* dst, int dst_cstride,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(src && dst && batch > 0 && c > 0 && h > 0 && w > 0 && dst_cstride >= c);
  const size_t total = (size_t)batch * c * h * w;
  nchw_to_padded_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, batch, c, h, w,
                                                                            static_cast<__nv_bfloat16*>(dst), dst_cstride);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_im2col_first(const float* img, int batch, int h, int w, void* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(img && dst && batch > 0 && h > 0 && w > 0);
  const size_t npix = (size_t)batch * h * w;
  im2col_first_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(img, batch, h, w, static_cast<__nv_bfloat16*>(dst));
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
