// First layer of the Darknet stacks (reference: model/models.py:45-74, block 0 of cfg/yolov3.cfg): fp32 NCHW image,
// cin = 3, 3x3 / stride 1 / pad 1, BN folded, leaky -> bf16 padded NHWC.  K = 27 is far too thin to stream through the
// TMA pipeline of conv.cu (its A operand would be a 27-channel im2col buffer twenty times the size of the image), and
// on the CUDA cores the 20 GFLOP of a batch-32 forward cost 0.84 ms; here the im2col row of each pixel is built in
// registers straight from the image and the contraction runs on the tensor pipe:
//
//   A tile [128 pixels x 64]: per pixel one 128-byte row  [hi(27) 0(5) lo(27) 0(5)]  of bf16, hi = bf16(v),
//   lo = bf16(v - hi): the image enters with 16 mantissa bits although the MMA operands are bf16;
//   B tile [COUT x 64]: [w(27) 0(5) w(27) 0(5)], w = bf16(weight * bn_scale) like every other layer's weights;
//   D = A * B^T: tcgen05.mma kind::f16, M = 128, N = COUT, 4 x K=16, fp32 accumulator in TMEM (32 columns).
//
// Both tiles are written by the threads themselves in the canonical K-major SWIZZLE_128B layout (16-byte chunk c of
// row r at chunk position c ^ (r & 7)), made visible to the async proxy with fence.proxy.async.  128 threads per CTA,
// thread = pixel = TMEM lane; several CTAs per SM overlap each other's load / MMA / store phases.  The kernel is bound by
// the write of the output (64 B per pixel): 0.76 GB at batch 32.
//
// With S2D the output goes directly into the space-to-depth buffer of the following 3x3/stride-2 layer
// (xs[b, y/2, x/2, ((y&1)*2 + (x&1))*COUT + c], see ryolo_space_to_depth), which removes one write and one read of
// the largest activation of the network.
#include "common.cuh"
#include "tc05.cuh"

namespace ryolo {

constexpr int FIRST_THREADS = 128;

struct FirstSmem {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBOffset = kABytes;          // up to 32 rows x 128 B
  static constexpr int kBarOffset = kBOffset + 32 * 128;
  static constexpr int kBiasOffset = kBarOffset + 16;
  static constexpr int kTotal = kBiasOffset + 32 * 4 + 1024;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int COUT, bool S2D>
__global__ void __launch_bounds__(FIRST_THREADS) conv_first_tc_kernel(const float* __restrict__ img, int batch, int h,
                                                                      int w, const float* __restrict__ weight,
                                                                      const float* __restrict__ bias, float slope,
                                                                      __nv_bfloat16* __restrict__ out, int out_cs) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_addr = smem_base, b_addr = smem_base + FirstSmem::kBOffset, bar = smem_base + FirstSmem::kBarOffset;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + FirstSmem::kBarOffset + 8);
  float* s_bias = reinterpret_cast<float*>(smem_gen + FirstSmem::kBiasOffset);
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- one-time: weights -> swizzled bf16 B tile, bias, barrier, TMEM ----
  for (int e = tid; e < COUT * 8; e += FIRST_THREADS) {
    const int n = e >> 3, c = e & 7;
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k0 = (c * 8 + 2 * j) & 31, k1 = k0 + 1;   // hi and lo halves carry the same weights
      const float w0 = k0 < 27 ? weight[n * 27 + k0] : 0.f;
      const float w1 = k1 < 27 ? weight[n * 27 + k1] : 0.f;
      pk[j] = pack_bf16(w0, w1);
    }
    *reinterpret_cast<uint4*>(smem_gen + FirstSmem::kBOffset + n * 128 + ((c ^ (n & 7)) << 4)) =
        make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (tid < COUT) s_bias[tid] = bias[tid];
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(32)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t npix = (size_t)batch * h * w;
  const int ntiles = (int)((npix + 127) / 128);
  const size_t plane = (size_t)h * w;
  uint32_t phase = 0;
  constexpr uint32_t idesc = make_idesc(128, COUT);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t pix = (size_t)tile * 128 + tid;
    const bool valid = pix < npix;
    int x = 0, y = 0, b = 0;
    float v[27];
    if (valid) {
      x = (int)(pix % w);
      y = (int)((pix / w) % h);
      b = (int)(pix / plane);
      const float* ib = img + (size_t)b * 3 * plane;
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
          for (int dx = 0; dx < 3; dx++) {
            const int yy = y + dy - 1, xx = x + dx - 1;
            v[c * 9 + dy * 3 + dx] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(ib + c * plane + (size_t)yy * w + xx) : 0.f;
          }
    } else {
#pragma unroll
      for (int k = 0; k < 27; k++) v[k] = 0.f;
    }
    // hi / lo split and the swizzled row store
    {
      float hi[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; k++) {
        hi[k] = k < 27 ? __bfloat162float(__float2bfloat16_rn(v[k < 27 ? k : 0])) : 0.f;
        lo[k] = k < 27 ? v[k < 27 ? k : 0] - hi[k] : 0.f;
      }
      unsigned char* row = smem_gen + tid * 128;
      const int sw = tid & 7;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          ph[j] = pack_bf16(hi[c * 8 + 2 * j], hi[c * 8 + 2 * j + 1]);
          pl[j] = pack_bf16(lo[c * 8 + 2 * j], lo[c * 8 + 2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(row + (((c + 4) ^ sw) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(b_addr);
#pragma unroll
      for (int k = 0; k < 4; k++) tc_mma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t acc[32];
    tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
    tc_wait_ld();
    if (valid) {
      __nv_bfloat16* o;
      if (S2D)
        o = out + (((size_t)b * (h / 2 + 2) + (y >> 1) + 1) * (w / 2 + 2) + (x >> 1) + 1) * out_cs +
            ((y & 1) * 2 + (x & 1)) * COUT;
      else
        o = out + (((size_t)b * (h + 2) + y + 1) * (w + 2) + x + 1) * out_cs;
#pragma unroll
      for (int g = 0; g < COUT / 8; g++) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float f0 = __uint_as_float(acc[g * 8 + 2 * e]) + s_bias[g * 8 + 2 * e];
          float f1 = __uint_as_float(acc[g * 8 + 2 * e + 1]) + s_bias[g * 8 + 2 * e + 1];
          f0 = f0 > 0.f ? f0 : slope * f0;
          f1 = f1 > 0.f ? f1 : slope * f1;
          pk[e] = pack_bf16(f0, f1);
        }
        reinterpret_cast<uint4*>(o)[g] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      if (!S2D)   // zero the padding channels [COUT, out_cs) (a wider buffer's 64-wide K chunk must read zeros)
        for (int co = COUT; co < out_cs; co += 8) *reinterpret_cast<uint4*>(o + co) = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();   // the next tile's MMA (after the next __syncthreads) overwrites the accumulator
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
  }
}

template <int COUT, bool S2D>
static int launch_first(const float* img, int batch, int h, int w, const float* weight, const float* bias, float slope,
                        void* y, int cs, cudaStream_t stream) {
  RYOLO_SMEM_OPT_IN((conv_first_tc_kernel<COUT, S2D>), FirstSmem::kTotal);
  const int sms = device_sm_count();
  const size_t npix = (size_t)batch * h * w;
  const long long ntiles = (long long)((npix + 127) / 128);
  RYOLO_ARG_CHECK(ntiles < (1ll << 31));
  const long long want = (long long)sms * 8;      // 8 co-resident CTAs per SM (22 KB smem, 32 TMEM columns each)
  const unsigned grid = (unsigned)(ntiles < want ? ntiles : want);
  conv_first_tc_kernel<COUT, S2D><<<grid, FIRST_THREADS, FirstSmem::kTotal, stream>>>(
      img, batch, h, w, weight, bias, slope, static_cast<__nv_bfloat16*>(y), cs);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

}  // namespace ryolo

using namespace ryolo;

extern "C" int ryolo_conv_first_fwd(const float* img, int batch, int h, int w, const float* weight, const float* bias,
                                    int cout, float slope, void* y, int cout_stride, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(img && weight && bias && y && batch > 0 && h > 0 && w > 0);
  RYOLO_ARG_CHECK(cout == 32 || cout == 16);
  RYOLO_ARG_CHECK(cout_stride >= cout && cout_stride % 8 == 0);
  RYOLO_ARG_CHECK((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  if (cout == 32) return launch_first<32, false>(img, batch, h, w, weight, bias, slope, y, cout_stride, stream);
  return launch_first<16, false>(img, batch, h, w, weight, bias, slope, y, cout_stride, stream);
}

extern "C" int ryolo_conv_first_s2d_fwd(const float* img, int batch, int h, int w, const float* weight,
                                        const float* bias, int cout, float slope, void* xs, int xs_cstride,
                                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(img && weight && bias && xs && batch > 0 && h > 0 && w > 0);
  RYOLO_ARG_CHECK(h % 2 == 0 && w % 2 == 0);
  RYOLO_ARG_CHECK(cout == 32 || cout == 16);
  RYOLO_ARG_CHECK(xs_cstride >= 4 * cout && xs_cstride % 8 == 0);
  RYOLO_ARG_CHECK((reinterpret_cast<uintptr_t>(xs) & 15) == 0);
  if (cout == 32) return launch_first<32, true>(img, batch, h, w, weight, bias, slope, xs, xs_cstride, stream);
  return launch_first<16, true>(img, batch, h, w, weight, bias, slope, xs, xs_cstride, stream);
}
