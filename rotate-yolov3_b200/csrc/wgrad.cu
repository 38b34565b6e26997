// Weight gradient of the conv blocks as an sm_100a tcgen05 GEMM whose K dimension is the PIXEL index:
//     dW[tap][co][ci] = sum_p dz[p, co] * x[p + off_tap, ci]          (flat padded-NHWC pixel index p, zero halos)
// the adjoint of the forward implicit GEMM in conv.cu (reference: autograd through nn.Conv2d, train.py:278-282).
// Both operands are read in their natural padded-NHWC layout -- 64 channels (128 B) contiguous per pixel row -- which
// makes them "MN-major" UMMA operands: TMA boxes [64 channels x 64 pixels] land in shared memory as 128B-swizzled
// rows = K, and the descriptors say so (a_major = b_major = MN).  No transposed copies of activations or gradients.
// The tap shift is again only a row-coordinate offset of the x box.  Split-K over CTAs, fp32 TMEM accumulation,
// fp32 atomics into the (pre-zeroed) dW.  Stride-2 layers pass the zero-inserted dz at input resolution.
#include "common.cuh"
#include "tc05.cuh"

namespace ryolo {

constexpr int WG_BM = 128;   // filters per tile (UMMA M)
constexpr int WG_BK = 64;    // pixels per k-step
constexpr int WG_THREADS = 256;
constexpr int WG_STAGES = 4;

struct WgradParams {
  int np, wp, taps;
  int cout_pad, cin_pad;
  int m_tiles, n_tiles, ksplit, ksteps_total, ksteps_per_split;
  float* dw;   // [taps][cout_pad][cin_pad] fp32, accumulated with atomics
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x, const WgradParams p) {
  constexpr int kABytes = WG_BM * WG_BK * 2;   // two [64 ch x 64 px] boxes
  constexpr int kBBytes = BN * WG_BK * 2;      // BN/64 boxes
  constexpr int kStageBytes = kABytes + kBBytes;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + WG_STAGES * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * WG_STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + WG_STAGES * kStageBytes + (2 * WG_STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile decode: blockIdx.x -> (split, nt, mt, tap)
  int t = blockIdx.x;
  const int split = t % p.ksplit; t /= p.ksplit;
  const int nt = t % p.n_tiles; t /= p.n_tiles;
  const int mt = t % p.m_tiles; t /= p.m_tiles;
  const int tap = t;
  const int k_begin = split * p.ksteps_per_split;
  const int k_end = min(p.ksteps_total, k_begin + p.ksteps_per_split);
  const int k_iters = k_end - k_begin;
  int off = 0;
  if (p.taps == 9) off = (tap / 3 - 1) * p.wp + (tap % 3 - 1);
  else if (p.taps == 4) off = (tap / 2 - 1) * p.wp + (tap % 2 - 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(BN < 32 ? 32 : BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (k_iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kk = 0; kk < k_iters; kk++) {
          const int prow = (k_begin + kk) * WG_BK;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * kStageBytes;
          mbar_expect_tx(full_bar(stage), kStageBytes);
#pragma unroll
          for (int g = 0; g < WG_BM / 64; g++) tma_load_2d(sa + g * 8192, &map_dz, full_bar(stage), mt * WG_BM + g * 64, prow);
#pragma unroll
          for (int g = 0; g < BN / 64; g++)
            tma_load_2d(sa + kABytes + g * 8192, &map_x, full_bar(stage), nt * BN + g * 64, prow + off);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_mn(WG_BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        for (int kk = 0; kk < k_iters; kk++) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
#pragma unroll
          for (int k = 0; k < WG_BK / 16; k++) {
            // 16 pixels = two 8-row groups = 2048 bytes along K
            const uint64_t adesc = make_smem_desc_mn(sa + k * 2048, 8192);
            const uint64_t bdesc = make_smem_desc_mn(sa + kABytes + k * 2048, 8192);
            tc_mma_f16(tmem_base, adesc, bdesc, idesc, (kk | k) != 0);
          }
          tc_commit(empty_bar(stage));
          if (kk == k_iters - 1) tc_commit(done_bar);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp >= 4) {
      const int q = warp & 3;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int co = mt * WG_BM + q * 32 + lane;
      float* drow = p.dw + ((size_t)tap * p.cout_pad + co) * p.cin_pad + nt * BN;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tc_ld32(t_row + c0, v);
        tc_wait_ld();
        if (co < p.cout_pad) {
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (nt * BN + c0 + j < p.cin_pad) atomicAdd(drow + c0 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN < 32 ? 32 : BN) : "memory");
  }
}

template <int BN>
static int launch_wgrad(const CUtensorMap& mdz, const CUtensorMap& mx, const WgradParams& p, cudaStream_t stream) {
  constexpr int smem = WG_STAGES * (WG_BM * WG_BK * 2 + BN * WG_BK * 2) + (2 * WG_STAGES + 1) * 8 + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    RYOLO_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = p.taps * p.m_tiles * p.n_tiles * p.ksplit;
  conv_wgrad_kernel<BN><<<grid, WG_THREADS, smem, stream>>>(mdz, mx, p);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

}  // namespace ryolo

using namespace ryolo;

extern "C" int ryolo_conv_wgrad(const void* dz, int dz_cstride, int cout_pad, const void* x, int x_cstride, int cin_pad,
                                int batch, int in_h, int in_w, int ksize, float* dw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(dz && x && dw && batch > 0 && in_h > 0 && in_w > 0);
  RYOLO_ARG_CHECK(ksize == 1 || ksize == 3 || ksize == 2);
  RYOLO_ARG_CHECK(cout_pad > 0 && cout_pad % 64 == 0 && cin_pad > 0 && cin_pad % 64 == 0);
  // strides narrower than the padded tile are fine: TMA zero-fills the box beyond the inner extent
  RYOLO_ARG_CHECK(dz_cstride > 0 && dz_cstride % 8 == 0 && x_cstride > 0 && x_cstride % 8 == 0);
  WgradParams p;
  const long long np = (long long)batch * (in_h + 2) * (in_w + 2);
  RYOLO_ARG_CHECK(np < (1ll << 31) - 4096);
  p.np = (int)np;
  p.wp = in_w + 2;
  p.taps = ksize * ksize;
  p.cout_pad = cout_pad;
  p.cin_pad = cin_pad;
  const int bn = cin_pad >= 192 ? 256 : cin_pad;   // 64, 128 or 256
  p.m_tiles = (cout_pad + WG_BM - 1) / WG_BM;
  p.n_tiles = (cin_pad + bn - 1) / bn;
  p.ksteps_total = (p.np + WG_BK - 1) / WG_BK;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    RYOLO_CUDA_TRY(cudaGetDevice(&dev));
    RYOLO_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = p.taps * p.m_tiles * p.n_tiles;
  int ksplit = (4 * num_sms + tiles - 1) / tiles;                  // ~4 waves
  const int max_split = (p.ksteps_total + 15) / 16;                // at least 16 k-steps per CTA
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  p.ksteps_per_split = (p.ksteps_total + ksplit - 1) / ksplit;
  p.ksplit = (p.ksteps_total + p.ksteps_per_split - 1) / p.ksteps_per_split;
  p.dw = dw;
  CUtensorMap mdz, mx;
  // inner extent = channel stride of the buffer (reads beyond cout_pad/cin_pad are masked at the store)
  int st = encode_map_2d(&mdz, dz, (uint64_t)dz_cstride, (uint64_t)p.np, (uint64_t)dz_cstride * 2, 64, WG_BK);
  if (st != RYOLO_OK) return st;
  st = encode_map_2d(&mx, x, (uint64_t)x_cstride, (uint64_t)p.np, (uint64_t)x_cstride * 2, 64, WG_BK);
  if (st != RYOLO_OK) return st;
  if (bn == 256) return launch_wgrad<256>(mdz, mx, p, stream);
  if (bn == 128) return launch_wgrad<128>(mdz, mx, p, stream);
  return launch_wgrad<64>(mdz, mx, p, stream);
}
