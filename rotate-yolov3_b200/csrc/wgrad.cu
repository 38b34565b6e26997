// Weight gradient of the conv blocks as an sm_100a tcgen05 GEMM whose K dimension is the PIXEL index:
//     dW[tap][co][ci] = sum_p dz[p, co] * x[p + off_tap, ci]          (flat padded-NHWC pixel index p, zero halos)
// the adjoint of the forward implicit GEMM in conv.cu (reference: autograd through nn.Conv2d, train.py:278-282).
// Both operands are read in their natural padded-NHWC layout -- 64 channels (128 B) contiguous per pixel row -- which
// makes them "MN-major" UMMA operands: TMA boxes [64 channels x 64 pixels] land in shared memory as 128B-swizzled
// rows = K, and the descriptors say so (a_major = b_major = MN).  No transposed copies of activations or gradients.
// The tap shift is again only a row-coordinate offset of the x box.  Split-K over CTAs, fp32 TMEM accumulation,
// fp32 atomics into the (pre-zeroed) dW.  Stride-2 layers pass the zero-inserted dz at input resolution.
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace ryolo {

constexpr int WG_BM = 128;   // filters per tile (UMMA M)
constexpr int WG_BK = 64;    // pixels per k-step
constexpr int WG_THREADS = 256;

struct WgradParams {
  int np, wp, taps;   // taps: 1, 4 (2x2, offsets -1..0) or 9 (3x3, offsets -1..1)
  int tap_groups;     // taps / TG: a CTA accumulates the TG taps of one kernel row (they share the dz tile)
  int cout_pad, cin_pad;
  int m_tiles, n_tiles, ksplit, ksteps_total, ksteps_per_split;
  float* dw;   // [taps][cout_pad][cin_pad] fp32, accumulated with atomics
};

// TG = taps per CTA.  For thin layers (cin <= 128) the per-tap GEMM is bound by the L2 -> SM operand stream (A 16 KB +
// B <= 16 KB per k-step against <= 325 cycles of MMA); the TG taps of one kernel row share the dz (A) tile, so a CTA
// that keeps TG accumulators in TMEM streams A once per row instead of once per tap.
template <int BN, int TG>
struct WgradCfg {
  static constexpr int kABytes = WG_BM * WG_BK * 2;   // two [64 ch x 64 px] boxes
  static constexpr int kBBytes = BN * WG_BK * 2;      // BN/64 boxes per tap
  static constexpr int kStageBytes = kABytes + TG * kBBytes;
  static constexpr int kStages = kStageBytes > 49152 ? 3 : 4;
  static constexpr int kTmemCols = TG * BN <= 32 ? 32 : (TG * BN <= 64 ? 64 : (TG * BN <= 128 ? 128 : (TG * BN <= 256 ? 256 : 512)));
  static constexpr int kSmem = kStages * kStageBytes + (2 * kStages + 1) * 8 + 16 + 1024;
};

template <int BN, int TG>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x, const WgradParams p) {
  using C = WgradCfg<BN, TG>;
  constexpr int kABytes = C::kABytes, kBBytes = C::kBBytes, kStageBytes = C::kStageBytes, WG_STAGES = C::kStages;
  static_assert(TG * BN <= 512, "accumulators must fit the 512 TMEM columns");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + WG_STAGES * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * WG_STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + WG_STAGES * kStageBytes + (2 * WG_STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile decode: blockIdx.x -> (tap group, nt, mt, split), tap group fastest: the CTAs that are resident together cover
  // ALL taps and channel tiles of a few pixel ranges, so dz / x stream from HBM once and the other reads hit L2 (with the
  // split index fastest every tap pass re-streamed both tensors from HBM: 9x the traffic on 3x3 layers)
  int t = blockIdx.x;
  const int tg = t % p.tap_groups; t /= p.tap_groups;
  const int nt = t % p.n_tiles; t /= p.n_tiles;
  const int mt = t % p.m_tiles; t /= p.m_tiles;
  const int split = t;
  const int k_begin = split * p.ksteps_per_split;
  const int k_end = min(p.ksteps_total, k_begin + p.ksteps_per_split);
  const int k_iters = k_end - k_begin;
  // pixel offset of tap j of this group: 3x3: (dy, dx) = (tg - 1, j - 1); 2x2: (tg - 1, j - 1); 1x1: 0.  With TG == 1
  // the group index IS the tap index.
  auto tap_of = [&](int j) { return TG == 1 ? tg : tg * TG + j; };
  auto tap_off = [&](int j) {
    const int tap = tap_of(j);
    if (p.taps == 9) return (tap / 3 - 1) * p.wp + (tap % 3 - 1);
    if (p.taps == 4) return (tap / 2 - 1) * p.wp + (tap % 2 - 1);
    return 0;
  };

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
                 "r"(C::kTmemCols)
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
          for (int j = 0; j < TG; j++) {
            const int off = tap_off(j);
#pragma unroll
            for (int g = 0; g < BN / 64; g++)
              tma_load_2d(sa + kABytes + j * kBBytes + g * 8192, &map_x, full_bar(stage), nt * BN + g * 64, prow + off);
          }
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
          for (int j = 0; j < TG; j++) {
#pragma unroll
            for (int k = 0; k < WG_BK / 16; k++) {
              // 16 pixels = two 8-row groups = 2048 bytes along K
              const uint64_t adesc = make_smem_desc_mn(sa + k * 2048, 8192);
              const uint64_t bdesc = make_smem_desc_mn(sa + kABytes + j * kBBytes + k * 2048, 8192);
              tc_mma_f16(tmem_base + j * BN, adesc, bdesc, idesc, (kk | k) != 0);
            }
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
#pragma unroll 1
      for (int j = 0; j < TG; j++) {
        float* drow = p.dw + ((size_t)tap_of(j) * p.cout_pad + co) * p.cin_pad + nt * BN;
        const uint32_t t_row = tmem_base + j * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tc_ld32(t_row + c0, v);
          tc_wait_ld();
          if (co < p.cout_pad && nt * BN + c0 < p.cin_pad) {   // cin_pad is a multiple of 64: whole chunks are in or out
#pragma unroll
            for (int e = 0; e < 32; e += 4)   // 16-byte vector reductions: 4x fewer L2 atomic transactions
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + e),
                           "f"(__uint_as_float(v[e])), "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])),
                           "f"(__uint_as_float(v[e + 3]))
                           : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
  }
}

template <int BN, int TG>
static int launch_wgrad(const CUtensorMap& mdz, const CUtensorMap& mx, const WgradParams& p, cudaStream_t stream) {
  using C = WgradCfg<BN, TG>;
  RYOLO_SMEM_OPT_IN((conv_wgrad_kernel<BN, TG>), C::kSmem);
  const int grid = p.tap_groups * p.m_tiles * p.n_tiles * p.ksplit;
  conv_wgrad_kernel<BN, TG><<<grid, WG_THREADS, C::kSmem, stream>>>(mdz, mx, p);
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
  const int num_sms = gemm_sm_count();
  // taps per CTA: the kernel row (3 or 2 taps) when the accumulators fit TMEM and the layer is thin enough to be
  // operand-stream bound; RYOLO_WGRAD_TG=1 forces one tap per CTA (measurement knob)
  static int force_tg = -1;
  if (force_tg < 0) {
    const char* e = getenv("RYOLO_WGRAD_TG");
    force_tg = e ? atoi(e) : 0;
  }
  int tg = 1;
  if (bn <= 128 && force_tg != 1) tg = p.taps == 9 ? 3 : (p.taps == 4 ? 2 : 1);
  p.tap_groups = p.taps / tg;
  const int tiles = p.tap_groups * p.m_tiles * p.n_tiles;
  // split-K: every (tile, split) CTA ends with tile-size fp32 reductions into dW and cannot overlap them with its
  // MMAs, so use the FEWEST waves that still fill the SMs (>= 90 % of the last wave), not a fixed multiple
  static int force_waves = -1;
  if (force_waves < 0) {
    const char* e = getenv("RYOLO_WGRAD_WAVES");   // measurement knob
    force_waves = e ? atoi(e) : 0;
  }
  int ksplit = 1;
  {
    double best_eff = 0.0;
    for (int waves = 1; waves <= 4; waves++) {
      if (force_waves > 0 && waves != force_waves) continue;
      int ks = waves * num_sms / tiles;
      if (ks < 1) ks = 1;
      const int ctas = tiles * ks;
      const double eff = (double)ctas / ((double)((ctas + num_sms - 1) / num_sms) * num_sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; ksplit = ks; }
      if (eff >= 0.9) break;
    }
  }
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
  if (bn == 256) return launch_wgrad<256, 1>(mdz, mx, p, stream);
  if (bn == 128) {
    if (tg == 3) return launch_wgrad<128, 3>(mdz, mx, p, stream);
    if (tg == 2) return launch_wgrad<128, 2>(mdz, mx, p, stream);
    return launch_wgrad<128, 1>(mdz, mx, p, stream);
  }
  if (tg == 3) return launch_wgrad<64, 3>(mdz, mx, p, stream);
  if (tg == 2) return launch_wgrad<64, 2>(mdz, mx, p, stream);
  return launch_wgrad<64, 1>(mdz, mx, p, stream);
}
