// Conv2d -> (folded) BatchNorm -> PReLU [-> + residual] as an sm_100a implicit GEMM.
//
// Replaces the nn.Sequential(Conv2d, BatchNorm2d, PReLU) blocks the reference builds in create_modules
// (model/models.py:49-66) plus the shortcut add (models.py:281-282) and the nearest x2 upsample (models.py:93-94)
// that follow some of them; BN is folded as in utils/torch_utils.py:45-69 (scale into the weights at pack time,
// shift into the epilogue bias).
//
// Layout.  Activations are bf16 "padded NHWC": [B][H+2][W+2][Cs] with a one-pixel ZERO halo (zeroed once at
// allocation, never written afterwards).  With the halo in memory, a 3x3 / stride-1 / pad-1 convolution is nine
// GEMMs that differ only by a constant shift of the FLAT pixel index  p = (b*(H+2) + y)*(W+2) + x :
//        out[p, :] = sum_{tap=(dy,dx)}  A[p + dy*(W+2) + dx, :] * W_tap            (valid for every interior p)
// so the A operand of tap t is a plain 2-D TMA box [128 pixels x 64 channels] whose row coordinate is shifted by
// the tap offset; rows that fall outside [0, NP) are zero-filled by TMA.  No im2col buffer, no gather.  The GEMM
// also produces values at halo positions; the epilogue simply does not store them.  Cost of that: (H+2)(W+2)/HW
// extra MMA rows (5 % at 76^2, 11 % at 38^2, 22 % at 19^2).
// Stride-2 3x3 layers run on a space-to-depth copy of their input as 2x2-tap stride-1 convs (ksize = 2); odd sizes
// fall back to the stride-1 GEMM on the input grid that stores only the even pixels.
//
// Kernel.  Persistent, warp-specialised, one CTA per SM, tile = 128 pixels x BN filters, BK = 64 channels:
//   warp 0   TMA producer: per k-step one A box (shifted by the tap offset) + one B box into a 128B-swizzled
//            shared-memory stage, completion on an mbarrier (expect_tx);
//   warp 1   MMA issuer: one elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16)
//            per stage, accumulating in TMEM; tcgen05.commit releases the stage / publishes the accumulator;
//   warp 2   TMEM allocator (2 accumulator stages x BN fp32 columns);
//   warps 4-11 epilogue, two teams of four warps that alternate over 64-filter groups of the accumulator:
//            tcgen05.ld -> +bias -> PReLU -> +residual -> bf16.  For the common case (bf16 output on the same padded
//            grid) the group is staged in a 128B-swizzled shared-memory buffer and leaves through a TMA bulk store
//            (whole 128-byte lines; halo rows staged as zeros), and the residual group ARRIVES in that same buffer
//            through a TMA load prefetched one group ahead, so the epilogue has no exposed global-memory latency
//            (measured: per-thread residual loads + 16-byte stores made the epilogue, not the MMA, the bottleneck of
//            every 3x3 body layer).  Up-sampling producers and the fp32 NCHW heads keep per-thread stores.
//            The epilogue overlaps the next tile's MMAs through the second accumulator stage.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "pack.cuh"
#include "tc05.cuh"

namespace ryolo {

constexpr int BM = 128;  // pixels per tile (UMMA M, TMEM lanes)
constexpr int BK = 64;   // channels per k-step = one 128-byte swizzle row of bf16
constexpr int CONV_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: two epilogue teams

// packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2): halves the instruction count of the epilogue math
__device__ __forceinline__ uint64_t f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// 32 accumulator columns of one row -> +bias -> PReLU -> +residual -> 32 bf16 (four 16-byte chunks).
// sb: the 32 biases (shared memory, warp-uniform address); res: the row's 128-byte staging line holding the residual
// group (chunk c at c ^ sw), hf selects its lower / upper 32 filters.  act: 0 none, 1 slope in [0, 1] (max(x, s*x),
// bit-identical to the select), 2 general slope.
template <bool RES>
__device__ __forceinline__ void epilogue_math32(const uint32_t (&v)[32], const float* sb, int act, float slope,
                                                const unsigned char* res, int hf, int sw, uint4 (&pk)[4]) {
  const uint64_t s2 = f32x2(slope, slope);
#pragma unroll
  for (int gg = 0; gg < 4; gg++) {
    const float4 b0 = *reinterpret_cast<const float4*>(sb + gg * 8), b1 = *reinterpret_cast<const float4*>(sb + gg * 8 + 4);
    const uint64_t bb[4] = {f32x2(b0.x, b0.y), f32x2(b0.z, b0.w), f32x2(b1.x, b1.y), f32x2(b1.z, b1.w)};
    uint4 rv = make_uint4(0, 0, 0, 0);
    if (RES) rv = *reinterpret_cast<const uint4*>(res + (((hf * 4 + gg) ^ sw) << 4));
    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      uint64_t x2 = add2(f32x2(__uint_as_float(v[gg * 8 + 2 * e]), __uint_as_float(v[gg * 8 + 2 * e + 1])), bb[e]);
      float lo, hi;
      if (act == 1) {
        const uint64_t y2 = mul2(x2, s2);
        float ylo, yhi;
        f32x2_unpack(x2, lo, hi);
        f32x2_unpack(y2, ylo, yhi);
        lo = fmaxf(lo, ylo);
        hi = fmaxf(hi, yhi);
        if (RES) x2 = f32x2(lo, hi);
      } else if (act == 2) {
        f32x2_unpack(x2, lo, hi);
        lo = lo > 0.f ? lo : slope * lo;
        hi = hi > 0.f ? hi : slope * hi;
        if (RES) x2 = f32x2(lo, hi);
      } else {
        f32x2_unpack(x2, lo, hi);
      }
      if (RES) {
        // bf16x2 word: low half = even filter.  bf16 -> fp32 is a 16-bit shift.
        x2 = add2(x2, f32x2(__uint_as_float(rw[e] << 16), __uint_as_float(rw[e] & 0xffff0000u)));
        f32x2_unpack(x2, lo, hi);
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
      o[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    pk[gg] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

struct ConvParams {
  int np;             // B * (H+2) * (W+2): GEMM M
  int hp, wp;         // padded input height / width
  int taps;           // 1, 9 (3x3, offsets -1..1) or 4 (2x2, offsets -1..0, or 0..1 when tap_flip)
  int tap_flip;
  int kchunks;        // cin_pad / 64
  int cout;           // real filters
  int cout_pad;       // rows per tap in the packed weight
  int n_tiles;        // cout_pad / BN
  int m_tiles;
  int stride;         // 1 or 2
  int oh, ow;         // output spatial size (unpadded)
  int out_cs;         // channel stride of the output buffer
  int store_cols;     // bf16 output: columns [0, store_cols) of the padded filter range are stored
  int res_cs;         // channel stride of the residual buffer
  int has_act, has_res, upsample2x, out_f32_nchw;
  int tma_store;      // bf16 stride-1 non-upsampled output: tiles leave (and residuals arrive) through TMA, 64-filter groups
  int arow_bo;        // A-row experiments: 1 = put the line shift into the descriptor's base-offset field as well
  int dbg;            // measurement knob RYOLO_CONV_DEBUG (scratch/conv_exp.py): timing experiments only, 0 in production
  float slope;
  const float* bias;               // [cout_pad]
  const __nv_bfloat16* residual;   // padded NHWC like the output, or null
  void* out;
  FastDiv d_ntiles, d_wp, d_hp;    // divisions of the epilogue's per-tile index arithmetic
};

// NBUF = group buffers per epilogue team.  2 lets a residual group prefetch while the other buffer is processed; the
// BN = 256 tile only has room for 4 pipeline stages with NBUF = 1, which is what its residual-free launches use.
//
// AROW ("one A load per kernel row").  The taps of one kernel row (dx = -1, 0, +1) read the SAME pixels shifted by one
// row of the flat pixel index, i.e. by one 128-byte line of the swizzled tile.  With AROW the producer loads ONE box of
// BM + 8 rows per (kernel row, k-chunk) and the MMA issuer addresses tap dx through a descriptor whose start is shifted
// by dx lines -- a third of the A traffic.  Motivation (ncu, profiles/r02_prof_thin_summary.csv): on the thin 304^2
// layers the tensor pipe is 21 % active and 43 % of all stall samples are epilogue warps waiting for the accumulator,
// i.e. the MMA warp starves on operands: the per-tap A boxes are DISTINCT data for every CTA (unlike the weights, which
// all CTAs request at the same time and L2 serves once), ~27 B/clk/SM of distinct L2 reads.  A and B live in separate
// rings (A stage = 18 KB every 3 taps, B stage = BN x 128 B every tap).
//
// BRES ("resident weights", thin layers).  With cout <= 64 there is ONE n-tile, so every tile of the CTA multiplies by the
// same <= 72 KB of packed weights; after AROW the per-tile weight re-load (9 x 8 KB) was the larger half of the L2 traffic
// of those layers.  With BRES the producer loads all (tap, k-chunk) weight blocks once per CTA into a resident region and
// the ring carries A only.
template <int BN, int NBUF, bool AROW, bool BRES = false>
struct ConvSmem {
  static constexpr int kStages = (BN == 256) ? (NBUF == 1 ? 4 : 3) : (BN == 128 ? 5 : 6);   // !AROW: (A, B) pairs
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // AROW rings
  static constexpr int kARowRows = BM + 8;
  static constexpr int kARowTx = kARowRows * 128;            // bytes one A-row box delivers
  static constexpr int kARowBytes = 18 * 1024;               // stage pitch (1024-byte aligned)
  static constexpr int kAStages = 3;
  static constexpr int kBStages = BRES ? 9 : ((BN == 256) ? (NBUF == 1 ? 4 : 3) : (BN == 128 ? 6 : 8));   // BRES: 9 resident blocks
  static constexpr int kTileBytes = AROW ? kAStages * kARowBytes + kBStages * kBBytes : kStages * kStageBytes;
  static constexpr int kBRingOffset = kAStages * kARowBytes;
  static constexpr int kGroupBytes = BM * 128;                        // one [128 pixels x 64 filters] bf16 group
  static constexpr int kStagingOffset = kTileBytes;                   // [team][buffer] group buffers
  static constexpr int kBarOffset = kStagingOffset + 2 * NBUF * kGroupBytes;
  // !AROW: full[kStages], empty[kStages];  AROW: afull[3], aempty[3], bfull[kBStages], bempty[kBStages];  then tfull[2],
  // tempty[2], rfull[4]
  static constexpr int kPipeBars = AROW ? 2 * kAStages + 2 * kBStages : 2 * kStages;
  static constexpr int kNumBars = kPipeBars + 8;
  static constexpr int kBiasOffset = kBarOffset + kNumBars * 8 + 16;  // [team][64] fp32: bias of the team's current group
  static constexpr int kTotal = kBiasOffset + 2 * 64 * 4 + 1024 /*alignment slack*/;
};

template <int BN, int NBUF, bool AROW, bool BRES = false>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_r, const ConvParams p) {
  using S = ConvSmem<BN, NBUF, AROW, BRES>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + S::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S::kStages + s); };
  // AROW rings (same barrier block, different carving)
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (S::kAStages + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * S::kAStages + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * S::kAStages + S::kBStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (S::kPipeBars + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (S::kPipeBars + 2 + a); };
  auto rfull_bar = [&](int i) { return bar_base + 8u * (S::kPipeBars + 4 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + S::kBarOffset + S::kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kPipeBars; s++) mbar_init(bar_base + 8u * s, 1);
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // one arrive per epilogue warp
    }
    for (int i = 0; i < 4; i++) mbar_init(rfull_bar(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // kernel-row decomposition of the taps (AROW): ky rows of kx taps; tap (dyi, dxi) has pixel offset
  // (dyi + d0) * wp + (dxi + d0), d0 = -1 (3x3, 2x2) or 0 (2x2 flipped)
  const int kx = p.taps == 9 ? 3 : 2, ky = kx;
  const int d0 = p.tap_flip ? 0 : -1;
  if (AROW && warp == 0) {
    // ===================== TMA producer, separate A-row / B rings =====================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      if (BRES) {
        // all weight blocks of the (single) n-tile, once: block index = tap * kchunks + kc, completion on bfull[0]
        const int nblk = p.taps * p.kchunks;
        mbar_expect_tx(bfull_bar(0), nblk * S::kBBytes);
        for (int blk = 0; blk < nblk; blk++)
          tma_load_2d(smem_base + S::kBRingOffset + blk * S::kBBytes, &map_b, bfull_bar(0), (blk % p.kchunks) * BK,
                      (blk / p.kchunks) * p.cout_pad);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int p0 = mt * BM;
        for (int dyi = 0; dyi < ky; dyi++) {
          for (int kc = 0; kc < p.kchunks; kc++) {
            mbar_wait(aempty_bar(as), aph ^ 1);
            mbar_expect_tx(afull_bar(as), S::kARowTx);
            tma_load_2d(smem_base + as * S::kARowBytes, &map_a, afull_bar(as), kc * BK, p0 + (dyi + d0) * p.wp + d0);
            if (++as == S::kAStages) { as = 0; aph ^= 1; }
            if (BRES) continue;
            for (int dxi = 0; dxi < kx; dxi++) {
              mbar_wait(bempty_bar(bs), bph ^ 1);
              mbar_expect_tx(bfull_bar(bs), S::kBBytes);
              tma_load_2d(smem_base + S::kBRingOffset + bs * S::kBBytes, &map_b, bfull_bar(bs), kc * BK,
                          (dyi * kx + dxi) * p.cout_pad + nt * BN);
              if (++bs == S::kBStages) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (AROW && warp == 1) {
    // ===================== MMA issuer, A-row form =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      if (BRES) {
        mbar_wait(bfull_bar(0), 0);          // the resident weights have landed
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accum = 0;
        for (int dyi = 0; dyi < ky; dyi++) {
          for (int kc = 0; kc < p.kchunks; kc++) {
            mbar_wait(afull_bar(as), aph);
            tc_fence_after();
            const uint32_t sa = smem_base + as * S::kARowBytes;
            for (int dxi = 0; dxi < kx; dxi++) {
              if (BRES) {
                bs = (dyi * kx + dxi) * p.kchunks + kc;
              } else {
                mbar_wait(bfull_bar(bs), bph);
                tc_fence_after();
              }
              // tap dxi = rows [dxi, dxi + 128) of the A-row box: the descriptor start is shifted by dxi 128-byte lines.
              // Measured on B200: the 128B swizzle is a function of the shared-memory ADDRESS bits (what TMA wrote at line
              // r is found at line r whatever the start), so the descriptor's base-offset field must stay 0 -- with
              // base offset = dxi every multi-tap layer is wrong (gpurun_out/r02_pytest_arow.log vs r02_pytest_arow_bo0.log)
              const uint64_t adesc = make_smem_desc(sa + dxi * 128) | (p.arow_bo ? ((uint64_t)dxi << 49) : 0ull);
              const uint64_t bdesc = make_smem_desc(smem_base + S::kBRingOffset + bs * S::kBBytes);
#pragma unroll
              for (int k = 0; k < BK / 16; k++) {
                tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                accum = 1;
              }
              if (!BRES) {
                tc_commit(bempty_bar(bs));
                if (++bs == S::kBStages) { bs = 0; bph ^= 1; }
              }
            }
            tc_commit(aempty_bar(as));
            if (++as == S::kAStages) { as = 0; aph ^= 1; }
          }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int p0 = mt * BM;
        for (int kk = 0; kk < k_iters; kk++) {
          const int tap = kk / p.kchunks, kc = kk % p.kchunks;
          int off = 0;
          if (p.taps == 9) off = (tap / 3 - 1) * p.wp + (tap % 3 - 1);
          else if (p.taps == 4) off = (tap / 2 - 1 + p.tap_flip) * p.wp + (tap % 2 - 1 + p.tap_flip);
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const bool skip_b = (p.dbg & 4) && !(tile == (int)blockIdx.x && kk < S::kStages);
          const bool skip_a = (p.dbg & 8) && !(tile == (int)blockIdx.x && kk < S::kStages);
          mbar_expect_tx(full_bar(stage), (skip_a ? 0 : S::kABytes) + (skip_b ? 0 : S::kBBytes));
          if (!skip_a) tma_load_2d(sa, &map_a, full_bar(stage), kc * BK, p0 + off);
          if (!skip_b) tma_load_2d(sa + S::kABytes, &map_b, full_bar(stage), kc * BK, tap * p.cout_pad + nt * BN);
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kk = 0; kk < k_iters; kk++) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) {
            // +32 bytes (16 bf16) along K inside the swizzle atom: start-address field += 2
            if (!(p.dbg & 16)) tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kk | k) != 0);
          }
          tc_commit(empty_bar(stage));  // frees the stage when these MMAs have read it
          if (kk == k_iters - 1) tc_commit(tfull_bar(acc));
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: two teams of four warps =====================
    // A tile's accumulator is drained in 64-filter GROUPS; group G = it * (BN/64) + g belongs to team G & 1, so the
    // two teams (one warp of each per SM sub-partition) overlap each other's TMEM-load / math / store latencies.
    constexpr int GPT = BN / 64;                  // groups per tile
    const int team = (warp - 4) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                // tile row = TMEM lane
    const int tid_t = threadIdx.x & 127;          // thread index within the team
    const bool issuer = tid_t == 0;               // first thread of the team issues its TMA traffic
    float* s_bias = reinterpret_cast<float*>(smem_gen + S::kBiasOffset) + team * 64;
    const int act = !p.has_act ? 0 : ((p.slope >= 0.f && p.slope <= 1.f) ? 1 : 2);
    const uint32_t bar_id = 1 + team;
    auto team_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    int acc = 0;
    uint32_t acc_phase = 0;
    // TMA-store mode state: two group buffers per team, used alternately.  A buffer first receives the residual
    // group (TMA load, prefetched one group ahead), is transformed in place by the team, and leaves through a TMA store.
    int buf = 0;                 // buffer of the next group to process
    uint32_t rphase[2] = {0, 0}; // parity of rfull[team*2 + b]
    int pit = GPT == 1 ? team : 0, pg = GPT == 1 ? 0 : team;   // prefetch cursor over this team's groups
    int pbuf = 0;
    auto group_buf = [&](int b) { return S::kStagingOffset + (team * NBUF + b) * S::kGroupBytes; };
    // issue the residual load of the next live group of this team.  Called when the team STARTS a group, for the group
    // after it (into the other buffer), so the load has a whole group's processing time to land.
    auto prefetch_residual = [&]() {
      for (;;) {
        const int tile = blockIdx.x + pit * gridDim.x;
        if (tile >= total_tiles) return;
        int mt, nt;
        fast_divmod(tile, p.d_ntiles, mt, nt);
        const int colbase = nt * BN + pg * 64;
        const bool live = colbase < p.store_cols;
        if (live && issuer) {
          // the buffer's previous TMA store (the team's latest one) must have finished READING it
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_expect_tx(rfull_bar(team * 2 + pbuf), S::kGroupBytes);
          tma_load_2d(smem_base + group_buf(pbuf), &map_r, rfull_bar(team * 2 + pbuf), colbase, mt * BM);
        }
        if (GPT == 1) pit += 2;
        else { pg += 2; if (pg >= GPT) { pg = team; pit += 1; } }
        if (live) { pbuf ^= (NBUF - 1); return; }
      }
    };
    const bool tma_res = NBUF == 2 && p.tma_store && p.has_res && !(p.dbg & 64);
    if (tma_res) prefetch_residual();

    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      int mt, nt;
      fast_divmod(tile, p.d_ntiles, mt, nt);
      const int pix = mt * BM + row;  // flat padded input-grid pixel of this thread's row
      bool valid = pix < p.np;
      int b = 0, yp = 0, xp = 0;
      if (valid) {
        int t;
        fast_divmod(pix, p.d_wp, t, xp);
        fast_divmod(t, p.d_hp, b, yp);
        valid = xp >= 1 && xp <= p.wp - 2 && yp >= 1 && yp <= p.hp - 2;
      }
      int oy = yp - 1, ox = xp - 1;
      if (p.stride == 2) {
        valid = valid && ((oy & 1) == 0) && ((ox & 1) == 0);
        oy >>= 1;
        ox >>= 1;
      }
      valid = valid && oy < p.oh && ox < p.ow;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const int n_valid = min(BN, p.cout - nt * BN);  // real filters in this n-tile (may be <= 0 for pure padding)

      if (p.tma_store) {
        // The tile's rows ARE rows of the output buffer (same padded geometry): groups are staged in shared memory
        // (128B-swizzled rows, conflict-free 16-byte accesses) and TMA writes whole 128-byte lines.  Halo rows are
        // staged as zeros, so the halo stays zero; rows beyond the tensor / filters beyond store_cols are clipped by
        // the tensor map.
#pragma unroll 1
        for (int g = (GPT == 1 ? 0 : team); g < GPT; g += (GPT == 1 ? 1 : 2)) {
          if (GPT == 1 && (it & 1) != team) break;
          const int colbase = nt * BN + g * 64;
          if (colbase >= p.store_cols || (p.dbg & 2)) break;
          const int halves = (p.store_cols - colbase) >= 64 ? 2 : 1;
          uint32_t v0[32], v1[32];
          tc_ld32(t_row + g * 64, v0);
          if (halves == 2) tc_ld32(t_row + g * 64 + 32, v1);
          unsigned char* srow = smem_gen + group_buf(buf) + row * 128;
          if (tid_t < 64) s_bias[tid_t] = __ldg(p.bias + colbase + tid_t);   // this group's biases
          if (tma_res) {
            prefetch_residual();                                 // next group -> other buffer
          } else if (issuer) {
            // the previous TMA store out of this buffer (NBUF groups ago) must have finished reading it
            if (NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          team_sync();
          if (tma_res) {
            mbar_wait(rfull_bar(team * 2 + buf), rphase[buf]);   // residual group landed in this buffer
            rphase[buf] ^= 1;
          }
          tc_wait_ld();
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            if (hf >= halves) break;
            uint4 pk[4];
            if (valid) {
              if (tma_res) epilogue_math32<true>(hf == 0 ? v0 : v1, s_bias + hf * 32, act, p.slope, srow, hf, row & 7, pk);
              else epilogue_math32<false>(hf == 0 ? v0 : v1, s_bias + hf * 32, act, p.slope, srow, hf, row & 7, pk);
            } else {
#pragma unroll
              for (int gg = 0; gg < 4; gg++) pk[gg] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int gg = 0; gg < 4; gg++) *reinterpret_cast<uint4*>(srow + (((hf * 4 + gg) ^ (row & 7)) << 4)) = pk[gg];
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          team_sync();
          if (issuer && !(p.dbg & 1)) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&map_o)),
                         "r"(smem_base + group_buf(buf)), "r"(colbase), "r"(mt * BM)
                         : "memory");
          }
          if (issuer) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          buf ^= (NBUF - 1);
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (p.dbg & 2) break;
          if (((c0 >> 5) & 1) != team) continue;   // 32-column chunks alternate between the teams
          if (!p.out_f32_nchw && nt * BN + c0 >= p.store_cols) continue;   // narrower output buffer than the padded tile
          uint32_t v[32];
          tc_ld32(t_row + c0, v);
          tc_wait_ld();
          if (!valid || (p.dbg & 1)) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j++) {
            float x = __uint_as_float(v[j]) + __ldg(p.bias + nt * BN + c0 + j);   // warp-uniform address
            if (p.has_act) x = x > 0.f ? x : p.slope * x;
            f[j] = x;
          }
          if (p.out_f32_nchw) {
            // heads: fp32 [B, cout, oh, ow]
            float* o = reinterpret_cast<float*>(p.out);
            const size_t plane = (size_t)p.oh * p.ow;
            const size_t base = ((size_t)b * p.cout) * plane + (size_t)oy * p.ow + ox;
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (c0 + j < n_valid) o[base + (size_t)(nt * BN + c0 + j) * plane] = f[j];
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out);
            const int ohp = (p.upsample2x ? 2 * p.oh : p.oh) + 2, owp = (p.upsample2x ? 2 * p.ow : p.ow) + 2;
            if (p.has_res) {
              const __nv_bfloat16* r =
                  p.residual + (((size_t)b * (p.oh + 2) + oy + 1) * (p.ow + 2) + ox + 1) * p.res_cs + nt * BN + c0;
              const uint4* r4 = reinterpret_cast<const uint4*>(r);
#pragma unroll
              for (int g = 0; g < 4; g++) {
                const uint4 rv = __ldg(r4 + g);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  const float2 ff = __bfloat1622float2(h[e]);
                  f[g * 8 + e * 2] += ff.x;
                  f[g * 8 + e * 2 + 1] += ff.y;
                }
              }
            }
            uint4 pk[4];
#pragma unroll
            for (int g = 0; g < 4; g++) {
              __nv_bfloat162 h[4];
#pragma unroll
              for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(f[g * 8 + e * 2], f[g * 8 + e * 2 + 1]);
              pk[g] = *reinterpret_cast<uint4*>(h);
            }
            const int reps = p.upsample2x ? 2 : 1;
            for (int ry = 0; ry < reps; ry++)
              for (int rx = 0; rx < reps; rx++) {
                const int yy = (p.upsample2x ? 2 * oy + ry : oy) + 1, xx = (p.upsample2x ? 2 * ox + rx : ox) + 1;
                uint4* o4 = reinterpret_cast<uint4*>(o + (((size_t)b * ohp + yy) * owp + xx) * p.out_cs + nt * BN + c0);
#pragma unroll
                for (int g = 0; g < 4; g++) o4[g] = pk[g];
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: [cout, cin, k, k] fp32 (* per-filter scale) -> bf16 [k*k][cout_pad][cin_pad], zero padded
// ---------------------------------------------------------------------------------------------------------------
// modes: see pack.cuh
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ scale, int cout, int cin,
                                         int ks, int cout_pad, int cin_pad, int mode, __nv_bfloat16* __restrict__ out) {
  const size_t total = (size_t)ks * ks * cout_pad * cin_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int co = (int)((i / cin_pad) % cout_pad);
    const int tap = (int)(i / ((size_t)cin_pad * cout_pad));
    float v = 0.f;
    if (ci < cin && co < cout) {
      v = pack_value(w, cout, cin, ks, mode, tap, co, ci);
      if (scale && (mode == 0 || mode == 2)) v *= scale[co];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// weight-gradient unpack: dw fp32 [taps][cout_pad][cin_pad] (wgrad output) -> grad [cout][cin][k][k] in nn.Conv2d layout.
// mode 0: plain k x k;  mode 2: space-to-depth form (taps 2x2, cin_eff = 4C) -> [cout][C][3][3]
__global__ void conv_unpack_wgrad_kernel(const float* __restrict__ dw, int cout_pad, int cin_pad, int mode, int cout, int cin,
                                         int ks, float* __restrict__ grad) {
  const size_t total = (size_t)cout * cin * ks * ks;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kw = (int)(i % ks), kh = (int)((i / ks) % ks);
  const int c = (int)((i / ((size_t)ks * ks)) % cin);
  const int co = (int)(i / ((size_t)ks * ks * cin));
  if (mode == 0) {
    grad[i] = dw[((size_t)(kh * ks + kw) * cout_pad + co) * cin_pad + c];
  } else {
    const int qy = kh == 0 ? 0 : 1, py = kh == 1 ? 0 : 1, qx = kw == 0 ? 0 : 1, px = kw == 1 ? 0 : 1;
    grad[i] = dw[((size_t)(qy * 2 + qx) * cout_pad + co) * cin_pad + (py * 2 + px) * cin + c];
  }
}

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct ConvGeom {
  int cin_pad, cout_pad, bn, taps;
};
static ConvGeom conv_geom(const ryolo_conv_desc* d) {
  ConvGeom g;
  g.cin_pad = round_up(d->cin, 64);
  g.bn = d->cout > 128 ? 256 : (d->cout > 64 ? 128 : 64);
  g.cout_pad = round_up(d->cout, g.bn);
  const int k = d->ksize < 0 ? -d->ksize : d->ksize;
  g.taps = k * k;
  return g;
}

// cuTensorMapEncodeTiled is a DRIVER entry point; it is resolved through the runtime at first use so that
// libryolo.so does not link libcuda.so.1 and still loads (for symbol / workspace queries) on a GPU-less host.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

int encode_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_err("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return RYOLO_E_CUDA;
  }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled failed: CUresult %d (inner=%llu rows=%llu stride=%llu box=%ux%u base=%p)", (int)r,
            (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner, box_rows,
            base);
    return RYOLO_E_CUDA;
  }
  return RYOLO_OK;
}

template <int BN, int NBUF, bool AROW, bool BRES = false>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mr,
                       const ConvParams& p, cudaStream_t stream) {
  using S = ConvSmem<BN, NBUF, AROW, BRES>;
  static_assert(S::kTotal <= 227 * 1024, "shared memory budget");
  RYOLO_SMEM_OPT_IN((conv_igemm_kernel<BN, NBUF, AROW, BRES>), S::kTotal);
  const int num_sms = gemm_sm_count();
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  conv_igemm_kernel<BN, NBUF, AROW, BRES><<<grid, CONV_THREADS, S::kTotal, stream>>>(ma, mb, mo, mr, p);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

}  // namespace ryolo

using namespace ryolo;

extern "C" size_t ryolo_conv_packed_weight_bytes(const ryolo_conv_desc* d) {
  if (!d) return 0;
  const ConvGeom g = conv_geom(d);
  return (size_t)g.taps * g.cout_pad * g.cin_pad * 2;
}

extern "C" int ryolo_conv_pack_weights_ex(const ryolo_conv_desc* d, const float* weight, const float* scale,
                                          void* packed_out, int mode, void* stream_);

extern "C" int ryolo_conv_pack_weights(const ryolo_conv_desc* d, const float* weight, const float* scale, void* packed_out,
                                       void* stream_) {
  return ryolo_conv_pack_weights_ex(d, weight, scale, packed_out, 0, stream_);
}

extern "C" int ryolo_conv_unpack_wgrad(const float* dw, int cout_pad, int cin_pad, int mode, int cout, int cin, int ksize,
                                       float* grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(dw && grad && cout > 0 && cin > 0 && (mode == 0 || mode == 2));
  RYOLO_ARG_CHECK(mode == 0 ? (ksize == 1 || ksize == 3) : ksize == 3);
  const size_t total = (size_t)cout * cin * ksize * ksize;
  conv_unpack_wgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dw, cout_pad, cin_pad, mode, cout, cin, ksize,
                                                                                grad);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_conv_pack_weights_ex(const ryolo_conv_desc* d, const float* weight, const float* scale,
                                          void* packed_out, int mode, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(d && weight && packed_out && mode >= 0 && mode <= 3);
  RYOLO_ARG_CHECK(mode < 2 || ((d->ksize == 2 || d->ksize == -2) && (mode == 2 ? d->cin % 4 == 0 : d->cout % 4 == 0)));
  RYOLO_ARG_CHECK(d->ksize == 1 || d->ksize == 3 || d->ksize == 2 || d->ksize == -2);
  const ConvGeom g = conv_geom(d);
  const size_t total = (size_t)g.taps * g.cout_pad * g.cin_pad;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  conv_pack_weights_kernel<<<blocks, 256, 0, stream>>>(weight, scale, d->cout, d->cin, d->ksize < 0 ? -d->ksize : d->ksize,
                                                        g.cout_pad, g.cin_pad, mode,
                                                        static_cast<__nv_bfloat16*>(packed_out));
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" size_t ryolo_conv_workspace_bytes(const ryolo_conv_desc* d) {
  (void)d;
  return 0;  // the implicit GEMM needs no scratch (no im2col buffer)
}

extern "C" int ryolo_conv_bn_act_fwd(const ryolo_conv_desc* d, const void* x, const void* packed_w, const float* bias,
                                     const void* residual, void* y, void* workspace, size_t workspace_bytes,
                                     void* stream_) {
  (void)workspace;
  (void)workspace_bytes;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(d && x && packed_w && bias && y);
  RYOLO_ARG_CHECK(d->batch > 0 && d->in_h > 0 && d->in_w > 0 && d->cin > 0 && d->cout > 0);
  RYOLO_ARG_CHECK(d->ksize == 1 || d->ksize == 3 || d->ksize == 2 || d->ksize == -2);
  RYOLO_ARG_CHECK(d->stride == 1 || (d->stride == 2 && d->ksize != 2 && d->ksize != -2));
  RYOLO_ARG_CHECK(!(d->stride == 2 && d->upsample2x));
  RYOLO_ARG_CHECK(!d->has_residual || (residual != nullptr && d->stride == 1 && !d->upsample2x &&
                                        d->res_stride % 8 == 0 && d->out_dtype == RYOLO_DT_BF16));
  const ConvGeom g = conv_geom(d);
  // a buffer narrower than the 64-wide K chunk is fine: TMA zero-fills the part of the box beyond the inner extent
  RYOLO_ARG_CHECK(d->cin_stride >= d->cin && d->cin_stride % 8 == 0);
  RYOLO_ARG_CHECK(d->out_dtype == RYOLO_DT_BF16 || d->out_dtype == RYOLO_DT_F32);
  RYOLO_ARG_CHECK(!d->has_residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0);
  if (d->out_dtype == RYOLO_DT_BF16) RYOLO_ARG_CHECK(d->cout_stride % 8 == 0 && d->cout_stride >= round_up(d->cout, 32));
  RYOLO_ARG_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0);

  ConvParams p;
  p.hp = d->in_h + 2;
  p.wp = d->in_w + 2;
  const long long np = (long long)d->batch * p.hp * p.wp;
  RYOLO_ARG_CHECK(np < (1ll << 31) - 4096);
  p.np = (int)np;
  p.taps = g.taps;
  p.tap_flip = d->ksize == -2 ? 1 : 0;
  p.kchunks = g.cin_pad / 64;
  p.cout = d->cout;
  p.cout_pad = g.cout_pad;
  p.n_tiles = g.cout_pad / g.bn;
  p.d_ntiles = make_fastdiv(p.n_tiles);
  p.d_wp = make_fastdiv(p.wp);
  p.d_hp = make_fastdiv(p.hp);
  p.m_tiles = (p.np + BM - 1) / BM;
  p.stride = d->stride;
  p.oh = d->stride == 2 ? (d->in_h + 1) / 2 : d->in_h;   // k=3,pad=1 (or k=1,pad=0) with stride 2: ceil(h/2)
  p.ow = d->stride == 2 ? (d->in_w + 1) / 2 : d->in_w;
  p.out_cs = d->cout_stride;
  p.store_cols = round_up(d->cout, 32);   // never touch channels beyond the 32-column chunk that holds the last filter
  p.res_cs = d->res_stride;
  p.has_act = d->has_act;
  p.has_res = d->has_residual;
  p.upsample2x = d->upsample2x;
  p.out_f32_nchw = d->out_dtype == RYOLO_DT_F32;
  p.slope = d->slope;
  p.bias = bias;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.out = y;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("RYOLO_CONV_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  p.dbg = dbg;

  // A-row sharing (one A box per kernel row instead of one per tap) on every multi-tap launch: RYOLO_CONV_AROW=1; =2 adds
  // resident weights on the thin (cout <= 64) layers; 0 = per-tap loads
  static int arow_env = -1;
  if (arow_env < 0) {
    const char* e = getenv("RYOLO_CONV_AROW");
    arow_env = e ? atoi(e) : 2;   // validated on B200: tests/test_conv_gpu.py etc. pass with 1 and 2 (gpurun_out/r02_pytest_arow2_bo0.log)
  }
  const bool arow = arow_env != 0 && g.taps > 1 && dbg == 0;
  static int arow_bo = -1;
  if (arow_bo < 0) {
    const char* e = getenv("RYOLO_CONV_AROW_BO");
    arow_bo = e ? atoi(e) : 0;    // 0 is the correct form (see the kernel comment); 1 kept as the record of the experiment
  }
  p.arow_bo = arow_bo;
  CUtensorMap ma, mb;
  const int a_inner = d->cin_stride < g.cin_pad ? d->cin_stride : g.cin_pad;
  int st = encode_map_2d(&ma, x, (uint64_t)a_inner, (uint64_t)p.np, (uint64_t)d->cin_stride * 2, BK, arow ? BM + 8 : BM);
  if (st != RYOLO_OK) return st;
  st = encode_map_2d(&mb, packed_w, (uint64_t)g.cin_pad, (uint64_t)g.taps * g.cout_pad, (uint64_t)g.cin_pad * 2, BK,
                     (uint32_t)g.bn);
  if (st != RYOLO_OK) return st;
  // output map for the TMA-store epilogue: rows = flat padded pixels (same geometry as the input when stride is 1),
  // inner extent = store_cols so that a wider (concat) buffer's neighbouring channels are never touched
  CUtensorMap mo = ma;
  p.tma_store = (d->out_dtype == RYOLO_DT_BF16 && d->stride == 1 && !d->upsample2x) ? 1 : 0;
  if (dbg & 32) p.tma_store = 0;
  if (p.tma_store) {
    st = encode_map_2d(&mo, y, (uint64_t)p.store_cols, (uint64_t)p.np, (uint64_t)d->cout_stride * 2, 64, BM);
    if (st != RYOLO_OK) return st;
  }
  CUtensorMap mr = ma;
  if (p.tma_store && p.has_res) {
    // residual groups arrive through TMA too (prefetched one group ahead); filters >= cout read as zero
    st = encode_map_2d(&mr, residual, (uint64_t)d->cout, (uint64_t)p.np, (uint64_t)d->res_stride * 2, 64, BM);
    if (st != RYOLO_OK) return st;
  }
  if (arow) {
    if (g.bn == 256) {
      if (p.tma_store && p.has_res) return launch_conv<256, 2, true>(ma, mb, mo, mr, p, stream);
      return launch_conv<256, 1, true>(ma, mb, mo, mr, p, stream);
    }
    if (g.bn == 128) return launch_conv<128, 2, true>(ma, mb, mo, mr, p, stream);
    if (arow_env >= 2 && p.n_tiles == 1 && g.taps * p.kchunks <= 9) return launch_conv<64, 2, true, true>(ma, mb, mo, mr, p, stream);
    return launch_conv<64, 2, true>(ma, mb, mo, mr, p, stream);
  }
  if (g.bn == 256) {
    if (p.tma_store && p.has_res) return launch_conv<256, 2, false>(ma, mb, mo, mr, p, stream);
    return launch_conv<256, 1, false>(ma, mb, mo, mr, p, stream);
  }
  if (g.bn == 128) return launch_conv<128, 2, false>(ma, mb, mo, mr, p, stream);
  return launch_conv<64, 2, false>(ma, mb, mo, mr, p, stream);
}

