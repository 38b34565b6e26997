// Parity-precision convolution: the same flat implicit GEMM as conv.cu, but with fp32-grade arithmetic on the bf16
// tensor pipe.  The reference runs fp32 nn.Conv2d (model/models.py:55-60); north_star asks for 1e-4 relative on conv
// activations, which 8-bit-mantissa operands cannot give through 75 layers.
//
// Split operands.  Every fp32 value v is carried as bf16 PLANES  p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1)
// (three planes = 24 significant bits, i.e. fp32 itself; two planes = 16 bits).  An activation buffer holds the planes side
// by side in the channel dimension, `plane_stride` channels apart; packed weights hold one K-range per TERM of
//        a*w ~= a2*w0 + a0*w2 + a1*w1 + a1*w0 + a0*w1 + a0*w0                    (small terms first)
// so one launch is a GEMM with K = terms x taps x channels whose k-chunks pick (activation plane, weight plane) from a
// term table (2 bits per term).  All products of two bf16 are exact in fp32; the dropped terms are <= 2^-24 relative.
// Measured (tests/test_parity_gpu.py): two planes / three terms leave 4.4e-6 rms per layer -- enough for the eval
// network (2.7e-5 of the head scale at 608 x 608) but not for training-mode BatchNorm, which amplifies it to 2e-4.
//
// Exact accumulation.  Measured on B200 (scratch/acc_precision.py, profiles/r02_acc_precision.txt): the tcgen05 fp32
// accumulator TRUNCATES on every update (relative error -6.6e-6, biased towards zero, after the 576 updates of a
// K = 9216 layer) -- a bias that adds up linearly over the network.  So the MMA warp closes the TMEM accumulator every
// SEG k-steps (<= 32 updates), the epilogue warps drain it and keep the running sum in REGISTERS with round-to-nearest
// fp32 adds, while the MMAs of the next segment run into the second TMEM stage.
//
// Output: raw fp32 sums on the same padded-NHWC grid (optionally accumulated into the existing contents: dgrad into a
// gradient buffer with several producers).  Bias / BN / PReLU / shortcut / re-splitting are fp32 elementwise kernels
// (parity.cu).  This path exists for the 1e-4 tolerance, not for speed: tile 128 x 64, one CTA per SM.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "pack.cuh"
#include "tc05.cuh"

namespace ryolo {

constexpr int PX_BM = 128, PX_BN = 64, PX_BK = 64;
constexpr int PX_STAGES = 6;
constexpr int PX_THREADS = 192;      // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: epilogue (one per TMEM lane quarter)
constexpr int PX_SEG = 8;            // k-steps per accumulator segment (32 MMA updates)
constexpr int PX_MAX_TERMS = 6;

struct PxParams {
  int np, hp, wp;
  int taps, tap_flip;
  int n;                 // 64-channel chunks per plane
  int nterms;
  int a_col0[PX_MAX_TERMS];   // first activation column (channels) of each term's plane
  int cout_pad, n_tiles, m_tiles;
  int out_cs, out_cols, accumulate;
  float* out;
};

struct PxSmem {
  static constexpr int kABytes = PX_BM * PX_BK * 2;
  static constexpr int kBBytes = PX_BN * PX_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = PX_STAGES * kStageBytes;
  static constexpr int kNumBars = 2 * PX_STAGES + 4;   // full, empty, tfull[2], tempty[2]
  static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
};

__global__ void __launch_bounds__(PX_THREADS, 1)
conv_px_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const PxParams p) {
  using S = PxSmem;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + S::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PX_STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + S::kBarOffset + S::kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int per_term = p.taps * p.n;
  const int k_iters = p.nterms * per_term;
  const int n_seg = (k_iters + PX_SEG - 1) / PX_SEG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PX_STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(2 * PX_BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int p0 = mt * PX_BM;
        for (int kk = 0; kk < k_iters; kk++) {
          const int term = kk / per_term, r = kk - term * per_term;
          const int tap = r / p.n, c = r - tap * p.n;
          int off = 0;
          if (p.taps == 9) off = (tap / 3 - 1) * p.wp + (tap % 3 - 1);
          else if (p.taps == 4) off = (tap / 2 - 1 + p.tap_flip) * p.wp + (tap % 2 - 1 + p.tap_flip);
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          mbar_expect_tx(full_bar(stage), S::kStageBytes);
          tma_load_2d(sa, &map_a, full_bar(stage), p.a_col0[term] + c * PX_BK, p0 + off);
          tma_load_2d(sa + S::kABytes, &map_b, full_bar(stage), (term * p.n + c) * PX_BK, tap * p.cout_pad + nt * PX_BN);
          if (++stage == PX_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(PX_BM, PX_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int seg = 0; seg < n_seg; seg++) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * PX_BN;
          const int k0 = seg * PX_SEG, k1 = min(k_iters, k0 + PX_SEG);
          for (int kk = k0; kk < k1; kk++) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * S::kStageBytes;
            const uint64_t adesc = make_smem_desc(sa);
            const uint64_t bdesc = make_smem_desc(sa + S::kABytes);
#pragma unroll
            for (int k = 0; k < PX_BK / 16; k++) tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kk > k0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == PX_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(tfull_bar(acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter this warp may read (warps 2,3,4,5 -> 2,3,0,1)
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const int pix = mt * PX_BM + row;
      bool valid = pix < p.np;
      if (valid) {
        const int xp = pix % p.wp;
        const int yp = (pix / p.wp) % p.hp;
        valid = xp >= 1 && xp <= p.wp - 2 && yp >= 1 && yp <= p.hp - 2;
      }
      float sum[PX_BN];
#pragma unroll
      for (int j = 0; j < PX_BN; j++) sum[j] = 0.f;
      for (int seg = 0; seg < n_seg; seg++) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * PX_BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          uint32_t v[32];
          tc_ld32(t_row + hf * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j++) sum[hf * 32 + j] = __fadd_rn(sum[hf * 32 + j], __uint_as_float(v[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (valid) {
        float* o = p.out + (size_t)pix * p.out_cs + nt * PX_BN;
#pragma unroll
        for (int j = 0; j < PX_BN; j += 4) {
          if (nt * PX_BN + j < p.out_cols) {
            float4 r = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
            if (p.accumulate) {
              const float4 e = *reinterpret_cast<const float4*>(o + j);
              r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
            }
            *reinterpret_cast<float4*>(o + j) = r;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * PX_BN) : "memory");
  }
}

// weights fp32 [cout, cin, k, k] -> bf16 [taps][cout_pad][nterms * n*64]; term t holds plane (plane_code >> 2t) & 3
__global__ void px_pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int ks, int cout_pad, int n,
                                       int nterms, int plane_code, int mode, __nv_bfloat16* __restrict__ out) {
  const int kp = n * 64;
  const size_t total = (size_t)ks * ks * cout_pad * kp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % kp);
    const int co = (int)((i / kp) % cout_pad);
    const int tap = (int)(i / ((size_t)kp * cout_pad));
    float v = 0.f;
    if (ci < cin && co < cout) v = pack_value(w, cout, cin, ks, mode, tap, co, ci);
    __nv_bfloat16 pl[3];
    pl[0] = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(pl[0]);
    pl[1] = __float2bfloat16_rn(r1);
    pl[2] = __float2bfloat16_rn(r1 - __bfloat162float(pl[1]));
    __nv_bfloat16* row = out + ((size_t)tap * cout_pad + co) * ((size_t)nterms * kp);
    for (int t = 0; t < nterms; t++) {
      const int pi = (plane_code >> (2 * t)) & 3;
      row[(size_t)t * kp + ci] = pi == 0 ? pl[0] : (pi == 1 ? pl[1] : pl[2]);
    }
  }
}

// wgrad of split operands: dw [taps][np*po][np*pi] fp32 holds the np x np (plane, plane) blocks; their sum is the
// gradient (small blocks first).
// mode 0: plain k x k -> [cout][cin][k][k];  mode 2: space-to-depth form (taps 2x2, pi = plane of 4C) -> [cout][C][3][3]
__global__ void px_unpack_wgrad_kernel(const float* __restrict__ dw, int po, int pi, int np, int mode, int cout, int cin,
                                       int ks, int ci_off, float* __restrict__ grad) {
  const size_t total = (size_t)cout * cin * ks * ks;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kw = (int)(i % ks), kh = (int)((i / ks) % ks);
  const int c = (int)((i / ((size_t)ks * ks)) % cin);
  const int co = (int)(i / ((size_t)ks * ks * cin));
  int tap, ci;
  if (mode == 0) {
    tap = kh * ks + kw;
    ci = ci_off + c;
  } else {
    const int qy = kh == 0 ? 0 : 1, py = kh == 1 ? 0 : 1, qx = kw == 0 ? 0 : 1, px = kw == 1 ? 0 : 1;
    tap = qy * 2 + qx;
    ci = ci_off + (py * 2 + px) * cin + c;
  }
  const size_t rs = (size_t)np * pi;
  const float* base = dw + (size_t)tap * ((size_t)np * po) * rs;
  float s = 0.f;
  for (int order = 2 * (np - 1); order >= 0; order--)      // plane index sum: larger = smaller magnitude
    for (int a = np - 1; a >= 0; a--) {
      const int b = order - a;
      if (b < 0 || b >= np) continue;
      s += base[(size_t)(a * po + co) * rs + (size_t)b * pi + ci];
    }
  grad[i] = s;
}

}  // namespace ryolo

using namespace ryolo;

static int px_round_up(int x, int m) { return (x + m - 1) / m * m; }

extern "C" size_t ryolo_px_packed_weight_bytes(int cout, int cin, int ksize, int nterms) {
  const int k = ksize < 0 ? -ksize : ksize;
  return (size_t)k * k * px_round_up(cout, 64) * ((size_t)nterms * px_round_up(cin, 64)) * 2;
}

extern "C" int ryolo_px_pack_weights(const float* weight, int cout, int cin, int ksize, int nterms, int w_plane_code,
                                     int mode, void* packed_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(weight && packed_out && cout > 0 && cin > 0 && nterms >= 1 && nterms <= PX_MAX_TERMS);
  RYOLO_ARG_CHECK(mode >= 0 && mode <= 3);
  RYOLO_ARG_CHECK(ksize == 1 || ksize == 3 || ksize == 2 || ksize == -2);
  RYOLO_ARG_CHECK(mode < 2 || ((ksize == 2 || ksize == -2) && (mode == 2 ? cin % 4 == 0 : cout % 4 == 0)));
  const int k = ksize < 0 ? -ksize : ksize;
  const int cout_pad = px_round_up(cout, 64), n = px_round_up(cin, 64) / 64;
  const size_t total = (size_t)k * k * cout_pad * n * 64;
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  px_pack_weights_kernel<<<blocks, 256, 0, stream>>>(weight, cout, cin, k, cout_pad, n, nterms, w_plane_code, mode,
                                                     static_cast<__nv_bfloat16*>(packed_out));
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_px_unpack_wgrad(const float* dw, int plane_out, int plane_in, int nplanes, int mode, int cout, int cin,
                                     int ksize, int cin_off, float* grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(dw && grad && cout > 0 && cin > 0 && (mode == 0 || mode == 2) && nplanes >= 1 && nplanes <= 3);
  RYOLO_ARG_CHECK(mode == 0 ? (ksize == 1 || ksize == 3) : ksize == 3);
  RYOLO_ARG_CHECK(cin_off >= 0 && plane_out >= cout && plane_in >= cin_off + (mode == 2 ? 4 * cin : cin));
  const size_t total = (size_t)cout * cin * ksize * ksize;
  px_unpack_wgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dw, plane_out, plane_in, nplanes, mode, cout, cin,
                                                                             ksize, cin_off, grad);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}

extern "C" int ryolo_px_conv(const void* x_base, int x_cstride, int x_ch_off, int plane_stride, int cin, const void* packed_w,
                             int nterms, int a_plane_code, int batch, int in_h, int in_w, int ksize, int cout, float* out,
                             int out_cstride, int out_cols, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RYOLO_ARG_CHECK(x_base && packed_w && out && batch > 0 && in_h > 0 && in_w > 0 && cin > 0 && cout > 0);
  RYOLO_ARG_CHECK(ksize == 1 || ksize == 3 || ksize == 2 || ksize == -2);
  RYOLO_ARG_CHECK(nterms >= 1 && nterms <= PX_MAX_TERMS);
  RYOLO_ARG_CHECK(x_cstride % 8 == 0 && x_ch_off % 8 == 0 && plane_stride % 8 == 0 && x_ch_off >= 0);
  RYOLO_ARG_CHECK((reinterpret_cast<uintptr_t>(x_base) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  RYOLO_ARG_CHECK(out_cstride % 4 == 0 && out_cols % 4 == 0 && out_cols > 0 && out_cols <= out_cstride);
  PxParams p;
  p.hp = in_h + 2;
  p.wp = in_w + 2;
  const long long np = (long long)batch * p.hp * p.wp;
  RYOLO_ARG_CHECK(np < (1ll << 31) - 4096);
  p.np = (int)np;
  const int k = ksize < 0 ? -ksize : ksize;
  p.taps = k * k;
  p.tap_flip = ksize == -2 ? 1 : 0;
  p.n = px_round_up(cin, 64) / 64;
  p.nterms = nterms;
  int max_plane = 0;
  for (int t = 0; t < PX_MAX_TERMS; t++) {
    const int pl = (a_plane_code >> (2 * t)) & 3;
    RYOLO_ARG_CHECK(pl <= 2);
    if (t < nterms && pl > max_plane) max_plane = pl;
    p.a_col0[t] = x_ch_off + pl * plane_stride;
  }
  // channels of a k-chunk beyond cin meet zero weights (and TMA zero-fills beyond the row), the planes themselves must fit
  RYOLO_ARG_CHECK(x_ch_off + max_plane * plane_stride + cin <= x_cstride);
  p.cout_pad = px_round_up(cout, 64);
  p.n_tiles = p.cout_pad / PX_BN;
  p.m_tiles = (p.np + PX_BM - 1) / PX_BM;
  p.out_cs = out_cstride;
  p.out_cols = out_cols;
  p.accumulate = accumulate;
  p.out = out;
  CUtensorMap ma, mb;
  int st = encode_map_2d(&ma, x_base, (uint64_t)x_cstride, (uint64_t)p.np, (uint64_t)x_cstride * 2, PX_BK, PX_BM);
  if (st != RYOLO_OK) return st;
  const uint64_t krow = (uint64_t)nterms * p.n * 64;
  st = encode_map_2d(&mb, packed_w, krow, (uint64_t)p.taps * p.cout_pad, krow * 2, PX_BK, PX_BN);
  if (st != RYOLO_OK) return st;
  RYOLO_SMEM_OPT_IN(conv_px_kernel, PxSmem::kTotal);
  const int sms = device_sm_count();
  const int total = p.m_tiles * p.n_tiles;
  conv_px_kernel<<<total < sms ? total : sms, PX_THREADS, PxSmem::kTotal, stream>>>(ma, mb, p);
  RYOLO_LAUNCH_CHECK();
  return RYOLO_OK;
}
