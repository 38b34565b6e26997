// PROTOTYPE (not part of libryolo.so): C[M,N] fp32 = A[M,K] * B[N,K]^T, bf16 operands, with the CTA-PAIR form of the
// tensor-core instruction (tcgen05.mma.cta_group::2, M = 256 across two SMs, N = 256, each CTA holding its 128 rows of
// A and HALF of the B tile) -- the mechanics DESIGN.md section 8 item (1) wants to bring into conv.cu: per 128x256x16
// MMA a single CTA reads 12 KB of operands from shared memory (measured 162 cycles per MMA); a pair reads 8 KB per SM.
// Build / run: scratch/proto_2cta.py (nvcc -> scratch/_proto/libproto2cta.so, checks against torch.matmul, times both
// forms).  Mechanics taken from the CUTLASS sm100 headers shipped in site-packages (cute/arch/copy_sm100_tma.hpp:78-102,
// cutlass/arch/barrier.h:811-863) and /opt/skills/guides/blackwell_cuda_programming.md section 3.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../rotate-yolov3_b200/csrc/tc05.cuh"

using namespace ryolo;

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> the pair's CTA 0

template <int PAIR>   // PAIR = 2: cta_group::2 (cluster of 2, B half per CTA); PAIR = 1: plain single-CTA reference
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBRows = BN / PAIR;            // B rows (filters) this CTA stages
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kSmem = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  // executed by both CTAs of the pair; the transaction bytes update the barrier named by `bar` (CTA 0's copy)
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

template <int PAIR>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ c,
            int m, int n, int k) {
  using S = Cfg<PAIR>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + S::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + S::kBarOffset + (2 * STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR == 2 ? cluster_ctarank() : 0;
  // tile of the pair: rows [m0, m0 + PAIR*128), cols [n0, n0 + 256); this CTA owns rows m0 + rank*128 ..
  const int tiles_n = n / BN;
  const int pair_id = blockIdx.x / PAIR;
  const int m0 = (pair_id / tiles_n) * (BM * PAIR) + (int)rank * BM;
  const int n0 = (pair_id % tiles_n) * BN;
  const int k_iters = k / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                   "r"(BN)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                   "r"(BN)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();   // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer (both CTAs) ----------------
    int stage = 0;
    uint32_t phase = 0;
    for (int kk = 0; kk < k_iters; kk++) {
      mbar_wait(empty_bar(stage), phase ^ 1);           // own copy: the pair's commit is multicast to both CTAs
      const uint32_t sa = smem_base + stage * S::kStageBytes;
      if (PAIR == 2) {
        const uint32_t fb = full_bar(stage) & kPeerMask;   // CTA 0's barrier collects the bytes of both CTAs
        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * S::kStageBytes);
        tma_load_2d_pair(sa, &map_a, fb, kk * BK, m0);
        tma_load_2d_pair(sa + S::kABytes, &map_b, fb, kk * BK, n0 + (int)rank * S::kBRows);
      } else {
        mbar_expect_tx(full_bar(stage), S::kStageBytes);
        tma_load_2d(sa, &map_a, full_bar(stage), kk * BK, m0);
        tma_load_2d(sa + S::kABytes, &map_b, full_bar(stage), kk * BK, n0);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    const uint32_t idesc = make_idesc(BM * PAIR, BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kk = 0; kk < k_iters; kk++) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t sa = smem_base + stage * S::kStageBytes;
      const uint64_t adesc = make_smem_desc(sa);
      const uint64_t bdesc = make_smem_desc(sa + S::kABytes);
#pragma unroll
      for (int q = 0; q < BK / 16; q++) {
        if (PAIR == 2) tc_mma_f16_pair(tmem_base, adesc + 2 * q, bdesc + 2 * q, idesc, (kk | q) != 0);
        else tc_mma_f16(tmem_base, adesc + 2 * q, bdesc + 2 * q, idesc, (kk | q) != 0);
      }
      if (PAIR == 2) {
        tc_commit_pair(empty_bar(stage));
        if (kk == k_iters - 1) tc_commit_pair(done_bar);
      } else {
        tc_commit(empty_bar(stage));
        if (kk == k_iters - 1) tc_commit(done_bar);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: every CTA drains its own 128 accumulator rows ----------------
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tc_ld32(t_row + c0, v);
      tc_wait_ld();
      if (row < m) {
        float4* o = reinterpret_cast<float4*>(c + (size_t)row * n + n0 + c0);
#pragma unroll
        for (int e = 0; e < 8; e++)
          o[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                             __uint_as_float(v[4 * e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) {
    tc_fence_after();
    if (PAIR == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  static EncodeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    fn = reinterpret_cast<EncodeFn>(p);
  }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : -2;
}

template <int PAIR>
int launch(const void* a, const void* b, float* c, int m, int n, int k, cudaStream_t stream) {
  using S = Cfg<PAIR>;
  if (m % (BM * PAIR) || n % BN || k % BK) return -3;
  CUtensorMap ma, mb;
  if (encode(&ma, a, (uint64_t)k, (uint64_t)m, BM)) return -4;
  if (encode(&mb, b, (uint64_t)k, (uint64_t)n, S::kBRows)) return -5;
  if (cudaFuncSetAttribute(gemm_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem) != cudaSuccess) return -6;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((m / BM) * (n / BN)), 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = S::kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<PAIR>, ma, mb, c, m, n, k);
  if (e != cudaSuccess) {
    fprintf(stderr, "proto launch: %s\n", cudaGetErrorString(e));
    return -7;
  }
  return 0;
}

}  // namespace

extern "C" int proto_gemm(const void* a, const void* b, float* c, int m, int n, int k, int pair, void* stream) {
  return pair == 2 ? launch<2>(a, b, c, m, n, k, static_cast<cudaStream_t>(stream))
                   : launch<1>(a, b, c, m, n, k, static_cast<cudaStream_t>(stream));
}
