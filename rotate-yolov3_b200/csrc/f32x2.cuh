// packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): one instruction for two fp32 lanes held in a 64-bit register.
// The element-wise BN / PReLU kernels run at ~55 % issue utilisation with scalar math -- instruction issue, not HBM, caps
// them at ~4.2 TB/s (ncu, profiles/r02_prof_bn_summary.csv); packed math removes a third of their instructions.
#pragma once
#include <stdint.h>

namespace ryolo {

__device__ __forceinline__ uint64_t f2pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// one bf16x2 word (low half = even element) -> (even, odd) as fp32x2: bf16 -> fp32 is a 16-bit shift
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t w) {
  return f2pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

}  // namespace ryolo
