// Shared helpers for libryolo.so (sm_100a only; no CPU fallback anywhere in this library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/ryolo.h"

namespace ryolo {

// thread-local error text returned by ryolo_last_error()
char* err_buf();
void set_err(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define RYOLO_CUDA_TRY(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::ryolo::set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RYOLO_E_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define RYOLO_LAUNCH_CHECK()                                                               \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::ryolo::set_err("%s:%d: kernel launch -> %s", __FILE__, __LINE__,                   \
                       cudaGetErrorString(_e));                                            \
      return RYOLO_E_CUDA;                                                                 \
    }                                                                                      \
    ::ryolo::count_launch();                                                               \
  } while (0)

#define RYOLO_ARG_CHECK(cond)                                                   \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ::ryolo::set_err("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);   \
      return RYOLO_E_ARG;                                                       \
    }                                                                           \
  } while (0)


// Per-device one-time setup.  Opt-in shared-memory sizes and SM counts are properties of a DEVICE: a process that
// uses several GPUs (model.to("cuda:1"), multi-device tests) must set / query them once per device, not once per
// process.  Racing threads may both run the (idempotent) setup; nobody launches before it has run on their device.
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
inline int device_sm_count() {
  static std::atomic<int> cache[64];
  const int d = current_device() & 63;
  int n = cache[d].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    if (n <= 0) n = 148;
    cache[d].store(n, std::memory_order_relaxed);
  }
  return n;
}
// SMs a persistent GEMM kernel may occupy: all of them, minus a reserve that a data-parallel training step keeps free for
// the NCCL all-reduce CTAs running concurrently on a side stream (ryolo_set_reserved_sms).  A persistent kernel with a
// static tile schedule that finds some SMs busy runs its surplus CTAs as a second wave -- twice the time; with a few SMs
// reserved both kernels are resident at once.
extern std::atomic<int> g_reserved_sms;
inline int gemm_sm_count() {
  const int n = device_sm_count(), r = g_reserved_sms.load(std::memory_order_relaxed);
  return n - r > 8 ? n - r : n;
}

#define RYOLO_SMEM_OPT_IN(kernel, bytes)                                                                      \
  do {                                                                                                        \
    static std::atomic<uint64_t> _mask{0};                                                                    \
    const uint64_t _bit = 1ull << (::ryolo::current_device() & 63);                                           \
    if (!(_mask.load(std::memory_order_acquire) & _bit)) {                                                    \
      RYOLO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _mask.fetch_or(_bit, std::memory_order_release);                                                        \
    }                                                                                                         \
  } while (0)

// Division of 0 <= n < 2^31 by a launch-invariant d >= 1 as a 64-bit multiply + shift: m = ceil(2^p / d), p = 31 + ceil(log2 d)
// (exact: the error m*d - 2^p is < d <= 2^(p-31)).  An integer division by a runtime value costs ~25 instructions; the conv
// epilogue did five per thread and tile -- 15 % of its instructions (ncu source page, profiles/r02_prof_conv_shapes_summary).
struct FastDiv {
  unsigned int m;
  int p;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  int l = 0;
  while ((1ll << l) < d) l++;
  f.p = 31 + l;
  f.m = (unsigned int)(((1ull << f.p) + (unsigned long long)d - 1) / (unsigned long long)d);
  f.d = d;
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
  return (int)(((unsigned long long)(unsigned int)n * f.m) >> f.p);
}
__device__ __forceinline__ void fast_divmod(int n, const FastDiv& f, int& q, int& r) {
  q = fast_div(n, f);
  r = n - q * f.d;
}
#endif

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-owned workspace
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
};

}  // namespace ryolo
