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
