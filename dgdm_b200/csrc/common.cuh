// Shared helpers for the dgdm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/dgdm_b200.h"

namespace dgdm {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define DGDM_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::dgdm::set_error(__VA_ARGS__);             \
      return DGDM_EINVAL;                         \
    }                                             \
  } while (0)

#define DGDM_CUDA(expr)                                                                             \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::dgdm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DGDM_ECUDA;                                                                            \
    }                                                                                               \
  } while (0)

#define DGDM_LAUNCH_CHECK()                                                                         \
  do {                                                                                              \
    ::dgdm::count_launch();                                                                         \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) {                                                                        \
      ::dgdm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DGDM_ECUDA;                                                                            \
    }                                                                                               \
  } while (0)

#define DGDM_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != DGDM_OK) return _r; \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap, off;
  bool ok;
  // the caller's pointer is rounded up to 256 bytes; the *_workspace_bytes() queries include that slack
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), ok(p != nullptr) {
    size_t mis = (size_t)((uintptr_t)p & 255);
    if (mis) { size_t adj = 256 - mis; if (adj > cap) { ok = false; } else { base += adj; cap -= adj; } }
  }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > cap) { ok = false; off += bytes; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

// ---- generic fp32 GEMM with row mapping (gemm.cu) ---------------------------------------------
// C[m,n] = epi( sum_k A(m,k) * W[n,k] + bias[n] )
//   A(m,k) = A[(m / a_lr)*a_ss + (m % a_lr)*a_rs + (k / a_ct)*a_ts + (k % a_ct)]
//   C(m,n) = C[(m / c_lr)*c_ss + (m % c_lr)*c_rs + n]           (mask / add use C's mapping)
// epi: +bias, act (0 none, 1 relu, 2 silu), then "mask": zero where mask(m,n) <= 0, then "+= add(m,n)".
struct GemmArgs {
  const float* A; const float* W; const float* bias; float* C;
  const float* mask; const float* add;
  int64_t M; int N, K;
  int64_t a_lr, a_ss, a_rs, a_ts; int a_ct;
  int64_t c_lr, c_ss, c_rs;
  int64_t m_lr, m_ss, m_rs;   // mapping for mask / add tensors
  int act;
};
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };
GemmArgs gemm_plain(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                    int64_t M, int N, int K, int act);
int gemm_f32(const GemmArgs& g, cudaStream_t s);

}  // namespace dgdm
