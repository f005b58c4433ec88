// Shared host-side plumbing of libxeq_b200: error reporting, launch checks, device info.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#include "../../include/xeq_b200.h"

namespace xeq {

void set_error(const char* fmt, ...);
int num_sms();
void count_launches(int n);  // feeds xeq_launch_count()

#define XEQ_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      xeq::set_error(__VA_ARGS__);        \
      return XEQ_ERR_INVALID;             \
    }                                     \
  } while (0)

#define XEQ_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      xeq::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return XEQ_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define XEQ_LAUNCH_CHECK() XEQ_CUDA(cudaGetLastError())
// after a batch of n kernel launches: check and count
#define XEQ_LAUNCHED(n)   \
  do {                    \
    XEQ_LAUNCH_CHECK();   \
    xeq::count_launches(n); \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve a typed region out of a caller-supplied workspace
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace xeq
