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

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A step of the model is ~600 small kernels; launched back to back (CUDA graph or stream) each one pays its launch
// latency and its prologue (TMEM allocation, barrier setup, smem carve-up) AFTER the previous kernel has drained.
// Kernels launched through launch_pdl() may start while their predecessor is still running; they call pdl_wait()
// before touching global memory (it returns once the predecessor has completed and its writes are visible), and
// pdl_trigger() as early as possible so that THEIR successor can be staged.  Correct with any predecessor: a kernel
// that never triggers releases its dependents when it exits.
bool pdl_enabled();  // XEQ_PDL=0 disables (debugging)

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

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
