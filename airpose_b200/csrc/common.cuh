// Shared host-side plumbing for libairpose_b200: error reporting and launch accounting.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/airpose_b200.h"

namespace airpose {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define AP_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      airpose::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,              \
                         cudaGetErrorString(_e));                                        \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define AP_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      airpose::set_error(__VA_ARGS__);                                                   \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

#define AP_LAUNCH_CHECK()                                                                \
  do {                                                                                   \
    airpose::count_launch();                                                             \
    AP_CHECK_CUDA(cudaGetLastError());                                                   \
  } while (0)

// Launch of a kernel that begins with griddepcontrol.wait / launch_dependents (ptx::grid_dep_wait / grid_dep_launch) with
// programmatic stream serialization: its CTAs become resident while the kernel before it drains.  For the chains of short dependent
// launches of the training step.  AIRPOSE_NO_CHAIN_PDL=1: plain launches (A/B runs).
inline bool chain_pdl() {
  static const bool on = getenv("AIRPOSE_NO_CHAIN_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_chain_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = chain_pdl() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e == cudaSuccess) count_launch();
  return e;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  return launch_chain_smem(kernel, grid, block, 0, st, args...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <class T>
int device_upload(T** dst, const T* src_host, size_t count) {
  AP_CHECK_CUDA(cudaMalloc((void**)dst, count * sizeof(T)));
  AP_CHECK_CUDA(cudaMemcpy(*dst, src_host, count * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace airpose
