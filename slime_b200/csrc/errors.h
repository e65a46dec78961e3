// Error codes + thread-local error message shared by every translation unit of libslime_b200.
#pragma once
#include <cuda_runtime.h>

#include "slime_b200.h"  // SLIME_OK / SLIME_E* codes (include/slime_b200.h)

void slime_set_error(const char* fmt, ...);
const char* slime_get_error();

// Launch accounting / optional CUDA-event profiling of the library's own kernels (errors.cu).
// class: 0 = tcgen05 GEMM, 1 = attention, 2 = everything else (HBM-bound kernels)
void slime_note_launch();
bool slime_prof_enabled();
void slime_prof_begin(int cls, double work, cudaStream_t stream);  // work = FLOPs (cls 0/1) or bytes (cls 2)
void slime_prof_end(cudaStream_t stream);

#define SLIME_AFTER_LAUNCH()                 \
  do {                                       \
    slime_note_launch();                     \
    SLIME_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

#define SLIME_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      slime_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,                 \
                      cudaGetErrorString(_e));                                              \
      return SLIME_ECUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define SLIME_REQUIRE(cond, ...)       \
  do {                                 \
    if (!(cond)) {                     \
      slime_set_error(__VA_ARGS__);    \
      return SLIME_EINVAL;             \
    }                                  \
  } while (0)

#define SLIME_PROPAGATE(expr)          \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != SLIME_OK) return _rc;   \
  } while (0)

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute (see pdl_wait() in common.cuh).
// SLIME_PDL=0 / slime_set_pdl_mode(0) launch everything the ordinary way.
bool slime_pdl_enabled();
// First launch of `kernel`: set its preferred shared-memory carve-out (SLIME_CARVEOUT_PCT, default 44 % = the 100 KB
// configuration; -1 leaves the driver's per-kernel choice).  The weight-streaming GEMM uses L1 as the landing buffer of
// its in-flight loads (forced to the maximum carve-out it drops from 3.9 to 6.5 ms per decode step), and under PDL an SM
// cannot be re-partitioned between overlapping kernels, so the whole chain asks for ONE modest configuration; a kernel
// that needs more (the GEMM with 16 staged rows: 132 KB) still gets it.
void slime_carveout_once(const void* kernel);
// Prefill chain (tcgen05 GEMMs, attention, the norm kernels between them): launched with the programmatic attribute unless
// SLIME_PREFILL_PDL=0 / slime_set_prefill_pdl(0) - the next kernel's prologue (barrier init, TMEM allocation, descriptor
// prefetch) then overlaps the tail of the previous one.  Every kernel launched this way executes pdl_wait() before its
// first access to global memory.  No carve-out preference here (these kernels size their own shared memory).
bool slime_prefill_pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t slime_launch_prefill(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = slime_prefill_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t slime_launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                bool pdl, Args&&... args) {
  slime_carveout_once(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && slime_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
