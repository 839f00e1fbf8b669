// Shared helpers for libcspe kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "cspe.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcspe is written for sm_100a (B200) only"
#endif

namespace cspe {

// ---- error plumbing (host) ------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();  // cached per device; <= 0 on failure

#define CSPE_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::cspe::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

#define CSPE_CUDA_OK(expr)                                                            \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::cspe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                        __FILE__, __LINE__);                                          \
      return CSPE_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

#define CSPE_LAUNCH_OK(name)                                                          \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::cspe::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
      return CSPE_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Launch `kernel` so that it may start as soon as the kernel before it in `stream` has executed
// griddepcontrol.launch_dependents in every CTA (or exited); after a kernel that never does, this
// is an ordinary serialised launch.  A kernel launched this way must execute griddepcontrol.wait
// before it touches anything the previous kernel writes.
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// RULE (measured on B200, round 2): whatever such a kernel reads BEFORE griddepcontrol.wait must bypass L1 — use
// __ldcg() / ld.global.cg or TMA, never __ldg() / plain loads / const __restrict__ dereferences.  A kernel that starts
// early as a programmatic dependent does not get the L1 invalidation an ordinary kernel boundary gives: lines an
// EARLIER kernel left in that SM's L1 survive, so an L1-cached load can return what the address held before the
// last host write (H2D copy) to it.  Seen as cspe_format_fixed6 formatting the PREVIOUS call's row count about once
// in 100 calls (tools/text_race_probe.py); loads issued after the wait were never seen stale.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---- device: mbarrier helpers for the TMA ring ---------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
  // make the inits visible to the async proxy before any bulk copy signals them
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#endif  // __CUDACC__

}  // namespace cspe
