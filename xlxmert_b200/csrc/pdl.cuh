// Programmatic dependent launch helpers shared by every translation unit that launches kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace xlx {

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A step is ~800 short launches on one stream; the gap between two of them (grid drain + launch latency + the next
// kernel's prologue) is a few µs each.  Kernels launched through launch_pdl() may become resident while their
// predecessor is still draining; they call pdl_wait() before touching global memory (it returns once every
// preceding grid has completed and its writes are visible) and pdl_trigger() right away so that their own successor
// can do the same.  Only kernels that contain pdl_wait() may be launched this way.  XLX_PDL=0 turns it off.
inline bool g_pdl_suspended = false;   // set while per-launch event timing is on: timed launches must not overlap
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("XLX_PDL"); return !(e && e[0] == '0'); }();
  return on && !g_pdl_suspended;
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace xlx
