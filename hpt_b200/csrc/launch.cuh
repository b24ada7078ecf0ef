// launch.cuh — programmatic dependent launch (PDL) for every kernel of the library.
//
// The hot-path kernels of the BASELINE configurations run for 15–100 µs; the 2–4 µs between two dependent
// launches on a stream (grid drain, launch latency, prologue of the next grid) is a visible fraction of that.
// Every kernel therefore (a) signals `griddepcontrol.launch_dependents` as its first instruction, so the NEXT
// kernel on the stream may start occupying SM slots as this grid's last CTAs retire, and (b) executes
// `griddepcontrol.wait` before its first global-memory access, which blocks until the PREVIOUS grid has
// completed and its writes are visible — stream order is unchanged, only launch latency and prologues overlap.
// Launches go through cudaLaunchKernelEx with cudaLaunchAttributeProgrammaticStreamSerialization.
// HPTB_NO_PDL=1 in the environment turns the attribute off (plain stream serialisation) for A/B measurements.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace hptb {

__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("HPTB_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

// the same with a thread-block cluster of `cluster_x` CTAs along x (distributed shared memory; ≤ 8 is portable)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_x,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

}  // namespace hptb
