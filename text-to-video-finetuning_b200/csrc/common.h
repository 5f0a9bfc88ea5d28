// Shared host-side helpers for the C ABI: error reporting and launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/t2v_b200.h"

namespace t2v {
// Records a formatted thread-local error message and returns `code` (negative).
int fail(int code, const char* fmt, ...);
// Maps a cudaError_t (as int) from a launch to the ABI convention; counts the launch on success.
int launch_checked(int cuda_err, const char* what);
void count_launch(int n = 1);
// false when T2V_NO_PDL is set in the environment (A/B switch for programmatic dependent launch)
bool pdl_enabled();

// Every kernel of this library is launched with programmatic stream serialization: its blocks may become resident while
// the previous kernel of the stream is still draining, and the kernel itself calls pdl_sync() before it touches global
// memory.  That hides the ~2 us launch latency between the ~3300 dependent kernels of a step (also inside CUDA graphs,
// where the attribute becomes a programmatic dependency edge).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
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
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// Let the next kernel of the stream start launching, then wait until every kernel before this one has completed and
// its writes are visible.  Must precede the first global-memory access of a kernel launched with launch_pdl.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
}  // namespace t2v
