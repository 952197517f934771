/*
 * isomc_launch.cuh -- programmatic dependent launch (PDL) for the kernel chain of one extract.
 *
 * sign -> count -> scan -> emit are four dependent launches on one stream; measured in the stream they take 37 us more than the sum
 * of the kernels (fbm512).  With the programmatic-stream-serialization attribute the next kernel's CTAs are placed as soon as the
 * CTAs of the previous one retire, run their prologue (table copies into shared memory) and then wait in
 * cudaGridDependencySynchronize() until the previous grid has completed and its writes are visible.  Every kernel of the chain
 * calls isomc_pdl_trigger() first thing (lets its successor be placed early) and isomc_pdl_wait() before it touches anything an
 * earlier kernel of the chain wrote.  Both are no-ops for a launch without the attribute.  ISOMC_PDL=0 switches the attribute off.
 */
#ifndef ISOMC_LAUNCH_CUH
#define ISOMC_LAUNCH_CUH

#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

__device__ __forceinline__ void isomc_pdl_trigger() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void isomc_pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

static inline bool isomc_pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char *p = getenv("ISOMC_PDL"); v = p ? (atoi(p) != 0) : 1; }
    return v != 0;
}

/* dependent = the kernel may be placed before its predecessor in the stream has completed (it calls isomc_pdl_wait()) */
template <class... KArgs, class... Args>
static inline cudaError_t isomc_launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, bool dependent,
                                       Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (dependent && isomc_pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#endif
