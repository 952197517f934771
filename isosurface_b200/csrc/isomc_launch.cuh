/*
 * isomc_launch.cuh -- programmatic dependent launch (PDL) for the kernel chain of one extract.
 *
 * sign -> count -> scan -> emit are four dependent launches on one stream; measured in the stream they take 37 us more than the sum
 * of the kernels (fbm512).  With the programmatic-stream-serialization attribute the next kernel's CTAs are placed as soon as the
 * CTAs of the previous one retire, run their prologue (table copies into shared memory) and then wait in
 * cudaGridDependencySynchronize() until the previous grid has completed and its writes are visible.  Every kernel of the chain
 * calls isomc_pdl_trigger() first thing (lets its successor be placed early) and isomc_pdl_wait() before it touches anything an
 * earlier kernel of the chain wrote.  Both are no-ops for a launch without the attribute.  ISOMC_PDL=0 / 1 forces it off / on.
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

static inline int isomc_pdl_mode() { /* ISOMC_PDL = 0: never, 1: always, unset: by lattice size */
    static int v = -2;
    if (v == -2) { const char *p = getenv("ISOMC_PDL"); v = p ? (atoi(p) != 0) : -1; }
    return v;
}
/* Measured (profiles/r02_history.md): the early placement saves 8-10 us per extract at 512^3 and below, and COSTS 25 us at
 * 1024^3 and 110 us at 2048^3 (the loss grows with the run time of the kernels the waiting CTAs sit behind), so it is used for
 * lattices of up to 3 * 10^8 samples only. */
static inline bool isomc_pdl_for(unsigned long long n_samples) {
    const int m = isomc_pdl_mode();
    return m < 0 ? n_samples <= 300000000ull : m != 0;
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
    cfg.numAttrs = dependent ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#endif
