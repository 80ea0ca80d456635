// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may become resident while its predecessor in the stream is still draining.  Every kernel that
// libssw launches that way starts with pdl_enter(): it releases ITS dependents (they will park at their own
// wait) and then blocks until all prerequisite grids have completed and their writes are visible.  Nothing
// is read or written before the wait, so the data dependencies between the passes are those of an
// ordinary in-order stream; what disappears is the launch gap and part of the drain/ramp between the
// many short kernels of one embed / extract step.  Launched without the attribute the two instructions
// are no-ops.
#pragma once

namespace ssw {

// 1: the short kernels (ordering, scatter/gather, similarity) release their dependents at their first instruction;
// 0: implicitly when they exit.  Written once per context (ssw_ctx_create), read-only for kernels.
__device__ int g_pdl_small_early = 1;

__device__ __forceinline__ void pdl_trigger() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__device__ __forceinline__ void pdl_enter() {
    if (g_pdl_small_early) pdl_trigger();
    pdl_wait();
}

}  // namespace ssw
