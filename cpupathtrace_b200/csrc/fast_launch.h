// Launchers of the production-math instantiations of the wavefront's arithmetic kernels (ptb_fast.cu).
//
// ptb.cu is compiled with -fmad=false, IEEE division / square root and the glibc libm restatement because validation
// mode must reproduce the reference's x86 arithmetic bit for bit (DESIGN.md "Numerics").  The production generator
// (PTB_RNG_COUNTER) cannot be compared with the reference sample by sample in the first place, so its generate / shade /
// accumulate kernels are compiled a second time in ptb_fast.cu with FMA contraction and the SFU approximations
// (-use_fast_math).  The traversal kernels are NOT rebuilt: hits stay bit-identical in both modes.
//
// The POD parameter blocks cross this boundary as opaque pointers: the two translation units compile the same structure
// definitions (kernels.cuh) in two different namespaces so that the linker never merges an exact and a fast instantiation.
#ifndef PTB_FAST_LAUNCH_H
#define PTB_FAST_LAUNCH_H

#include <cuda_runtime.h>
#include <stdint.h>

namespace ptb_fast_api {

    // generateKernel<CounterRng>
    void launchGenerate(const void *pool, const void *params, const void *src, uint32_t count, uint32_t *queue, uint32_t *counters, int queue_slot,
                        cudaStream_t stream);
    // shadeKernel<CounterRng>
    void launchShade(const void *scene, const void *pool, const void *params, const uint32_t *queue, uint32_t *counters, int queue_slot, uint32_t *shadow_queue,
                     int grid, cudaStream_t stream);
    // accumulateKernel<CounterRng>
    void launchAccumulate(const void *pool, const void *params, const void *src, const uint32_t *queue, uint32_t *counters, int queue_slot, uint32_t *next_queue,
                          int next_slot, float4 *samples, unsigned long long *work_cursor, int grid, cudaStream_t stream);

}

#endif
