// Production-math build of the wavefront's arithmetic kernels (see fast_launch.h).  Compiled with
//   -fmad=true -use_fast_math -DPTB_FAST_MATH=1
// Everything kernels.cuh declares lands in namespace ptb_fast here, so these instantiations never collide with the exact
// ones of ptb.cu at link time.
#ifndef PTB_FAST_TU_EXACT // (experiments only: the same translation unit with the exact arithmetic)
#define PTB_FAST_MATH 1
#endif
#define ptb ptb_fast
#include "kernels.cuh"
#undef ptb

#include "fast_launch.h"

namespace ptb_fast_api {

    using namespace ptb_fast;

    void launchGenerate(const void *pool, const void *params, const void *src, uint32_t count, uint32_t *queue, uint32_t *counters, int queue_slot,
                        cudaStream_t stream) {
        generateKernel<CounterRng><<<(count + kBlock - 1) / kBlock, kBlock, 0, stream>>>(*static_cast<const PathPool *>(pool), *static_cast<const RenderParams *>(params),
                                                                                        *static_cast<const PathSource *>(src), count, queue, counters, queue_slot);
    }

    void launchShade(const void *scene, const void *pool, const void *params, const uint32_t *queue, uint32_t *counters, int queue_slot, uint32_t *shadow_queue,
                     int grid, cudaStream_t stream) {
        shadeKernel<CounterRng><<<grid, kFlatBlock, 0, stream>>>(*static_cast<const DeviceScene *>(scene), *static_cast<const PathPool *>(pool),
                                                             *static_cast<const RenderParams *>(params), queue, counters, queue_slot, shadow_queue);
    }

    void launchAccumulate(const void *pool, const void *params, const void *src, const uint32_t *queue, uint32_t *counters, int queue_slot, uint32_t *next_queue,
                          int next_slot, float4 *samples, unsigned long long *work_cursor, int grid, cudaStream_t stream) {
        accumulateKernel<CounterRng><<<grid, kFlatBlock, 0, stream>>>(*static_cast<const PathPool *>(pool), *static_cast<const RenderParams *>(params),
                                                                  *static_cast<const PathSource *>(src), queue, counters, queue_slot, next_queue, next_slot, samples,
                                                                  work_cursor);
    }

}
