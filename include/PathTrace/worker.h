// Render entry points (API of the reference's include/PathTrace/worker.h).
//
// processItem / processJob keep their signatures; the work runs as a wavefront path tracer on the GPU
// (ptb_render).  Device selection and multi-GPU use are controlled by the environment, see ptb::RenderControl.
#ifndef PATHTRACE_WORKER_H
#define PATHTRACE_WORKER_H

#include <PathTrace/base.h>
#include <PathTrace/camera.h>
#include <PathTrace/image/image.h>
#include <PathTrace/scene/scene.h>

#include <cstddef>
#include <cstdint>
#include <functional>

struct RenderOptions {
    int image_width;
    int image_height;
    //! per pixel: samples always taken before the adaptive acceptance test may stop sampling
    int min_sample_count;
    //! per pixel: upper bound of samples
    int max_sample_count;

    //! ray offset and distance-comparison tolerance
    float epsilon;

    //! kept for source compatibility; the reference never reads it either
    bool allow_bias = false;
};

struct FrameRenderJob {
    const Camera &camera;
    const Scene &scene;
    const RenderOptions &options;
};

//! A rectangle of the image
struct WorkItem {
    const FrameRenderJob *job;
    int offset_x;
    int offset_y;
    int width;
    int height;

    WorkItem() noexcept;
    WorkItem(const FrameRenderJob *job, int offset_x, int offset_y, int width, int height) noexcept;
};

/**
 * Renders one tile.  Two words are drawn from `re` to key the device's counter-based generator, so that equal engine
 * states give equal tiles.
 */
Image<> processItem(const WorkItem &item, RandomEngine &re);

/**
 * Renders the whole frame on the GPU.  `progress_callback(done, total)` is called from the calling thread, once per
 * tile of the reference's tile grid, in order.  `worker_count` is accepted for source compatibility; parallelism is
 * the GPU's.
 */
Image<> processJob(
  const FrameRenderJob &job, const std::function<void(int, int)> &progress_callback = [](int, int) {}, int worker_count = 0);

namespace ptb {

    //! Knobs without a counterpart in RenderOptions.  The three query options are result-neutral (validation-mode samples
    //! are bit-identical with and without them, tests/test_parity_gpu.py) and therefore ON by default; each can be
    //! switched off to trace exactly the reference's rays in exactly the reference's order.
    struct RenderControl {
        int max_depth = 0;            //!< 0 = unlimited, paths end by Russian roulette only
        bool any_hit_shadows = true;  //!< stop shadow rays at the first occluder
        bool skip_null_shadows = true; //!< do not trace shadow rays whose contribution is always zero (glass, mirror, surfaces facing away)
        bool certified_closest = true; //!< closest hits on the SAH hierarchy where a certificate proves the reference's result, else re-traced
        bool relaxed_guard = true;    //!< processJob / processItem (counter-based generator) skip the certified walk's guard table; renderSamples never does
        int devices = 1;              //!< processJob: GPUs of this process to render on (ptb_render_multi; capped by the devices present)
        uint64_t fixed_seed = 0;      //!< processJob: non-zero replaces std::random_device
        int shard_index = 0;          //!< multi-GPU, one process per GPU: this process renders tile k of the frame's tile grid iff
        int shard_count = 1;          //!< k % shard_count == shard_index and leaves the other pixels 0 (sum-reduce the images)
    };

    //! process-wide control block; initialised ONCE (first use) from PTB_MAX_DEPTH, PTB_ANY_HIT_SHADOWS, PTB_SKIP_NULL_SHADOWS,
    //! PTB_CERTIFIED_CLOSEST, PTB_CERTIFIED_RELAXED, PTB_SEED, PTB_SHARD_INDEX, PTB_SHARD_COUNT, PTB_DEVICES; later changes go through this reference
    RenderControl &renderControl();

    /**
     * Validation entry: one path per (pixel, seed) with the reference engine RandomEngine(seed), i.e. exactly what
     * processItem computes for a 1x1 item at 1 spp with that engine.  pixels: 2 ints each; out: RGBA per sample.
     */
    void renderSamples(const FrameRenderJob &job, std::size_t count, const int *pixels, const uint64_t *seeds, float *out_rgba);

}

#endif // PATHTRACE_WORKER_H
