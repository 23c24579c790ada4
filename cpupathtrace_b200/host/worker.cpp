// Render entry points on top of ptb_render (the wavefront path tracer).
//
// Mapping of the reference's control flow (src/worker.cpp:328-424):
//   processJob   tile grid + thread pool + per-thread engines  ->  one ptb_render over the whole frame (the GPU grid
//                is the worker pool), or ptb_render_multi over several GPUs of this process; the progress callback fires
//                once per tile, in order, WHILE the frame renders (as retired samples reach each tile's share)
//   processItem  sequential per-pixel sampling with one engine ->  ptb_render over the tile's rectangle; the engine
//                supplies the job key of the device's counter-based generator
// The per-pixel statistics (batch Welford, adaptive acceptance, candidate merge) run in the resolve kernel.
#include "device.h"

#include <PathTrace/worker.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>

namespace {

    long envLong(const char *name, long fallback) {
        const char *v = std::getenv(name);
        return (v == nullptr || *v == '\0') ? fallback : std::atol(v);
    }

    // `production`: processJob / processItem (counter-based generator); false: the validation entry renderSamples
    ptb_render_opts makeOpts(const RenderOptions &options, uint64_t seed, uint32_t rng_mode, bool production) {
        const ptb::RenderControl &control = ptb::renderControl();
        ptb_render_opts opts{};
        opts.image_width = options.image_width;
        opts.image_height = options.image_height;
        opts.min_sample_count = options.min_sample_count;
        opts.max_sample_count = options.max_sample_count;
        opts.epsilon = options.epsilon;
        opts.max_depth = control.max_depth;
        opts.rng_mode = rng_mode;
        opts.flags = (control.any_hit_shadows ? PTB_FLAG_ANY_HIT_SHADOWS : 0U) | (control.skip_null_shadows ? PTB_FLAG_SKIP_NULL_SHADOWS : 0U) |
                     (control.certified_closest ? PTB_FLAG_CERTIFIED_CLOSEST : 0U) |
                     ((control.certified_closest && control.relaxed_guard && production) ? PTB_FLAG_CERTIFIED_RELAXED : 0U);
        opts.seed = seed;
        opts.shard_count = std::max(control.shard_count, 1);
        opts.shard_index = std::min(std::max(control.shard_index, 0), opts.shard_count - 1);
        return opts;
    }

    ptb_camera lowerCamera(const Camera &camera) {
        ptb_camera pod;
        if(!camera.lower(pod)) {
            throw std::logic_error("PathTrace (B200): the camera uses a user ApertureSampler subclass, which cannot run on the GPU");
        }
        return pod;
    }

    // Turns the render core's progress reports (pixel-samples retired so far) into the reference's callback protocol:
    // progress_callback(k, total_tiles) for k = 1 .. total_tiles, in order, never concurrently (worker.h:76-79).  Tile k is
    // reported as soon as k / total_tiles of the frame's samples are retired; the last tile only when the image is complete.
    struct TileProgress {
        const std::function<void(int, int)> *callback;
        int total_tiles;
        int reported;

        static void onSamples(void *user, uint64_t done, uint64_t total) {
            auto *self = static_cast<TileProgress *>(user);
            if(total == 0) {
                return;
            }
            const bool finished = done >= total;
            int target = static_cast<int>((static_cast<long double>(done) / static_cast<long double>(total)) * self->total_tiles);
            target = finished ? self->total_tiles : std::min(target, self->total_tiles - 1);
            while(self->reported < target) {
                self->reported++;
                (*self->callback)(self->reported, self->total_tiles);
            }
        }
    };

    Image<> renderRect(const FrameRenderJob &job, int x0, int y0, int w, int h, int tile_size, uint64_t seed, bool whole_job, TileProgress *progress) {
        Image<> image(std::max(w, 0), std::max(h, 0));
        if(w <= 0 || h <= 0) {
            return image;
        }
        static_assert(sizeof(Color<float>) == 4 * sizeof(float), "Color<float> must be four packed floats");
        const ptb_camera camera = lowerCamera(job.camera);
        ptb_render_opts opts = makeOpts(job.options, seed, PTB_RNG_COUNTER, true);
        opts.tile_size = tile_size;
        const int devices = whole_job && opts.shard_count <= 1 ? std::max(ptb::renderControl().devices, 1) : 1;
        if(!whole_job) {
            // processItem renders exactly the tile it was asked for; only processJob splits the frame
            opts.shard_index = 0;
            opts.shard_count = 1;
        }
        const ptb_progress_fn report = progress != nullptr ? &TileProgress::onSamples : nullptr;
        if(devices > 1) {
            const std::vector<ptb_scene *> replicas = job.scene.deviceScenes(devices);
            ptb::host::check(ptb_render_multi(replicas.data(), static_cast<int32_t>(replicas.size()), &camera, &opts, x0, y0, w, h, reinterpret_cast<float *>(image.data()),
                                              nullptr, report, progress),
                             "render on several GPUs");
        }
        else {
            ptb::host::check(ptb_render_with_progress(job.scene.deviceScene(), &camera, &opts, x0, y0, w, h, reinterpret_cast<float *>(image.data()), nullptr, report, progress),
                             "render");
        }
        return image;
    }

}

namespace ptb {

    RenderControl &renderControl() {
        static RenderControl control = [] {
            RenderControl c;
            c.max_depth = static_cast<int>(envLong("PTB_MAX_DEPTH", 0));
            c.any_hit_shadows = envLong("PTB_ANY_HIT_SHADOWS", 1) != 0;
            c.skip_null_shadows = envLong("PTB_SKIP_NULL_SHADOWS", 1) != 0;
            c.certified_closest = envLong("PTB_CERTIFIED_CLOSEST", 1) != 0;
            c.relaxed_guard = envLong("PTB_CERTIFIED_RELAXED", 1) != 0;
            c.shard_index = static_cast<int>(envLong("PTB_SHARD_INDEX", 0));
            c.shard_count = static_cast<int>(std::max(1L, envLong("PTB_SHARD_COUNT", 1)));
            c.fixed_seed = static_cast<uint64_t>(envLong("PTB_SEED", 0));
            const char *devices = std::getenv("PTB_DEVICES");
            if(devices != nullptr && std::strcmp(devices, "all") == 0) {
                int count = 1;
                c.devices = (ptb_device_count(&count) == PTB_OK && count > 0) ? count : 1;
            }
            else {
                c.devices = static_cast<int>(std::max(1L, envLong("PTB_DEVICES", 1)));
            }
            return c;
        }();
        return control;
    }

    void renderSamples(const FrameRenderJob &job, std::size_t count, const int *pixels, const uint64_t *seeds, float *out_rgba) {
        const ptb_camera camera = lowerCamera(job.camera);
        const ptb_render_opts opts = makeOpts(job.options, 0, PTB_RNG_REFERENCE_XORSHIFT, false);
        host::check(ptb_render_samples(job.scene.deviceScene(), &camera, &opts, count, pixels, seeds, out_rgba, nullptr), "renderSamples");
    }

}

WorkItem::WorkItem() noexcept : job(nullptr), offset_x(0), offset_y(0), width(0), height(0) {}

WorkItem::WorkItem(const FrameRenderJob *job, int offset_x, int offset_y, int width, int height) noexcept :
  job(job), offset_x(offset_x), offset_y(offset_y), width(width), height(height) {}

Image<> processItem(const WorkItem &item, RandomEngine &re) {
    const uint64_t low = re();
    const uint64_t high = re();
    const uint64_t seed = (high << 32) | low;
    // one tile: a single group of pixels, no further subdivision
    return renderRect(*item.job, item.offset_x, item.offset_y, item.width, item.height, std::max(std::max(item.width, item.height), 1), seed, false, nullptr);
}

Image<> processJob(const FrameRenderJob &job, const std::function<void(int, int)> &progress_callback, int /*worker_count*/) {
    const int width = std::max(job.options.image_width, 0);
    const int height = std::max(job.options.image_height, 0);
    if(width == 0 || height == 0) {
        return Image<>(width, height);
    }

    uint64_t seed = ptb::renderControl().fixed_seed;
    if(seed == 0) {
        std::random_device device;
        seed = (static_cast<uint64_t>(device()) << 32) | device();
    }

    // the reference's tile grid (worker.cpp:398-402) decides how many progress callbacks fire
    const int tile_size = std::max(std::min(std::min(width, height) / 4, 32), 1);
    const int horizontal_tiles = (width + tile_size - 1) / tile_size;
    const int vertical_tiles = (height + tile_size - 1) / tile_size;
    const int total_tiles = horizontal_tiles * vertical_tiles;

    if(ptb::renderControl().shard_count > 1) {
        static bool warned = false;
        if(!warned) {
            warned = true;
            std::fprintf(stderr, "PathTrace (B200): processJob renders shard %d of %d of the tile grid (ptb::RenderControl / PTB_SHARD_*); the image holds zeros in the other "
                                 "shards' tiles until the shards are summed\n", ptb::renderControl().shard_index, ptb::renderControl().shard_count);
        }
    }
    TileProgress progress{&progress_callback, total_tiles, 0};
    Image<> image = renderRect(job, 0, 0, width, height, tile_size, seed, true, &progress);
    TileProgress::onSamples(&progress, 1, 1); // whatever the render core did not report (e.g. an image without samples)
    return image;
}
