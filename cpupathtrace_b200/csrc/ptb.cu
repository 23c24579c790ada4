// C-ABI implementation (include/ptb.h): context, scene upload and the host driver of the wavefront pipeline.
// Everything that computes goes through the kernels in kernels.cuh; there is no host implementation of the hot path.
#include "../../include/ptb.h"

#include "bvh_build.h"
#include "cert_guard.h"
#include "fast_launch.h"
#include "host_math.h"
#include "kernels.cuh"
#include "lbvh.cuh"
#include "gpu_build.cuh"
#include "post_process.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace {

    thread_local std::string g_last_error;

    int fail(int status, const std::string &message) {
        g_last_error = message;
        return status;
    }

#define PTB_CUDA(call)                                                                                                                                        \
    do {                                                                                                                                                      \
        cudaError_t err__ = (call);                                                                                                                           \
        if(err__ != cudaSuccess) {                                                                                                                            \
            cudaGetLastError();                                                                                                                               \
            return fail(err__ == cudaErrorMemoryAllocation ? PTB_ERR_OUT_OF_MEMORY : PTB_ERR_CUDA,                                                            \
                        std::string(#call) + ": " + cudaGetErrorString(err__));                                                                               \
        }                                                                                                                                                     \
    } while(0)

    double nowSeconds() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    long envLong(const char *name, long fallback) {
        const char *v = std::getenv(name);
        if(v == nullptr || *v == '\0') {
            return fallback;
        }
        return std::atol(v);
    }

    // A device allocation that grows on demand and is reused across calls
    struct Buffer {
        void *ptr = nullptr;
        size_t bytes = 0;

        int reserve(size_t wanted) {
            if(wanted <= bytes) {
                return PTB_OK;
            }
            if(ptr != nullptr) {
                cudaFree(ptr);
                ptr = nullptr;
                bytes = 0;
            }
            PTB_CUDA(cudaMalloc(&ptr, wanted));
            bytes = wanted;
            return PTB_OK;
        }

        void release() {
            if(ptr != nullptr) {
                cudaFree(ptr);
            }
            ptr = nullptr;
            bytes = 0;
        }

        template<typename T>
        T *as() const {
            return static_cast<T *>(ptr);
        }
    };

    constexpr int kEventPairs = 2048;
    constexpr int kMaxIterationsPerSync = 16;

}

struct ptb_context {
    std::mutex mutex; // entry points serialise on their context: its workspace and stream are shared state
    int device = 0;
    int sm_count = 0;
    ptb::VoteParams vote{12, 12, 2, 4};        // closest-hit kernels
    ptb::VoteParams vote_shadow{16, 12, 2, 4}; // any-hit / shadow kernels (shorter rays: refill in larger batches)
    int trace_blocks_per_sm = 16;
    int shade_blocks_per_sm = PTB_SHADE_MIN_BLOCKS; // grid of the shade kernel = its resident blocks (PTB_SHADE_BLOCKS_PER_SM)
    bool log_iterations = false; // PTB_LOG_ITERATIONS=1: one stderr line per bounce iteration
    bool production_math = true; // PTB_RNG_COUNTER renders use the FMA / SFU build of generate, shade and accumulate (PTB_PRODUCTION_MATH=0: the exact build)
    int iterations_per_sync = 8; // bounce iterations per batch; two batches are in flight (PTB_ITERATIONS_PER_SYNC).  Measured, six timed frames each: 4 per batch
                                 // without double buffering 2-3 outliers of +40...80 ms, 4 with double buffering 0-1, 8 with double buffering 0
    cudaStream_t stream = nullptr;

    // wavefront workspace
    Buffer pool_mem;
    Buffer queue_a;
    Buffer queue_b;
    Buffer shadow_queue;
    Buffer redo_queue; // paths / rays the certified closest-hit walk handed back for a re-trace on the reference tree
    Buffer counters;
    Buffer visits;
    Buffer work_cursor;
    Buffer samples;
    Buffer pixel_list;
    Buffer pixel_states;  // adaptive sampling (min != max): parked processItem loop state per pixel, the two active-pixel lists and their counters
    Buffer active_lists;
    Buffer adaptive_counters;
    bool adaptive_rounds = true; // PTB_ADAPTIVE_ROUNDS=0: trace all max_sample_count samples of every pixel and let the resolve discard the surplus
    long long pixel_list_key[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1}; // what the device-side pixel list currently holds
    Buffer io_a; // staging for host<->device bulk arrays
    Buffer io_b;
    Buffer io_c;
    Buffer io_d;
    Buffer sort_keys;   // batch queries: ray sort keys (in | out), ray numbers (in | out), cub workspace
    Buffer sort_ids;
    Buffer sort_temp;
    bool sort_rays = true; // PTB_SORT_RAYS=0: trace batches in the caller's order
    int sort_dir_bits = 1; // bits per axis of the ray direction in the sort key (PTB_SORT_DIR_BITS; 1 = the octant)
    // PTB_STREAMS > 1: large renders are split between `streams` contexts of this device (this one and streams - 1 siblings
    // with their own stream and workspace) that run concurrently, so that the drain phase of one share's persistent trace
    // launch and the launch-bound tail of its bounce loop overlap with the other share's kernels.  Off by default: the gain
    // depends on the phase the two shares happen to fall into (bench scene, 256 spp: 778 -> 737-745 ms per frame in most
    // runs, 795-811 in others; 1024 spp with depth 16: 2 % slower, and one end-to-end run 50 % slower) -- DESIGN.md section 4.
    int streams = 1;
    bool is_sibling = false;
    std::vector<ptb_context *> siblings;
    std::mutex call_mutex;      // serialises split calls on this context
    Buffer split_image;         // split call with a host result: the image the shares write into
    cudaEvent_t split_start = nullptr;
    cudaEvent_t split_stop = nullptr;
    int budget_divisor = 1;     // shares of a split call divide the per-sample buffer budget between them
    Buffer build_arena;   // scene setup on the device: every temporary of a tree build is carved from this one allocation
    Buffer multi_image;   // ptb_render_multi: this replica's share of the frame (its tiles, zeros elsewhere)
    Buffer multi_staging; // ptb_render_multi on the first replica: copies of the other replicas' images when peers cannot map each other
    uint32_t *host_counters = nullptr; // pinned: two batches of bounce iterations (double-buffered)
    cudaEvent_t batch_done[2] = {nullptr, nullptr};
    bool log_batches = false;      // PTB_LOG_BATCHES=1: one stderr line per batch of bounce iterations (device time since the previous batch ended)
    bool pipelined_batches = true; // PTB_PIPELINED_BATCHES=0: the host waits for every batch before enqueuing the next
    unsigned long long *host_cursor = nullptr; // pinned: the work cursor as of the last batch of bounce iterations

    // profiling events: [pair][0 = start, 1 = stop], class 0 = closest-hit trace, 2 = shadow trace, 1 = everything else
    cudaEvent_t events[kEventPairs][2];
    int event_class[kEventPairs];
    int events_used = 0;
    bool events_ready = false;
    bool profile_all = false; // this call times every launch (PTB_FLAG_PROFILE_ALL), not only the closest-hit trace
    cudaEvent_t call_start = nullptr;
    cudaEvent_t call_stop = nullptr;
};

struct ptb_scene {
    ptb_context *ctx = nullptr;
    ptb::DeviceScene dev{};
    Buffer nodes;
    Buffer occ_nodes;
    Buffer geom;
    Buffer shade;
    Buffer mats;
    Buffer lights;
    Buffer emis;
    Buffer cdf;
    Buffer slot_to_prim;
    ptb_scene_info info{};
    uint32_t shadow_stride = 1;
    ptb_guard::CertGuard guard{}; // guard table of the certified closest-hit walk (passed to its kernels by value)
    std::vector<ptb_scene *> aliases; // the same device arrays seen from the sibling contexts (no ownership)
};

namespace {

    using namespace ptb;

    int useDevice(const ptb_context *ctx) {
        PTB_CUDA(cudaSetDevice(ctx->device));
        return PTB_OK;
    }

    int gridFor(const ptb_context *ctx, int blocks_per_sm) {
        return std::max(1, ctx->sm_count) * blocks_per_sm;
    }

    // ---- profiling helpers: every launch is bracketed by an event pair on the launching stream

    struct LaunchTimer {
        ptb_context *ctx;
        int pair;
        LaunchTimer(ptb_context *c, int klass) : ctx(c), pair(-1) {
            if(ctx->events_ready && (klass == 0 || ctx->profile_all) && ctx->events_used < kEventPairs) {
                pair = ctx->events_used++;
                ctx->event_class[pair] = klass;
                cudaEventRecord(ctx->events[pair][0], ctx->stream);
            }
        }
        ~LaunchTimer() {
            if(pair >= 0) {
                cudaEventRecord(ctx->events[pair][1], ctx->stream);
            }
        }
    };

    void collectTimers(ptb_context *ctx, ptb_render_stats *stats) {
        for(int i = 0; i < ctx->events_used; i++) {
            float ms = 0.0F;
            if(cudaEventElapsedTime(&ms, ctx->events[i][0], ctx->events[i][1]) == cudaSuccess && stats != nullptr) {
                if(ctx->event_class[i] == 0) {
                    stats->device_ms_trace += ms;
                }
                else if(ctx->event_class[i] == 2) {
                    stats->device_ms_trace += ms;
                    stats->device_ms_trace_shadow += ms;
                }
                else {
                    stats->device_ms_shade += ms;
                }
            }
        }
        ctx->events_used = 0;
    }

    // When the event pool runs dry mid-call the remaining launches go untimed; callers that need complete
    // per-kernel sums (bench.py) keep calls short enough (see PTB_POOL_PATHS) or read device_ms_total.

    // HBM this context may plan with: what is free now plus what its own workspace already holds from earlier calls
    uint64_t plannableBytes(const ptb_context *ctx) {
        size_t free_bytes = 0;
        size_t total_bytes = 0;
        if(cudaMemGetInfo(&free_bytes, &total_bytes) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return static_cast<uint64_t>(free_bytes) + ctx->pool_mem.bytes + ctx->queue_a.bytes + ctx->queue_b.bytes + ctx->shadow_queue.bytes + ctx->redo_queue.bytes +
               ctx->samples.bytes;
    }

    // Paths kept in flight.  The trace kernels are persistent and end every launch with a drain phase in which each warp
    // finishes its last rays at low lane occupancy; its length does not depend on the queue length, so the pool is made
    // as large as memory comfortably allows (measured on the bench scene at 256 spp, round 1: 4 Mi paths 321, 32 Mi 438,
    // 128 Mi 469 Msamples/s; round 2: 128 Mi 769, 160 Mi 756, 192 Mi 751, 256 Mi 745 ms per frame).  Default: 256 Mi paths
    // (~63 GB with two shadow slots), at most three eighths of the HBM at hand.
    uint64_t poolLimit(const ptb_context *ctx, uint32_t shadow_stride) {
        const long forced = envLong("PTB_POOL_PATHS", 0);
        if(forced > 0) {
            return static_cast<uint64_t>(forced);
        }
        const uint64_t bytes_per_path = 112ULL + 48ULL * shadow_stride + 4ULL * (3ULL + shadow_stride);
        const uint64_t by_memory = plannableBytes(ctx) * 3ULL / 8ULL / static_cast<uint64_t>(std::max(ctx->budget_divisor, 1)) / bytes_per_path;
        return std::max<uint64_t>(1ULL << 16, std::min<uint64_t>(1ULL << 28, by_memory));
    }

    // Per-sample buffer budget of ptb_render (16 B per pixel-sample): PTB_SAMPLE_BUFFER_MB, else 40 % of the HBM at hand.
    uint64_t sampleBufferBudget(const ptb_context *ctx) {
        const long forced = envLong("PTB_SAMPLE_BUFFER_MB", 0);
        if(forced > 0) {
            return static_cast<uint64_t>(forced) << 20;
        }
        return std::max<uint64_t>(64ULL << 20, plannableBytes(ctx) * 2ULL / 5ULL / static_cast<uint64_t>(std::max(ctx->budget_divisor, 1)));
    }

    // Pool capacity for a call of `total` work items: never more than half of them (above 2 Mi) so that path regeneration
    // keeps the queues full for the first half of the call instead of starting everything in one wave that then thins
    // out over dozens of iterations (8 GPUs x 66 M samples each: one wave 3.40, half 3.64 Gsamples/s).
    uint32_t poolCapacity(uint64_t total, uint64_t limit) {
        const uint64_t wanted = total <= (2ULL << 20) ? total : std::max<uint64_t>(2ULL << 20, (total + 1) / 2);
        return static_cast<uint32_t>(std::min<uint64_t>(wanted, limit));
    }

    int carvePool(ptb_context *ctx, uint32_t capacity, uint32_t shadow_stride, PathPool &pool) {
        const size_t n = capacity;
        const size_t ns = n * shadow_stride;
        size_t offset = 0;
        auto take = [&](size_t bytes) {
            const size_t at = offset;
            offset += (bytes + 255) & ~static_cast<size_t>(255);
            return at;
        };
        const size_t o_ray_o = take(n * sizeof(float4));
        const size_t o_ray_d = take(n * sizeof(float4));
        const size_t o_hit = take(n * sizeof(float2));
        const size_t o_thr = take(n * sizeof(float4));
        const size_t o_rad = take(n * sizeof(float4));
        const size_t o_div = take(n * sizeof(double));
        const size_t o_bpd = take(n * sizeof(double));
        const size_t o_con = take(n * sizeof(float));
        const size_t o_state = take(n * sizeof(uint32_t));
        const size_t o_rng = take(n * sizeof(uint64_t));
        const size_t o_dest = take(n * sizeof(uint32_t));
        const size_t o_so = take(ns * sizeof(float4));
        const size_t o_sd = take(ns * sizeof(float4));
        const size_t o_sc = take(ns * sizeof(float4));

        int status = ctx->pool_mem.reserve(offset);
        if(status != PTB_OK) {
            return status;
        }
        char *base = ctx->pool_mem.as<char>();
        pool.ray_o = reinterpret_cast<float4 *>(base + o_ray_o);
        pool.ray_d = reinterpret_cast<float4 *>(base + o_ray_d);
        pool.hit = reinterpret_cast<float2 *>(base + o_hit);
        pool.throughput = reinterpret_cast<float4 *>(base + o_thr);
        pool.radiance = reinterpret_cast<float4 *>(base + o_rad);
        pool.divisor = reinterpret_cast<double *>(base + o_div);
        pool.bounce_pd = reinterpret_cast<double *>(base + o_bpd);
        pool.contribution = reinterpret_cast<float *>(base + o_con);
        pool.state = reinterpret_cast<uint32_t *>(base + o_state);
        pool.rng = reinterpret_cast<uint64_t *>(base + o_rng);
        pool.dest = reinterpret_cast<uint32_t *>(base + o_dest);
        pool.shadow_o = reinterpret_cast<float4 *>(base + o_so);
        pool.shadow_d = reinterpret_cast<float4 *>(base + o_sd);
        pool.shadow_c = reinterpret_cast<float4 *>(base + o_sc);
        pool.shadow_stride = shadow_stride;
        pool.capacity = capacity;

        if((status = ctx->queue_a.reserve(n * sizeof(uint32_t))) != PTB_OK) {
            return status;
        }
        if((status = ctx->queue_b.reserve(n * sizeof(uint32_t))) != PTB_OK) {
            return status;
        }
        if((status = ctx->shadow_queue.reserve(ns * sizeof(uint32_t))) != PTB_OK) {
            return status;
        }
        if((status = ctx->redo_queue.reserve(n * sizeof(uint32_t))) != PTB_OK) {
            return status;
        }
        if((status = ctx->counters.reserve(kCounterSlots * sizeof(uint32_t))) != PTB_OK) {
            return status;
        }
        if((status = ctx->visits.reserve(2 * sizeof(VisitCounters))) != PTB_OK) {
            return status;
        }
        if((status = ctx->work_cursor.reserve(sizeof(unsigned long long))) != PTB_OK) {
            return status;
        }
        return PTB_OK;
    }

    RenderParams makeParams(const ptb_camera &camera, const ptb_render_opts &opts) {
        RenderParams p{};
        p.camera = camera;
        p.image_width = opts.image_width;
        p.image_height = opts.image_height;
        p.epsilon = opts.epsilon;
        p.max_depth = opts.max_depth;
        p.rng_xorshift = opts.rng_mode == PTB_RNG_REFERENCE_XORSHIFT ? 1U : 0U;
        p.any_hit_shadows = (opts.flags & PTB_FLAG_ANY_HIT_SHADOWS) != 0U ? 1U : 0U;
        p.skip_null_shadows = (opts.flags & PTB_FLAG_SKIP_NULL_SHADOWS) != 0U ? 1U : 0U;
        p.seed = opts.seed;
        return p;
    }

    // Starts the first `first_wave` work items of `src` in the pool and runs bounce iterations until every work item of
    // the call has been retired into `samples` (retired slots are refilled by the accumulate kernel).
    // How closest-hit queries are answered for a call's `flags`: the certified walk needs the query hierarchy and, unless
    // relaxed, a scene the guard table covers.
    struct ClosestMode {
        bool certified;
        int guarded;
    };

    ClosestMode closestMode(const ptb_scene *scene, uint32_t flags) {
        ClosestMode m{};
        const bool relaxed = (flags & PTB_FLAG_CERTIFIED_RELAXED) != 0U;
        m.certified = (flags & PTB_FLAG_CERTIFIED_CLOSEST) != 0U && scene->dev.occ_nodes != nullptr && (relaxed || scene->guard.certifiable != 0U);
        m.guarded = relaxed ? 0 : 1;
        return m;
    }

    // Where progress reports of a frame go: `done_before` pixel-samples were finished by earlier pixel groups of the call.
    struct ProgressSink {
        ptb_progress_fn fn;
        void *user;
        uint64_t done_before;
        uint64_t total;
    };

    int runBounces(ptb_scene *scene, const PathPool &pool, const RenderParams &params, const PathSource &src, float4 *samples, bool count_visits,
                   ClosestMode closest, ptb_render_stats *stats, const ProgressSink *progress = nullptr) {
        ptb_context *ctx = scene->ctx; // the calling entry point holds ctx->mutex
        uint32_t *counters = ctx->counters.as<uint32_t>();
        uint32_t *queues[2] = {ctx->queue_a.as<uint32_t>(), ctx->queue_b.as<uint32_t>()};
        uint32_t *shadow_queue = ctx->shadow_queue.as<uint32_t>();
        uint32_t *redo_queue = ctx->redo_queue.as<uint32_t>();
        VisitCounters *visits = ctx->visits.as<VisitCounters>();
        const bool certified = closest.certified;

        const int trace_grid = gridFor(ctx, ctx->trace_blocks_per_sm);
        const uint32_t first_wave = static_cast<uint32_t>(std::min<unsigned long long>(pool.capacity, src.total));
        unsigned long long *work_cursor = ctx->work_cursor.as<unsigned long long>();
        {
            const unsigned long long started = first_wave;
            PTB_CUDA(cudaMemcpyAsync(work_cursor, &started, sizeof(started), cudaMemcpyHostToDevice, ctx->stream));
            PTB_CUDA(cudaStreamSynchronize(ctx->stream)); // `started` lives on this stack frame
            LaunchTimer timer(ctx, 1);
            if(params.rng_xorshift != 0U) {
                generateKernel<ReferenceRng><<<(first_wave + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(pool, params, src, first_wave, queues[0], counters,
                                                                                                             kCountQueueA);
            }
            else if(ctx->production_math) {
                ptb_fast_api::launchGenerate(&pool, &params, &src, first_wave, queues[0], counters, kCountQueueA, ctx->stream);
            }
            else {
                generateKernel<CounterRng><<<(first_wave + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(pool, params, src, first_wave, queues[0], counters,
                                                                                                           kCountQueueA);
            }
        }
        PTB_CUDA(cudaGetLastError());
        if(stats != nullptr) {
            stats->samples += src.total;
            stats->kernel_launches += 1;
        }
        int cur = 0;
        uint32_t n_cur = first_wave;
        const int queue_slot[2] = {kCountQueueA, kCountQueueB}; // where the length of queues[0] / queues[1] lives

        // Bounce iterations are launched in batches without a host round trip in between: queue lengths live in device
        // memory, the ping-pong order is known in advance, and an iteration on an empty queue is five launches that return
        // at once.  The host reads the counters of every iteration of the batch afterwards (statistics, termination); a
        // descheduled host thread then delays one batch boundary instead of every iteration.  Queues never grow, so the
        // length known at the start of a batch bounds the grids of all its iterations.
        // closest-hit trace of the paths in `queue` (length counters[queue_slot]) with the walk `mode`
        auto launch_closest = [&](auto mode, const uint32_t *queue, int queue_slot, int cursor_slot) {
            constexpr int kMode = decltype(mode)::value;
            uint32_t *redo_out = kMode == kTraceCertified ? redo_queue : nullptr; // only the certified walk hands rays back
            if(count_visits) {
                traceClosestKernel<kMode, true><<<trace_grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote, pool, queue, counters, queue_slot, cursor_slot, redo_out, visits,
                                                                                        scene->guard, closest.guarded);
            }
            else {
                traceClosestKernel<kMode, false><<<trace_grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote, pool, queue, counters, queue_slot, cursor_slot, redo_out, visits,
                                                                                         scene->guard, closest.guarded);
            }
        };

        // Batches are double-buffered: batch b + 1 is enqueued BEFORE the host waits for batch b's counters, so the GPU never
        // idles at a batch boundary while the host thread reads counters, reports progress or is descheduled (measured: up to
        // 26 ms of a 790 ms frame were such gaps).  Termination is noticed one batch late; the surplus iterations run on empty
        // queues.
        const int batch = ctx->iterations_per_sync;
        struct Inflight {
            int launched;
            int cur_after; // `cur` after the batch was enqueued
        };
        Inflight inflight[2] = {};
        auto enqueue_batch = [&](int buffer) -> int {
            uint32_t *host_counters = ctx->host_counters + buffer * kMaxIterationsPerSync * kCounterSlots;
            int launched = 0;
            for(; launched < batch; launched++) {
                const int nxt = cur ^ 1;
                // zero: next queue length; shadow queue length, both fetch cursors and the per-iteration statistics (slots 2..6)
                PTB_CUDA(cudaMemsetAsync(counters + queue_slot[nxt], 0, sizeof(uint32_t), ctx->stream));
                PTB_CUDA(cudaMemsetAsync(counters + kCountShadow, 0, kPerIterationCounters * sizeof(uint32_t), ctx->stream));

                const int flat_grid = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n_cur) + kFlatBlock - 1) / kFlatBlock, static_cast<uint64_t>(gridFor(ctx, 32 * kBlock / kFlatBlock))));
                // the shade kernel splits its queue statically between blocks that all run the whole launch: a grid of exactly
                // the resident blocks (6 per SM at 80 registers) has no last partial wave (32 per SM = 5.33 waves left a third
                // of the slots empty for the last sixth of every launch: shade + accumulate 303 -> 291 ms per frame)
                const int shade_grid = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n_cur) + kFlatBlock - 1) / kFlatBlock, static_cast<uint64_t>(gridFor(ctx, ctx->shade_blocks_per_sm))));
                {
                    LaunchTimer timer(ctx, 0);
                    if(certified) {
                        // SAH walk with certificate, then the handed-back rays on the reference tree (usually a handful)
                        launch_closest(std::integral_constant<int, kTraceCertified>{}, queues[cur], queue_slot[cur], kCountFetchClosest);
                        launch_closest(std::integral_constant<int, kTraceClosest>{}, redo_queue, kCountRedo, kCountFetchRedo);
                    }
                    else {
                        launch_closest(std::integral_constant<int, kTraceClosest>{}, queues[cur], queue_slot[cur], kCountFetchClosest);
                    }
                }
                {
                    LaunchTimer timer(ctx, 1);
                    if(params.rng_xorshift != 0U) {
                        shadeKernel<ReferenceRng><<<shade_grid, kFlatBlock, 0, ctx->stream>>>(scene->dev, pool, params, queues[cur], counters, queue_slot[cur], shadow_queue);
                    }
                    else if(ctx->production_math) {
                        ptb_fast_api::launchShade(&scene->dev, &pool, &params, queues[cur], counters, queue_slot[cur], shadow_queue, shade_grid, ctx->stream);
                    }
                    else {
                        shadeKernel<CounterRng><<<shade_grid, kFlatBlock, 0, ctx->stream>>>(scene->dev, pool, params, queues[cur], counters, queue_slot[cur], shadow_queue);
                    }
                }
                {
                    LaunchTimer timer(ctx, 2);
                    if(count_visits) {
                        traceShadowKernel<true><<<trace_grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote_shadow, pool, shadow_queue, counters, params.any_hit_shadows, visits + 1);
                    }
                    else {
                        traceShadowKernel<false><<<trace_grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote_shadow, pool, shadow_queue, counters, params.any_hit_shadows, visits + 1);
                    }
                }
                {
                    LaunchTimer timer(ctx, 1);
                    if(params.rng_xorshift != 0U) {
                        accumulateKernel<ReferenceRng><<<flat_grid, kFlatBlock, 0, ctx->stream>>>(pool, params, src, queues[cur], counters, queue_slot[cur], queues[nxt],
                                                                                            queue_slot[nxt], samples, work_cursor);
                    }
                    else if(ctx->production_math) {
                        ptb_fast_api::launchAccumulate(&pool, &params, &src, queues[cur], counters, queue_slot[cur], queues[nxt], queue_slot[nxt], samples, work_cursor, flat_grid,
                                                       ctx->stream);
                    }
                    else {
                        accumulateKernel<CounterRng><<<flat_grid, kFlatBlock, 0, ctx->stream>>>(pool, params, src, queues[cur], counters, queue_slot[cur], queues[nxt],
                                                                                          queue_slot[nxt], samples, work_cursor);
                    }
                }
                PTB_CUDA(cudaMemcpyAsync(host_counters + launched * kCounterSlots, counters, kCounterSlots * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
                cur = nxt;
            }
            if(progress != nullptr) {
                PTB_CUDA(cudaMemcpyAsync(ctx->host_cursor, work_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
            }
            PTB_CUDA(cudaEventRecord(ctx->batch_done[buffer], ctx->stream));
            PTB_CUDA(cudaGetLastError());
            inflight[buffer] = Inflight{launched, cur};
            return PTB_OK;
        };

        int status = PTB_OK;
        int oldest = 0;
        bool ahead = false; // a batch beyond `oldest` is in flight
        if(n_cur > 0U && (status = enqueue_batch(0)) != PTB_OK) {
            return status;
        }
        while(n_cur > 0U) {
            if(!ahead && ctx->pipelined_batches) {
                if((status = enqueue_batch(oldest ^ 1)) != PTB_OK) {
                    return status;
                }
                ahead = true;
            }
            PTB_CUDA(cudaEventSynchronize(ctx->batch_done[oldest]));
            if(ctx->log_batches) {
                float since_start = 0.0F;
                cudaEventElapsedTime(&since_start, ctx->call_start, ctx->batch_done[oldest]);
                std::fprintf(stderr, "[ptb] batch of %d iterations done %.1f ms after the call started (%u paths entered it)\n", inflight[oldest].launched, since_start, n_cur);
            }

            // walk the oldest batch in launch order
            const Inflight &done = inflight[oldest];
            const uint32_t *host_counters = ctx->host_counters + oldest * kMaxIterationsPerSync * kCounterSlots;
            int slot_in = (done.launched % 2 == 0) ? done.cur_after : (done.cur_after ^ 1);
            uint32_t n_in = n_cur;
            for(int j = 0; j < done.launched; j++) {
                const uint32_t *hc = host_counters + j * kCounterSlots;
                const int slot_out = slot_in ^ 1;
                if(n_in > 0U) {
                    if(stats != nullptr) {
                        stats->closest_rays += n_in;
                        stats->shadow_rays += hc[kCountShadow];
                        stats->shadow_rays_skipped += hc[kCountSkippedShadows];
                        stats->closest_rays_retraced += hc[kCountRedo];
                        stats->path_vertices += hc[kCountVertices];
                        stats->bounce_iterations += 1;
                        stats->kernel_launches += certified ? 5 : 4;
                    }
                    if(ctx->log_iterations) {
                        std::fprintf(stderr, "[ptb] bounce iteration: %u paths, %u shadow rays, %u retraced, %u continue\n", n_in, hc[kCountShadow], hc[kCountRedo],
                                     hc[queue_slot[slot_out]]);
                    }
                }
                n_in = hc[queue_slot[slot_out]];
                slot_in = slot_out;
            }
            n_cur = n_in;
            if(progress != nullptr && progress->fn != nullptr && n_cur > 0U) {
                // started work items minus the paths still in flight = samples retired into the per-sample buffer
                const uint64_t started = std::min<uint64_t>(*ctx->host_cursor, src.total);
                const uint64_t retired = started > n_cur ? started - n_cur : 0;
                progress->fn(progress->user, progress->done_before + retired, progress->total);
            }
            if(ahead) {
                oldest ^= 1;
                ahead = false;
                if(n_cur == 0U) {
                    PTB_CUDA(cudaEventSynchronize(ctx->batch_done[oldest])); // the surplus batch: empty queues throughout
                }
            }
            else if(n_cur > 0U) {
                if((status = enqueue_batch(oldest)) != PTB_OK) {
                    return status;
                }
            }
        }
        PTB_CUDA(cudaGetLastError());
        collectTimers(ctx, stats);
        return PTB_OK;
    }

    // One pixel group of an adaptive (min < max) render: rounds of samples for the pixels still sampling (kernels.cuh,
    // "adaptive sampling in rounds").  ctx->pixel_list holds the group's rp.n_pixels pixels, ctx->samples has room for the
    // longest round.  A round's `samples` entries a pixel's loop does not consume (it ended inside the round) are the only
    // work traced beyond what the reference draws.
    int renderAdaptiveGroup(ptb_scene *scene, const RenderParams &params, const ResolveParams &rp, const ResolveConsts &rc, int first_round, int later_round,
                            uint64_t pool_limit, bool count_visits, ClosestMode closest, float4 *d_out, ptb_render_stats *stats, ProgressSink *sink) {
        ptb_context *ctx = scene->ctx;
        const uint32_t n_pixels = rp.n_pixels;
        int status;
        if((status = ctx->pixel_states.reserve(static_cast<size_t>(n_pixels) * sizeof(PixelStats))) != PTB_OK ||
           (status = ctx->active_lists.reserve(2 * static_cast<size_t>(n_pixels) * sizeof(uint32_t))) != PTB_OK ||
           (status = ctx->adaptive_counters.reserve(4 * sizeof(unsigned long long))) != PTB_OK) {
            return status;
        }
        PixelStats *states = ctx->pixel_states.as<PixelStats>();
        uint32_t *lists[2] = {ctx->active_lists.as<uint32_t>(), ctx->active_lists.as<uint32_t>() + n_pixels};
        unsigned long long *used = ctx->adaptive_counters.as<unsigned long long>();   // samples the pixels' loops consumed
        uint32_t *next_count = reinterpret_cast<uint32_t *>(used + 1);                // length of the list being written
        PTB_CUDA(cudaMemsetAsync(used, 0, 4 * sizeof(unsigned long long), ctx->stream));
        const unsigned pixel_grid = (n_pixels + kBlock - 1) / kBlock;
        {
            LaunchTimer timer(ctx, 1);
            adaptiveInitKernel<<<pixel_grid, kBlock, 0, ctx->stream>>>(rc, n_pixels, states, lists[0]);
        }
        PTB_CUDA(cudaGetLastError());

        uint32_t n_active = n_pixels;
        int cur = 0;
        int sample_base = 0;
        const uint64_t done_at_entry = sink != nullptr ? sink->done_before : 0;
        while(n_active > 0U && sample_base < rc.max_samples) {
            const int count = std::min(sample_base == 0 ? first_round : later_round, rc.max_samples - sample_base);
            const uint64_t total = static_cast<uint64_t>(n_active) * static_cast<uint64_t>(count);
            PathPool pool{};
            if((status = carvePool(ctx, poolCapacity(total, pool_limit), scene->shadow_stride, pool)) != PTB_OK) {
                return status;
            }
            PathSource src{};
            src.pixel_list = ctx->pixel_list.as<uint32_t>();
            src.active = lists[cur];
            src.n_pixels = n_active;
            src.sample_base = static_cast<uint32_t>(sample_base);
            src.total = total;
            if((status = runBounces(scene, pool, params, src, ctx->samples.as<float4>(), count_visits, closest, stats, sink)) != PTB_OK) {
                return status;
            }
            if(sink != nullptr) {
                sink->done_before += total;
            }
            PTB_CUDA(cudaMemsetAsync(next_count, 0, sizeof(uint32_t), ctx->stream));
            {
                LaunchTimer timer(ctx, 1);
                adaptiveAdvanceKernel<<<(n_active + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(rc, ctx->samples.as<float4>(), lists[cur], n_active, count, states, lists[cur ^ 1],
                                                                                                 next_count);
            }
            PTB_CUDA(cudaGetLastError());
            uint32_t n_next = 0;
            PTB_CUDA(cudaMemcpyAsync(&n_next, next_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            PTB_CUDA(cudaStreamSynchronize(ctx->stream));
            if(stats != nullptr) {
                stats->kernel_launches += 1;
                stats->adaptive_rounds += 1;
            }
            n_active = n_next;
            cur ^= 1;
            sample_base += count;
        }
        {
            LaunchTimer timer(ctx, 1);
            adaptiveFinishKernel<<<pixel_grid, kBlock, 0, ctx->stream>>>(rp, states, ctx->pixel_list.as<uint32_t>(), d_out, used);
        }
        PTB_CUDA(cudaGetLastError());
        unsigned long long used_host = 0;
        PTB_CUDA(cudaMemcpyAsync(&used_host, used, sizeof(used_host), cudaMemcpyDeviceToHost, ctx->stream));
        PTB_CUDA(cudaStreamSynchronize(ctx->stream));
        if(stats != nullptr) {
            stats->kernel_launches += 2;
            stats->samples_used += used_host;
        }
        if(sink != nullptr) {
            sink->done_before = done_at_entry; // the caller advances it by the group's nominal size
        }
        return PTB_OK;
    }

    int beginCall(ptb_context *ctx, ptb_render_stats *stats, uint32_t flags = 0U) {
        ctx->profile_all = (flags & PTB_FLAG_PROFILE_ALL) != 0U;
        int status = useDevice(ctx);
        if(status != PTB_OK) {
            return status;
        }
        // the arena of the last scene build (kept so that building several scenes in a row does not go through the
        // allocator each time) is handed back before a query or render sizes its workspace from the free memory
        ctx->build_arena.release();
        if(stats != nullptr) {
            std::memset(stats, 0, sizeof(*stats));
        }
        ctx->events_used = 0;
        PTB_CUDA(cudaEventRecord(ctx->call_start, ctx->stream));
        return PTB_OK;
    }

    int endCall(ptb_context *ctx, ptb_render_stats *stats) {
        PTB_CUDA(cudaEventRecord(ctx->call_stop, ctx->stream));
        PTB_CUDA(cudaStreamSynchronize(ctx->stream));
        PTB_CUDA(cudaGetLastError());
        if(stats != nullptr) {
            float ms = 0.0F;
            cudaEventElapsedTime(&ms, ctx->call_start, ctx->call_stop);
            stats->device_ms_total = ms;
        }
        collectTimers(ctx, stats);
        return PTB_OK;
    }

    // Builds the query hierarchy on the device (lbvh.cuh).  `boxes` holds 6 floats per slot in the reference tree's leaf
    // order.  On success `records` holds n - 1 inner records, the root is record 0 and `height` the number of inner
    // levels (what the traversal stack must hold).
    int buildQueryBvhOnDevice(ptb_context *ctx, const float *boxes_on_device, uint32_t n, const float root_lo[3], const float root_hi[3], Buffer &records,
                              uint32_t &height, double &device_ms) {
        height = 0;
        device_ms = 0.0;
        if(n < 2) {
            return PTB_OK;
        }
        Buffer keys_a, keys_b, slots_a, slots_b, leaf_parent, node_parent, children, arrivals, node_box, node_height, node_count, sort_temp, d_height;
        auto release_all = [&]() {
            for(Buffer *b : {&keys_a, &keys_b, &slots_a, &slots_b, &leaf_parent, &node_parent, &children, &arrivals, &node_box, &node_height, &node_count, &sort_temp,
                             &d_height}) {
                b->release();
            }
        };
        const size_t inner = static_cast<size_t>(n) - 1;
        int status = PTB_OK;
        auto reserve = [&](Buffer &b, size_t bytes) {
            if(status == PTB_OK) {
                status = b.reserve(std::max<size_t>(bytes, 16));
            }
        };
        reserve(keys_a, n * sizeof(uint64_t));
        reserve(keys_b, n * sizeof(uint64_t));
        reserve(slots_a, n * sizeof(uint32_t));
        reserve(slots_b, n * sizeof(uint32_t));
        reserve(leaf_parent, n * sizeof(int32_t));
        reserve(node_parent, inner * sizeof(int32_t));
        reserve(children, inner * sizeof(int2));
        reserve(arrivals, inner * sizeof(uint32_t));
        reserve(node_box, inner * 6 * sizeof(float));
        reserve(node_height, inner * sizeof(uint32_t));
        reserve(node_count, inner * sizeof(uint32_t));
        reserve(d_height, sizeof(uint32_t));
        reserve(records, inner * sizeof(NodeRecord));
        size_t temp_bytes = 0;
        if(status == PTB_OK &&
           cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), slots_a.as<uint32_t>(), slots_b.as<uint32_t>(), static_cast<int>(n), 0,
                                           63, ctx->stream) != cudaSuccess) {
            cudaGetLastError();
            status = fail(PTB_ERR_CUDA, "buildQueryBvhOnDevice: radix sort sizing failed");
        }
        reserve(sort_temp, temp_bytes);
        if(status != PTB_OK) {
            release_all();
            return status;
        }

        cudaEvent_t start = nullptr;
        cudaEvent_t stop = nullptr;
        cudaEventCreate(&start);
        cudaEventCreate(&stop);
        auto finish = [&](int st) {
            cudaEventDestroy(start);
            cudaEventDestroy(stop);
            release_all();
            return st;
        };
        cudaEventRecord(start, ctx->stream);

        LbvhWorkspace w{};
        w.boxes = boxes_on_device;
        w.keys = keys_b.as<uint64_t>();
        w.slots = slots_b.as<uint32_t>();
        w.leaf_parent = leaf_parent.as<int32_t>();
        w.node_parent = node_parent.as<int32_t>();
        w.children = children.as<int2>();
        w.arrivals = arrivals.as<uint32_t>();
        w.node_box = node_box.as<float>();
        w.node_height = node_height.as<uint32_t>();
        w.node_count = node_count.as<uint32_t>();
        w.records = records.as<float4>();
        w.n = n;
        for(int c = 0; c < 3; c++) {
            w.root_lo[c] = root_lo[c];
            w.root_hi[c] = root_hi[c];
        }
        const unsigned leaf_grid = (n + 255U) / 256U;
        mortonKernel<<<leaf_grid, 256, 0, ctx->stream>>>(w, keys_a.as<uint64_t>(), slots_a.as<uint32_t>());
        cub::DeviceRadixSort::SortPairs(sort_temp.ptr, temp_bytes, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), slots_a.as<uint32_t>(), slots_b.as<uint32_t>(), static_cast<int>(n), 0, 63,
                                        ctx->stream);
        cudaMemsetAsync(arrivals.ptr, 0, inner * sizeof(uint32_t), ctx->stream);
        cudaMemsetAsync(d_height.ptr, 0, sizeof(uint32_t), ctx->stream);
        hierarchyKernel<<<(n - 1U + 255U) / 256U, 256, 0, ctx->stream>>>(w);
        fitKernel<<<leaf_grid, 256, 0, ctx->stream>>>(w, d_height.as<uint32_t>());
        cudaEventRecord(stop, ctx->stream);
        uint32_t h = 0;
        if(cudaMemcpyAsync(&h, d_height.ptr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
           cudaGetLastError() != cudaSuccess) {
            cudaGetLastError();
            return finish(fail(PTB_ERR_CUDA, "buildQueryBvhOnDevice: build kernels failed"));
        }
        float ms = 0.0F;
        cudaEventElapsedTime(&ms, start, stop);
        device_ms = ms;
        height = h;
        return finish(PTB_OK);
    }

    struct SubShard {
        int index;
        int count;
        bool clear;
    };

    // a piece of a larger device allocation
    struct Carved {
        void *ptr = nullptr;
        size_t bytes = 0;
        template<typename T>
        T *as() const {
            return static_cast<T *>(ptr);
        }
    };

    // ---- scene setup on the device (gpu_build.cuh)
    //
    // Builds one hierarchy over `n` boxes (6 floats per id, on the device) level by level.
    //   reference_split: impl::constructBVH's topology (ids = the caller's primitive numbers); `order_out` receives the
    //                    leaf order (slot -> primitive), leaf refs of the records are ~slot.
    //   otherwise:       full-sweep SAH query tree (ids = leaf slots); leaf refs are ~id.
    // `depth` = deepest leaf with the root at 1; 0 when the tree came out deeper than `max_depth` (caller falls back).
    int buildTreeOnDevice(ptb_context *ctx, const float *d_boxes, uint32_t n, bool reference_split, uint32_t max_depth, Buffer &records, Buffer *order_out, uint32_t &depth,
                          float root_box[6], double &device_ms, size_t *arena_bytes_only = nullptr) {
        depth = 0;
        device_ms = 0.0;
        if(n < 2U) {
            return fail(PTB_ERR_INVALID_ARGUMENT, "buildTreeOnDevice: fewer than two primitives");
        }
        const size_t inner = static_cast<size_t>(n) - 1;
        const int n_lists = reference_split ? 4 : 3;
        // all temporaries of a build are carved from one arena (ctx->build_arena): thirty cudaMalloc / cudaFree pairs cost
        // ten times the build itself
        Carved lists_a, lists_b, keys, keys_sorted, ids, node_pos_a, node_pos_b, seg_begin, seg_count, children, node_parent, leaf_parent, side, flags, ranks, any_active, cut,
            group_keys, best_cost, best_key, seq_a, seq_b, costs, temp, arrivals, node_box, node_height, node_count, d_height;
        std::vector<Carved *> carved;
        auto release_all = [&]() {};
        int status = PTB_OK;
        auto carve = [&](Carved &c, size_t bytes) {
            c.bytes = (std::max<size_t>(bytes, 16) + 255) & ~static_cast<size_t>(255);
            carved.push_back(&c);
        };
        auto reserve = [&](Buffer &b, size_t bytes) {
            if(status == PTB_OK) {
                status = b.reserve(std::max<size_t>(bytes, 16));
            }
        };
        carve(lists_a, static_cast<size_t>(n_lists) * n * sizeof(uint32_t));
        carve(lists_b, static_cast<size_t>(n_lists) * n * sizeof(uint32_t));
        carve(keys, 3 * static_cast<size_t>(n) * sizeof(float));
        carve(keys_sorted, static_cast<size_t>(n) * sizeof(float));
        carve(ids, static_cast<size_t>(n) * sizeof(uint32_t));
        carve(node_pos_a, static_cast<size_t>(n) * sizeof(int32_t));
        carve(node_pos_b, static_cast<size_t>(n) * sizeof(int32_t));
        carve(seg_begin, inner * sizeof(uint32_t));
        carve(seg_count, inner * sizeof(uint32_t));
        carve(children, inner * sizeof(int2));
        carve(node_parent, inner * sizeof(int32_t));
        carve(leaf_parent, static_cast<size_t>(n) * sizeof(int32_t));
        carve(side, static_cast<size_t>(n) * sizeof(uint32_t));
        carve(flags, 3 * (static_cast<size_t>(n) + 1) * sizeof(uint32_t));
        carve(ranks, 3 * (static_cast<size_t>(n) + 1) * sizeof(uint32_t));
        carve(any_active, sizeof(uint32_t));
        if(reference_split) {
            carve(cut, 3 * inner * sizeof(float));
            carve(group_keys, 36 * inner * sizeof(uint32_t));
        }
        else {
            carve(best_cost, inner * sizeof(uint32_t));
            carve(best_key, inner * sizeof(unsigned long long));
            carve(seq_a, 6 * static_cast<size_t>(n) * sizeof(SweepBox));
            carve(seq_b, 6 * static_cast<size_t>(n) * sizeof(SweepBox));
            carve(costs, 3 * static_cast<size_t>(n) * sizeof(float));
        }
        carve(arrivals, inner * sizeof(uint32_t));
        carve(node_box, inner * 6 * sizeof(float));
        carve(node_height, inner * sizeof(uint32_t));
        carve(node_count, inner * sizeof(uint32_t));
        carve(d_height, sizeof(uint32_t));
        if(arena_bytes_only == nullptr) {
            reserve(records, inner * sizeof(NodeRecord));
            if(order_out != nullptr) {
                reserve(*order_out, static_cast<size_t>(n) * sizeof(uint32_t));
            }
        }

        size_t sort_bytes = 0;
        size_t sum_bytes = 0;
        size_t sweep_bytes = 0;
        const int scan_items = static_cast<int>(3 * (static_cast<size_t>(n) + 1));
        if(status == PTB_OK) {
            cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys.as<float>(), keys_sorted.as<float>(), ids.as<uint32_t>(), lists_a.as<uint32_t>(),
                                                             static_cast<int>(n), 0, 32, ctx->stream);
            cudaError_t e2 = cub::DeviceScan::ExclusiveSum(nullptr, sum_bytes, flags.as<uint32_t>(), ranks.as<uint32_t>(), scan_items, ctx->stream);
            cudaError_t e3 = cudaSuccess;
            if(!reference_split) {
                e3 = cub::DeviceScan::InclusiveScan(nullptr, sweep_bytes, seq_a.as<SweepBox>(), seq_b.as<SweepBox>(), SweepMerge{}, static_cast<int>(6 * static_cast<size_t>(n)),
                                                    ctx->stream);
            }
            if(e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
                cudaGetLastError();
                status = fail(PTB_ERR_CUDA, "buildTreeOnDevice: cub workspace sizing failed");
            }
        }
        size_t temp_bytes = std::max(sort_bytes, std::max(sum_bytes, sweep_bytes));
        carve(temp, temp_bytes);
        size_t arena_bytes = 0;
        for(const Carved *c : carved) {
            arena_bytes += c->bytes;
        }
        if(arena_bytes_only != nullptr) {
            *arena_bytes_only = arena_bytes; // dry run: the caller sizes one arena for several builds
            return status;
        }
        reserve(ctx->build_arena, arena_bytes);
        if(status == PTB_OK) {
            char *at = ctx->build_arena.as<char>();
            for(Carved *c : carved) {
                c->ptr = at;
                at += c->bytes;
            }
        }
        if(status != PTB_OK) {
            release_all();
            return status;
        }
        if(static_cast<size_t>(n) * 6 > static_cast<size_t>(std::numeric_limits<int>::max()) - 8) {
            release_all();
            return fail(PTB_ERR_UNSUPPORTED, "buildTreeOnDevice: too many primitives for the device builder");
        }

        cudaEvent_t start = nullptr;
        cudaEvent_t stop = nullptr;
        cudaEventCreate(&start);
        cudaEventCreate(&stop);
        auto finish = [&](int st) {
            cudaEventDestroy(start);
            cudaEventDestroy(stop);
            release_all();
            return st;
        };
        cudaEventRecord(start, ctx->stream);

        BuildState st{};
        st.n = n;
        st.boxes = d_boxes;
        for(int a = 0; a < n_lists; a++) {
            st.list[a] = lists_a.as<uint32_t>() + static_cast<size_t>(a) * n;
            st.list_next[a] = lists_b.as<uint32_t>() + static_cast<size_t>(a) * n;
        }
        st.node_of_pos = node_pos_a.as<int32_t>();
        st.node_of_pos_next = node_pos_b.as<int32_t>();
        st.seg_begin = seg_begin.as<uint32_t>();
        st.seg_count = seg_count.as<uint32_t>();
        st.children = children.as<int2>();
        st.node_parent = node_parent.as<int32_t>();
        st.leaf_parent = leaf_parent.as<int32_t>();
        st.side = side.as<uint32_t>();
        st.flags = flags.as<uint32_t>();
        st.ranks = ranks.as<uint32_t>();
        st.any_active = any_active.as<uint32_t>();
        st.cut = cut.as<float>();
        st.group_keys = group_keys.as<uint32_t>();
        st.best_cost = best_cost.as<uint32_t>();
        st.best_key = best_key.as<unsigned long long>();

        const unsigned grid = (n + kBuildBlock - 1U) / kBuildBlock;
        const unsigned grid_plus = (n + 1U + kBuildBlock - 1U) / kBuildBlock; // kernels that also write the sentinel at position n
        const dim3 grid3(grid, 3);
        const dim3 grid3_plus(grid_plus, 3);
        axisKeysKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(d_boxes, n, reference_split ? 0 : 1, keys.as<float>(), ids.as<uint32_t>());
        for(int a = 0; a < 3; a++) {
            cub::DeviceRadixSort::SortPairs(temp.ptr, sort_bytes, keys.as<float>() + static_cast<size_t>(a) * n, keys_sorted.as<float>(), ids.as<uint32_t>(), st.list[a],
                                            static_cast<int>(n), 0, 32, ctx->stream);
        }
        if(reference_split) {
            iotaKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(st.list[3], n);
        }
        buildInitKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(st);

        uint32_t levels = 0;
        for(;;) {
            if(levels >= max_depth) {
                cudaStreamSynchronize(ctx->stream);
                cudaGetLastError();
                return finish(PTB_OK); // depth stays 0: deeper than the traversal stack allows
            }
            cudaMemsetAsync(any_active.ptr, 0, sizeof(uint32_t), ctx->stream);
            if(reference_split) {
                refCutKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(st);
                refGroupKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(st);
                refChooseKernel<<<grid_plus, kBuildBlock, 0, ctx->stream>>>(st);
                cub::DeviceScan::ExclusiveSum(temp.ptr, sum_bytes, st.flags, st.ranks, static_cast<int>(n + 1U), ctx->stream);
                refPartitionKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(st);
            }
            else {
                sweepFillKernel<<<grid3, kBuildBlock, 0, ctx->stream>>>(st, seq_a.as<SweepBox>());
                cub::DeviceScan::InclusiveScan(temp.ptr, sweep_bytes, seq_a.as<SweepBox>(), seq_b.as<SweepBox>(), SweepMerge{}, static_cast<int>(6 * static_cast<size_t>(n)), ctx->stream);
                sweepCostKernel<<<grid3, kBuildBlock, 0, ctx->stream>>>(st, seq_b.as<SweepBox>(), costs.as<float>());
                sweepPickKernel<<<grid3, kBuildBlock, 0, ctx->stream>>>(st, costs.as<float>());
                sweepSideKernel<<<grid3, kBuildBlock, 0, ctx->stream>>>(st);
            }
            sideFlagsKernel<<<grid3_plus, kBuildBlock, 0, ctx->stream>>>(st);
            cub::DeviceScan::ExclusiveSum(temp.ptr, sum_bytes, st.flags, st.ranks, scan_items, ctx->stream);
            listPartitionKernel<<<grid3, kBuildBlock, 0, ctx->stream>>>(st, reference_split ? 0 : 1);
            uint32_t more = 0;
            if(cudaMemcpyAsync(&more, any_active.ptr, sizeof(more), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
               cudaGetLastError() != cudaSuccess) {
                cudaGetLastError();
                return finish(fail(PTB_ERR_CUDA, "buildTreeOnDevice: a level of the build failed"));
            }
            for(int a = 0; a < n_lists; a++) {
                std::swap(st.list[a], st.list_next[a]);
            }
            std::swap(st.node_of_pos, st.node_of_pos_next);
            levels++;
            if(more == 0U) {
                break;
            }
        }

        LbvhWorkspace w{};
        w.boxes = d_boxes;
        w.slots = reference_split ? nullptr : st.list[0];
        w.box_ids = reference_split ? st.list[3] : nullptr;
        w.leaf_parent = st.leaf_parent;
        w.node_parent = st.node_parent;
        w.children = st.children;
        w.arrivals = arrivals.as<uint32_t>();
        w.node_box = node_box.as<float>();
        w.node_height = node_height.as<uint32_t>();
        w.node_count = node_count.as<uint32_t>();
        w.records = records.as<float4>();
        w.n = n;
        cudaMemsetAsync(arrivals.ptr, 0, inner * sizeof(uint32_t), ctx->stream);
        cudaMemsetAsync(d_height.ptr, 0, sizeof(uint32_t), ctx->stream);
        fitKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(w, d_height.as<uint32_t>());
        if(order_out != nullptr) {
            cudaMemcpyAsync(order_out->ptr, st.list[3], static_cast<size_t>(n) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
        }
        cudaEventRecord(stop, ctx->stream);
        uint32_t height = 0;
        if(cudaMemcpyAsync(&height, d_height.ptr, sizeof(height), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
           cudaMemcpyAsync(root_box, node_box.ptr, 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
           cudaGetLastError() != cudaSuccess) {
            cudaGetLastError();
            return finish(fail(PTB_ERR_CUDA, "buildTreeOnDevice: box fit failed"));
        }
        float ms = 0.0F;
        cudaEventElapsedTime(&ms, start, stop);
        device_ms = ms;
        depth = height + 1U;
        return finish(PTB_OK);
    }

    constexpr uint32_t kSortThreshold = 1U << 16; // smaller batches are traced in the caller's order
    constexpr uint64_t kBigSceneBytes = 512ULL << 20; // batch queries on scenes beyond this keep the bottom of the traversal stack in shared memory

    // Fills ctx->sort_ids (second half) with the numbers of the chunk's n rays in sort-key order; returns the device pointer
    // through `order`.  stride_floats: 6 (closest-hit rays) or 7 (rays with a limit).
    int sortRays(ptb_scene *scene, const float *d_rays, uint32_t stride_floats, uint32_t n, const uint32_t **order) {
        ptb_context *ctx = scene->ctx;
        *order = nullptr;
        size_t temp_bytes = 0;
        if(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, static_cast<const uint32_t *>(nullptr), static_cast<uint32_t *>(nullptr), static_cast<const uint32_t *>(nullptr),
                                           static_cast<uint32_t *>(nullptr), static_cast<int>(n), 0, 30, ctx->stream) != cudaSuccess) {
            cudaGetLastError();
            return fail(PTB_ERR_CUDA, "sortRays: radix sort sizing failed");
        }
        int status;
        if((status = ctx->sort_keys.reserve(2 * static_cast<size_t>(n) * sizeof(uint32_t))) != PTB_OK || (status = ctx->sort_ids.reserve(2 * static_cast<size_t>(n) * sizeof(uint32_t))) != PTB_OK ||
           (status = ctx->sort_temp.reserve(std::max<size_t>(temp_bytes, 16))) != PTB_OK) {
            return status;
        }
        uint32_t *keys = ctx->sort_keys.as<uint32_t>();
        uint32_t *ids = ctx->sort_ids.as<uint32_t>();
        rayKeyKernel<<<(n + 255U) / 256U, 256, 0, ctx->stream>>>(scene->dev, d_rays, stride_floats, n, static_cast<uint32_t>(ctx->sort_dir_bits), keys, ids);
        PTB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->sort_temp.ptr, temp_bytes, keys, keys + n, ids, ids + n, static_cast<int>(n), 0, 30, ctx->stream));
        *order = ids + n;
        return PTB_OK;
    }

    int finishStats(ptb_context *ctx, bool count_visits, ptb_render_stats *stats) {
        if(stats == nullptr) {
            return PTB_OK;
        }
        if(count_visits) {
            VisitCounters v[2] = {};
            PTB_CUDA(cudaMemcpy(v, ctx->visits.ptr, sizeof(v), cudaMemcpyDeviceToHost));
            stats->inner_visits = v[0].inner + v[1].inner;
            stats->leaf_visits = v[0].leaf + v[1].leaf;
            stats->shadow_inner_visits = v[1].inner;
            stats->shadow_leaf_visits = v[1].leaf;
            stats->certified_suspect_hits = v[0].suspect + v[1].suspect;
        }
        return PTB_OK;
    }

}

extern "C" {

int ptb_abi_version(void) {
    return PTB_ABI_VERSION;
}

const char *ptb_last_error(void) {
    return g_last_error.c_str();
}

int ptb_camera_init(ptb_camera *out, const float origin[3], const float look_at[3], const float up[3], float focal_length, float height, float aspect_ratio,
                    float aperture_width, float aperture_height, uint32_t aperture_kind, float hexagon_horizontal_ratio, float focal_plane_dist) {
    if(out == nullptr || origin == nullptr || look_at == nullptr || up == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_camera_init: null argument");
    }
    ptb::cameraInit(out, origin, look_at, up, focal_length, height, aspect_ratio, aperture_width, aperture_height, aperture_kind, hexagon_horizontal_ratio,
                    focal_plane_dist);
    return PTB_OK;
}

int ptb_context_create(int device, ptb_context **out) {
    if(out == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_context_create: out is null");
    }
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if(err != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(PTB_ERR_NO_DEVICE, std::string("no CUDA device available (") + (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                                         "); this library has no CPU fallback");
    }
    if(device < 0) {
        device = static_cast<int>(envLong("PTB_DEVICE", envLong("LOCAL_RANK", 0)));
    }
    if(device >= count) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_context_create: device index out of range");
    }
    PTB_CUDA(cudaSetDevice(device));

    auto *ctx = new(std::nothrow) ptb_context();
    if(ctx == nullptr) {
        return fail(PTB_ERR_OUT_OF_MEMORY, "ptb_context_create: host allocation failed");
    }
    ctx->device = device;
    for(auto &pair : ctx->events) {
        pair[0] = pair[1] = nullptr;
    }
    // every failure below frees what was created so far (ptb_context_destroy copes with a partly built context)
    struct Guard {
        ptb_context *ctx;
        ~Guard() {
            if(ctx != nullptr) {
                const std::string keep = g_last_error;
                ptb_context_destroy(ctx);
                g_last_error = keep;
            }
        }
    } guard{ctx};
    cudaDeviceProp prop{};
    PTB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if(prop.major < 10) {
        return fail(PTB_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100-class; the kernels are built for sm_100a only");
    }
    PTB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    PTB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ctx->host_counters), 2 * kMaxIterationsPerSync * kCounterSlots * sizeof(uint32_t)));
    PTB_CUDA(cudaEventCreate(&ctx->batch_done[0]));
    PTB_CUDA(cudaEventCreate(&ctx->batch_done[1]));
    ctx->log_batches = envLong("PTB_LOG_BATCHES", 0) != 0;
    ctx->pipelined_batches = envLong("PTB_PIPELINED_BATCHES", 1) != 0;
    PTB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ctx->host_cursor), sizeof(unsigned long long)));
    for(auto &pair : ctx->events) {
        PTB_CUDA(cudaEventCreate(&pair[0]));
        PTB_CUDA(cudaEventCreate(&pair[1]));
    }
    PTB_CUDA(cudaEventCreate(&ctx->call_start));
    PTB_CUDA(cudaEventCreate(&ctx->call_stop));
    guard.ctx = nullptr; // fully built: ownership passes to the caller
    ctx->events_ready = envLong("PTB_PROFILE", 1) != 0;
    // tuned on the bench scene with the 128 Mi-path pool (with the earlier 4 Mi pool the drain phases dominated and smaller votes won)
    ctx->vote.refill = static_cast<int>(std::min(32L, std::max(1L, envLong("PTB_REFILL_VOTE", 12))));
    ctx->vote.leaf = static_cast<int>(std::min(32L, std::max(1L, envLong("PTB_LEAF_VOTE", 12))));
    ctx->vote.leaf_burst = static_cast<int>(std::min(8L, std::max(1L, envLong("PTB_LEAF_BURST", 2))));
    ctx->vote.inner_burst = static_cast<int>(std::min(8L, std::max(1L, envLong("PTB_INNER_BURST", 4))));
    ctx->vote_shadow = ctx->vote;
    ctx->vote_shadow.refill = static_cast<int>(std::min(32L, std::max(1L, envLong("PTB_SHADOW_REFILL_VOTE", envLong("PTB_REFILL_VOTE", 16)))));
    ctx->vote_shadow.leaf = static_cast<int>(std::min(32L, std::max(1L, envLong("PTB_SHADOW_LEAF_VOTE", envLong("PTB_LEAF_VOTE", 12)))));
    ctx->trace_blocks_per_sm = static_cast<int>(std::max(1L, envLong("PTB_TRACE_BLOCKS_PER_SM", 16)));
    ctx->shade_blocks_per_sm = static_cast<int>(std::max(1L, envLong("PTB_SHADE_BLOCKS_PER_SM", PTB_SHADE_MIN_BLOCKS)));
    ctx->log_iterations = envLong("PTB_LOG_ITERATIONS", 0) != 0;
    ctx->production_math = envLong("PTB_PRODUCTION_MATH", 1) != 0;
    ctx->sort_rays = envLong("PTB_SORT_RAYS", 1) != 0;
    ctx->sort_dir_bits = static_cast<int>(std::min(4L, std::max(1L, envLong("PTB_SORT_DIR_BITS", 1))));
    ctx->adaptive_rounds = envLong("PTB_ADAPTIVE_ROUNDS", 1) != 0;
    ctx->streams = static_cast<int>(std::min(4L, std::max(1L, envLong("PTB_STREAMS", 1))));
    PTB_CUDA(cudaEventCreate(&ctx->split_start));
    PTB_CUDA(cudaEventCreate(&ctx->split_stop));
    ctx->iterations_per_sync = static_cast<int>(std::min<long>(kMaxIterationsPerSync, std::max(1L, envLong("PTB_ITERATIONS_PER_SYNC", 8))));
    *out = ctx;
    return PTB_OK;
}

int ptb_context_destroy(ptb_context *ctx) {
    if(ctx == nullptr) {
        return PTB_OK;
    }
    for(ptb_context *sibling : ctx->siblings) {
        ptb_context_destroy(sibling);
    }
    ctx->siblings.clear();
    cudaSetDevice(ctx->device);
    if(ctx->stream != nullptr) {
        cudaStreamSynchronize(ctx->stream);
    }
    ctx->split_image.release();
    for(cudaEvent_t *e : {&ctx->split_start, &ctx->split_stop}) {
        if(*e != nullptr) {
            cudaEventDestroy(*e);
            *e = nullptr;
        }
    }
    for(Buffer *b : {&ctx->pool_mem, &ctx->queue_a, &ctx->queue_b, &ctx->shadow_queue, &ctx->redo_queue, &ctx->counters, &ctx->visits, &ctx->work_cursor, &ctx->samples, &ctx->pixel_list, &ctx->pixel_states, &ctx->active_lists, &ctx->adaptive_counters, &ctx->build_arena,
                     &ctx->io_a, &ctx->io_b, &ctx->io_c, &ctx->io_d, &ctx->multi_image, &ctx->multi_staging, &ctx->sort_keys, &ctx->sort_ids, &ctx->sort_temp}) {
        b->release();
    }
    for(cudaEvent_t &e : ctx->batch_done) {
        if(e != nullptr) {
            cudaEventDestroy(e);
            e = nullptr;
        }
    }
    if(ctx->host_counters != nullptr) {
        cudaFreeHost(ctx->host_counters);
    }
    if(ctx->host_cursor != nullptr) {
        cudaFreeHost(ctx->host_cursor);
    }
    for(auto &pair : ctx->events) {
        if(pair[0] != nullptr) {
            cudaEventDestroy(pair[0]);
        }
        if(pair[1] != nullptr) {
            cudaEventDestroy(pair[1]);
        }
    }
    if(ctx->call_start != nullptr) {
        cudaEventDestroy(ctx->call_start);
    }
    if(ctx->call_stop != nullptr) {
        cudaEventDestroy(ctx->call_stop);
    }
    if(ctx->stream != nullptr) {
        cudaStreamDestroy(ctx->stream);
    }
    cudaGetLastError();
    delete ctx;
    return PTB_OK;
}

int ptb_context_device(const ptb_context *ctx, int *device_out) {
    if(ctx == nullptr || device_out == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_context_device: null argument");
    }
    *device_out = ctx->device;
    return PTB_OK;
}

int ptb_context_synchronize(ptb_context *ctx) {
    if(ctx == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_context_synchronize: null context");
    }
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

// ---------------------------------------------------------------------------------------------------- scene

int ptb_scene_create(ptb_context *ctx, const ptb_scene_desc *desc, ptb_scene **out) {
    if(ctx == nullptr || desc == nullptr || out == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_create: null argument");
    }
    *out = nullptr;
    if(desc->n_prims > 0 && (desc->prims == nullptr || desc->materials == nullptr || desc->n_materials == 0)) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_create: primitives need a material table");
    }
    if(desc->n_lights > 0 && desc->lights == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_create: n_lights > 0 but lights is null");
    }
    if(desc->n_prims >= (1ULL << 31) - 1) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_scene_create: more than 2^31 - 2 primitives");
    }
    if(desc->bvh_mode != PTB_BVH_REFERENCE && desc->bvh_mode != PTB_BVH_REFERENCE_GPU_QUERY_TREE) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_scene_create: unknown bvh_mode");
    }
    for(uint64_t i = 0; i < desc->n_prims; i++) {
        const ptb_prim &prim = desc->prims[i];
        if(prim.kind > PTB_PRIM_NULL) {
            return fail(PTB_ERR_UNSUPPORTED, "ptb_scene_create: unknown primitive kind");
        }
        if(prim.material >= desc->n_materials) {
            return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_create: material index out of range");
        }
    }
    for(uint32_t i = 0; i < desc->n_materials; i++) {
        if(desc->materials[i].bsdf > PTB_BSDF_MIRROR) {
            return fail(PTB_ERR_UNSUPPORTED, "ptb_scene_create: unknown BSDF kind");
        }
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }

    const double t0 = nowSeconds();
    const int threads = static_cast<int>(envLong("PTB_BUILD_THREADS", 0));
    const uint64_t n = desc->n_prims;
    // Scene setup runs on the device (gpu_build.cuh): bounding boxes, the reference-topology tree, the leaf-order geometry
    // and shading records, and the query tree.  PTB_DEVICE_BUILD=0 (and scenes of fewer than two primitives) take the
    // host builders of bvh_build.cpp, which produce the same parity tree bit for bit.
    const bool device_build = n >= 2 && envLong("PTB_DEVICE_BUILD", 1) != 0;
    const bool want_query_tree = n > 1 && envLong("PTB_OCCLUSION_BVH", 1) != 0;
    // query tree: "sweep" full-sweep SAH on the device (default with the device build), "lbvh" Morton-order linear BVH on
    // the device, "host" binned SAH on the host (default with the host build)
    const char *query_env = std::getenv("PTB_QUERY_TREE");
    std::string query_kind = query_env != nullptr ? query_env : "";
    if(desc->bvh_mode == PTB_BVH_REFERENCE_GPU_QUERY_TREE || envLong("PTB_GPU_BVH", 0) != 0) {
        query_kind = "lbvh";
    }
    if(query_kind != "sweep" && query_kind != "lbvh" && query_kind != "host") {
        query_kind = device_build ? "sweep" : "host";
    }

    auto *scene = new(std::nothrow) ptb_scene();
    if(scene == nullptr) {
        return fail(PTB_ERR_OUT_OF_MEMORY, "ptb_scene_create: host allocation failed");
    }
    scene->ctx = ctx;
    auto abandon = [&](int st) {
        ptb_scene_destroy(scene);
        return st;
    };
    auto upload = [&](Buffer &buffer, const void *src, size_t bytes) -> int {
        int st = buffer.reserve(std::max<size_t>(bytes, 16));
        if(st != PTB_OK) {
            return st;
        }
        if(bytes > 0) {
            PTB_CUDA(cudaMemcpyAsync(buffer.ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
            PTB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        return PTB_OK;
    };

    const bool log_build = envLong("PTB_LOG_BUILD", 0) != 0;
    double t_mark = t0;
    auto mark = [&](const char *what) {
        if(log_build) {
            const double now = nowSeconds();
            std::fprintf(stderr, "[ptb] scene setup: %-34s %8.2f ms\n", what, (now - t_mark) * 1e3);
            t_mark = now;
        }
    };
    FlatBvh bvh; // host path: the whole tree; device path: slot_to_prim, root, depth (the records stay on the device)
    size_t n_inner_nodes = 0;
    Buffer boxes_by_slot; // 6 floats per leaf slot, on the device: input of the device-side query-tree builders
    bool query_tree_on_device = false;
    int32_t query_root_ref = -1;
    double query_tree_device_ms = 0.0;
    double reference_tree_device_ms = 0.0;
    double upload_seconds = 0.0;

    if(device_build) {
        Buffer d_prims, boxes_by_prim;
        auto cleanup = [&](int st) {
            d_prims.release();
            boxes_by_prim.release();
            boxes_by_slot.release();
            return st != PTB_OK ? abandon(st) : st;
        };
        const double tu = nowSeconds();
        if((status = upload(d_prims, desc->prims, n * sizeof(ptb_prim))) != PTB_OK) {
            return cleanup(status);
        }
        upload_seconds += nowSeconds() - tu;
        mark("upload of the primitives");
        const uint32_t n32 = static_cast<uint32_t>(n);
        const unsigned grid = (n32 + kBuildBlock - 1U) / kBuildBlock;
        if((status = boxes_by_prim.reserve(6 * n * sizeof(float))) != PTB_OK || (status = boxes_by_slot.reserve(6 * n * sizeof(float))) != PTB_OK ||
           (status = scene->geom.reserve(kGeomLanes * n * sizeof(float4))) != PTB_OK || (status = scene->shade.reserve(3 * n * sizeof(float4))) != PTB_OK) {
            return cleanup(status);
        }
        {
            // one arena for both tree builds (growing it between them would free and re-allocate a few hundred MB)
            size_t for_reference = 0;
            size_t for_query = 0;
            uint32_t unused_depth = 0;
            double unused_ms = 0.0;
            float unused_box[6];
            Buffer unused;
            buildTreeOnDevice(ctx, nullptr, n32, true, 0U, unused, nullptr, unused_depth, unused_box, unused_ms, &for_reference);
            if(want_query_tree && query_kind == "sweep") {
                buildTreeOnDevice(ctx, nullptr, n32, false, 0U, unused, nullptr, unused_depth, unused_box, unused_ms, &for_query);
            }
            if((status = ctx->build_arena.reserve(std::max(for_reference, for_query))) != PTB_OK) {
                return cleanup(status);
            }
        }
        primBoundsKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(d_prims.as<ptb_prim>(), n32, boxes_by_prim.as<float>());
        float root_box[6] = {};
        // depth limit: the traversal stack holds one deferred sibling per level
        status = buildTreeOnDevice(ctx, boxes_by_prim.as<float>(), n32, true, static_cast<uint32_t>(kStackCapacity), scene->nodes, &scene->slot_to_prim, bvh.depth, root_box,
                                   reference_tree_device_ms);
        if(status != PTB_OK) {
            return cleanup(status);
        }
        if(bvh.depth == 0U) {
            return cleanup(fail(PTB_ERR_UNSUPPORTED, "ptb_scene_create: BVH deeper than the traversal stack"));
        }
        mark("boxes + reference-topology tree");
        packSlotsKernel<<<grid, kBuildBlock, 0, ctx->stream>>>(d_prims.as<ptb_prim>(), scene->slot_to_prim.as<uint32_t>(), boxes_by_prim.as<float>(), n32, scene->geom.as<float4>(),
                                                              scene->shade.as<float4>(), boxes_by_slot.as<float>());
        bvh.slot_to_prim.resize(n);
        if(cudaMemcpyAsync(bvh.slot_to_prim.data(), scene->slot_to_prim.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
           cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            cudaGetLastError();
            return cleanup(fail(PTB_ERR_CUDA, "ptb_scene_create: packing the leaf records failed"));
        }
        bvh.root_ref = 0;
        for(int c = 0; c < 3; c++) {
            bvh.root_low[c] = root_box[c];
            bvh.root_high[c] = root_box[3 + c];
        }
        n_inner_nodes = n - 1;
        d_prims.release();
        boxes_by_prim.release();
        mark("leaf records + slot table");
    }
    else {
        bvh = buildReferenceBvh(desc->prims, desc->n_prims, threads);
        if(bvh.depth > static_cast<uint32_t>(kStackCapacity)) {
            return abandon(fail(PTB_ERR_UNSUPPORTED, "ptb_scene_create: BVH deeper than the traversal stack (" + std::to_string(bvh.depth) + ")"));
        }
        n_inner_nodes = bvh.nodes.size();

        std::vector<float4> geom(kGeomLanes * n, make_float4(0.0F, 0.0F, 0.0F, 0.0F));
        std::vector<float4> shade(3 * n);
        for(uint64_t slot = 0; slot < n; slot++) {
            const ptb_prim &prim = desc->prims[bvh.slot_to_prim[slot]];
            const float *p = prim.p;
            uint32_t flags = prim.kind & kKindMask;
            if(prim.kind == PTB_PRIM_TRIANGLE && prim.cull_backface != 0U) {
                flags |= kCullBit;
            }
            float4 *g = &geom[kGeomLanes * slot];
            float4 *s = &shade[3 * slot];
            if(prim.kind == PTB_PRIM_TRIANGLE) {
                // edges are differenced on the host exactly as Triangle::getIntersection does per call (object.cpp:149-150)
                g[0] = make_float4(p[0], p[1], p[2], 0.0F);
                g[1] = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0F);
                g[2] = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0F);
                s[0] = make_float4(p[9], p[10], p[11], 0.0F);
                s[1] = make_float4(p[12], p[13], p[14], 0.0F);
                s[2] = make_float4(p[15], p[16], p[17], 0.0F);
            }
            else if(prim.kind == PTB_PRIM_SPHERE) {
                g[0] = make_float4(p[0], p[1], p[2], 0.0F);
                g[1] = make_float4(p[3], p[3] * p[3], 0.0F, 0.0F);
                g[2] = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
                s[0] = s[1] = s[2] = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
            }
            else {
                g[0] = g[1] = g[2] = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
                s[0] = s[1] = s[2] = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
            }
            std::memcpy(&g[0].w, &flags, sizeof(flags));
            std::memcpy(&s[0].w, &prim.material, sizeof(uint32_t));
        }
        const double tu = nowSeconds();
        status = upload(scene->nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(NodeRecord));
        status = status != PTB_OK ? status : upload(scene->geom, geom.data(), geom.size() * sizeof(float4));
        status = status != PTB_OK ? status : upload(scene->shade, shade.data(), shade.size() * sizeof(float4));
        status = status != PTB_OK ? status : upload(scene->slot_to_prim, bvh.slot_to_prim.data(), bvh.slot_to_prim.size() * sizeof(uint32_t));
        if(status == PTB_OK && want_query_tree && query_kind != "host") {
            std::vector<float> boxes(6 * n);
            for(uint64_t slot = 0; slot < n; slot++) {
                primBounds(desc->prims[bvh.slot_to_prim[slot]], &boxes[6 * slot], &boxes[6 * slot + 3]);
            }
            status = upload(boxes_by_slot, boxes.data(), boxes.size() * sizeof(float));
        }
        upload_seconds += nowSeconds() - tu;
        if(status != PTB_OK) {
            boxes_by_slot.release();
            return abandon(status);
        }
    }

    // query hierarchy for any-hit and certified closest-hit queries over the same primitives, leaf refs in reference slots
    if(want_query_tree && query_kind != "host") {
        uint32_t levels = 0;
        if(query_kind == "lbvh") {
            status = buildQueryBvhOnDevice(ctx, boxes_by_slot.as<float>(), static_cast<uint32_t>(n), bvh.root_low, bvh.root_high, scene->occ_nodes, levels, query_tree_device_ms);
        }
        else {
            float unused_root[6];
            uint32_t depth = 0;
            // (PTB_QUERY_TREE_MAX_LEVELS below 64 exists to exercise the fallback: a sweep tree deeper than the stack needs
            // surface areas growing a hundredfold per primitive, which float coordinates cannot hold for 64 levels)
            const uint32_t max_levels = static_cast<uint32_t>(std::min<long>(kStackCapacity, std::max(1L, envLong("PTB_QUERY_TREE_MAX_LEVELS", kStackCapacity))));
            status = buildTreeOnDevice(ctx, boxes_by_slot.as<float>(), static_cast<uint32_t>(n), false, max_levels, scene->occ_nodes, nullptr, depth, unused_root,
                                       query_tree_device_ms);
            levels = depth > 0U ? depth - 1U : 0U;
        }
        if(status != PTB_OK) {
            boxes_by_slot.release();
            return abandon(status);
        }
        // many coincident centres can chain into a tree deeper than the traversal stack: then the host builder takes over
        query_tree_on_device = levels >= 1U && levels <= static_cast<uint32_t>(kStackCapacity);
        if(query_tree_on_device) {
            query_root_ref = 0;
        }
        else {
            scene->occ_nodes.release();
        }
    }
    boxes_by_slot.release();
    mark("query tree (device)");
    if(want_query_tree && !query_tree_on_device) {
        std::vector<uint32_t> prim_to_slot(n);
        for(uint64_t slot = 0; slot < n; slot++) {
            prim_to_slot[bvh.slot_to_prim[slot]] = static_cast<uint32_t>(slot);
        }
        FlatBvh occlusion = buildOcclusionBvh(desc->prims, n, prim_to_slot.data(), threads);
        if(occlusion.depth <= static_cast<uint32_t>(kStackCapacity) && !occlusion.nodes.empty()) { // else (pathological input) keep using the reference tree
            const double tu = nowSeconds();
            if((status = upload(scene->occ_nodes, occlusion.nodes.data(), occlusion.nodes.size() * sizeof(NodeRecord))) != PTB_OK) {
                return abandon(status);
            }
            upload_seconds += nowSeconds() - tu;
            query_root_ref = occlusion.root_ref;
        }
    }
    const bool have_query_tree = scene->occ_nodes.ptr != nullptr && (query_tree_on_device || query_root_ref != -1);
    mark("query tree (host)");

    std::vector<float4> mats(3 * static_cast<size_t>(desc->n_materials));
    for(uint32_t i = 0; i < desc->n_materials; i++) {
        const ptb_material &m = desc->materials[i];
        mats[3 * i] = make_float4(m.diffuse[0], m.diffuse[1], m.diffuse[2], m.diffuse[3]);
        mats[3 * i + 1] = make_float4(m.emission[0], m.emission[1], m.emission[2], m.emission[3]);
        float4 misc = make_float4(m.refractive_index, 0.0F, 0.0F, 0.0F);
        std::memcpy(&misc.y, &m.bsdf, sizeof(uint32_t));
        const uint32_t one_way = m.one_way != 0U ? 1U : 0U;
        std::memcpy(&misc.z, &one_way, sizeof(uint32_t));
        mats[3 * i + 2] = misc;
    }

    std::vector<float4> lights(2 * static_cast<size_t>(desc->n_lights));
    for(uint32_t i = 0; i < desc->n_lights; i++) {
        const ptb_point_light &l = desc->lights[i];
        lights[2 * i] = make_float4(l.pos[0], l.pos[1], l.pos[2], 0.0F);
        lights[2 * i + 1] = make_float4(l.rgba[0], l.rgba[1], l.rgba[2], l.rgba[3]);
    }

    EmissiveTable emissive = buildEmissiveTable(desc->prims, desc->materials, bvh.slot_to_prim.data(), n);
    std::vector<float4> emis(3 * emissive.slots.size());
    for(size_t i = 0; i < emissive.slots.size(); i++) {
        const uint32_t slot = emissive.slots[i];
        const ptb_prim &prim = desc->prims[bvh.slot_to_prim[slot]];
        const float *p = prim.p;
        float4 e0;
        float4 e1;
        float4 e2;
        if(prim.kind == PTB_PRIM_TRIANGLE) {
            e0 = make_float4(p[0], p[1], p[2], 0.0F);
            e1 = make_float4(p[3], p[4], p[5], 0.0F);
            e2 = make_float4(p[6], p[7], p[8], 0.0F);
        }
        else {
            e0 = make_float4(p[0], p[1], p[2], 0.0F);
            e1 = make_float4(p[3], p[3] * p[3], 0.0F, 0.0F);
            e2 = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
        }
        std::memcpy(&e0.w, &slot, sizeof(uint32_t));
        e1.w = primSampleDensity(prim);
        emis[3 * i] = e0;
        emis[3 * i + 1] = e1;
        emis[3 * i + 2] = e2;
    }
    const double t1 = nowSeconds();
    mark("materials, lights, emissive table");

    status = upload(scene->mats, mats.data(), mats.size() * sizeof(float4));
    status = status != PTB_OK ? status : upload(scene->lights, lights.data(), lights.size() * sizeof(float4));
    status = status != PTB_OK ? status : upload(scene->emis, emis.data(), emis.size() * sizeof(float4));
    status = status != PTB_OK ? status : upload(scene->cdf, emissive.cdf.data(), emissive.cdf.size() * sizeof(float));
    if(status != PTB_OK) {
        return abandon(status);
    }
    const double t2 = nowSeconds();

    DeviceScene &d = scene->dev;
    d.nodes = scene->nodes.as<float4>();
    d.occ_nodes = have_query_tree ? scene->occ_nodes.as<float4>() : nullptr;
    d.occ_root_ref = query_root_ref;
    d.geom = scene->geom.as<float4>();
    d.shade = scene->shade.as<float4>();
    d.mats = scene->mats.as<float4>();
    d.lights = scene->lights.as<float4>();
    d.emis = scene->emis.as<float4>();
    d.cdf = scene->cdf.as<float>();
    d.slot_to_prim = scene->slot_to_prim.as<uint32_t>();
    d.n_prims = static_cast<uint32_t>(n);
    d.n_lights = desc->n_lights;
    d.n_emissive = static_cast<uint32_t>(emissive.slots.size());
    d.object_sample_count = emissive.object_sample_count;
    d.root_ref = bvh.root_ref;
    for(int c = 0; c < 3; c++) {
        d.root_lo[c] = bvh.root_low[c];
        d.root_hi[c] = bvh.root_high[c];
    }

    scene->shadow_stride = std::max<uint32_t>(1U, desc->n_lights + emissive.object_sample_count);
    mark("upload of the small tables");
    ptb_guard::buildCertGuard(desc->prims, desc->n_prims, &scene->guard);
    mark("guard table of the certified walk");

    ptb_scene_info &info = scene->info;
    info.n_prims = n;
    info.n_inner_nodes = n_inner_nodes;
    info.bvh_depth = bvh.depth;
    info.n_emissive = d.n_emissive;
    info.object_sample_count = d.object_sample_count;
    info.n_lights = desc->n_lights;
    info.device_bytes = scene->nodes.bytes + scene->occ_nodes.bytes + scene->geom.bytes + scene->shade.bytes + scene->mats.bytes + scene->lights.bytes + scene->emis.bytes +
                        scene->cdf.bytes + scene->slot_to_prim.bytes;
    info.build_seconds = (t1 - t0) - upload_seconds;
    info.query_tree_on_device = query_tree_on_device ? 1U : 0U;
    info.certifiable = scene->guard.certifiable;
    info.query_tree_device_ms = query_tree_device_ms;
    info.upload_seconds = (t2 - t1) + upload_seconds;
    info.built_on_device = device_build ? 1U : 0U;
    info.query_tree_kind = !have_query_tree ? 0U : (!query_tree_on_device ? 1U : (query_kind == "lbvh" ? 2U : 3U));
    info.reference_tree_device_ms = reference_tree_device_ms;
    for(int c = 0; c < 3; c++) {
        info.root_low[c] = bvh.root_low[c];
        info.root_high[c] = bvh.root_high[c];
    }

    *out = scene;
    return PTB_OK;
}

int ptb_scene_destroy(ptb_scene *scene) {
    if(scene == nullptr) {
        return PTB_OK;
    }
    for(ptb_scene *alias : scene->aliases) {
        delete alias; // owns nothing: its device arrays are this scene's
    }
    scene->aliases.clear();
    if(scene->ctx != nullptr) {
        cudaSetDevice(scene->ctx->device);
        cudaStreamSynchronize(scene->ctx->stream);
    }
    for(Buffer *b : {&scene->nodes, &scene->occ_nodes, &scene->geom, &scene->shade, &scene->mats, &scene->lights, &scene->emis, &scene->cdf, &scene->slot_to_prim}) {
        b->release();
    }
    delete scene;
    return PTB_OK;
}

int ptb_scene_read(const ptb_scene *scene, uint32_t array, void *out, uint64_t bytes) {
    if(scene == nullptr || (bytes > 0 && out == nullptr)) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_read: null argument");
    }
    const Buffer *source = nullptr;
    size_t valid = 0; // buffers may be larger than the array they hold
    const size_t n = scene->dev.n_prims;
    const size_t inner = n > 0 ? n - 1 : 0;
    switch(array) {
        case PTB_SCENE_NODES: source = &scene->nodes; valid = inner * sizeof(NodeRecord); break;
        case PTB_SCENE_QUERY_NODES: source = &scene->occ_nodes; valid = scene->dev.occ_nodes != nullptr ? inner * sizeof(NodeRecord) : 0; break;
        case PTB_SCENE_GEOM: source = &scene->geom; valid = kGeomLanes * n * sizeof(float4); break;
        case PTB_SCENE_SHADE: source = &scene->shade; valid = 3 * n * sizeof(float4); break;
        case PTB_SCENE_SLOT_TO_PRIM: source = &scene->slot_to_prim; valid = n * sizeof(uint32_t); break;
        default: return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_read: unknown array");
    }
    if(bytes > valid) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_read: more bytes requested than the array holds");
    }
    if(bytes == 0) {
        return PTB_OK;
    }
    ptb_context *ctx = scene->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(out, source->ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_scene_get_info(const ptb_scene *scene, ptb_scene_info *out) {
    if(scene == nullptr || out == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_get_info: null argument");
    }
    *out = scene->info;
    return PTB_OK;
}

// ---------------------------------------------------------------------------------------------------- intersect

int ptb_intersect(ptb_scene *scene, const float *rays, uint64_t n_rays, float *t_out, int32_t *prim_out, uint32_t flags, ptb_render_stats *stats) {
    if(scene == nullptr || (n_rays > 0 && (rays == nullptr || t_out == nullptr || prim_out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_intersect: null argument");
    }
    ptb_context *ctx = scene->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = beginCall(ctx, stats, flags);
    if(status != PTB_OK) {
        return status;
    }
    if(n_rays == 0) {
        return endCall(ctx, stats);
    }
    const bool device_io = (flags & PTB_FLAG_DEVICE_IO) != 0U;
    const bool count_visits = (flags & PTB_FLAG_COUNT_VISITS) != 0U;

    const float *d_rays = rays;
    float *d_t = t_out;
    int32_t *d_prim = prim_out;
    if(!device_io) {
        if((status = ctx->io_a.reserve(n_rays * 6 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n_rays * sizeof(float))) != PTB_OK ||
           (status = ctx->io_c.reserve(n_rays * sizeof(int32_t))) != PTB_OK) {
            return status;
        }
        PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, rays, n_rays * 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        d_rays = ctx->io_a.as<float>();
        d_t = ctx->io_b.as<float>();
        d_prim = ctx->io_c.as<int32_t>();
    }
    if((status = ctx->counters.reserve(kCounterSlots * sizeof(uint32_t))) != PTB_OK || (status = ctx->visits.reserve(2 * sizeof(VisitCounters))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, kCounterSlots * sizeof(uint32_t), ctx->stream));
    PTB_CUDA(cudaMemsetAsync(ctx->visits.ptr, 0, 2 * sizeof(VisitCounters), ctx->stream));

    const ClosestMode closest = closestMode(scene, flags);
    const bool certified = closest.certified;
    constexpr uint64_t kChunk = 1ULL << 26;
    if(certified && (status = ctx->redo_queue.reserve(std::min<uint64_t>(kChunk, n_rays) * sizeof(uint32_t))) != PTB_OK) {
        return status;
    }
    uint32_t *counters = ctx->counters.as<uint32_t>();
    uint32_t *redo = ctx->redo_queue.as<uint32_t>();
    VisitCounters *visits = ctx->visits.as<VisitCounters>();
    uint64_t retraced = 0;
    for(uint64_t first = 0; first < n_rays; first += kChunk) {
        const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(kChunk, n_rays - first));
        const int grid = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n) + kBlock - 1) / kBlock, static_cast<uint64_t>(gridFor(ctx, 16))));
        PTB_CUDA(cudaMemsetAsync(counters, 0, kCounterSlots * sizeof(uint32_t), ctx->stream));
        const float *chunk_rays = d_rays + 6 * first;
        const uint32_t *order = nullptr; // large batches are traced in sort-key order (rayKeyKernel); results go by ray number
        if(ctx->sort_rays && n >= kSortThreshold && (status = sortRays(scene, chunk_rays, 6, n, &order)) != PTB_OK) {
            return status;
        }
        LaunchTimer timer(ctx, 0);
        // scenes far beyond the 126 MB L2 keep the bottom of the traversal stack in shared memory (traverse.cuh)
        const bool big_scene = scene->info.device_bytes > kBigSceneBytes;
        auto launch_intersect = [&](auto mode, const uint32_t *index, const uint32_t *index_count, uint32_t *cursor, uint32_t *redo_queue, uint32_t *redo_count) {
            constexpr int kMode = decltype(mode)::value;
            if(count_visits) {
                intersectKernel<kMode, true, 0><<<grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote, chunk_rays, index, index_count, n, d_t + first, d_prim + first, cursor, redo_queue,
                                                                                  redo_count, visits, scene->guard, closest.guarded);
            }
            else if(big_scene) {
                intersectKernel<kMode, false, kQuerySmemClosest><<<grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote, chunk_rays, index, index_count, n, d_t + first, d_prim + first,
                                                                                                   cursor, redo_queue, redo_count, visits, scene->guard, closest.guarded);
            }
            else {
                intersectKernel<kMode, false, 0><<<grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote, chunk_rays, index, index_count, n, d_t + first, d_prim + first, cursor, redo_queue,
                                                                                   redo_count, visits, scene->guard, closest.guarded);
            }
        };
        if(certified) {
            launch_intersect(std::integral_constant<int, kTraceCertified>{}, order, nullptr, counters + kCountFetchClosest, redo, counters + kCountRedo);
            launch_intersect(std::integral_constant<int, kTraceClosest>{}, redo, counters + kCountRedo, counters + kCountFetchRedo, nullptr, nullptr);
            if(stats != nullptr) {
                PTB_CUDA(cudaMemcpyAsync(ctx->host_counters, counters, kCounterSlots * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
                PTB_CUDA(cudaStreamSynchronize(ctx->stream));
                retraced += ctx->host_counters[kCountRedo];
            }
        }
        else {
            launch_intersect(std::integral_constant<int, kTraceClosest>{}, order, nullptr, counters + kCountFetchClosest, nullptr, nullptr);
        }
    }
    PTB_CUDA(cudaGetLastError());
    if(!device_io) {
        PTB_CUDA(cudaMemcpyAsync(t_out, d_t, n_rays * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        PTB_CUDA(cudaMemcpyAsync(prim_out, d_prim, n_rays * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    status = endCall(ctx, stats);
    if(status != PTB_OK) {
        return status;
    }
    if(stats != nullptr) {
        stats->closest_rays = n_rays;
        stats->closest_rays_retraced = retraced;
        stats->kernel_launches = (certified ? 2 : 1) * ((n_rays + kChunk - 1) / kChunk);
    }
    return finishStats(ctx, count_visits, stats);
}

int ptb_occluded(ptb_scene *scene, const float *rays, uint64_t n_rays, uint8_t *occluded_out, uint32_t flags, ptb_render_stats *stats) {
    if(scene == nullptr || (n_rays > 0 && (rays == nullptr || occluded_out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_occluded: null argument");
    }
    ptb_context *ctx = scene->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = beginCall(ctx, stats, flags);
    if(status != PTB_OK) {
        return status;
    }
    if(n_rays == 0) {
        return endCall(ctx, stats);
    }
    const bool device_io = (flags & PTB_FLAG_DEVICE_IO) != 0U;
    const bool count_visits = (flags & PTB_FLAG_COUNT_VISITS) != 0U;

    const float *d_rays = rays;
    uint8_t *d_out = occluded_out;
    if(!device_io) {
        if((status = ctx->io_a.reserve(n_rays * 7 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n_rays)) != PTB_OK) {
            return status;
        }
        PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, rays, n_rays * 7 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        d_rays = ctx->io_a.as<float>();
        d_out = ctx->io_b.as<uint8_t>();
    }
    if((status = ctx->counters.reserve(kCounterSlots * sizeof(uint32_t))) != PTB_OK || (status = ctx->visits.reserve(2 * sizeof(VisitCounters))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, kCounterSlots * sizeof(uint32_t), ctx->stream));
    PTB_CUDA(cudaMemsetAsync(ctx->visits.ptr, 0, 2 * sizeof(VisitCounters), ctx->stream));

    constexpr uint64_t kChunk = 1ULL << 26;
    for(uint64_t first = 0; first < n_rays; first += kChunk) {
        const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(kChunk, n_rays - first));
        const int grid = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n) + kBlock - 1) / kBlock, static_cast<uint64_t>(gridFor(ctx, 16))));
        PTB_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, sizeof(uint32_t), ctx->stream));
        const uint32_t *order = nullptr;
        if(ctx->sort_rays && n >= kSortThreshold && (status = sortRays(scene, d_rays + 7 * first, 7, n, &order)) != PTB_OK) {
            return status;
        }
        LaunchTimer timer(ctx, 0); // the only kernel of this entry: always timed
        if(count_visits) {
            occludedKernel<true, 0><<<grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote_shadow, d_rays + 7 * first, order, n, d_out + first, ctx->counters.as<uint32_t>(),
                                                                       ctx->visits.as<VisitCounters>() + 1);
        }
        else if(scene->info.device_bytes > kBigSceneBytes) {
            occludedKernel<false, kQuerySmemAnyHit><<<grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote_shadow, d_rays + 7 * first, order, n, d_out + first,
                                                                                      ctx->counters.as<uint32_t>(), ctx->visits.as<VisitCounters>() + 1);
        }
        else {
            occludedKernel<false, 0><<<grid, kBlock, 0, ctx->stream>>>(scene->dev, ctx->vote_shadow, d_rays + 7 * first, order, n, d_out + first, ctx->counters.as<uint32_t>(),
                                                                        ctx->visits.as<VisitCounters>() + 1);
        }
    }
    PTB_CUDA(cudaGetLastError());
    if(!device_io) {
        PTB_CUDA(cudaMemcpyAsync(occluded_out, d_out, n_rays, cudaMemcpyDeviceToHost, ctx->stream));
    }
    status = endCall(ctx, stats);
    if(status != PTB_OK) {
        return status;
    }
    if(stats != nullptr) {
        stats->shadow_rays = n_rays;
        stats->kernel_launches = 1;
    }
    return finishStats(ctx, count_visits, stats);
}

// ---------------------------------------------------------------------------------------------------- render

int ptb_render_samples(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, uint64_t n, const int32_t *pixels, const uint64_t *seeds,
                       float *out_rgba, ptb_render_stats *stats) {
    if(scene == nullptr || camera == nullptr || opts == nullptr || (n > 0 && (pixels == nullptr || seeds == nullptr || out_rgba == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_samples: null argument");
    }
    if(opts->image_width <= 0 || opts->image_height <= 0) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_samples: empty image");
    }
    ptb_context *ctx = scene->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = beginCall(ctx, stats, opts->flags);
    if(status != PTB_OK) {
        return status;
    }
    if(n == 0) {
        return endCall(ctx, stats);
    }
    const bool device_io = (opts->flags & PTB_FLAG_DEVICE_IO) != 0U;
    const bool count_visits = (opts->flags & PTB_FLAG_COUNT_VISITS) != 0U;
    const uint32_t capacity = poolCapacity(n, poolLimit(ctx, scene->shadow_stride));

    PathPool pool{};
    if((status = carvePool(ctx, capacity, scene->shadow_stride, pool)) != PTB_OK) {
        return status;
    }
    const int32_t *d_pixels = pixels;
    const uint64_t *d_seeds = seeds;
    float4 *d_out = reinterpret_cast<float4 *>(out_rgba);
    if(!device_io) {
        if((status = ctx->io_a.reserve(n * 2 * sizeof(int32_t))) != PTB_OK || (status = ctx->io_b.reserve(n * sizeof(uint64_t))) != PTB_OK ||
           (status = ctx->io_c.reserve(n * sizeof(float4))) != PTB_OK) {
            return status;
        }
        PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, pixels, n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        PTB_CUDA(cudaMemcpyAsync(ctx->io_b.ptr, seeds, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        d_pixels = ctx->io_a.as<int32_t>();
        d_seeds = ctx->io_b.as<uint64_t>();
        d_out = ctx->io_c.as<float4>();
    }
    PTB_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, kCounterSlots * sizeof(uint32_t), ctx->stream));
    PTB_CUDA(cudaMemsetAsync(ctx->visits.ptr, 0, 2 * sizeof(VisitCounters), ctx->stream));

    const RenderParams params = makeParams(*camera, *opts);
    if(n > 0xFFFFFFFFULL) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_samples: more than 2^32 - 1 samples in one call");
    }
    PathSource src{};
    src.pixels = d_pixels;
    src.seeds = d_seeds;
    src.explicit_samples = 1U;
    src.total = n;
    if((status = runBounces(scene, pool, params, src, d_out, count_visits, closestMode(scene, opts->flags), stats)) != PTB_OK) {
        return status;
    }
    if(!device_io) {
        PTB_CUDA(cudaMemcpyAsync(out_rgba, d_out, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    }
    status = endCall(ctx, stats);
    if(status != PTB_OK) {
        return status;
    }
    return finishStats(ctx, count_visits, stats);
}

int ptb_render(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h, float *out_rgba,
               ptb_render_stats *stats) {
    return ptb_render_with_progress(scene, camera, opts, x0, y0, w, h, out_rgba, stats, nullptr, nullptr);
}

// One share of a render on one context: the tiles of shard (opts->shard_index, opts->shard_count), and of those every
// sub.count-th one starting at sub.index (several contexts of ONE device splitting a call between them, see
// ptb_render_with_progress).  sub.clear: zero the output first (false when the shares of a call write into one image).
static int renderShare(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h, float *out_rgba,
                       ptb_render_stats *stats, ptb_progress_fn progress, void *user, SubShard sub) {
    if(scene == nullptr || camera == nullptr || opts == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render: null argument");
    }
    if(w < 0 || h < 0 || x0 < 0 || y0 < 0 || x0 + w > 65535 || y0 + h > 65535) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render: rectangle out of range (coordinates are limited to 16 bits)");
    }
    if(opts->image_width <= 0 || opts->image_height <= 0) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render: empty image");
    }
    ptb_context *ctx = scene->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = beginCall(ctx, stats, opts->flags);
    if(status != PTB_OK) {
        return status;
    }
    if(w == 0 || h == 0) {
        return endCall(ctx, stats);
    }
    if(out_rgba == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render: out_rgba is null");
    }
    const bool device_io = (opts->flags & PTB_FLAG_DEVICE_IO) != 0U;
    const bool count_visits = (opts->flags & PTB_FLAG_COUNT_VISITS) != 0U;
    const int spp = std::max(opts->max_sample_count, 0);

    // tile grid over the rectangle; the reference's tile size when the rectangle is the whole image (worker.cpp:398)
    int tile = opts->tile_size;
    if(tile <= 0) {
        tile = std::max(std::min(std::min(w, h) / 4, 32), 1);
    }
    const int tiles_x = (w + tile - 1) / tile;
    const int tiles_y = (h + tile - 1) / tile;
    const int shard_count = std::max(opts->shard_count, 1);
    const int shard_index = std::min(std::max(opts->shard_index, 0), shard_count - 1);

    std::vector<int> owned;
    for(int t = 0, ordinal = 0; t < tiles_x * tiles_y; t++) {
        if(t % shard_count == shard_index) {
            if(ordinal % sub.count == sub.index) {
                owned.push_back(t);
            }
            ordinal++;
        }
    }

    const size_t out_bytes = static_cast<size_t>(w) * h * sizeof(float4);
    float4 *d_out = reinterpret_cast<float4 *>(out_rgba);
    if(!device_io) {
        if((status = ctx->io_d.reserve(out_bytes)) != PTB_OK) {
            return status;
        }
        d_out = ctx->io_d.as<float4>();
    }
    // pixels of tiles owned by other shards (and everything when spp == 0) stay 0
    if(sub.clear) {
        PTB_CUDA(cudaMemsetAsync(d_out, 0, out_bytes, ctx->stream));
    }

    // Pixel groups: consecutive runs of the owned tiles' pixels (tile order) whose per-sample buffer fits the budget and
    // whose sample count fits the 32-bit destination index of a path -- a single tile larger than either (processItem
    // passes tile_size = max(w, h); a caller may pass any tile_size) is split between groups like everything else.
    const uint64_t pool_limit = poolLimit(ctx, scene->shadow_stride);
    const uint64_t budget_bytes = sampleBufferBudget(ctx);
    std::vector<uint64_t> tile_first(owned.size() + 1, 0); // prefix sums of the owned tiles' pixel counts
    for(size_t k = 0; k < owned.size(); k++) {
        const int t = owned[k];
        const int tx0 = (t % tiles_x) * tile;
        const int ty0 = (t / tiles_x) * tile;
        tile_first[k + 1] = tile_first[k] + static_cast<uint64_t>(std::min(tx0 + tile, w) - tx0) * static_cast<uint64_t>(std::min(ty0 + tile, h) - ty0);
    }
    const uint64_t n_owned_pixels = tile_first.back();
    const uint64_t max_group_samples = std::min<uint64_t>(std::max<uint64_t>(budget_bytes / sizeof(float4), 1), 0xFFFFFFFFULL);

    // Adaptive sampling (min < max, worker.cpp:236-260): samples are traced in rounds and only for the pixels whose loop
    // has not ended.  The first round is as long as the shortest loop the reference can run (the acceptance test needs
    // check_sample_count consecutive passes, the first one no earlier than two full batches and min samples), later rounds
    // cover a quarter of the remaining range (PTB_ADAPTIVE_ROUND_DIVISOR; measured on the bench scene, 32..256 / 16..1024 spp:
    // 2 -> 540 / 3078 ms, 4 -> 562 / 3032, 8 -> 586 / 3182, 16 -> 650 / 3149), at least one run of checks.
    const ResolveConsts rc = resolveConsts(opts->min_sample_count, spp);
    // candidates a pixel's loop can open (worker.cpp:208-219): one per candidate_batch_count batches plus the remainder; the
    // per-pixel state holds kMaxCandidates of them (at most 6 for any option set: candidate_batch_count >= max / 4 batches)
    if(spp > 0 && (spp / rc.stats_sample_count) / rc.candidate_batch_count + 1 > kMaxCandidates) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_render: sample counts that open more than " + std::to_string(kMaxCandidates) + " candidates per pixel");
    }
    const bool adaptive = ctx->adaptive_rounds && spp > 0 && opts->min_sample_count < spp;
    int first_round = spp;
    int later_round = spp;
    if(adaptive) {
        const int batch = rc.stats_sample_count;
        const int first_check = ((std::max(std::max(rc.min_samples, 2), 2 * batch) + batch - 1) / batch) * batch;
        first_round = std::min(spp, first_check + (std::max(rc.check_sample_count, 1) - 1) * batch);
        const int divisor = static_cast<int>(std::min(64L, std::max(1L, envLong("PTB_ADAPTIVE_ROUND_DIVISOR", 4))));
        const int eighth = (((spp - first_round) / divisor + batch - 1) / batch) * batch;
        later_round = std::max(std::max(eighth, std::max(rc.check_sample_count, 1) * batch), 1);
    }
    const int longest_round = adaptive ? std::max(first_round, std::min(later_round, std::max(spp - first_round, 1))) : std::max(spp, 1);
    const uint64_t max_group_pixels = std::max<uint64_t>(1, max_group_samples / static_cast<uint64_t>(longest_round));
    if(static_cast<uint64_t>(std::max(spp, 1)) > 0xFFFFFFFFULL) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_render: more than 2^32 - 1 samples per pixel");
    }

    const RenderParams params = makeParams(*camera, *opts);
    if((status = ctx->counters.reserve(kCounterSlots * sizeof(uint32_t))) != PTB_OK || (status = ctx->visits.reserve(2 * sizeof(VisitCounters))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemsetAsync(ctx->visits.ptr, 0, 2 * sizeof(VisitCounters), ctx->stream));
    PTB_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, kCounterSlots * sizeof(uint32_t), ctx->stream));

    ProgressSink sink{progress, user, 0, n_owned_pixels * static_cast<uint64_t>(std::max(spp, 0))};
    std::vector<uint32_t> pixel_list;
    for(uint64_t group_begin = 0; group_begin < n_owned_pixels && spp > 0; group_begin += max_group_pixels) {
        const uint64_t group_end = std::min<uint64_t>(n_owned_pixels, group_begin + max_group_pixels);
        // pixels of the group in tile order; frames rendered repeatedly reuse the list already on the device
        const long long key[9] = {x0, y0, w, h, tile, shard_index + 65536LL * sub.index, shard_count + 65536LL * sub.count, static_cast<long long>(group_begin),
                                  static_cast<long long>(group_end)};
        const bool list_cached = std::equal(key, key + 9, ctx->pixel_list_key);
        const uint32_t n_pixels = static_cast<uint32_t>(group_end - group_begin);
        const uint64_t total = static_cast<uint64_t>(n_pixels) * (adaptive ? longest_round : spp);
        if((status = ctx->pixel_list.reserve(n_pixels * sizeof(uint32_t))) != PTB_OK || (status = ctx->samples.reserve(total * sizeof(float4))) != PTB_OK) {
            return status;
        }
        if(!list_cached) {
            pixel_list.clear();
            pixel_list.reserve(n_pixels);
            size_t k = static_cast<size_t>(std::upper_bound(tile_first.begin(), tile_first.end(), group_begin) - tile_first.begin()) - 1;
            for(uint64_t at = group_begin; at < group_end; k++) {
                const int t = owned[k];
                const int tx0 = (t % tiles_x) * tile;
                const int ty0 = (t / tiles_x) * tile;
                const int tw = std::min(tx0 + tile, w) - tx0;
                const uint64_t tile_end = std::min<uint64_t>(tile_first[k + 1], group_end);
                for(uint64_t q = at - tile_first[k]; q < tile_end - tile_first[k]; q++) {
                    const int x = tx0 + static_cast<int>(q % static_cast<uint64_t>(tw));
                    const int y = ty0 + static_cast<int>(q / static_cast<uint64_t>(tw));
                    pixel_list.push_back(static_cast<uint32_t>(x0 + x) | (static_cast<uint32_t>(y0 + y) << 16));
                }
                at = tile_end;
            }
            std::fill(ctx->pixel_list_key, ctx->pixel_list_key + 9, -1LL);
            // the list is consumed by kernels of this group only; a pageable copy on the stream is ordered before them
            PTB_CUDA(cudaMemcpyAsync(ctx->pixel_list.ptr, pixel_list.data(), n_pixels * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
            PTB_CUDA(cudaStreamSynchronize(ctx->stream));
            std::copy(key, key + 9, ctx->pixel_list_key);
        }

        ResolveParams rp{};
        rp.min_sample_count = opts->min_sample_count;
        rp.max_sample_count = spp;
        rp.n_pixels = n_pixels;
        rp.rect_x0 = x0;
        rp.rect_y0 = y0;
        rp.rect_w = w;

        if(adaptive) {
            if((status = renderAdaptiveGroup(scene, params, rp, rc, first_round, later_round, pool_limit, count_visits, closestMode(scene, opts->flags), d_out, stats, &sink)) != PTB_OK) {
                return status;
            }
            sink.done_before = group_end * static_cast<uint64_t>(spp);
            if(progress != nullptr && group_end < n_owned_pixels) {
                progress(user, sink.done_before, sink.total);
            }
            continue;
        }

        const uint32_t capacity = poolCapacity(total, pool_limit);
        PathPool pool{};
        if((status = carvePool(ctx, capacity, scene->shadow_stride, pool)) != PTB_OK) {
            return status;
        }

        PathSource src{};
        src.pixel_list = ctx->pixel_list.as<uint32_t>();
        src.n_pixels = n_pixels;
        src.explicit_samples = 0U;
        src.total = total;
        if((status = runBounces(scene, pool, params, src, ctx->samples.as<float4>(), count_visits, closestMode(scene, opts->flags), stats, &sink)) != PTB_OK) {
            return status;
        }
        sink.done_before += total;
        if(stats != nullptr) {
            stats->samples_used += total; // refined below for min != max without rounds: unknown, reported as traced
        }
        {
            LaunchTimer timer(ctx, 1);
            resolveKernel<<<(n_pixels + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(rp, ctx->samples.as<float4>(), ctx->pixel_list.as<uint32_t>(), d_out);
        }
        PTB_CUDA(cudaGetLastError());
        if(stats != nullptr) {
            stats->kernel_launches += 1;
        }
        if(progress != nullptr && group_end < n_owned_pixels) {
            progress(user, sink.done_before, sink.total);
        }
    }

    if(!device_io) {
        PTB_CUDA(cudaMemcpyAsync(out_rgba, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    status = endCall(ctx, stats);
    if(status != PTB_OK) {
        return status;
    }
    if(progress != nullptr) {
        progress(user, sink.total, sink.total);
    }
    return finishStats(ctx, count_visits, stats);
}

// ---------------------------------------------------------------------------------------------------- unit entries

int ptb_camera_shoot(ptb_context *ctx, const ptb_camera *camera, uint64_t n, const float *xy, float pixel_width, float pixel_height,
                     uint64_t *engine_states, float *rays_out) {
    if(ctx == nullptr || camera == nullptr || (n > 0 && (xy == nullptr || engine_states == nullptr || rays_out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_camera_shoot: null argument");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n == 0) {
        return status;
    }
    if((status = ctx->io_a.reserve(n * 2 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n * sizeof(uint64_t))) != PTB_OK ||
       (status = ctx->io_c.reserve(n * 6 * sizeof(float))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, xy, n * 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(ctx->io_b.ptr, engine_states, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    cameraKernel<<<static_cast<unsigned>((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(*camera, n, ctx->io_a.as<float>(), pixel_width, pixel_height,
                                                                                                 ctx->io_b.as<uint64_t>(), ctx->io_c.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(rays_out, ctx->io_c.ptr, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(engine_states, ctx->io_b.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_aperture_sample(ptb_context *ctx, uint32_t aperture_kind, float hexagon_horizontal_ratio, uint64_t n, uint64_t *engine_states, float *out) {
    if(ctx == nullptr || (n > 0 && (engine_states == nullptr || out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_aperture_sample: null argument");
    }
    if(aperture_kind != PTB_APERTURE_CIRCULAR && aperture_kind != PTB_APERTURE_HEXAGONAL) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_aperture_sample: unknown aperture kind");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n == 0) {
        return status;
    }
    if((status = ctx->io_b.reserve(n * sizeof(uint64_t))) != PTB_OK || (status = ctx->io_c.reserve(n * 2 * sizeof(float))) != PTB_OK) {
        return status;
    }
    ptb_camera camera{};
    camera.aperture_kind = aperture_kind;
    camera.hexagon_horizontal_ratio = std::min(std::max(hexagon_horizontal_ratio, 0.0F), 1.0F);
    PTB_CUDA(cudaMemcpyAsync(ctx->io_b.ptr, engine_states, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    apertureKernel<<<static_cast<unsigned>((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(camera, n, ctx->io_b.as<uint64_t>(), ctx->io_c.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(out, ctx->io_c.ptr, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(engine_states, ctx->io_b.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_sample_lights(ptb_scene *scene, const float pos[3], uint64_t *engine_state, uint32_t max_out, float *out, uint32_t *n_out) {
    if(scene == nullptr || pos == nullptr || engine_state == nullptr || n_out == nullptr || (max_out > 0 && out == nullptr)) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_sample_lights: null argument");
    }
    ptb_context *ctx = scene->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    const size_t floats = 4 + 8 * static_cast<size_t>(max_out);
    if((status = ctx->io_a.reserve(floats * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(sizeof(uint64_t))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_b.ptr, engine_state, sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    sampleLightsKernel<<<1, 32, 0, ctx->stream>>>(scene->dev, pos[0], pos[1], pos[2], ctx->io_b.as<uint64_t>(), max_out, ctx->io_a.as<float>());
    PTB_CUDA(cudaGetLastError());
    std::vector<float> host(floats);
    PTB_CUDA(cudaMemcpyAsync(host.data(), ctx->io_a.ptr, floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(engine_state, ctx->io_b.ptr, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t n = 0;
    std::memcpy(&n, &host[0], sizeof(uint32_t));
    *n_out = n;
    const uint32_t copy = std::min(n, max_out);
    if(copy > 0) {
        std::memcpy(out, &host[4], 8 * static_cast<size_t>(copy) * sizeof(float));
    }
    return PTB_OK;
}

int ptb_aabb_intersect(ptb_context *ctx, const float low[3], const float high[3], uint64_t n_rays, const float *rays, float *t_out) {
    if(ctx == nullptr || low == nullptr || high == nullptr || (n_rays > 0 && (rays == nullptr || t_out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_aabb_intersect: null argument");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n_rays == 0) {
        return status;
    }
    if((status = ctx->io_a.reserve(n_rays * 6 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n_rays * sizeof(float))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, rays, n_rays * 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    aabbKernel<<<static_cast<unsigned>((n_rays + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(low[0], low[1], low[2], high[0], high[1], high[2],
                                                                                                    ctx->io_a.as<float>(), n_rays, ctx->io_b.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(t_out, ctx->io_b.ptr, n_rays * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_prim_intersect(ptb_context *ctx, const ptb_prim *prim, uint64_t n_rays, const float *rays, float *t_out) {
    if(ctx == nullptr || prim == nullptr || (n_rays > 0 && (rays == nullptr || t_out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_prim_intersect: null argument");
    }
    if(prim->kind > PTB_PRIM_NULL) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_prim_intersect: unknown primitive kind");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n_rays == 0) {
        return status;
    }
    const float *p = prim->p;
    float4 geom[4] = {};
    uint32_t flags = prim->kind & kKindMask;
    if(prim->kind == PTB_PRIM_TRIANGLE) {
        if(prim->cull_backface != 0U) {
            flags |= kCullBit;
        }
        geom[0] = make_float4(p[0], p[1], p[2], 0.0F);
        geom[1] = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0F);
        geom[2] = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0F);
    }
    else {
        geom[0] = make_float4(p[0], p[1], p[2], 0.0F);
        geom[1] = make_float4(p[3], p[3] * p[3], 0.0F, 0.0F);
        geom[2] = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
    }
    std::memcpy(&geom[0].w, &flags, sizeof(flags));
    if((status = ctx->io_a.reserve(n_rays * 6 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n_rays * sizeof(float))) != PTB_OK ||
       (status = ctx->io_c.reserve(sizeof(geom))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, rays, n_rays * 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(ctx->io_c.ptr, geom, sizeof(geom), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    DeviceScene one{};
    one.geom = ctx->io_c.as<float4>();
    one.n_prims = 1;
    primKernel<<<static_cast<unsigned>((n_rays + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(one, ctx->io_a.as<float>(), n_rays, ctx->io_b.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(t_out, ctx->io_b.ptr, n_rays * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

namespace {

    // geometry lanes (differenced edges), shading lanes and un-differenced lanes of one primitive
    struct PrimLanes {
        float4 geom[4];
        float4 shade[3];
        float4 raw[3];
        uint32_t flags;
    };

    PrimLanes lanesOf(const ptb_prim &prim) {
        PrimLanes l{};
        const float *p = prim.p;
        l.flags = prim.kind & kKindMask;
        if(prim.kind == PTB_PRIM_TRIANGLE) {
            if(prim.cull_backface != 0U) {
                l.flags |= kCullBit;
            }
            l.geom[0] = make_float4(p[0], p[1], p[2], 0.0F);
            l.geom[1] = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0F);
            l.geom[2] = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0F);
            l.shade[0] = make_float4(p[9], p[10], p[11], 0.0F);
            l.shade[1] = make_float4(p[12], p[13], p[14], 0.0F);
            l.shade[2] = make_float4(p[15], p[16], p[17], 0.0F);
            l.raw[0] = make_float4(p[0], p[1], p[2], 0.0F);
            l.raw[1] = make_float4(p[3], p[4], p[5], primSampleDensity(prim));
            l.raw[2] = make_float4(p[6], p[7], p[8], 0.0F);
        }
        else if(prim.kind == PTB_PRIM_SPHERE) {
            l.geom[0] = make_float4(p[0], p[1], p[2], 0.0F);
            l.geom[1] = make_float4(p[3], p[3] * p[3], 0.0F, 0.0F);
            l.raw[0] = l.geom[0];
            l.raw[1] = make_float4(p[3], p[3] * p[3], 0.0F, primSampleDensity(prim));
        }
        std::memcpy(&l.geom[0].w, &l.flags, sizeof(uint32_t));
        return l;
    }

}

int ptb_prim_normal(ptb_context *ctx, const ptb_prim *prim, uint64_t n, const float *positions, float *normals_out) {
    if(ctx == nullptr || prim == nullptr || (n > 0 && (positions == nullptr || normals_out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_prim_normal: null argument");
    }
    if(prim->kind > PTB_PRIM_NULL) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_prim_normal: unknown primitive kind");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n == 0) {
        return status;
    }
    const PrimLanes lanes = lanesOf(*prim);
    if((status = ctx->io_a.reserve(n * 3 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n * 3 * sizeof(float))) != PTB_OK ||
       (status = ctx->io_c.reserve(8 * sizeof(float4))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, positions, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(ctx->io_c.ptr, lanes.geom, 4 * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(ctx->io_c.as<float4>() + 4, lanes.shade, 3 * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    DeviceScene one{};
    one.geom = ctx->io_c.as<float4>();
    one.shade = ctx->io_c.as<float4>() + 4;
    one.n_prims = 1;
    primNormalKernel<<<static_cast<unsigned>((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(one, n, ctx->io_a.as<float>(), ctx->io_b.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(normals_out, ctx->io_b.ptr, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_prim_sample(ptb_context *ctx, const ptb_prim *prim, uint64_t n, uint64_t *engine_states, float *out) {
    if(ctx == nullptr || prim == nullptr || (n > 0 && (engine_states == nullptr || out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_prim_sample: null argument");
    }
    if(prim->kind > PTB_PRIM_NULL) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_prim_sample: unknown primitive kind");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n == 0) {
        return status;
    }
    const PrimLanes lanes = lanesOf(*prim);
    if((status = ctx->io_a.reserve(n * sizeof(uint64_t))) != PTB_OK || (status = ctx->io_b.reserve(n * 5 * sizeof(float))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, engine_states, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    primSampleKernel<<<static_cast<unsigned>((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(lanes.raw[0], lanes.raw[1], lanes.raw[2], lanes.flags, n,
                                                                                                     ctx->io_a.as<uint64_t>(), ctx->io_b.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(out, ctx->io_b.ptr, n * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(engine_states, ctx->io_a.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_bsdf_propagate(ptb_context *ctx, const ptb_material *material, float epsilon, uint64_t n, const float *in, uint64_t *engine_states, float *out) {
    if(ctx == nullptr || material == nullptr || (n > 0 && (in == nullptr || engine_states == nullptr || out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_bsdf_propagate: null argument");
    }
    if(material->bsdf > PTB_BSDF_MIRROR) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_bsdf_propagate: unknown BSDF kind");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n == 0) {
        return status;
    }
    if((status = ctx->io_a.reserve(n * 9 * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(n * sizeof(uint64_t))) != PTB_OK ||
       (status = ctx->io_c.reserve(n * 8 * sizeof(float))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, in, n * 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(ctx->io_b.ptr, engine_states, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    bsdfPropagateKernel<<<static_cast<unsigned>((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(*material, epsilon, n, ctx->io_a.as<float>(),
                                                                                                        ctx->io_b.as<uint64_t>(), ctx->io_c.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(out, ctx->io_c.ptr, n * 8 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaMemcpyAsync(engine_states, ctx->io_b.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_bsdf_spectrum(ptb_context *ctx, const ptb_material *material, uint32_t synthetic, uint64_t n, const float *in, float *out) {
    if(ctx == nullptr || material == nullptr || (n > 0 && (in == nullptr || out == nullptr))) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_bsdf_spectrum: null argument");
    }
    if(material->bsdf > PTB_BSDF_MIRROR) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_bsdf_spectrum: unknown BSDF kind");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK || n == 0) {
        return status;
    }
    if((status = ctx->io_a.reserve(n * 13 * sizeof(float))) != PTB_OK || (status = ctx->io_c.reserve(n * 6 * sizeof(float))) != PTB_OK) {
        return status;
    }
    PTB_CUDA(cudaMemcpyAsync(ctx->io_a.ptr, in, n * 13 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    bsdfSpectrumKernel<<<static_cast<unsigned>((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(*material, synthetic, n, ctx->io_a.as<float>(),
                                                                                                       ctx->io_c.as<float>());
    PTB_CUDA(cudaGetLastError());
    PTB_CUDA(cudaMemcpyAsync(out, ctx->io_c.ptr, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}

int ptb_post_process(ptb_context *ctx, float *rgba, int32_t width, int32_t height, uint32_t mode, float gamma, uint32_t flags) {
    if(ctx == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_post_process: null context");
    }
    if(mode > PTB_POST_BOTH) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_post_process: unknown mode");
    }
    if(width < 0 || height < 0 || static_cast<int64_t>(width) * height > 0x7FFFFFFFLL) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_post_process: image size out of range");
    }
    const uint32_t n = static_cast<uint32_t>(width) * static_cast<uint32_t>(height);
    if(n == 0U) {
        return PTB_OK;
    }
    if(rgba == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_post_process: null image");
    }
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    const bool device_io = (flags & PTB_FLAG_DEVICE_IO) != 0U;
    const size_t image_bytes = static_cast<size_t>(n) * sizeof(float4);
    float4 *pixels = reinterpret_cast<float4 *>(rgba);
    if(!device_io) {
        if((status = ctx->io_d.reserve(image_bytes)) != PTB_OK) {
            return status;
        }
        pixels = ctx->io_d.as<float4>();
        PTB_CUDA(cudaMemcpyAsync(pixels, rgba, image_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    const unsigned blocks = (n + 255U) / 256U;

    if(mode == PTB_POST_TONE_MAP || mode == PTB_POST_BOTH) {
        // segment weights and the data-independent part of the ceilings recurrence (post_processing.cpp:92-128):
        // which sorted element each segment's ceiling picks depends on the pixel count and the weights only
        const int pixel_count = static_cast<int>(n);
        const int segments = std::min(1024, pixel_count);
        std::vector<float> weights(static_cast<size_t>(segments));
        float total_weight = 0.0F;
        for(int i = 0; i < segments; i++) {
            float x = (static_cast<float>(i) + 0.5F) / static_cast<float>(segments);
            x = 2.0F * (x - 0.5F);
            constexpr float pi = static_cast<float>(M_PI);
            const float fac = 1.0F / (std::sqrt(2 * pi));
            const float exponent_part = (x - 0.0F) / (0.3F);
            const float gaussian = fac * std::exp(-(exponent_part * exponent_part) / 2.0F) / 0.3F;
            weights[i] = 0.1F + gaussian;
            total_weight += weights[i];
        }
        std::vector<int32_t> pick(static_cast<size_t>(segments), -1);
        int previous_index = 0;
        float missed = 0.0F;
        for(int i = 0; i < segments - 1; i++) {
            const int item_count = static_cast<int>(std::round(weights[i] * static_cast<float>(pixel_count) / total_weight + missed));
            if(item_count > 0) {
                pick[i] = std::min(previous_index + item_count - 1, pixel_count - 1);
                previous_index += item_count;
                missed = 0.0F;
            }
            else {
                missed += weights[i] * static_cast<float>(pixel_count) / total_weight;
            }
        }

        size_t sort_bytes = 0;
        size_t reduce_bytes = 0;
        PTB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, static_cast<const float *>(nullptr), static_cast<float *>(nullptr), static_cast<int>(n), 0, 32,
                                                ctx->stream));
        PTB_CUDA(cub::DeviceReduce::Min(nullptr, reduce_bytes, static_cast<const float *>(nullptr), static_cast<float *>(nullptr), static_cast<int>(n), ctx->stream));
        const size_t temp_bytes = std::max(sort_bytes, reduce_bytes) + 256;
        const size_t floats = static_cast<size_t>(n);
        // io_a: brightness | sorted ; io_b: cub temp ; io_c: range(2) + ceilings + pick
        if((status = ctx->io_a.reserve(2 * floats * sizeof(float))) != PTB_OK || (status = ctx->io_b.reserve(temp_bytes)) != PTB_OK ||
           (status = ctx->io_c.reserve((2 + 2 * static_cast<size_t>(segments)) * sizeof(float) + 64)) != PTB_OK) {
            return status;
        }
        float *brightness = ctx->io_a.as<float>();
        float *sorted = brightness + floats;
        float *range = ctx->io_c.as<float>();
        float *ceilings = range + 2;
        int32_t *d_pick = reinterpret_cast<int32_t *>(ceilings + segments);
        PTB_CUDA(cudaMemcpyAsync(d_pick, pick.data(), static_cast<size_t>(segments) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        PTB_CUDA(cudaStreamSynchronize(ctx->stream)); // `pick` is pageable host memory owned by this frame

        brightnessKernel<<<blocks, 256, 0, ctx->stream>>>(pixels, n, brightness);
        size_t bytes = temp_bytes;
        PTB_CUDA(cub::DeviceReduce::Min(ctx->io_b.ptr, bytes, brightness, range, static_cast<int>(n), ctx->stream));
        bytes = temp_bytes;
        PTB_CUDA(cub::DeviceReduce::Max(ctx->io_b.ptr, bytes, brightness, range + 1, static_cast<int>(n), ctx->stream));
        rangeKernel<<<1, 32, 0, ctx->stream>>>(range);
        bytes = temp_bytes;
        PTB_CUDA(cub::DeviceRadixSort::SortKeys(ctx->io_b.ptr, bytes, brightness, sorted, static_cast<int>(n), 0, 32, ctx->stream));
        ceilingsKernel<<<1, 32, 0, ctx->stream>>>(sorted, d_pick, segments, range, ceilings);
        toneMapKernel<<<blocks, 256, 0, ctx->stream>>>(pixels, n, ceilings, segments, range);
        PTB_CUDA(cudaGetLastError());
    }
    if(mode == PTB_POST_GAMMA || mode == PTB_POST_BOTH) {
        const float exponent = 1.0F / gamma - 1.0F;
        gammaKernel<<<blocks, 256, 0, ctx->stream>>>(pixels, n, exponent);
        PTB_CUDA(cudaGetLastError());
    }
    if(!device_io) {
        PTB_CUDA(cudaMemcpyAsync(rgba, pixels, image_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    PTB_CUDA(cudaGetLastError());
    return PTB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------- several GPUs

namespace {

    constexpr int kMaxReplicas = 16;

    struct GatherParams {
        const float4 *src[kMaxReplicas];
        int32_t n;
        int32_t w;
        int32_t h;
        int32_t tile;
        int32_t tiles_x;
    };

    // Assembles the frame from the replicas' images: pixel (x, y) belongs to tile (y / tile) * tiles_x + x / tile, which
    // replica (tile index % n) rendered.  The sources are peer-device pointers (read over NVLink) or staged local copies.
    __global__ void gatherTilesKernel(GatherParams g, float4 *__restrict__ dst) {
        const int64_t count = static_cast<int64_t>(g.w) * g.h;
        for(int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < count; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            const int x = static_cast<int>(p % g.w);
            const int y = static_cast<int>(p / g.w);
            const int owner = ((y / g.tile) * g.tiles_x + x / g.tile) % g.n;
            dst[p] = g.src[owner][p];
        }
    }

    struct MultiProgress {
        std::mutex mutex;
        std::atomic<uint64_t> done[kMaxReplicas];
        std::atomic<uint64_t> total[kMaxReplicas];
        ptb_progress_fn fn = nullptr;
        void *user = nullptr;
        int n = 0;
    };

    struct ReplicaProgress {
        MultiProgress *all;
        int index;
    };

    void replicaProgress(void *user, uint64_t done, uint64_t total) {
        auto *self = static_cast<ReplicaProgress *>(user);
        MultiProgress *all = self->all;
        all->done[self->index].store(done);
        all->total[self->index].store(total);
        std::lock_guard<std::mutex> lock(all->mutex); // the caller's callback never runs concurrently (reference include/PathTrace/worker.h:76-79)
        uint64_t sum_done = 0;
        uint64_t sum_total = 0;
        bool all_known = true;
        for(int i = 0; i < all->n; i++) {
            sum_done += all->done[i].load();
            const uint64_t t = all->total[i].load();
            all_known = all_known && t != ~0ULL;
            sum_total += t != ~0ULL ? t : 0;
        }
        if(all_known && sum_done < sum_total) { // the final report (done == total) is issued once, after the gather
            all->fn(all->user, sum_done, sum_total);
        }
    }

}

extern "C" {

int ptb_render_with_progress(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h,
                             float *out_rgba, ptb_render_stats *stats, ptb_progress_fn progress, void *user) {
    if(scene == nullptr || camera == nullptr || opts == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render: null argument");
    }
    ptb_context *ctx = scene->ctx;
    // split between the contexts of this device?  Only renders large enough for the second workspace and the threads to pay.
    const int64_t share_samples = static_cast<int64_t>(std::max(w, 0)) * std::max(h, 0) / std::max(opts->shard_count, 1) * std::max(opts->max_sample_count, 0);
    const int parts = (ctx->is_sibling || out_rgba == nullptr || share_samples < (1LL << 26) || (opts->flags & PTB_FLAG_SINGLE_STREAM) != 0U) ? 1 : ctx->streams;
    if(parts <= 1) {
        return renderShare(scene, camera, opts, x0, y0, w, h, out_rgba, stats, progress, user, SubShard{0, 1, true});
    }
    if(w < 0 || h < 0 || x0 < 0 || y0 < 0 || x0 + w > 65535 || y0 + h > 65535) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render: rectangle out of range (coordinates are limited to 16 bits)");
    }

    std::lock_guard<std::mutex> call_lock(ctx->call_mutex);
    const double t_call = nowSeconds();
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    while(static_cast<int>(ctx->siblings.size()) < parts - 1) {
        ptb_context *sibling = nullptr;
        if((status = ptb_context_create(ctx->device, &sibling)) != PTB_OK) {
            return status;
        }
        sibling->is_sibling = true;
        ctx->siblings.push_back(sibling);
    }
    while(static_cast<int>(scene->aliases.size()) < parts - 1) {
        auto *alias = new(std::nothrow) ptb_scene();
        if(alias == nullptr) {
            return fail(PTB_ERR_OUT_OF_MEMORY, "ptb_render: host allocation failed");
        }
        alias->ctx = ctx->siblings[scene->aliases.size()];
        alias->dev = scene->dev;
        alias->info = scene->info;
        alias->shadow_stride = scene->shadow_stride;
        alias->guard = scene->guard;
        scene->aliases.push_back(alias);
    }
    for(int k = 0; k < parts - 1; k++) {
        scene->aliases[static_cast<size_t>(k)]->ctx = ctx->siblings[static_cast<size_t>(k)];
        ctx->siblings[static_cast<size_t>(k)]->budget_divisor = parts;
    }
    ctx->budget_divisor = parts;
    struct RestoreBudget { // every way out of this call gives the context its whole budget back
        ptb_context *context;
        ~RestoreBudget() { context->budget_divisor = 1; }
    } restore_budget{ctx};

    const bool device_io = (opts->flags & PTB_FLAG_DEVICE_IO) != 0U;
    const size_t image_bytes = static_cast<size_t>(w) * h * sizeof(float4);
    float4 *d_out = reinterpret_cast<float4 *>(out_rgba);
    {
        std::lock_guard<std::mutex> lock(ctx->mutex);
        if(!device_io) {
            if((status = ctx->split_image.reserve(image_bytes)) != PTB_OK) {
                return status;
            }
            d_out = ctx->split_image.as<float4>();
        }
        // the shares write their own pixels only; everything else (other shards' tiles) stays 0
        PTB_CUDA(cudaMemsetAsync(d_out, 0, image_bytes, ctx->stream));
        PTB_CUDA(cudaEventRecord(ctx->split_start, ctx->stream));
        PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    }

    MultiProgress all;
    all.fn = progress;
    all.user = user;
    all.n = parts;
    for(int i = 0; i < kMaxReplicas; i++) {
        all.done[i].store(0);
        all.total[i].store(~0ULL);
    }
    std::vector<ReplicaProgress> sinks(static_cast<size_t>(parts));
    std::vector<ptb_render_stats> part_stats(static_cast<size_t>(parts));
    std::vector<int> statuses(static_cast<size_t>(parts), PTB_OK);
    std::vector<std::string> errors(static_cast<size_t>(parts));
    auto render_part = [&](int k) {
        ptb_render_opts mine = *opts;
        mine.flags |= PTB_FLAG_DEVICE_IO;
        sinks[static_cast<size_t>(k)] = ReplicaProgress{&all, k};
        ptb_scene *target = k == 0 ? scene : scene->aliases[static_cast<size_t>(k) - 1];
        const double t0 = nowSeconds();
        statuses[static_cast<size_t>(k)] = renderShare(target, camera, &mine, x0, y0, w, h, reinterpret_cast<float *>(d_out), &part_stats[static_cast<size_t>(k)],
                                                       progress != nullptr ? replicaProgress : nullptr, &sinks[static_cast<size_t>(k)], SubShard{k, parts, false});
        if(envLong("PTB_LOG_SPLIT", 0) != 0) {
            const ptb_render_stats &st = part_stats[static_cast<size_t>(k)];
            std::fprintf(stderr, "[ptb] split render: share %d of %d: %.1f ms wall (started %.1f ms after the call), %.1f ms on its stream, %llu samples, %llu iterations, %llu launches\n", k,
                         parts, (nowSeconds() - t0) * 1e3, (t0 - t_call) * 1e3, st.device_ms_total, static_cast<unsigned long long>(st.samples),
                         static_cast<unsigned long long>(st.bounce_iterations), static_cast<unsigned long long>(st.kernel_launches));
        }
        if(statuses[static_cast<size_t>(k)] != PTB_OK) {
            errors[static_cast<size_t>(k)] = g_last_error;
        }
    };
    {
        std::vector<std::thread> workers;
        for(int k = 1; k < parts; k++) {
            workers.emplace_back(render_part, k);
        }
        render_part(0);
        for(std::thread &t : workers) {
            t.join();
        }
    }
    for(int k = 0; k < parts; k++) {
        if(statuses[static_cast<size_t>(k)] != PTB_OK) {
            return fail(statuses[static_cast<size_t>(k)], "ptb_render: share " + std::to_string(k) + ": " + errors[static_cast<size_t>(k)]);
        }
    }
    float elapsed_ms = 0.0F;
    {
        std::lock_guard<std::mutex> lock(ctx->mutex);
        if((status = useDevice(ctx)) != PTB_OK) {
            return status;
        }
        // every share has synchronised its own stream: this event closes the call on the device's timeline
        PTB_CUDA(cudaEventRecord(ctx->split_stop, ctx->stream));
        if(!device_io) {
            PTB_CUDA(cudaMemcpyAsync(out_rgba, d_out, image_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        PTB_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&elapsed_ms, ctx->split_start, ctx->split_stop);
    }
    if(stats != nullptr) {
        std::memset(stats, 0, sizeof(*stats));
        for(const ptb_render_stats &part : part_stats) {
            stats->samples += part.samples;
            stats->closest_rays += part.closest_rays;
            stats->shadow_rays += part.shadow_rays;
            stats->shadow_rays_skipped += part.shadow_rays_skipped;
            stats->path_vertices += part.path_vertices;
            stats->inner_visits += part.inner_visits;
            stats->leaf_visits += part.leaf_visits;
            stats->bounce_iterations = std::max(stats->bounce_iterations, part.bounce_iterations);
            stats->kernel_launches += part.kernel_launches;
            // kernel times of concurrent shares overlap on the device: their sums exceed the call's own duration
            stats->device_ms_trace += part.device_ms_trace;
            stats->device_ms_shade += part.device_ms_shade;
            stats->device_ms_trace_shadow += part.device_ms_trace_shadow;
            stats->shadow_inner_visits += part.shadow_inner_visits;
            stats->shadow_leaf_visits += part.shadow_leaf_visits;
            stats->closest_rays_retraced += part.closest_rays_retraced;
            stats->certified_suspect_hits += part.certified_suspect_hits;
            stats->samples_used += part.samples_used;
            stats->adaptive_rounds = std::max(stats->adaptive_rounds, part.adaptive_rounds);
        }
        stats->device_ms_total = elapsed_ms;
    }
    if(progress != nullptr) {
        uint64_t total = 0;
        for(int k = 0; k < parts; k++) {
            const uint64_t t = all.total[k].load();
            total += t != ~0ULL ? t : 0;
        }
        progress(user, total, total);
    }
    return PTB_OK;
}

int ptb_device_count(int *count_out) {
    if(count_out == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_device_count: null argument");
    }
    int count = 0;
    const cudaError_t err = cudaGetDeviceCount(&count);
    if(err != cudaSuccess) {
        cudaGetLastError();
        count = 0;
    }
    *count_out = count;
    return PTB_OK;
}

int ptb_scene_clone(const ptb_scene *scene, ptb_context *ctx, ptb_scene **out) {
    if(scene == nullptr || ctx == nullptr || out == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_scene_clone: null argument");
    }
    *out = nullptr;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    auto *copy = new(std::nothrow) ptb_scene();
    if(copy == nullptr) {
        return fail(PTB_ERR_OUT_OF_MEMORY, "ptb_scene_clone: host allocation failed");
    }
    copy->ctx = ctx;
    copy->info = scene->info;
    copy->shadow_stride = scene->shadow_stride;
    copy->guard = scene->guard;
    copy->dev = scene->dev;
    const int src_device = scene->ctx->device;
    struct Pair {
        const Buffer *from;
        Buffer *to;
    };
    const Pair pairs[] = {{&scene->nodes, &copy->nodes}, {&scene->occ_nodes, &copy->occ_nodes}, {&scene->geom, &copy->geom}, {&scene->shade, &copy->shade}, {&scene->mats, &copy->mats},
                          {&scene->lights, &copy->lights}, {&scene->emis, &copy->emis}, {&scene->cdf, &copy->cdf}, {&scene->slot_to_prim, &copy->slot_to_prim}};
    for(const Pair &pair : pairs) {
        if(pair.from->ptr == nullptr) {
            continue;
        }
        if((status = pair.to->reserve(pair.from->bytes)) != PTB_OK) {
            ptb_scene_destroy(copy);
            return status;
        }
        const cudaError_t err = cudaMemcpyPeer(pair.to->ptr, ctx->device, pair.from->ptr, src_device, pair.from->bytes);
        if(err != cudaSuccess) {
            cudaGetLastError();
            ptb_scene_destroy(copy);
            return fail(PTB_ERR_CUDA, std::string("ptb_scene_clone: cudaMemcpyPeer: ") + cudaGetErrorString(err));
        }
    }
    DeviceScene &d = copy->dev;
    d.nodes = copy->nodes.as<float4>();
    d.occ_nodes = scene->dev.occ_nodes != nullptr ? copy->occ_nodes.as<float4>() : nullptr;
    d.geom = copy->geom.as<float4>();
    d.shade = copy->shade.as<float4>();
    d.mats = copy->mats.as<float4>();
    d.lights = copy->lights.as<float4>();
    d.emis = copy->emis.as<float4>();
    d.cdf = copy->cdf.as<float>();
    d.slot_to_prim = copy->slot_to_prim.as<uint32_t>();
    *out = copy;
    return PTB_OK;
}

int ptb_render_multi(ptb_scene *const *replicas, int32_t n, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h,
                     float *out_rgba, ptb_render_stats *stats, ptb_progress_fn progress, void *user) {
    if(replicas == nullptr || n < 1 || camera == nullptr || opts == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_multi: null argument");
    }
    if(n > kMaxReplicas) {
        return fail(PTB_ERR_UNSUPPORTED, "ptb_render_multi: more than 16 replicas");
    }
    for(int i = 0; i < n; i++) {
        if(replicas[i] == nullptr) {
            return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_multi: null replica");
        }
        for(int j = 0; j < i; j++) {
            if(replicas[j]->ctx == replicas[i]->ctx) {
                return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_multi: two replicas share a context");
            }
        }
    }
    if(opts->shard_count > 1) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_multi: the call shards the frame itself; shard_count must be <= 1");
    }
    if(n == 1) {
        return ptb_render_with_progress(replicas[0], camera, opts, x0, y0, w, h, out_rgba, stats, progress, user);
    }
    if(w < 0 || h < 0) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_multi: negative rectangle");
    }
    if(w == 0 || h == 0) {
        return PTB_OK;
    }
    if(out_rgba == nullptr) {
        return fail(PTB_ERR_INVALID_ARGUMENT, "ptb_render_multi: out_rgba is null");
    }
    const size_t image_bytes = static_cast<size_t>(w) * h * sizeof(float4);
    for(int i = 0; i < n; i++) {
        ptb_context *ctx = replicas[i]->ctx;
        std::lock_guard<std::mutex> lock(ctx->mutex);
        int status = useDevice(ctx);
        if(status != PTB_OK || (status = ctx->multi_image.reserve(image_bytes)) != PTB_OK) {
            return status;
        }
    }

    MultiProgress all;
    all.fn = progress;
    all.user = user;
    all.n = n;
    for(int i = 0; i < kMaxReplicas; i++) {
        all.done[i].store(0);
        all.total[i].store(~0ULL);
    }
    std::vector<ReplicaProgress> sinks(static_cast<size_t>(n));
    std::vector<int> statuses(static_cast<size_t>(n), PTB_OK);
    std::vector<std::string> errors(static_cast<size_t>(n));
    const bool log_multi = envLong("PTB_LOG_MULTI", 0) != 0;
    const long stagger_ms = envLong("PTB_MULTI_STAGGER_MS", 0);
    const double t_begin = nowSeconds();
    std::vector<ptb_render_stats> own_stats(static_cast<size_t>(n));
    std::vector<double> wall(static_cast<size_t>(n), 0.0);
    auto render_share = [&](int i) {
        ptb_render_opts mine = *opts;
        mine.shard_index = i;
        mine.shard_count = n;
        mine.flags |= PTB_FLAG_DEVICE_IO;
        sinks[i] = ReplicaProgress{&all, i};
        if(stagger_ms > 0 && i > 0) {
            std::this_thread::sleep_for(std::chrono::microseconds(static_cast<long>(stagger_ms) * 1000L * i)); // experiment: phase offset between replicas
        }
        const double t0 = nowSeconds();
        statuses[i] = ptb_render_with_progress(replicas[i], camera, &mine, x0, y0, w, h, replicas[i]->ctx->multi_image.as<float>(), stats != nullptr ? &stats[i] : &own_stats[i],
                                               progress != nullptr ? replicaProgress : nullptr, &sinks[i]);
        wall[i] = nowSeconds() - t0;
        if(statuses[i] != PTB_OK) {
            errors[i] = g_last_error; // thread-local: carried back to the caller's thread below
        }
    };
    std::vector<std::thread> workers;
    for(int i = 1; i < n; i++) {
        workers.emplace_back(render_share, i);
    }
    render_share(0);
    for(std::thread &t : workers) {
        t.join();
    }
    for(int i = 0; i < n; i++) {
        if(statuses[i] != PTB_OK) {
            return fail(statuses[i], "ptb_render_multi: replica " + std::to_string(i) + ": " + errors[i]);
        }
    }

    // gather on the first replica's device
    ptb_context *ctx = replicas[0]->ctx;
    std::lock_guard<std::mutex> lock(ctx->mutex);
    int status = useDevice(ctx);
    if(status != PTB_OK) {
        return status;
    }
    const bool device_io = (opts->flags & PTB_FLAG_DEVICE_IO) != 0U;
    float4 *d_out = reinterpret_cast<float4 *>(out_rgba);
    if(!device_io) {
        if((status = ctx->io_d.reserve(image_bytes)) != PTB_OK) {
            return status;
        }
        d_out = ctx->io_d.as<float4>();
    }
    int tile = opts->tile_size;
    if(tile <= 0) {
        tile = std::max(std::min(std::min(w, h) / 4, 32), 1);
    }
    GatherParams g{};
    g.n = n;
    g.w = w;
    g.h = h;
    g.tile = tile;
    g.tiles_x = (w + tile - 1) / tile;
    g.src[0] = ctx->multi_image.as<float4>();
    size_t staged = 0;
    for(int i = 1; i < n; i++) {
        ptb_context *peer = replicas[i]->ctx;
        int can_access = 0;
        if(peer->device != ctx->device && cudaDeviceCanAccessPeer(&can_access, ctx->device, peer->device) == cudaSuccess && can_access != 0) {
            const cudaError_t err = cudaDeviceEnablePeerAccess(peer->device, 0);
            if(err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) {
                can_access = 0;
            }
            cudaGetLastError();
        }
        if(peer->device == ctx->device || can_access != 0) {
            g.src[i] = peer->multi_image.as<float4>(); // same device, or mapped peer memory: the kernel reads it directly
        }
        else {
            if((status = ctx->multi_staging.reserve(static_cast<size_t>(n - 1) * image_bytes)) != PTB_OK) {
                return status;
            }
            char *slot = ctx->multi_staging.as<char>() + staged * image_bytes;
            PTB_CUDA(cudaMemcpyPeerAsync(slot, ctx->device, peer->multi_image.ptr, peer->device, image_bytes, ctx->stream));
            g.src[i] = reinterpret_cast<const float4 *>(slot);
            staged++;
        }
    }
    const int64_t pixels = static_cast<int64_t>(w) * h;
    const int grid = static_cast<int>(std::min<int64_t>((pixels + 255) / 256, static_cast<int64_t>(gridFor(ctx, 8))));
    gatherTilesKernel<<<grid, 256, 0, ctx->stream>>>(g, d_out);
    PTB_CUDA(cudaGetLastError());
    if(!device_io) {
        PTB_CUDA(cudaMemcpyAsync(out_rgba, d_out, image_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PTB_CUDA(cudaStreamSynchronize(ctx->stream));
    PTB_CUDA(cudaGetLastError());
    if(log_multi) {
        for(int i = 0; i < n; i++) {
            const ptb_render_stats &st = stats != nullptr ? stats[i] : own_stats[static_cast<size_t>(i)];
            std::fprintf(stderr, "[ptb] render_multi: replica %d on device %d: %.1f ms wall, %.1f ms on the device, %llu samples, %llu bounce iterations\n", i,
                         replicas[i]->ctx->device, wall[static_cast<size_t>(i)] * 1e3, st.device_ms_total, static_cast<unsigned long long>(st.samples),
                         static_cast<unsigned long long>(st.bounce_iterations));
        }
        std::fprintf(stderr, "[ptb] render_multi: whole call %.1f ms\n", (nowSeconds() - t_begin) * 1e3);
    }
    if(progress != nullptr) {
        uint64_t total = 0;
        for(int i = 0; i < n; i++) {
            const uint64_t t = all.total[i].load();
            total += t != ~0ULL ? t : 0;
        }
        progress(user, total, total);
    }
    return PTB_OK;
}

} // extern "C"
