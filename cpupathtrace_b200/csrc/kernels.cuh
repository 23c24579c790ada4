// Wavefront path-tracing kernels (sm_100a).
//
// One bounce iteration of impl::getSample (reference src/worker.cpp:44-138) for a whole pool of paths is five
// launches over structure-of-arrays path state in HBM:
//
//   generate        primary rays for a batch of (pixel, sample) pairs            Camera::shootRay
//   trace_closest   closest hit for every queued path                            Scene::getIntersection
//   shade           emission, Russian roulette, light sampling -> shadow slots,  getSample body, sampleLights,
//                   BSDF sampling -> next ray                                     BSDF::propagateRay/getSpectrum
//   trace_shadow    visibility of the queued shadow candidates                   the shadow query, worker.cpp:80-86
//   accumulate      adds the visible next-event contributions IN LIGHT ORDER, retires finished paths into the
//                   per-sample buffer and compacts survivors into the next queue (warp-aggregated atomics)
//
// and, once all samples of a pixel group exist,
//
//   resolve         processItem's per-pixel statistics (worker.cpp:158-317) run sequentially over the samples
//
// Queues hold path indices; their lengths live in device memory so no launch depends on a host round trip.
#ifndef PTB_KERNELS_CUH
#define PTB_KERNELS_CUH

#include "shading.cuh"
#include "traverse.cuh"

namespace ptb {

#ifndef PTB_FLAT_BLOCK
#define PTB_FLAT_BLOCK 128 // threads per CTA of the shade and accumulate kernels (one global atomic per CTA and trip)
#endif
    constexpr int kBlock = 128;
    constexpr int kFlatBlock = PTB_FLAT_BLOCK;
    static_assert(kBlock == kTraceBlock, "the traversal's shared-memory stack is laid out for kBlock threads per CTA");
#ifndef PTB_SHADE_MIN_BLOCKS
#define PTB_SHADE_MIN_BLOCKS 6 // 80 registers: +4 % shade throughput over the unconstrained 93 (8 CTAs / 64 registers spill and gain nothing)
#endif
#ifndef PTB_CLOSEST_MIN_BLOCKS
#define PTB_CLOSEST_MIN_BLOCKS 10 // closest-hit kernels of the path tracer: 48 registers, no spills (round 2: 8.1 vs 7.5 Grays/s at 12 CTAs / 40 registers)
#endif
#ifndef PTB_TRACE_MIN_BLOCKS
#define PTB_TRACE_MIN_BLOCKS 12 // resident 128-thread CTAs per SM the trace kernels are compiled for (register cap 65536 / (12 * 128) = 42)
#endif

    // Device-side counters.  Every counter sits on a 128-byte line of its own: they are the targets of the wavefront's
    // atomics, and atomics on one L2 line are served at about one per clock.  Round 1 kept all of them in 48 adjacent
    // bytes and issued three per warp trip from the shade kernel (queue append + two statistics): 12.6 M operations on one
    // line per 134 M-path launch = 6.6 ms at 1.9 GHz, which WAS the kernel's duration (6.7 ms, with 42 % of its stall
    // samples on the returning atomic, profiles/r02_ncu_shade_atomics.md).  Now: one line per counter, statistics kept in
    // registers until the end of the kernel, queue appends aggregated per thread block.
    constexpr int kCounterStride = 32; // 32-bit words between two counters
    enum CounterSlot : int {
        kCountQueueA = 0 * kCounterStride,
        kCountQueueB = 1 * kCounterStride,
        kCountShadow = 2 * kCounterStride,
        kCountFetchClosest = 3 * kCounterStride,
        kCountFetchShadow = 4 * kCounterStride,
        kCountSkippedShadows = 5 * kCounterStride,
        kCountVertices = 6 * kCounterStride,
        kCountRedo = 7 * kCounterStride,      // rays the certified closest-hit walk handed back
        kCountFetchRedo = 8 * kCounterStride, // fetch cursor of their re-trace on the reference tree
        kCounterSlots = 12 * kCounterStride
    };
    constexpr int kPerIterationCounters = kCountFetchRedo - kCountShadow + 1; // words zeroed before every bounce iteration (slots 2..8 and their padding)

    constexpr uint32_t kFlagTerminated = 1U;
    constexpr uint32_t kFlagXorshift = 2U;

    // radiance[i].w, written by the shade kernel for the accumulate kernel: everything accumulate needs to know about
    // path i sits in the 16 bytes it has to read anyway (round 1 read shadow_count[i] and state[i] on top: two more
    // dependent 32-byte sectors per path in a kernel that is a pure latency chain).
    constexpr uint32_t kRadianceShadowMask = 0xFFU;   // shadow candidates stored for the current vertex
    constexpr uint32_t kRadianceTerminated = 0x100U;  // the path ends at this vertex
    constexpr uint32_t kRadianceCollected = 0x200U;   // path_length > 0 (worker.cpp:141-143)

    // Structure-of-arrays path pool.  Every array has `capacity` entries (shadow arrays capacity * shadow_stride).
    struct PathPool {
        float4 *ray_o;       // origin xyz
        float4 *ray_d;       // direction xyz
        float2 *hit;         // (t, slot bits)
        float4 *throughput;  // sample_spectrum
        float4 *radiance;    // out_spectrum rgb; w = bookkeeping word of the accumulate kernel (kRadiance* below)
        double *divisor;     // sample_divisor
        double *bounce_pd;   // sample_bounce_pd
        float *contribution; // contribution_unweighted
        uint32_t *state;     // path_length << 8 | flags
        uint64_t *rng;       // xorshift state or counter key
        uint32_t *dest;      // index into the per-sample buffer
        float4 *shadow_o;    // (origin, limit)
        float4 *shadow_d;    // (direction, 0)
        float4 *shadow_c;    // (contribution rgb, visible flag)
        uint32_t shadow_stride;
        uint32_t capacity;
    };

    struct RenderParams {
        ptb_camera camera;
        int32_t image_width;
        int32_t image_height;
        float epsilon;
        int32_t max_depth;
        uint32_t rng_xorshift;
        uint32_t any_hit_shadows;
        uint32_t skip_null_shadows;
        uint64_t seed;
    };

    PTB_DEV uint32_t laneId() {
        return threadIdx.x & 31U;
    }

    // Appends one element per participating lane with a single atomic per group of lanes that arrive together
    // (normally the whole warp; see "Warp collectives and convergence" in traverse.cuh for why that is not assumed).
    PTB_DEV uint32_t warpAppend(uint32_t *counter, bool participate) {
        const uint32_t mask = __ballot_sync(__activemask(), participate);
        if(!participate) {
            return 0U;
        }
        const uint32_t leader = __ffs(mask) - 1U;
        uint32_t base = 0U;
        if(laneId() == leader) {
            base = atomicAdd(counter, __popc(mask));
        }
        base = __shfl_sync(mask, base, leader);
        return base + __popc(mask & ((1U << laneId()) - 1U));
    }

    // The alpha lane of the throughput and radiance spectra is never observable: impl::getSample overwrites the
    // output alpha with the "collected" flag (worker.cpp:141-143) and getContribution reads rgb only (worker.cpp:12-14).
    // It is therefore not carried through the pool (constant lanes let the compiler drop the alpha arithmetic).
    template<typename RNG>
    PTB_DEV void loadPath(const PathPool &pool, uint32_t i, PathRegs<RNG> &p, uint32_t &flags) {
        const float4 o = sld(&pool.ray_o[i]);
        const float4 d = sld(&pool.ray_d[i]);
        p.ray_o = mk3(o.x, o.y, o.z);
        p.ray_d = mk3(d.x, d.y, d.z);
        const float4 thr = sld(&pool.throughput[i]);
        const float4 rad = sld(&pool.radiance[i]);
        p.throughput = V4{thr.x, thr.y, thr.z, 1.0F};
        p.radiance = V4{rad.x, rad.y, rad.z, 0.0F};
        p.divisor = sld(&pool.divisor[i]);
        p.bounce_pd = sld(&pool.bounce_pd[i]);
        p.contribution_unweighted = sld(&pool.contribution[i]);
        const uint32_t st = sld(&pool.state[i]);
        p.path_length = static_cast<int>(st >> 8);
        flags = st & 0xFFU;
        p.rng.state = sld(&pool.rng[i]);
        p.rng.counter = 0U;
    }

    template<typename RNG>
    PTB_DEV void storePath(const PathPool &pool, uint32_t i, const PathRegs<RNG> &p, uint32_t flags, uint32_t radiance_word) {
        sst(&pool.ray_o[i], make_float4(p.ray_o.x, p.ray_o.y, p.ray_o.z, 0.0F));
        sst(&pool.ray_d[i], make_float4(p.ray_d.x, p.ray_d.y, p.ray_d.z, 0.0F));
        sst(&pool.throughput[i], make_float4(p.throughput.x, p.throughput.y, p.throughput.z, 1.0F));
        sst(&pool.radiance[i], make_float4(p.radiance.x, p.radiance.y, p.radiance.z, __uint_as_float(radiance_word)));
        sst(&pool.divisor[i], p.divisor);
        sst(&pool.bounce_pd[i], p.bounce_pd);
        sst(&pool.contribution[i], p.contribution_unweighted);
        sst(&pool.state[i], (static_cast<uint32_t>(p.path_length) << 8) | (flags & 0xFFU));
        sst(&pool.rng[i], p.rng.state);
    }

    // ------------------------------------------------------------------------------------------------ generate

    // Where new paths come from.  Work item g of a call is
    //   frame mode     : sample = g / n_pixels of pixel pixel_list[g % n_pixels] (x | y << 16), keyed on (seed, x, y, sample)
    //   validation mode: the explicit (pixel, seed) pair g, engine RandomEngine(seeds[g])
    // and its result goes to samples[g].
    struct PathSource {
        const uint32_t *pixel_list;
        const uint32_t *active; // adaptive rounds: work item g renders pixel pixel_list[active[g % n_pixels]] (n_pixels = active pixels); else null
        const int32_t *pixels;
        const uint64_t *seeds;
        uint32_t n_pixels;
        uint32_t explicit_samples;
        uint32_t sample_base; // index of the first sample of this round
        unsigned long long total;
    };

    // Camera::shootRay for work item g into pool slot i (worker.cpp:27-34, 168-170)
    template<typename RNG>
    PTB_DEV void generatePath(const PathPool &pool, const RenderParams &params, const PathSource &src, uint32_t i, unsigned long long g) {
        int px;
        int py;
        uint64_t key;
        if(src.explicit_samples != 0U) {
            px = src.pixels[2 * g];
            py = src.pixels[2 * g + 1];
            key = src.seeds[g];
        }
        else {
            const uint32_t sample = src.sample_base + static_cast<uint32_t>(g / src.n_pixels);
            const uint32_t slot = static_cast<uint32_t>(g % src.n_pixels);
            const uint32_t packed = src.pixel_list[src.active != nullptr ? src.active[slot] : slot];
            px = static_cast<int>(packed & 0xFFFFU);
            py = static_cast<int>(packed >> 16);
            key = counterKey(params.seed, static_cast<uint32_t>(px), static_cast<uint32_t>(py), sample);
        }

        PathRegs<RNG> p;
        initPath(p);
        p.rng.counter = 0U;
        p.rng.state = RNG::kXorshift ? xorshiftSeed(key) : key;

        float x_camera;
        float y_camera;
        pixelToCamera(px, py, params.image_width, params.image_height, x_camera, y_camera);
        shootRay(params.camera, x_camera, y_camera, 1.0F / static_cast<float>(params.image_width), 1.0F / static_cast<float>(params.image_height), p.rng,
                 p.ray_o, p.ray_d);

        storePath(pool, i, p, RNG::kXorshift ? kFlagXorshift : 0U, 0U);
        pool.dest[i] = static_cast<uint32_t>(g);
    }

    // Fills pool slots [0, count) with work items [0, count) and queues them; later work items are started by the
    // accumulate kernel in the slots of retired paths (path regeneration), so the pool stays full until the call's
    // work runs out and only the very last bounce iterations of a call run on a thin queue.
    template<typename RNG>
    __global__ void __launch_bounds__(kBlock) generateKernel(PathPool pool, RenderParams params, PathSource src, uint32_t count, uint32_t *__restrict__ queue,
                                                             uint32_t *__restrict__ counters, int queue_slot) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i == 0U) {
            counters[queue_slot] = count;
        }
        if(i >= count) {
            return;
        }
        generatePath<RNG>(pool, params, src, i, i);
        queue[i] = i;
    }

    // ------------------------------------------------------------------------------------------------ trace

    // Any-hit queries walk the occlusion hierarchy when the scene has one (visibility is order-independent).
    PTB_DEV DeviceScene occlusionView(const DeviceScene &scene) {
        DeviceScene view = scene;
        if(scene.occ_nodes != nullptr) {
            view.nodes = scene.occ_nodes;
            view.root_ref = scene.occ_root_ref;
        }
        return view;
    }

    // Persistent warps over the device-side queue; scheduling by warp votes, see warpTrace in traverse.cuh.
    // MODE = kTraceClosest walks the reference tree; MODE = kTraceCertified walks the SAH hierarchy and appends the
    // paths whose result carries no certificate to `redo_queue` (length counters[kCountRedo]), which a second launch
    // of the kTraceClosest instantiation then serves (queue = redo_queue, queue_slot = kCountRedo).
    template<int MODE, bool COUNT>
    __global__ void __launch_bounds__(kBlock, PTB_CLOSEST_MIN_BLOCKS) traceClosestKernel(DeviceScene scene, VoteParams vote, PathPool pool, const uint32_t *__restrict__ queue,
                                                                 uint32_t *__restrict__ counters, int queue_slot, int cursor_slot, uint32_t *__restrict__ redo_queue,
                                                                 VisitCounters *visits, const __grid_constant__ ptb_guard::CertGuard guard, int guarded) {
        const uint32_t count = counters[queue_slot];
        warpTrace<MODE, COUNT>(
          MODE == kTraceCertified ? occlusionView(scene) : scene, vote, &counters[cursor_slot], count,
          [&](uint32_t k, V3 &o, V3 &d, float &limit) {
              const uint32_t i = sld_t(&queue[k]);
              const float4 ro = sld_t(&pool.ray_o[i]);
              const float4 rd = sld_t(&pool.ray_d[i]);
              o = mk3(ro.x, ro.y, ro.z);
              d = mk3(rd.x, rd.y, rd.z);
              limit = 0.0F;
              return i;
          },
          [&](uint32_t i, const Hit &h, bool certain) {
              if(MODE == kTraceCertified && !certain) {
                  redo_queue[atomicAdd(&counters[kCountRedo], 1U)] = i;
              }
              else {
                  sst_t(&pool.hit[i], make_float2(h.t, __int_as_float(h.slot)));
              }
          },
          visits, (MODE == kTraceCertified && guarded != 0) ? &guard : nullptr);
    }

    template<bool COUNT>
    __global__ void __launch_bounds__(kBlock, PTB_TRACE_MIN_BLOCKS) traceShadowKernel(DeviceScene scene, VoteParams vote, PathPool pool, const uint32_t *__restrict__ shadow_queue,
                                                                uint32_t *__restrict__ counters, uint32_t any_hit, VisitCounters *visits) {
        const uint32_t count = counters[kCountShadow];
        auto fetch = [&](uint32_t k, V3 &o, V3 &d, float &limit) {
            const uint32_t slot = sld_t(&shadow_queue[k]);
            const float4 so = sld_t(&pool.shadow_o[slot]);
            const float4 sd = sld_t(&pool.shadow_d[slot]);
            o = mk3(so.x, so.y, so.z);
            d = mk3(sd.x, sd.y, sd.z);
            limit = so.w;
            return slot;
        };
        if(any_hit != 0U) {
            warpTrace<kTraceAnyHit, COUNT>(
              occlusionView(scene), vote, &counters[kCountFetchShadow], count, fetch,
              [&](uint32_t slot, const Hit &h, bool) { sst_t(&pool.shadow_c[slot].w, h.slot < 0 ? 1.0F : 0.0F); }, visits);
        }
        else {
            // the reference's full closest-hit query (worker.cpp:84-86): unoccluded iff t < 0 or t >= |to_light| - epsilon
            warpTrace<kTraceClosest, COUNT>(
              scene, vote, &counters[kCountFetchShadow], count, fetch,
              [&](uint32_t slot, const Hit &h, bool) {
                  const float limit = pool.shadow_o[slot].w;
                  pool.shadow_c[slot].w = (h.t < 0.0F || h.t >= limit) ? 1.0F : 0.0F;
              },
              visits);
        }
    }

    // ------------------------------------------------------------------------------------------------ shade

    template<typename RNG>
    __global__ void __launch_bounds__(kFlatBlock, PTB_SHADE_MIN_BLOCKS) shadeKernel(DeviceScene scene, PathPool pool, RenderParams params, const uint32_t *__restrict__ queue,
                                                          uint32_t *__restrict__ counters, int queue_slot, uint32_t *__restrict__ shadow_queue) {
        const uint32_t count = counters[queue_slot];
        const uint32_t stride = gridDim.x * blockDim.x;
        // Block-level aggregation of the shadow-queue append (two trips in flight: parity p = trip & 1)
        __shared__ uint32_t block_total[2];
        __shared__ uint32_t block_base[2];
        if(threadIdx.x == 0U) {
            block_total[0] = 0U;
            block_total[1] = 0U;
        }
        __syncthreads();
        uint32_t stat_vertices = 0U;
        uint32_t stat_skipped = 0U;
        uint32_t parity = 0U;
        // whole blocks iterate together (the barriers below need every thread of the block in every trip)
        for(uint32_t first = blockIdx.x * blockDim.x; first < count; first += stride, parity ^= 1U) {
            const uint32_t k = first + threadIdx.x;
            const bool active = k < count;
            uint32_t i = 0U;
            PathRegs<RNG> p;
            uint32_t flags = 0U;
            bool hit_surface = false;
            float t = -1.0F;
            uint32_t slot = 0U;
            if(active) {
                i = sld(&queue[k]);
                loadPath(pool, i, p, flags);
                const float2 h = sld(&pool.hit[i]);
                t = h.x;
                slot = static_cast<uint32_t>(__float_as_int(h.y));
                hit_surface = !(t < 0.0F);
            }

            uint32_t n_shadow = 0U;
            bool continues = false;
            const uint32_t shadow_base = i * pool.shadow_stride;
            if(hit_surface) {
                stat_vertices++;
                continues = shadeVertex(scene, params.epsilon, params.max_depth, p, t, slot, [&](const ShadowCandidate &c, bool is_null) {
                    // candidates without weight are traced only to reproduce the reference's ray count
                    if(is_null && params.skip_null_shadows != 0U) {
                        stat_skipped++;
                        return;
                    }
                    if(n_shadow < pool.shadow_stride) {
                        const uint32_t s = shadow_base + n_shadow;
                        sst(&pool.shadow_o[s], make_float4(c.o.x, c.o.y, c.o.z, c.limit));
                        sst(&pool.shadow_d[s], make_float4(c.d.x, c.d.y, c.d.z, 0.0F));
                        sst(&pool.shadow_c[s], make_float4(c.contribution.x, c.contribution.y, c.contribution.z, 0.0F));
                        n_shadow++;
                    }
                });
            }

            if(active) {
                uint32_t word = n_shadow | (p.path_length > 0 ? kRadianceCollected : 0U);
                if(!continues) {
                    flags |= kFlagTerminated;
                    word |= kRadianceTerminated;
                }
                storePath(pool, i, p, flags, word);
            }

            // queue the shadow rays: exclusive scan of the per-lane counts over the lanes that arrive together (normally
            // the whole warp; one ballot per bit of the count stays correct for any group of lanes), one SHARED-memory
            // atomic per group, one global atomic per block and trip.
            const uint32_t present = __activemask();
            const uint32_t below_me = (1U << laneId()) - 1U;
            uint32_t exclusive = 0U;
            uint32_t group_total = 0U;
            for(uint32_t bit = 0U; (pool.shadow_stride >> bit) != 0U; bit++) {
                const uint32_t votes = __ballot_sync(present, ((n_shadow >> bit) & 1U) != 0U);
                exclusive += static_cast<uint32_t>(__popc(votes & below_me)) << bit;
                group_total += static_cast<uint32_t>(__popc(votes)) << bit;
            }
            if(group_total != 0U) {
                const uint32_t leader = static_cast<uint32_t>(__ffs(static_cast<int>(present))) - 1U;
                uint32_t group_base = 0U;
                if(laneId() == leader) {
                    group_base = atomicAdd(&block_total[parity], group_total);
                }
                exclusive += __shfl_sync(present, group_base, static_cast<int>(leader));
            }
            __syncthreads();
            if(threadIdx.x == 0U) {
                const uint32_t total = block_total[parity];
                block_base[parity] = total != 0U ? atomicAdd(&counters[kCountShadow], total) : 0U;
                block_total[parity ^ 1U] = 0U; // the next trip's accumulator; nobody touches it between these two barriers
            }
            __syncthreads();
            const uint32_t base = block_base[parity] + exclusive;
            for(uint32_t j = 0U; j < n_shadow; j++) {
                sst(&shadow_queue[base + j], shadow_base + j);
            }
        }

        // statistics: kept in registers over the whole kernel, one atomic per warp (group) at the end
        const uint32_t present = __activemask();
        const uint32_t vertices = __reduce_add_sync(present, stat_vertices);
        const uint32_t skipped = __reduce_add_sync(present, stat_skipped);
        if(laneId() == static_cast<uint32_t>(__ffs(static_cast<int>(present))) - 1U) {
            if(vertices != 0U) {
                atomicAdd(&counters[kCountVertices], vertices);
            }
            if(skipped != 0U) {
                atomicAdd(&counters[kCountSkippedShadows], skipped);
            }
        }
    }

    // ------------------------------------------------------------------------------------------------ accumulate

    template<typename RNG>
    __global__ void __launch_bounds__(kFlatBlock) accumulateKernel(PathPool pool, RenderParams params, PathSource src, const uint32_t *__restrict__ queue,
                                                               uint32_t *__restrict__ counters, int queue_slot, uint32_t *__restrict__ next_queue, int next_slot,
                                                               float4 *__restrict__ samples, unsigned long long *__restrict__ work_cursor) {
        const uint32_t count = counters[queue_slot];
        const uint32_t stride = gridDim.x * blockDim.x;
        // Block-level aggregation of the two reservations a trip needs (work items for the retired slots, places in the
        // next queue): shared-memory atomics per group of lanes, then thread 0 issues the two global atomics of the block.
        __shared__ uint32_t block_retired[2];
        __shared__ uint32_t block_survivors[2];
        __shared__ unsigned long long block_work_base[2];
        __shared__ uint32_t block_queue_base[2];
        if(threadIdx.x == 0U) {
            block_retired[0] = block_retired[1] = 0U;
            block_survivors[0] = block_survivors[1] = 0U;
        }
        __syncthreads();
        uint32_t parity = 0U;
        for(uint32_t first = blockIdx.x * blockDim.x; first < count; first += stride, parity ^= 1U) {
            const uint32_t k = first + threadIdx.x;
            const bool active = k < count;
            bool survives = false;
            bool retired = false;
            uint32_t i = 0U;
            if(active) {
                i = queue[k];
                float4 radiance = sld(&pool.radiance[i]);
                const uint32_t word = __float_as_uint(radiance.w);
                const uint32_t n_shadow = word & kRadianceShadowMask;
                if(pool.shadow_stride == 2U) {
                    // the common case (two emissive-object samples per vertex, no point lights): both candidates of the
                    // path are one aligned 32-byte record, loaded together with the radiance instead of after it
                    float4 c0;
                    float4 c1;
                    ld256cg(pool.shadow_c + 2U * static_cast<size_t>(i), c0, c1);
                    if(n_shadow > 0U && c0.w != 0.0F) {
                        radiance.x = radiance.x + c0.x;
                        radiance.y = radiance.y + c0.y;
                        radiance.z = radiance.z + c0.z;
                    }
                    if(n_shadow > 1U && c1.w != 0.0F) {
                        radiance.x = radiance.x + c1.x;
                        radiance.y = radiance.y + c1.y;
                        radiance.z = radiance.z + c1.z;
                    }
                }
                else {
                    const size_t base = static_cast<size_t>(i) * pool.shadow_stride;
                    for(uint32_t j = 0U; j < n_shadow; j++) {
                        const float4 c = sld(&pool.shadow_c[base + j]);
                        if(c.w != 0.0F) {
                            radiance.x = radiance.x + c.x;
                            radiance.y = radiance.y + c.y;
                            radiance.z = radiance.z + c.z;
                        }
                    }
                }
                if((word & kRadianceTerminated) != 0U) {
                    // out_color[3] = sample_collected ? 1 : 0 (worker.cpp:141-143)
                    samples[pool.dest[i]] = make_float4(radiance.x, radiance.y, radiance.z, (word & kRadianceCollected) != 0U ? 1.0F : 0.0F);
                    retired = true;
                }
                else {
                    if(n_shadow > 0U) {
                        pool.radiance[i] = radiance; // the shade kernel rewrites w at the next vertex
                    }
                    survives = true;
                }
            }

            // ranks inside the block: every group of lanes that arrives together adds its counts to the block's shared
            // accumulators and learns where its retired lanes / survivors start
            const uint32_t present = __activemask();
            const uint32_t below_me = (1U << laneId()) - 1U;
            const uint32_t retired_mask = __ballot_sync(present, retired);
            const uint32_t survivor_mask = __ballot_sync(present, survives);
            const uint32_t leader = static_cast<uint32_t>(__ffs(static_cast<int>(present))) - 1U;
            uint32_t retired_rank = 0U;
            uint32_t survivor_rank = 0U;
            if(laneId() == leader) {
                if(retired_mask != 0U) {
                    retired_rank = atomicAdd(&block_retired[parity], static_cast<uint32_t>(__popc(retired_mask)));
                }
                if(survivor_mask != 0U) {
                    survivor_rank = atomicAdd(&block_survivors[parity], static_cast<uint32_t>(__popc(survivor_mask)));
                }
            }
            retired_rank = __shfl_sync(present, retired_rank, static_cast<int>(leader)) + static_cast<uint32_t>(__popc(retired_mask & below_me));
            survivor_rank = __shfl_sync(present, survivor_rank, static_cast<int>(leader)) + static_cast<uint32_t>(__popc(survivor_mask & below_me));
            __syncthreads();
            if(threadIdx.x == 0U) {
                // path regeneration: the retired slots of the block take the next unstarted work items of the call; those
                // that still get one join the survivors in the next queue (behind them)
                const uint32_t n_retired = block_retired[parity];
                const uint32_t n_survivors = block_survivors[parity];
                unsigned long long work_base = src.total;
                uint32_t regenerated = 0U;
                if(n_retired != 0U) {
                    work_base = atomicAdd(work_cursor, static_cast<unsigned long long>(n_retired));
                    if(work_base < src.total) {
                        const unsigned long long left = src.total - work_base;
                        regenerated = left < static_cast<unsigned long long>(n_retired) ? static_cast<uint32_t>(left) : n_retired;
                    }
                }
                block_work_base[parity] = work_base;
                block_queue_base[parity] = (n_survivors + regenerated) != 0U ? atomicAdd(&counters[next_slot], n_survivors + regenerated) : 0U;
                block_retired[parity ^ 1U] = 0U;
                block_survivors[parity ^ 1U] = 0U;
            }
            __syncthreads();
            const uint32_t queue_base = block_queue_base[parity];
            if(retired) {
                const unsigned long long g = block_work_base[parity] + static_cast<unsigned long long>(retired_rank);
                if(g < src.total) {
                    generatePath<RNG>(pool, params, src, i, g);
                    next_queue[queue_base + block_survivors[parity] + retired_rank] = i;
                }
            }
            else if(survives) {
                next_queue[queue_base + survivor_rank] = i;
            }
        }
    }

#if !defined(PTB_FAST_MATH) && !defined(PTB_FAST_TU_EXACT) // the remaining kernels exist in the exact build only

    // ------------------------------------------------------------------------------------------------ resolve

    struct ResolveParams {
        int32_t min_sample_count;
        int32_t max_sample_count;
        uint32_t n_pixels; // pixels in the group; samples[s * n_pixels + q]
        int32_t rect_x0;
        int32_t rect_y0;
        int32_t rect_w;
    };

    constexpr int kMaxCandidates = 8; // processItem opens a candidate every candidate_batch_count batches: at most 6 for any legal option set
                                      // (candidate_batch_count >= max(min, max / 4) / stats, i.e. at most ~4 candidates + remainder); the host asserts it

    PTB_DEV V4 sub4(V4 a, V4 b) {
        return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
    }

    // The constants processItem derives from RenderOptions (worker.cpp:158-164)
    struct ResolveConsts {
        int min_samples;
        int max_samples;
        int stats_sample_count;
        int candidate_batch_count;
        int check_sample_count;
    };

    __host__ __device__ inline ResolveConsts resolveConsts(int min_samples, int max_samples) {
        ResolveConsts c;
        c.min_samples = min_samples;
        c.max_samples = max_samples;
        const int quarter = min_samples / 4;
        c.stats_sample_count = quarter < 1 ? 1 : (quarter > 64 ? 64 : quarter);
        const int a = max_samples / 4;
        const int b = (min_samples > a ? min_samples : a) / c.stats_sample_count;
        c.candidate_batch_count = b > 2 ? b : 2;
        const int half = min_samples / 2;
        const int eighth = (max_samples - min_samples) / 8;
        const int lo = half > eighth ? half : eighth;
        const int floor8 = c.stats_sample_count > 8 ? c.stats_sample_count : 8;
        int checks = lo > floor8 ? lo : floor8;
        checks = checks < 1024 ? checks : 1024;
        c.check_sample_count = checks / c.stats_sample_count;
        return c;
    }

    // The per-pixel state of processItem's sampling loop (worker.cpp:172-260), resumable: the one-shot resolve keeps it
    // in registers / local memory, the adaptive path parks it in HBM between rounds of samples.
    struct PixelStats {
        V4 pixel_value;
        V4 contribution_mean;
        V4 contribution_m2;
        V4 sample_aggregate;
        V4 candidate_mean;
        V4 candidate_m2;
        V4 candidate_means[kMaxCandidates];
        V4 candidate_m2s[kMaxCandidates];
        int candidate_counts[kMaxCandidates];
        int collected_sample_count;
        int contribution_count;
        int stats_sample_index;
        int candidate_count;
        int n_candidates;
        int remaining_checks;
        int accepted_candidate;
        int next_sample; // pixel_sample of the reference's loop: samples consumed so far
    };

    PTB_DEV void pixelBegin(PixelStats &st, const ResolveConsts &c) {
        const V4 zero = V4{0.0F, 0.0F, 0.0F, 0.0F};
        st.pixel_value = zero;
        st.contribution_mean = zero;
        st.contribution_m2 = zero;
        st.sample_aggregate = zero;
        st.candidate_mean = zero;
        st.candidate_m2 = zero;
        st.collected_sample_count = 0;
        st.contribution_count = 0;
        st.stats_sample_index = 0;
        st.candidate_count = 0;
        st.n_candidates = 0;
        st.remaining_checks = c.check_sample_count;
        st.accepted_candidate = 0;
        st.next_sample = 0;
    }

    // One iteration of the reference's per-pixel loop with the sample `raw` (rgb, collected flag).  Returns true when the
    // loop ends with this sample (the adaptive acceptance test fired: `break` at worker.cpp:251).
    PTB_DEV bool pixelAdd(PixelStats &st, const ResolveConsts &c, float4 raw) {
        const V4 zero = V4{0.0F, 0.0F, 0.0F, 0.0F};
        st.next_sample++;
        if(raw.w == 0.0F) {
            return false; // sample not collected: the primary ray missed everything
        }
        const V4 color = V4{raw.x, raw.y, raw.z, 1.0F};

        st.contribution_count++;
        st.stats_sample_index++;
        st.sample_aggregate = st.sample_aggregate + color;

        if(st.stats_sample_index == c.stats_sample_count) {
            st.sample_aggregate = st.sample_aggregate / static_cast<float>(c.stats_sample_count);

            const V4 delta = sub4(st.sample_aggregate, st.contribution_mean);
            st.contribution_mean = st.contribution_mean + delta / static_cast<float>(st.contribution_count / c.stats_sample_count);
            const V4 delta2 = sub4(st.sample_aggregate, st.contribution_mean);
            st.contribution_m2 = st.contribution_m2 + delta * delta2;

            if(st.candidate_count == c.candidate_batch_count) {
                if(st.n_candidates < kMaxCandidates) {
                    st.candidate_means[st.n_candidates] = st.candidate_mean;
                    st.candidate_m2s[st.n_candidates] = st.candidate_m2;
                    st.candidate_counts[st.n_candidates] = st.candidate_count;
                    st.n_candidates++;
                }
                st.candidate_mean = zero;
                st.candidate_m2 = zero;
                st.candidate_count = 0;
            }

            st.candidate_count++;
            const V4 candidate_delta = sub4(st.sample_aggregate, st.candidate_mean);
            st.candidate_mean = st.candidate_mean + candidate_delta / static_cast<float>(st.candidate_count);
            const V4 candidate_delta2 = sub4(st.sample_aggregate, st.candidate_mean);
            st.candidate_m2 = st.candidate_m2 + candidate_delta * candidate_delta2;

            st.stats_sample_index = 0;
            st.sample_aggregate = zero;
        }

        st.pixel_value = st.pixel_value + color;
        st.collected_sample_count++;

        if(st.stats_sample_index == 0 && st.collected_sample_count >= max(c.min_samples, 2)) {
            bool passed_check = false;
            if(st.contribution_count / c.stats_sample_count >= 2) {
                const V4 m2_weighted = st.contribution_m2 / static_cast<float>(st.contribution_count / c.stats_sample_count - 1);
                const float stddev = sqrtf((m2_weighted.x + m2_weighted.y) + m2_weighted.z);
                const float mean_contribution = ((st.contribution_mean.x + st.contribution_mean.y) + st.contribution_mean.z) / 3.0F;
                // `stddev / (3 * 3 * getContribution(mean) + 1E-5) < 0.2F` is evaluated in double (worker.cpp:243)
                const double relative = static_cast<double>(stddev) / (static_cast<double>(9.0F * mean_contribution) + 1E-5);
                if(stddev < 1E-4F || relative < static_cast<double>(0.2F)) {
                    passed_check = true;
                    st.remaining_checks--;
                    if(st.remaining_checks <= 0) {
                        st.accepted_candidate = 1;
                        return true;
                    }
                }
            }
            if(!passed_check) {
                st.remaining_checks = c.check_sample_count;
            }
        }
        return false;
    }

    // What processItem writes for the pixel once its loop has ended (worker.cpp:262-317)
    PTB_DEV V4 pixelFinish(PixelStats &st, const ResolveConsts &c) {
        V4 pixel_value = st.pixel_value;
        if(st.collected_sample_count > 0) {
            pixel_value = pixel_value * (1.0F / static_cast<float>(st.collected_sample_count));
        }

        if(st.candidate_count > 0 && st.n_candidates < kMaxCandidates) {
            st.candidate_means[st.n_candidates] = st.candidate_mean;
            st.candidate_m2s[st.n_candidates] = st.candidate_m2;
            st.candidate_counts[st.n_candidates] = st.candidate_count;
            st.n_candidates++;
        }

        if(st.accepted_candidate == 0) {
            V4 colors[kMaxCandidates];
            float stddevs[kMaxCandidates];
            int n = 0;
            const int needed = max((c.candidate_batch_count * 3) / 4, 2);
            for(int k = 0; k < st.n_candidates; k++) {
                if(st.candidate_counts[k] < needed) {
                    continue;
                }
                const V4 m2_weighted = st.candidate_m2s[k] / static_cast<float>(st.candidate_counts[k]);
                const float stddev = sqrtf((m2_weighted.x + m2_weighted.y) + m2_weighted.z);
                // stable insertion (std::sort on fewer than 16 elements is a stable insertion sort in libstdc++)
                int at = n;
                while(at > 0 && stddev < stddevs[at - 1]) {
                    stddevs[at] = stddevs[at - 1];
                    colors[at] = colors[at - 1];
                    at--;
                }
                stddevs[at] = stddev;
                colors[at] = st.candidate_means[k];
                n++;
            }
            if(n > 0) {
                pixel_value = colors[0];
                float stddev = stddevs[0];
                for(int k = 1; k < n; k++) {
                    const float other = stddevs[k];
                    if(other < stdmax(stddev + 0.005F, stddev * 1.01F)) {
                        pixel_value = pixel_value + sub4(colors[k], pixel_value) / static_cast<float>(k + 1);
                        stddev = other;
                    }
                    else {
                        break;
                    }
                }
            }
        }
        return pixel_value;
    }

    PTB_DEV void storePixel(const ResolveParams &rp, const uint32_t *pixel_list, uint32_t q, V4 value, float4 *out) {
        const uint32_t packed = pixel_list[q];
        const int px = static_cast<int>(packed & 0xFFFFU) - rp.rect_x0;
        const int py = static_cast<int>(packed >> 16) - rp.rect_y0;
        out[static_cast<size_t>(py) * rp.rect_w + px] = f4(value);
    }

    // processItem's per-pixel loop (worker.cpp:172-319) over the already-computed samples of one pixel: all max_sample_count
    // samples exist (the fixed-spp path, min == max, where the loop cannot end early).
    __global__ void __launch_bounds__(kBlock) resolveKernel(ResolveParams rp, const float4 *__restrict__ samples, const uint32_t *__restrict__ pixel_list,
                                                            float4 *__restrict__ out) {
        const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
        if(q >= rp.n_pixels) {
            return;
        }
        const ResolveConsts c = resolveConsts(rp.min_sample_count, rp.max_sample_count);
        PixelStats st;
        pixelBegin(st, c);
        for(int pixel_sample = 0; pixel_sample < c.max_samples; pixel_sample++) {
            if(pixelAdd(st, c, samples[static_cast<size_t>(pixel_sample) * rp.n_pixels + q])) {
                break;
            }
        }
        storePixel(rp, pixel_list, q, pixelFinish(st, c), out);
    }

    // ---- adaptive sampling (min != max) in rounds: only the pixels whose loop has not ended get more samples.
    //
    // Round r traces samples [first, first + count) of every pixel still active and stores them as
    // samples[(s - first) * n_active + a] for the a-th active pixel; this kernel feeds them to the pixel's parked state in
    // sample order -- exactly the sequence the reference's loop sees -- and keeps the pixel active iff the loop neither
    // ended (acceptance) nor ran out of samples.  Samples of a round traced beyond the end of a pixel's loop are ignored,
    // as the reference never draws them.
    __global__ void __launch_bounds__(kBlock) adaptiveInitKernel(ResolveConsts c, uint32_t n_pixels, PixelStats *__restrict__ states, uint32_t *__restrict__ active) {
        const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
        if(q >= n_pixels) {
            return;
        }
        PixelStats st;
        pixelBegin(st, c);
        for(int k = 0; k < kMaxCandidates; k++) {
            st.candidate_means[k] = V4{0.0F, 0.0F, 0.0F, 0.0F};
            st.candidate_m2s[k] = V4{0.0F, 0.0F, 0.0F, 0.0F};
            st.candidate_counts[k] = 0;
        }
        states[q] = st;
        active[q] = q;
    }

    __global__ void __launch_bounds__(kBlock) adaptiveAdvanceKernel(ResolveConsts c, const float4 *__restrict__ samples, const uint32_t *__restrict__ active, uint32_t n_active,
                                                                    int round_count, PixelStats *__restrict__ states, uint32_t *__restrict__ next_active,
                                                                    uint32_t *__restrict__ next_count) {
        const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
        if(a >= n_active) {
            return;
        }
        const uint32_t q = active[a];
        PixelStats st = states[q];
        bool ended = false;
        for(int s = 0; s < round_count && !ended; s++) {
            ended = pixelAdd(st, c, samples[static_cast<size_t>(s) * n_active + a]);
        }
        states[q] = st;
        if(!ended && st.next_sample < c.max_samples) {
            next_active[atomicAdd(next_count, 1U)] = q;
        }
    }

    __global__ void __launch_bounds__(kBlock) adaptiveFinishKernel(ResolveParams rp, PixelStats *__restrict__ states, const uint32_t *__restrict__ pixel_list,
                                                                   float4 *__restrict__ out, unsigned long long *__restrict__ samples_used) {
        const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
        if(q >= rp.n_pixels) {
            return;
        }
        const ResolveConsts c = resolveConsts(rp.min_sample_count, rp.max_sample_count);
        PixelStats st = states[q];
        storePixel(rp, pixel_list, q, pixelFinish(st, c), out);
        const unsigned long long used = __reduce_add_sync(__activemask(), static_cast<uint32_t>(st.next_sample));
        if(laneId() == static_cast<uint32_t>(__ffs(static_cast<int>(__activemask()))) - 1U) {
            atomicAdd(samples_used, used);
        }
    }

    // ------------------------------------------------------------------------------------------------ unit kernels

    // Sort key of a ray for batch queries: direction octant (3 bits) over a 27-bit Morton code of the origin inside the
    // scene's root box.  Batches of incoherent rays are traced in key order (a pure reordering: results are written by
    // ray number), so that the warps of the persistent traversal kernel hold rays that start in the same region and head
    // the same way and their node fetches hit the same cache lines (configs[2]: 16 Mi-triangle soup, HBM-bound).
    PTB_DEV uint32_t spread3(uint32_t v) { // 9 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FFU;
        v = (v | (v << 8)) & 0x0300F00FU;
        v = (v | (v << 4)) & 0x030C30C3U;
        v = (v | (v << 2)) & 0x09249249U;
        return v;
    }

    // dir_bits = b: the top 3 b bits of the key are the direction quantised to b bits per axis (b = 1: the octant), the
    // remaining 30 - 3 b bits the Morton code of the origin at 10 - b bits per axis.  Measured on the soups (round 2,
    // incoherent rays, 16 Mi / 4 Mi triangles): b = 1 307 / 612 Mrays/s, 2 -> 224 / 517, 3 -> 218 / 469, 4 -> 207 / 450: where
    // a ray starts matters more than where exactly it heads; the octant stays.
    __global__ void rayKeyKernel(DeviceScene scene, const float *__restrict__ rays, uint32_t stride_floats, uint32_t n, uint32_t dir_bits, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ ids) {
        const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        const float *p = rays + static_cast<size_t>(stride_floats) * k;
        const uint32_t cell_bits = 10U - dir_bits;
        uint32_t cell[3];
        uint32_t heading[3];
        for(int c = 0; c < 3; c++) {
            const float extent = scene.root_hi[c] - scene.root_lo[c];
            float u = extent > 0.0F ? (p[c] - scene.root_lo[c]) / extent : 0.0F;
            u = fminf(fmaxf(u, 0.0F), 0.99999F);
            cell[c] = static_cast<uint32_t>(u * static_cast<float>(1U << cell_bits));
            // negative directions first within every bit, like the octant of dir_bits = 1
            const float v = fminf(fmaxf(0.5F - 0.5F * p[3 + c], 0.0F), 0.99999F);
            heading[c] = static_cast<uint32_t>(v * static_cast<float>(1U << dir_bits));
        }
        const uint32_t origin = (spread3(cell[0]) << 2) | (spread3(cell[1]) << 1) | spread3(cell[2]);
        const uint32_t direction = (spread3(heading[0]) << 2) | (spread3(heading[1]) << 1) | spread3(heading[2]);
        keys[k] = (direction << (3U * cell_bits)) | origin;
        ids[k] = k;
    }

    // Batch closest-hit query.  `index` (may be null) selects the rays of a re-trace pass: work item k is ray index[k].
    // MODE = kTraceCertified appends the rays without a certificate to `redo` (length *redo_count) instead of
    // writing their result.
    template<int MODE, bool COUNT, int SMEM>
    __global__ void __launch_bounds__(kBlock, PTB_TRACE_MIN_BLOCKS) intersectKernel(DeviceScene scene, VoteParams vote, const float *__restrict__ rays, const uint32_t *__restrict__ index,
                                                              const uint32_t *__restrict__ index_count, uint32_t n, float *__restrict__ t_out, int32_t *__restrict__ prim_out,
                                                              uint32_t *__restrict__ cursor, uint32_t *__restrict__ redo, uint32_t *__restrict__ redo_count,
                                                              VisitCounters *visits, const __grid_constant__ ptb_guard::CertGuard guard, int guarded) {
        const uint32_t count = (index != nullptr && index_count != nullptr) ? *index_count : n; // (an index without a count: a permutation of all n rays)
        warpTrace<MODE, COUNT, SMEM>(
          MODE == kTraceCertified ? occlusionView(scene) : scene, vote, cursor, count,
          [&](uint32_t k, V3 &o, V3 &d, float &limit) {
              const uint32_t ray = index != nullptr ? index[k] : k;
              const float *p = rays + 6 * static_cast<size_t>(ray);
              o = mk3(p[0], p[1], p[2]);
              d = mk3(p[3], p[4], p[5]);
              limit = 0.0F;
              return ray;
          },
          [&](uint32_t ray, const Hit &h, bool certain) {
              if(MODE == kTraceCertified && !certain) {
                  redo[atomicAdd(redo_count, 1U)] = ray;
              }
              else {
                  t_out[ray] = h.t;
                  prim_out[ray] = (h.slot >= 0 && h.t >= 0.0F) ? static_cast<int32_t>(scene.slot_to_prim[h.slot]) : -1;
              }
          },
          visits, (MODE == kTraceCertified && guarded != 0) ? &guard : nullptr);
    }

    template<bool COUNT, int SMEM>
    __global__ void __launch_bounds__(kBlock, PTB_TRACE_MIN_BLOCKS) occludedKernel(DeviceScene scene, VoteParams vote, const float *__restrict__ rays, const uint32_t *__restrict__ index,
                                                             uint32_t n, uint8_t *__restrict__ out, uint32_t *__restrict__ cursor, VisitCounters *visits) {
        warpTrace<kTraceAnyHit, COUNT, SMEM>(
          occlusionView(scene), vote, cursor, n,
          [&](uint32_t k, V3 &o, V3 &d, float &limit) {
              const uint32_t ray = index != nullptr ? index[k] : k;
              const float *p = rays + 7 * static_cast<size_t>(ray);
              o = mk3(p[0], p[1], p[2]);
              d = mk3(p[3], p[4], p[5]);
              limit = p[6];
              return ray;
          },
          [&](uint32_t k, const Hit &h, bool) { out[k] = h.slot >= 0 ? 1 : 0; }, visits);
    }

    __global__ void aabbKernel(float lox, float loy, float loz, float hix, float hiy, float hiz, const float *__restrict__ rays, uint64_t n,
                               float *__restrict__ t_out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        const float *p = rays + 6 * k;
        const RayInv r = makeRay(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]));
        t_out[k] = slab(r, lox, loy, loz, hix, hiy, hiz);
    }

    __global__ void primKernel(DeviceScene scene, const float *__restrict__ rays, uint64_t n, float *__restrict__ t_out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        const float *p = rays + 6 * k;
        const RayInv r = makeRay(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]));
        t_out[k] = hitSlot(scene, r, 0U);
    }

    PTB_DEV ReferenceRng engineFromState(uint64_t state) {
        ReferenceRng rng;
        rng.counter = 0U;
        rng.state = state;
        return rng;
    }

    __global__ void cameraKernel(ptb_camera camera, uint64_t n, const float *__restrict__ xy, float pixel_width, float pixel_height,
                                 uint64_t *__restrict__ states, float *__restrict__ out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        ReferenceRng rng = engineFromState(states[k]);
        V3 o;
        V3 d;
        shootRay(camera, xy[2 * k], xy[2 * k + 1], pixel_width, pixel_height, rng, o, d);
        states[k] = rng.state;
        float *w = out + 6 * k;
        w[0] = o.x;
        w[1] = o.y;
        w[2] = o.z;
        w[3] = d.x;
        w[4] = d.y;
        w[5] = d.z;
    }

    __global__ void apertureKernel(ptb_camera camera, uint64_t n, uint64_t *__restrict__ states, float *__restrict__ out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        ReferenceRng rng = engineFromState(states[k]);
        float sx;
        float sy;
        sampleAperture(camera, rng, sx, sy);
        states[k] = rng.state;
        out[2 * k] = sx;
        out[2 * k + 1] = sy;
    }

    // out[0] = number of samples; samples follow at out + 4, 8 floats each; *state is advanced
    __global__ void sampleLightsKernel(DeviceScene scene, float px, float py, float pz, uint64_t *__restrict__ state, uint32_t max_out,
                                       float *__restrict__ out) {
        if(blockIdx.x != 0U || threadIdx.x != 0U) {
            return;
        }
        ReferenceRng rng = engineFromState(*state);
        uint32_t n = 0U;
        sampleLights(scene, mk3(px, py, pz), rng, [&](const LightSample &ls) {
            if(n < max_out) {
                float *w = out + 4 + 8 * n;
                w[0] = ls.pos.x;
                w[1] = ls.pos.y;
                w[2] = ls.pos.z;
                w[3] = ls.spectrum.x;
                w[4] = ls.spectrum.y;
                w[5] = ls.spectrum.z;
                w[6] = ls.spectrum.w;
                w[7] = ls.pd;
            }
            n++;
        });
        *state = rng.state;
        out[0] = __uint_as_float(n);
    }

    // `scene` holds exactly one primitive in slot 0 (geometry + shading lanes)
    __global__ void primNormalKernel(DeviceScene scene, uint64_t n, const float *__restrict__ positions, float *__restrict__ out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        uint32_t material;
        const V3 normal = surfaceNormal(scene, 0U, mk3(positions[3 * k], positions[3 * k + 1], positions[3 * k + 2]), material);
        out[3 * k] = normal.x;
        out[3 * k + 1] = normal.y;
        out[3 * k + 2] = normal.z;
    }

    // lanes: the three un-differenced lanes of the primitive (emissive-table layout)
    __global__ void primSampleKernel(float4 e0, float4 e1, float4 e2, uint32_t flags, uint64_t n, uint64_t *__restrict__ states, float *__restrict__ out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        ReferenceRng rng = engineFromState(states[k]);
        V3 pos;
        float density;
        bool cull;
        samplePrimSurface(e0, e1, e2, flags, rng, pos, density, cull);
        states[k] = rng.state;
        float *w = out + 5 * k;
        w[0] = pos.x;
        w[1] = pos.y;
        w[2] = pos.z;
        w[3] = density;
        w[4] = cull ? 1.0F : 0.0F;
    }

    PTB_DEV Material materialFromPod(const ptb_material &m) {
        Material out;
        out.diffuse = V4{m.diffuse[0], m.diffuse[1], m.diffuse[2], m.diffuse[3]};
        out.emission = V4{m.emission[0], m.emission[1], m.emission[2], m.emission[3]};
        out.ior = m.refractive_index;
        out.bsdf = m.bsdf;
        out.one_way = m.one_way != 0U;
        return out;
    }

    __global__ void bsdfPropagateKernel(ptb_material material, float epsilon, uint64_t n, const float *__restrict__ in, uint64_t *__restrict__ states,
                                        float *__restrict__ out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        const float *p = in + 9 * k;
        ReferenceRng rng = engineFromState(states[k]);
        V3 o;
        V3 d;
        float factor;
        float pd;
        propagateRay(materialFromPod(material), mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), epsilon, rng, o, d, factor, pd);
        states[k] = rng.state;
        float *w = out + 8 * k;
        w[0] = o.x;
        w[1] = o.y;
        w[2] = o.z;
        w[3] = d.x;
        w[4] = d.y;
        w[5] = d.z;
        w[6] = factor;
        w[7] = pd;
    }

    __global__ void bsdfSpectrumKernel(ptb_material material, uint32_t synthetic, uint64_t n, const float *__restrict__ in, float *__restrict__ out) {
        const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if(k >= n) {
            return;
        }
        const float *p = in + 13 * k;
        V4 spectrum;
        float shade;
        float pd;
        bsdfSpectrum(materialFromPod(material), mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), V4{p[9], p[10], p[11], p[12]},
                     synthetic != 0U, spectrum, shade, pd);
        float *w = out + 6 * k;
        w[0] = spectrum.x;
        w[1] = spectrum.y;
        w[2] = spectrum.z;
        w[3] = spectrum.w;
        w[4] = shade;
        w[5] = pd;
    }

#endif // !PTB_FAST_MATH && !PTB_FAST_TU_EXACT

}

#endif
