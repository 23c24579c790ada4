/*
 * ptb.h — C-ABI of the B200-native render core ("PathTrace on B200").
 *
 * This is the drop-in boundary for ONE hot path of johannesschaeufele/CPUPathTrace:
 *
 *     processJob -> processItem -> impl::getSample -> Scene::getIntersection / Scene::sampleLights / BSDF
 *
 * The reference has no plugin/FFI layer; its boundary is the public C++ API in include/PathTrace (reference
 * include/PathTrace/worker.h:69,83-84, include/PathTrace/scene/scene.h:32,41,54, include/PathTrace/camera.h:92,108,123).
 * The C++ host layer of this repository (include/PathTrace + cpupathtrace_b200/host) re-provides exactly that
 * API and forwards the hot path through the entry points declared below.  A binding from any other language
 * (ctypes, cgo, JNI, N-API) binds these functions directly; see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, POD only, no exceptions cross this boundary; every function returns a ptb_status (0 = ok) and
 *     ptb_last_error() returns a thread-local, human readable description of the last failure;
 *   - the caller owns every host buffer it passes; a ptb_scene owns its device memory;
 *   - there is NO CPU fallback: without a CUDA device (or without the sm_100a kernels) every compute entry point
 *     fails with PTB_ERR_NO_DEVICE / PTB_ERR_CUDA;
 *   - pointers are host pointers unless the call's `flags` carry PTB_FLAG_DEVICE_IO, in which case the bulk
 *     input/output arrays are device pointers (same process, same device) and no host<->device copy is made.
 *
 * Environment (read when a context is created / a scene is built / a frame is rendered; none changes a result)
 *   PTB_DEVICE, LOCAL_RANK        device of ptb_context_create(-1)
 *   PTB_POOL_PATHS                paths in flight (default: 256 Mi, at most 3/8 of the free HBM, at most half of a call's samples)
 *   PTB_SAMPLE_BUFFER_MB          per-sample buffer budget of ptb_render (default: 40 % of the free HBM)
 *   PTB_GPU_BVH=1                 build the query hierarchy on the GPU (PTB_BVH_REFERENCE_GPU_QUERY_TREE); PTB_OCCLUSION_BVH=0: none
 *   PTB_BUILD_THREADS             host threads of the BVH builders
 *   PTB_REFILL_VOTE, PTB_LEAF_VOTE, PTB_SHADOW_REFILL_VOTE, PTB_SHADOW_LEAF_VOTE, PTB_LEAF_BURST, PTB_INNER_BURST,
 *   PTB_TRACE_BLOCKS_PER_SM, PTB_ITERATIONS_PER_SYNC    scheduling of the traversal kernels and of the bounce loop
 *   PTB_PROFILE=0                 no per-launch CUDA events; PTB_LOG_ITERATIONS=1: one stderr line per bounce iteration
 */
#ifndef PTB_H
#define PTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 2 /* 2: ptb_scene_info and ptb_render_stats grew at the end (round 2); compare with ptb_abi_version() */

typedef enum ptb_status {
    PTB_OK = 0,
    PTB_ERR_INVALID_ARGUMENT = 1,
    PTB_ERR_NO_DEVICE = 2,      /* no usable CUDA device: the library never falls back to the CPU */
    PTB_ERR_CUDA = 3,           /* a CUDA runtime call or kernel failed; see ptb_last_error()       */
    PTB_ERR_OUT_OF_MEMORY = 4,
    PTB_ERR_UNSUPPORTED = 5     /* e.g. a primitive / BSDF kind the device code does not know       */
} ptb_status;

typedef struct ptb_context ptb_context; /* one per (host thread, GPU) */
typedef struct ptb_scene ptb_scene;     /* immutable device-resident scene: BVH + primitives + materials + lights */

/* ------------------------------------------------------------------------------------------------ scene POD */

/* replaces the virtual Object hierarchy (reference include/PathTrace/scene/object.h:54-169) */
typedef enum ptb_prim_kind {
    PTB_PRIM_TRIANGLE = 0, /* Triangle  (object.h:145-169, src/scene/object.cpp:118-207) */
    PTB_PRIM_SPHERE = 1,   /* Sphere    (object.h:127-143, src/scene/object.cpp:68-116)  */
    PTB_PRIM_NULL = 2      /* NullObject (object.h:112-122): never hit, no area           */
} ptb_prim_kind;

typedef struct ptb_prim {
    uint32_t kind;          /* ptb_prim_kind */
    uint32_t material;      /* index into ptb_scene_desc.materials */
    uint32_t cull_backface; /* Triangle::cull_backface (object.h:157) */
    uint32_t reserved;
    /* triangle: a, b, c, normal_a, normal_b, normal_c (object.h:149-154); sphere: origin xyz, radius */
    float p[18];
} ptb_prim;

/* replaces BSDF subclasses (reference include/PathTrace/scene/propagation.h:55-104) */
typedef enum ptb_bsdf_kind {
    PTB_BSDF_LAMBERT = 0, /* LambertianBRDF (src/scene/propagation.cpp:87-116)  */
    PTB_BSDF_GLASS = 1,   /* GlassBDF       (src/scene/propagation.cpp:118-176) */
    PTB_BSDF_MIRROR = 2   /* MirrorBRDF     (src/scene/propagation.cpp:178-217) */
} ptb_bsdf_kind;

/* replaces ConstantMaterialHandler + ConstantMaterial + BSDF (object.h:35-47, material.h:62-77); the specular
 * colour is the Material default, white (src/scene/material.cpp:15-17) */
typedef struct ptb_material {
    float diffuse[4];
    float emission[4];
    float refractive_index;
    uint32_t bsdf;    /* ptb_bsdf_kind */
    uint32_t one_way; /* MirrorBRDF::one_way (propagation.h:86-87) */
    uint32_t reserved;
} ptb_material;

/* replaces PointLightSource (reference include/PathTrace/scene/light.h:53-66) */
typedef struct ptb_point_light {
    float pos[3];
    float rgba[4];
} ptb_point_light;

typedef enum ptb_bvh_mode {
    /* topology identical to impl::constructBVH (src/scene/scene.cpp:12-102): needed for bit-exact closest-hit parity */
    PTB_BVH_REFERENCE = 0,
    /* the same reference tree (it defines the answers), but the second hierarchy that any-hit and certified closest-hit
     * queries walk is built on the GPU as a linear BVH (Morton sort + Karras hierarchy, csrc/lbvh.cuh) instead of the
     * host's binned-SAH builder: milliseconds instead of tenths of a second, results unchanged (they do not depend on
     * that tree's shape), somewhat more node fetches per ray.  $PTB_GPU_BVH=1 selects it for PTB_BVH_REFERENCE scenes. */
    PTB_BVH_REFERENCE_GPU_QUERY_TREE = 1
} ptb_bvh_mode;

typedef struct ptb_scene_desc {
    const ptb_prim *prims; /* in the order of the `objects` vector given to Scene::Scene (scene.h:32) */
    uint64_t n_prims;
    const ptb_material *materials;
    uint32_t n_materials;
    const ptb_point_light *lights; /* in the order of the `light_sources` vector */
    uint32_t n_lights;
    uint32_t bvh_mode; /* ptb_bvh_mode */
    uint32_t reserved;
} ptb_scene_desc;

typedef struct ptb_scene_info {
    uint64_t n_prims;
    uint64_t n_inner_nodes;       /* 64-byte two-child records */
    uint32_t bvh_depth;           /* leaf depth maximum, root = 1 */
    uint32_t n_emissive;          /* Scene::object_light_sources (scene.h:19) */
    uint32_t object_sample_count; /* scene.cpp:226 */
    uint32_t n_lights;
    uint64_t device_bytes;        /* scene data resident in HBM */
    double build_seconds;         /* host BVH build + flatten */
    double upload_seconds;
    float root_low[3];
    float root_high[3];
    uint32_t query_tree_on_device; /* 1: the any-hit / certified-closest hierarchy was built by the GPU builder */
    uint32_t certifiable;          /* 1: the guard table of the certified walk covers every large triangle / sphere of the  */
                                   /* scene (csrc/cert_guard.h); 0: guarded certified queries walk the reference tree        */
    double query_tree_device_ms;   /* CUDA-event time of that build (Morton codes, sort, hierarchy, box fit) */
    uint32_t built_on_device;      /* 1: boxes, the reference-topology tree (impl::constructBVH, scene.cpp:12-102) and the leaf  */
                                   /* records were built by the GPU (csrc/gpu_build.cuh); 0: by the host builders                 */
    uint32_t query_tree_kind;      /* 0 none, 1 host binned SAH, 2 device linear BVH, 3 device full-sweep SAH (default)          */
    double reference_tree_device_ms; /* CUDA-event time of the device build of the reference-topology tree                       */
} ptb_scene_info;

/* ------------------------------------------------------------------------------------------------ camera POD */

typedef enum ptb_aperture_kind {
    PTB_APERTURE_NONE = 0,
    PTB_APERTURE_CIRCULAR = 1, /* CircularApertureSampler  (src/camera.cpp:7-19)  */
    PTB_APERTURE_HEXAGONAL = 2 /* HexagonalApertureSampler (src/camera.cpp:21-49) */
} ptb_aperture_kind;

/* The private state of Camera after its constructor ran (reference include/PathTrace/camera.h:68-78,
 * src/camera.cpp:54-76): `forward` is already scaled by the focal length, `up`/`right` by the half extents. */
typedef struct ptb_camera {
    float origin[3];
    float forward[3];
    float up[3];
    float right[3];
    float aperture_width_half;
    float aperture_height_half;
    uint32_t aperture_kind; /* ptb_aperture_kind */
    float hexagon_horizontal_ratio;
    float focal_plane_dist;
} ptb_camera;

/* Camera::Camera (src/camera.cpp:54-76) on the host, for bindings that have no C++ Camera object */
int ptb_camera_init(ptb_camera *out, const float origin[3], const float look_at[3], const float up[3], float focal_length, float height,
                    float aspect_ratio, float aperture_width, float aperture_height, uint32_t aperture_kind, float hexagon_horizontal_ratio,
                    float focal_plane_dist);

/* ------------------------------------------------------------------------------------------------ options */

typedef enum ptb_rng_mode {
    /* production: stateless counter-based generator keyed on (job seed, pixel, sample), counter = (bounce, draw) */
    PTB_RNG_COUNTER = 0,
    /* validation: one reference engine per (pixel, sample) — xorshift (reference include/PathTrace/base.h:24-41)
     * seeded by the caller, consumed through libstdc++'s distribution arithmetic, draw for draw as the reference */
    PTB_RNG_REFERENCE_XORSHIFT = 1
} ptb_rng_mode;

#define PTB_FLAG_DEVICE_IO 0x1u      /* bulk in/out arrays are device pointers                                          */
#define PTB_FLAG_ANY_HIT_SHADOWS 0x2u /* shadow rays stop at the first occluder (result-identical up to ulp-level box/   */
                                      /* primitive disagreement) instead of the reference's full closest-hit query       */
#define PTB_FLAG_SKIP_NULL_SHADOWS 0x4u /* do not trace shadow rays that cannot change the radiance: BSDF pd 0 for       */
                                        /* synthetic rays (Glass, Mirror: worker.cpp:84-92 traces them and discards the   */
                                        /* result) or a contribution of +-0 in all channels (surface facing away)         */
#define PTB_FLAG_COUNT_VISITS 0x8u   /* count BVH node / primitive fetches (slower; feeds the bytes-per-ray figure)      */
#define PTB_FLAG_CERTIFIED_CLOSEST 0x10u /* closest-hit queries walk the SAH hierarchy and keep the result only when it    */
                                     /* carries a certificate that Scene::getIntersection returns the same primitive in  */
                                     /* any visiting order (strictly nearest, no rival within its leaf-box entry); the    */
                                     /* remaining rays (ties on shared edges/vertices, near-ties) are re-traced on the    */
                                     /* reference-topology tree.  See csrc/traverse.cuh "certified closest hit".          */
                                     /* Guarded by default: rays for which fp32 rounding in a LARGE triangle's or a nearby  */
                                     /* sphere's own intersection routine could matter are sent to the reference walk up    */
                                     /* front (csrc/cert_guard.h); scenes the guard table cannot cover are not certified.   */
#define PTB_FLAG_SINGLE_STREAM 0x80u /* do not split this render between several contexts of the device (a large render is    */
                                     /* otherwise shared by PTB_STREAMS concurrent contexts: same image, ~5 % sooner); with the */
                                     /* flag the per-kernel event times of ptb_render_stats are those of kernels running alone  */
#define PTB_FLAG_PROFILE_ALL 0x40u   /* CUDA events around EVERY launch (device_ms_shade / device_ms_trace_shadow are filled);  */
                                     /* without it only the dominant kernel, the closest-hit trace, is timed per launch     */
                                     /* (device_ms_trace = closest-hit trace only): the extra ~300 event pairs per frame    */
                                     /* cost 4-5 % of a frame                                                                */
#define PTB_FLAG_CERTIFIED_RELAXED 0x20u /* with PTB_FLAG_CERTIFIED_CLOSEST: no guard.  Identical results except where the  */
                                     /* reference's own answer is rounding noise of a grazing hit on a large triangle       */
                                     /* (tests/stress_cases.py); meant for production renders (PTB_RNG_COUNTER)             */

/* RenderOptions (reference include/PathTrace/worker.h:14-31) + the knobs that exist only on this side */
typedef struct ptb_render_opts {
    int32_t image_width;
    int32_t image_height;
    int32_t min_sample_count;
    int32_t max_sample_count;
    float epsilon;
    int32_t max_depth;    /* 0 = unlimited like the reference (paths end by Russian roulette, worker.cpp:67-70) */
    uint32_t rng_mode;    /* ptb_rng_mode */
    uint32_t flags;       /* PTB_FLAG_* */
    uint64_t seed;        /* job key for PTB_RNG_COUNTER */
    /* multi-GPU sharding by interleaved tiles (reference tile grid: worker.cpp:398-414): tile k (row-major) is
     * rendered iff k % shard_count == shard_index; pixels of other tiles are written as 0 so that a sum-reduce of
     * the per-rank images is the full image.  shard_count <= 1 renders everything. */
    int32_t tile_size;    /* 0 = the reference's clamp(min(W,H)/4, 1, 32) */
    int32_t shard_index;
    int32_t shard_count;
    uint32_t reserved;
} ptb_render_opts;

typedef struct ptb_render_stats {
    uint64_t samples;        /* pixel-samples started (primary rays)                     */
    uint64_t closest_rays;   /* Scene::getIntersection-equivalent closest-hit queries    */
    uint64_t shadow_rays;    /* shadow queries traced                                     */
    uint64_t shadow_rays_skipped; /* shadow queries the reference traces but PTB_FLAG_SKIP_NULL_SHADOWS dropped */
    uint64_t path_vertices;  /* surface hits shaded                                       */
    uint64_t inner_visits;   /* 64-byte inner records fetched (PTB_FLAG_COUNT_VISITS)    */
    uint64_t leaf_visits;    /* 48-byte primitive records fetched (PTB_FLAG_COUNT_VISITS) */
    uint64_t bounce_iterations;
    uint64_t kernel_launches;
    double device_ms_total;  /* CUDA-event time of the whole call on the context's stream */
    double device_ms_trace;  /* closest-hit traversal kernels (+ shadow traversal with PTB_FLAG_PROFILE_ALL) */
    double device_ms_shade;  /* generate + shade + accumulate + resolve (PTB_FLAG_PROFILE_ALL only)           */
    double device_ms_trace_shadow;  /* the shadow-ray share of device_ms_trace (PTB_FLAG_PROFILE_ALL only)    */
    uint64_t shadow_inner_visits;   /* the shadow-ray share of inner_visits               */
    uint64_t shadow_leaf_visits;    /* the shadow-ray share of leaf_visits                */
    uint64_t closest_rays_retraced; /* PTB_FLAG_CERTIFIED_CLOSEST: queries without a certificate, re-traced on the reference tree */
    uint64_t certified_suspect_hits; /* with PTB_FLAG_COUNT_VISITS: primitive tests of the certified walk that reported a hit */
                                     /* more than 2^-8 in front of the primitive's own bounding box -- the only situation in */
                                     /* which a primitive the walk never reaches could change the reference's answer         */
    uint64_t samples_used;    /* ptb_render: pixel-samples the per-pixel loops of processItem consumed (worker.cpp:172-260).   */
                              /* min == max: equals `samples`.  min < max: the loop of a pixel ends early once the acceptance */
                              /* test fires; samples - samples_used were traced ahead of a loop that had already ended        */
    uint64_t adaptive_rounds; /* min < max: rounds of samples traced (only pixels still sampling take part in a round)        */
} ptb_render_stats;

/* ------------------------------------------------------------------------------------------------ entry points */

int ptb_abi_version(void);
const char *ptb_last_error(void);

/* device < 0 selects $PTB_DEVICE, else $LOCAL_RANK, else 0 */
int ptb_context_create(int device, ptb_context **out);
int ptb_context_destroy(ptb_context *ctx);
int ptb_context_device(const ptb_context *ctx, int *device_out);
int ptb_context_synchronize(ptb_context *ctx);

/* Scene::Scene (src/scene/scene.cpp:153-181): BVH build, emissive registration, CDF; then flatten + upload */
int ptb_scene_create(ptb_context *ctx, const ptb_scene_desc *desc, ptb_scene **out);
int ptb_scene_destroy(ptb_scene *scene);
int ptb_scene_get_info(const ptb_scene *scene, ptb_scene_info *out);

/* Copies one of the scene's device arrays to the host (inspection, tests, serialisation): `bytes` must not exceed the
 * array's size.  PTB_SCENE_NODES / PTB_SCENE_QUERY_NODES: 64-byte inner records (csrc/bvh_build.h: left box, right box,
 * left ref, right ref, leaf count, parent; ref >= 0 inner record, < 0 ~leaf slot); PTB_SCENE_GEOM: 64 bytes per leaf slot;
 * PTB_SCENE_SHADE: 48 bytes per leaf slot; PTB_SCENE_SLOT_TO_PRIM: uint32 per leaf slot (index into ptb_scene_desc.prims). */
typedef enum ptb_scene_array {
    PTB_SCENE_NODES = 0,
    PTB_SCENE_QUERY_NODES = 1,
    PTB_SCENE_GEOM = 2,
    PTB_SCENE_SHADE = 3,
    PTB_SCENE_SLOT_TO_PRIM = 4
} ptb_scene_array;
int ptb_scene_read(const ptb_scene *scene, uint32_t array, void *out, uint64_t bytes);

/* Scene::getIntersection (src/scene/scene.cpp:210-220) for a batch.
 * rays: 6 floats each (origin xyz, unit direction xyz).  t_out[i] < 0 = miss (prim_out[i] = -1), else the
 * distance and the index of the primitive in ptb_scene_desc.prims. */
int ptb_intersect(ptb_scene *scene, const float *rays, uint64_t n_rays, float *t_out, int32_t *prim_out, uint32_t flags, ptb_render_stats *stats);

/* The shadow query of impl::getSample (src/worker.cpp:80-86) as an any-hit test.
 * rays: 7 floats each (origin, unit direction, limit); occluded_out[i] = 1 iff some primitive is hit with
 * 0 <= t < limit. */
int ptb_occluded(ptb_scene *scene, const float *rays, uint64_t n_rays, uint8_t *occluded_out, uint32_t flags, ptb_render_stats *stats);

/* processItem / processJob (src/worker.cpp:149-326, 389-424) for the pixel rectangle [x0, x0+w) x [y0, y0+h):
 * out_rgba receives w*h*4 floats, row-major, the per-pixel value processItem computes (incl. the Welford batch
 * statistics, adaptive acceptance and candidate merge when min != max). */
int ptb_render(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h, float *out_rgba,
               ptb_render_stats *stats);

/* ptb_render with progress reports: `progress(user, samples_done, samples_total)` is called from the calling thread while
 * the frame renders (between batches of bounce iterations: samples_done = pixel-samples retired so far) and once more
 * when everything is done.  It is what processJob's progress_callback (src/worker.cpp:354-360: fired as tiles complete)
 * is driven by.  progress may be NULL. */
typedef void (*ptb_progress_fn)(void *user, uint64_t samples_done, uint64_t samples_total);
int ptb_render_with_progress(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h,
                             float *out_rgba, ptb_render_stats *stats, ptb_progress_fn progress, void *user);

/* ---- several GPUs, one process, one call (processJob's contract: one call -> the whole image, src/worker.cpp:389-424)
 *
 * ptb_scene_clone copies a scene's device arrays to the device of `ctx` (no BVH rebuild).  ptb_render_multi renders the
 * rectangle on n replicas of one scene at once -- one host thread per replica, tile k of the tile grid on replica
 * k % n, exactly the tiles ptb_render renders with shard_index = replica -- then ONE kernel on the first replica's device
 * reads every replica's owned tiles over NVLink (peer access; staged copies where peers cannot map each other) into the
 * final image, which is copied to out_rgba (host pointer, or a pointer on the first replica's device with
 * PTB_FLAG_DEVICE_IO).  Bit-identical to ptb_render on one device.  stats: n entries (one per replica) or NULL.
 * progress may be called from any of the n threads, never concurrently. */
int ptb_device_count(int *count_out);
int ptb_scene_clone(const ptb_scene *scene, ptb_context *ctx, ptb_scene **out);
int ptb_render_multi(ptb_scene *const *replicas, int32_t n, const ptb_camera *camera, const ptb_render_opts *opts, int32_t x0, int32_t y0, int32_t w, int32_t h,
                     float *out_rgba, ptb_render_stats *stats, ptb_progress_fn progress, void *user);

/* Validation entry: one impl::getSample (src/worker.cpp:26-146) per (pixel, seed) with rng_mode
 * PTB_RNG_REFERENCE_XORSHIFT == RandomEngine(seed).  pixels: 2 ints each; out_rgba: 4 floats each (alpha = collected). */
int ptb_render_samples(ptb_scene *scene, const ptb_camera *camera, const ptb_render_opts *opts, uint64_t n, const int32_t *pixels,
                       const uint64_t *seeds, float *out_rgba, ptb_render_stats *stats);

/* ------------------------------------------------------------------------------------------------ unit entries
 *
 * The reference exposes every stage of the path as a public (virtual) method that host code may call one element at a
 * time: Camera::shootRay, ApertureSampler::sampleAperture, AABB::getIntersection, Object::getIntersection /
 * getSurfaceNormal / sampleSurface, Scene::sampleLights, BSDF::propagateRay / getSpectrum.  The host layer answers
 * those calls with the batch entries below, which launch the SAME device functions the wavefront kernels use.
 *
 * engine_states: raw 64-bit xorshift states (reference include/PathTrace/base.h:24-41; a fresh RandomEngine(seed) has
 * state seed ^ (~seed << 32)), one per element, updated in place to the state after the element's draws.
 */

/* Camera::shootRay (src/camera.cpp:78-113); xy: 2 floats, rays_out: 6 floats per element */
int ptb_camera_shoot(ptb_context *ctx, const ptb_camera *camera, uint64_t n, const float *xy, float pixel_width, float pixel_height,
                     uint64_t *engine_states, float *rays_out);

/* ApertureSampler::sampleAperture (src/camera.cpp:7-49); out: 2 floats per element */
int ptb_aperture_sample(ptb_context *ctx, uint32_t aperture_kind, float hexagon_horizontal_ratio, uint64_t n, uint64_t *engine_states, float *out);

/* Scene::sampleLights (src/scene/scene.cpp:222-289) at one position.
 * out: 8 floats per sample (pos xyz, spectrum rgba, pd); *n_out = number of samples produced (may exceed max_out). */
int ptb_sample_lights(ptb_scene *scene, const float pos[3], uint64_t *engine_state, uint32_t max_out, float *out, uint32_t *n_out);

/* AABB::getIntersection (src/scene/bounding_box.cpp:38-73) for one box and a batch of rays (6 floats each) */
int ptb_aabb_intersect(ptb_context *ctx, const float low[3], const float high[3], uint64_t n_rays, const float *rays, float *t_out);

/* Object::getIntersection for one primitive and a batch of rays (Triangle: object.cpp:146-182, Sphere: :72-84) */
int ptb_prim_intersect(ptb_context *ctx, const ptb_prim *prim, uint64_t n_rays, const float *rays, float *t_out);

/* Object::getSurfaceNormal (Triangle: object.cpp:126-144, Sphere: :86-88); positions / normals_out: 3 floats each */
int ptb_prim_normal(ptb_context *ctx, const ptb_prim *prim, uint64_t n, const float *positions, float *normals_out);

/* Object::sampleSurface (Triangle: object.cpp:192-207, Sphere: :101-116); out: 5 floats (pos xyz, density, cull) */
int ptb_prim_sample(ptb_context *ctx, const ptb_prim *prim, uint64_t n, uint64_t *engine_states, float *out);

/* BSDF::propagateRay (src/scene/propagation.cpp:89-99, 120-160, 180-204).
 * in: 9 floats (incoming direction, position, normal); out: 8 floats (origin, direction, factor, density) */
int ptb_bsdf_propagate(ptb_context *ctx, const ptb_material *material, float epsilon, uint64_t n, const float *in, uint64_t *engine_states, float *out);

/* BSDF::getSpectrum (src/scene/propagation.cpp:101-116, 162-176, 206-217).
 * in: 13 floats (from-camera direction, to-light direction, normal, light rgba); out: 6 floats (rgba, shade, density) */
int ptb_bsdf_spectrum(ptb_context *ctx, const ptb_material *material, uint32_t synthetic, uint64_t n, const float *in, float *out);

/* ------------------------------------------------------------------------------------------------ post-processing
 *
 * toneMap / gammaCorrect / postProcess (reference src/post_processing.cpp:32-163, 165-177, 179-182) in place on a
 * width x height RGBA float image (host pointer, or a device pointer with PTB_FLAG_DEVICE_IO so that a frame rendered
 * into HBM by ptb_render can be post-processed without leaving the device).  Alpha is left untouched. */
typedef enum ptb_post_mode {
    PTB_POST_TONE_MAP = 0,
    PTB_POST_GAMMA = 1,
    PTB_POST_BOTH = 2 /* tone map, then gamma */
} ptb_post_mode;

int ptb_post_process(ptb_context *ctx, float *rgba, int32_t width, int32_t height, uint32_t mode, float gamma, uint32_t flags);

#ifdef __cplusplus
}
#endif

#endif /* PTB_H */
