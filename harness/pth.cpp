// pth — a C-callable test harness written ONLY against the public PathTrace C++ API
// (include/PathTrace/**).  The same source is compiled twice:
//
//   * against /root/reference/{include,src}  -> oracle/_ref/libpth_ref*.so   (the oracle)
//   * against this repo's include/ + host lib -> cpupathtrace_b200/lib/libpth_b200.so
//
// so that the Python test-suite can build identical scenes on both sides and compare results.
// It plays the role of the reference's own callers (demo/main.cpp, benchmark/main.cpp, test/**):
// if this file compiles against our headers, those callers do too.
//
// Nothing in here is product code; nothing in here touches the GPU directly.  The few
// `#ifdef PATHTRACE_B200` blocks use batch extensions of the new headers where the reference API
// only offers one-ray-at-a-time calls (a GPU round trip per ray would be pointless to measure).

#include <PathTrace/base.h>
#include <PathTrace/camera.h>
#include <PathTrace/post_processing.h>
#include <PathTrace/worker.h>
#include <PathTrace/scene/bounding_box.h>
#include <PathTrace/scene/light.h>
#include <PathTrace/scene/mesh.h>
#include <PathTrace/scene/object.h>
#include <PathTrace/scene/scene.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

    struct Builder {
        std::vector<std::unique_ptr<Object>> objects;
        std::vector<std::unique_ptr<LightSource>> lights;
        std::vector<std::shared_ptr<MaterialHandler>> handlers;
    };

    struct SceneBox {
        std::unique_ptr<Scene> scene;
        std::unordered_map<const Object *, int> ids;
    };

    double g_last_job_timeline[3] = {0.0, -1.0, -1.0};

    vec3<float> v3(const float *p) {
        return vec3<float>{p[0], p[1], p[2]};
    }

    void applyMaterial(Builder *b, Object &object, int material) {
        if(material >= 0 && material < static_cast<int>(b->handlers.size())) {
            object.setMaterialHandler(b->handlers[material]);
        }
    }

    int pushTriangles(Builder *b, std::vector<Triangle> &triangles, int material) {
        for(auto &triangle : triangles) {
            applyMaterial(b, triangle, material);
        }
        moveObjects(b->objects, triangles);
        return static_cast<int>(triangles.size());
    }

}

extern "C" {

const char *pth_impl_name() {
#ifdef PATHTRACE_B200
    return "b200";
#else
    return "reference";
#endif
}

// ---------------------------------------------------------------- builder

void *pth_builder_new() {
    return new Builder();
}

void pth_builder_free(void *builder) {
    delete static_cast<Builder *>(builder);
}

int pth_builder_object_count(void *builder) {
    return static_cast<int>(static_cast<Builder *>(builder)->objects.size());
}

// bsdf: 0 = LambertianBRDF, 1 = GlassBDF, 2 = MirrorBRDF(one_way)
int pth_add_material(void *builder, const float *diffuse, float ior, const float *emission, int bsdf, int one_way) {
    auto *b = static_cast<Builder *>(builder);

    auto material = std::make_shared<ConstantMaterial>(Color<float>(diffuse[0], diffuse[1], diffuse[2], diffuse[3]), ior,
                                                       Spectrum(Color<float>{emission[0], emission[1], emission[2], emission[3]}));
    std::shared_ptr<BSDF> f;
    switch(bsdf) {
        case 1:
            f = std::make_shared<GlassBDF>();
            break;
        case 2:
            f = std::make_shared<MirrorBRDF>(one_way != 0);
            break;
        default:
            f = std::make_shared<LambertianBRDF>();
            break;
    }
    b->handlers.push_back(std::make_shared<ConstantMaterialHandler>(material, f));

    return static_cast<int>(b->handlers.size()) - 1;
}

// verts: 9 floats per triangle (a, b, c); normals: 9 floats per triangle or NULL (face normal)
// material < 0 keeps the library's default (white Lambertian) handler
int pth_add_triangles(void *builder, int count, const float *verts, const float *normals, int cull, int material) {
    auto *b = static_cast<Builder *>(builder);
    b->objects.reserve(b->objects.size() + count);
    for(int i = 0; i < count; i++) {
        const float *p = verts + 9 * static_cast<size_t>(i);
        auto triangle = std::make_unique<Triangle>(v3(p), v3(p + 3), v3(p + 6), cull != 0);
        if(normals != nullptr) {
            const float *n = normals + 9 * static_cast<size_t>(i);
            triangle->normal_a = v3(n);
            triangle->normal_b = v3(n + 3);
            triangle->normal_c = v3(n + 6);
        }
        applyMaterial(b, *triangle, material);
        b->objects.emplace_back(std::move(triangle));
    }
    return count;
}

// spheres: 4 floats each (origin, radius)
int pth_add_spheres(void *builder, int count, const float *spheres, int material) {
    auto *b = static_cast<Builder *>(builder);
    for(int i = 0; i < count; i++) {
        const float *p = spheres + 4 * static_cast<size_t>(i);
        auto sphere = std::make_unique<Sphere>(v3(p), p[3]);
        applyMaterial(b, *sphere, material);
        b->objects.emplace_back(std::move(sphere));
    }
    return count;
}

int pth_add_plane(void *builder, const float *a, const float *b_, int cull, int material) {
    auto triangles = makePlane(v3(a), v3(b_), cull != 0);
    return pushTriangles(static_cast<Builder *>(builder), triangles, material);
}

// transform: optional row-major 4x4 applied to every vertex of the box before the triangles are re-made
// (the demo does exactly that: Triangle(transformation * a, ...) which also recomputes face normals)
int pth_add_box(void *builder, const float *a, const float *b_, int cull, const float *transform, int material) {
    auto triangles = makeBox(v3(a), v3(b_), cull != 0);
    if(transform != nullptr) {
        mat4<float> m{vec4<float>{transform[0], transform[1], transform[2], transform[3]}, //
                      vec4<float>{transform[4], transform[5], transform[6], transform[7]}, //
                      vec4<float>{transform[8], transform[9], transform[10], transform[11]}, //
                      vec4<float>{transform[12], transform[13], transform[14], transform[15]}};
        std::vector<Triangle> transformed;
        transformed.reserve(triangles.size());
        for(auto &triangle : triangles) {
            transformed.emplace_back(m * triangle.a, m * triangle.b, m * triangle.c, cull != 0);
        }
        triangles = std::move(transformed);
    }
    return pushTriangles(static_cast<Builder *>(builder), triangles, material);
}

int pth_add_mesh_obj(void *builder, const char *text, long length, const float *transform, int cull, int smooth, int material) {
    std::istringstream stream(std::string(text, static_cast<size_t>(length)));
    std::vector<Triangle> triangles;
    if(transform != nullptr) {
        mat4<float> m{vec4<float>{transform[0], transform[1], transform[2], transform[3]}, //
                      vec4<float>{transform[4], transform[5], transform[6], transform[7]}, //
                      vec4<float>{transform[8], transform[9], transform[10], transform[11]}, //
                      vec4<float>{transform[12], transform[13], transform[14], transform[15]}};
        triangles = io::loadMesh(stream, m, cull != 0, smooth != 0);
    }
    else {
        triangles = io::loadMesh(stream, mat4_identity<float>, cull != 0, smooth != 0);
    }
    return pushTriangles(static_cast<Builder *>(builder), triangles, material);
}

int pth_add_mesh_file(void *builder, const char *path, const float *transform, int cull, int smooth, int material) {
    mat4<float> m = mat4_identity<float>;
    if(transform != nullptr) {
        m = mat4<float>{vec4<float>{transform[0], transform[1], transform[2], transform[3]}, //
                        vec4<float>{transform[4], transform[5], transform[6], transform[7]}, //
                        vec4<float>{transform[8], transform[9], transform[10], transform[11]}, //
                        vec4<float>{transform[12], transform[13], transform[14], transform[15]}};
    }
    auto triangles = io::loadMesh(std::filesystem::path(path), m, cull != 0, smooth != 0);
    return pushTriangles(static_cast<Builder *>(builder), triangles, material);
}

void pth_add_point_light(void *builder, const float *pos, const float *rgba) {
    auto *b = static_cast<Builder *>(builder);
    b->lights.emplace_back(std::make_unique<PointLightSource>(v3(pos), Spectrum(Color<float>{rgba[0], rgba[1], rgba[2], rgba[3]})));
}

// Reads back triangles [first, first+count) of the builder: 18 floats each (a, b, c, na, nb, nc);
// non-triangle objects are written as NaN.  Returns the number of triangles written.
int pth_builder_get_triangles(void *builder, int first, int count, float *out) {
    auto *b = static_cast<Builder *>(builder);
    int written = 0;
    for(int i = 0; i < count; i++) {
        float *o = out + 18 * static_cast<size_t>(i);
        const auto *triangle = dynamic_cast<const Triangle *>(b->objects[first + i].get());
        if(triangle == nullptr) {
            for(int k = 0; k < 18; k++) {
                o[k] = std::numeric_limits<float>::quiet_NaN();
            }
            continue;
        }
        const vec3<float> *fields[6] = {&triangle->a, &triangle->b, &triangle->c, &triangle->normal_a, &triangle->normal_b, &triangle->normal_c};
        for(int f = 0; f < 6; f++) {
            for(int k = 0; k < 3; k++) {
                o[3 * f + k] = (*fields[f])[k];
            }
        }
        written++;
    }
    return written;
}

// Per-object scalar queries through the virtual Object interface: area and bounding volume (7 floats each)
void pth_builder_get_object_info(void *builder, int first, int count, float *out) {
    auto *b = static_cast<Builder *>(builder);
    for(int i = 0; i < count; i++) {
        const auto &object = b->objects[first + i];
        auto area = object->getBoundingVolume();
        float *o = out + 7 * static_cast<size_t>(i);
        o[0] = object->getSurfaceArea();
        for(int k = 0; k < 3; k++) {
            o[1 + k] = area.low[k];
            o[4 + k] = area.high[k];
        }
    }
}

// ---------------------------------------------------------------- scene

// Consumes the builder's objects and lights (the builder stays valid but empty).
void *pth_scene_new(void *builder) {
    auto *b = static_cast<Builder *>(builder);
    auto *box = new SceneBox();
    box->ids.reserve(b->objects.size());
    for(size_t i = 0; i < b->objects.size(); i++) {
        box->ids.emplace(b->objects[i].get(), static_cast<int>(i));
    }
    box->scene = std::make_unique<Scene>(std::move(b->objects), std::move(b->lights));
    b->objects.clear();
    b->lights.clear();
    return box;
}

void pth_scene_free(void *scene) {
    delete static_cast<SceneBox *>(scene);
}

// rays: 6 floats each (origin, dir).  t_out[i] = distance or negative; id_out[i] = index of the object in the
// builder's insertion order, or -1 when the library returned no object.
void pth_scene_intersect(void *scene, long count, const float *rays, float *t_out, int *id_out) {
    auto *box = static_cast<SceneBox *>(scene);
#ifdef PATHTRACE_B200
    std::vector<Ray> batch(static_cast<size_t>(count));
    for(long i = 0; i < count; i++) {
        batch[i] = Ray{v3(rays + 6 * i), v3(rays + 6 * i + 3)};
    }
    std::vector<const Object *> objects(static_cast<size_t>(count));
    box->scene->getIntersections(batch.data(), static_cast<size_t>(count), t_out, objects.data());
    for(long i = 0; i < count; i++) {
        auto it = box->ids.find(objects[i]);
        id_out[i] = (objects[i] != nullptr && t_out[i] >= 0.0F && it != box->ids.end()) ? it->second : -1;
    }
#else
    for(long i = 0; i < count; i++) {
        Ray ray{v3(rays + 6 * i), v3(rays + 6 * i + 3)};
        auto [t, object] = box->scene->getIntersection(ray);
        t_out[i] = t;
        int id = -1;
        if(object != nullptr && t >= 0.0F) {
            auto it = box->ids.find(object);
            if(it != box->ids.end()) {
                id = it->second;
            }
        }
        id_out[i] = id;
    }
#endif
}

// One ray through the plain single-ray API (both builds)
void pth_scene_intersect_one(void *scene, const float *ray6, float *t_out, int *id_out) {
    auto *box = static_cast<SceneBox *>(scene);
    Ray ray{v3(ray6), v3(ray6 + 3)};
    auto [t, object] = box->scene->getIntersection(ray);
    *t_out = t;
    *id_out = -1;
    if(object != nullptr && t >= 0.0F) {
        auto it = box->ids.find(object);
        if(it != box->ids.end()) {
            *id_out = it->second;
        }
    }
}

// Scene::sampleLights at one position with RandomEngine(seed).  out: 8 floats per light sample
// (pos xyz, spectrum rgba, pd).  Returns the number of samples (at most max_out are written).
int pth_scene_sample_lights(void *scene, const float *pos, const float *n, uint64_t seed, int max_out, float *out) {
    auto *box = static_cast<SceneBox *>(scene);
    RandomEngine re(seed);
    auto lights = box->scene->sampleLights(v3(pos), v3(n), re);
    int index = 0;
    for(const auto &[light_pos, spectrum, pd] : lights) {
        if(index < max_out) {
            float *o = out + 8 * static_cast<size_t>(index);
            auto color = spectrum.getColor();
            o[0] = light_pos[0];
            o[1] = light_pos[1];
            o[2] = light_pos[2];
            o[3] = color[0];
            o[4] = color[1];
            o[5] = color[2];
            o[6] = color[3];
            o[7] = pd;
        }
        index++;
    }
    return index;
}

// AABB::getIntersection on a leaf node wrapping a unit sphere's box scaled to [low, high] (the slab-test KAT)
void pth_aabb_intersect(const float *low, const float *high, long count, const float *rays, float *t_out) {
    AABBArea area{v3(low), v3(high)};
    AABB aabb(area, std::make_unique<Sphere>(vec3<float>(0.0F, 0.0F, 0.0F), 1.0F));
    for(long i = 0; i < count; i++) {
        Ray ray{v3(rays + 6 * i), v3(rays + 6 * i + 3)};
        t_out[i] = aabb.getIntersection(ray);
    }
}

// ---------------------------------------------------------------- one-element virtual methods
//
// Object::getSurfaceNormal / sampleSurface and BSDF::propagateRay / getSpectrum called through the public virtual
// interface on an object / material of a builder, one element at a time like a host-side caller would.  Engines are
// RandomEngine(seed); `next_draw` receives the engine's next output after the call (the reference engine has no state
// accessor), which pins how many draws the call consumed.

// out: 3 floats per position
void pth_object_normal(void *builder, int index, long count, const float *positions, float *out) {
    auto *b = static_cast<Builder *>(builder);
    const Object &object = *b->objects[index];
    for(long i = 0; i < count; i++) {
        const auto n = object.getSurfaceNormal(v3(positions + 3 * i));
        out[3 * i] = n[0];
        out[3 * i + 1] = n[1];
        out[3 * i + 2] = n[2];
    }
}

// out: 5 floats per seed (pos xyz, density, cull)
void pth_object_sample(void *builder, int index, long count, const uint64_t *seeds, float *out, uint32_t *next_draw) {
    auto *b = static_cast<Builder *>(builder);
    const Object &object = *b->objects[index];
    for(long i = 0; i < count; i++) {
        RandomEngine re(seeds[i]);
        const auto [pos, density, cull] = object.sampleSurface(re);
        float *o = out + 5 * i;
        o[0] = pos[0];
        o[1] = pos[1];
        o[2] = pos[2];
        o[3] = density;
        o[4] = cull ? 1.0F : 0.0F;
        next_draw[i] = re();
    }
}

// in: 9 floats (incoming direction, position, normal); out: 8 floats (origin, direction, factor, density)
void pth_bsdf_propagate(void *builder, int material, float epsilon, long count, const float *in, const uint64_t *seeds, float *out, uint32_t *next_draw) {
    auto *b = static_cast<Builder *>(builder);
    const MaterialHandler &handler = *b->handlers[material];
    for(long i = 0; i < count; i++) {
        const float *p = in + 9 * i;
        const auto pos = v3(p + 3);
        RandomEngine re(seeds[i]);
        const auto [ray, factor, density] = handler.getBSDF(pos)->propagateRay(Ray{pos, v3(p)}, pos, v3(p + 6), epsilon, re, handler.getMaterial(pos));
        float *o = out + 8 * i;
        o[0] = ray.origin[0];
        o[1] = ray.origin[1];
        o[2] = ray.origin[2];
        o[3] = ray.dir[0];
        o[4] = ray.dir[1];
        o[5] = ray.dir[2];
        o[6] = factor;
        o[7] = density;
        next_draw[i] = re();
    }
}

// in: 13 floats (from-camera direction, to-light direction, normal, light rgba); out: 6 floats (rgba, shade, density)
void pth_bsdf_spectrum(void *builder, int material, int synthetic, long count, const float *in, float *out) {
    auto *b = static_cast<Builder *>(builder);
    const MaterialHandler &handler = *b->handlers[material];
    const vec3<float> pos{0.0F, 0.0F, 0.0F};
    for(long i = 0; i < count; i++) {
        const float *p = in + 13 * i;
        const auto [spectrum, shade, density] = handler.getBSDF(pos)->getSpectrum(Ray{pos, v3(p)}, Ray{pos, v3(p + 3)}, pos, v3(p + 6),
                                                                                  Spectrum(Color<float>{p[9], p[10], p[11], p[12]}), handler.getMaterial(pos), synthetic != 0);
        const auto color = spectrum.getColor();
        float *o = out + 6 * i;
        o[0] = color[0];
        o[1] = color[1];
        o[2] = color[2];
        o[3] = color[3];
        o[4] = shade;
        o[5] = density;
    }
}

// ---------------------------------------------------------------- camera

// sampler: 0 = none (pinhole ctor), 1 = circular, 2 = hexagonal(hex_ratio)
void *pth_camera_new(const float *origin, const float *look_at, const float *up, float focal_length, float height, float aspect_ratio,
                     float aperture_width, float aperture_height, int sampler, float hex_ratio, float focal_plane_dist) {
    if(sampler == 0) {
        return new Camera(v3(origin), v3(look_at), v3(up), focal_length, height, aspect_ratio);
    }
    std::unique_ptr<ApertureSampler> aperture;
    if(sampler == 1) {
        aperture = std::make_unique<CircularApertureSampler>();
    }
    else {
        aperture = std::make_unique<HexagonalApertureSampler>(hex_ratio);
    }
    return new Camera(v3(origin), v3(look_at), v3(up), focal_length, height, aspect_ratio, aperture_width, aperture_height, std::move(aperture),
                      focal_plane_dist);
}

void pth_camera_free(void *camera) {
    delete static_cast<Camera *>(camera);
}

// xy: 2 floats per ray (camera-space centre), one RandomEngine(seed) per ray; out: 6 floats per ray
void pth_camera_shoot(void *camera, long count, const float *xy, float pixel_width, float pixel_height, const uint64_t *seeds, float *out) {
    auto *cam = static_cast<Camera *>(camera);
    for(long i = 0; i < count; i++) {
        RandomEngine re(seeds[i]);
        Ray ray = cam->shootRay(xy[2 * i], xy[2 * i + 1], pixel_width, pixel_height, re);
        for(int k = 0; k < 3; k++) {
            out[6 * i + k] = ray.origin[k];
            out[6 * i + 3 + k] = ray.dir[k];
        }
    }
}

// ---------------------------------------------------------------- render

// The per-(pixel, sample) oracle: processItem on a 1x1 WorkItem with min = max = 1 spp and RandomEngine(seed)
// returns exactly one impl::getSample.  pixels: 2 ints each; out: 4 floats each.
void pth_render_samples(void *scene, void *camera, int width, int height, float epsilon, long count, const int *pixels, const uint64_t *seeds,
                        float *out) {
    auto *box = static_cast<SceneBox *>(scene);
    auto *cam = static_cast<Camera *>(camera);
    RenderOptions options{width, height, 1, 1, epsilon};
    FrameRenderJob job{*cam, *box->scene, options};
#ifdef PATHTRACE_B200
    ptb::renderSamples(job, static_cast<size_t>(count), pixels, seeds, out);
#else
    for(long i = 0; i < count; i++) {
        WorkItem item(&job, pixels[2 * i], pixels[2 * i + 1], 1, 1);
        RandomEngine re(seeds[i]);
        Image<> tile = processItem(item, re);
        auto value = tile(0, 0);
        for(int k = 0; k < 4; k++) {
            out[4 * i + k] = value[k];
        }
    }
#endif
}

// processItem on an arbitrary tile with RandomEngine(seed); out: 4 floats per pixel, row-major, tile-sized
void pth_process_item(void *scene, void *camera, int width, int height, int min_samples, int max_samples, float epsilon, int offset_x, int offset_y,
                      int tile_width, int tile_height, uint64_t seed, float *out) {
    auto *box = static_cast<SceneBox *>(scene);
    auto *cam = static_cast<Camera *>(camera);
    RenderOptions options{width, height, min_samples, max_samples, epsilon};
    FrameRenderJob job{*cam, *box->scene, options};
    WorkItem item(&job, offset_x, offset_y, tile_width, tile_height);
    RandomEngine re(seed);
    Image<> tile = processItem(item, re);
    std::memcpy(out, tile.data(), sizeof(float) * 4 * tile.size());
}

// processJob; out: 4 floats per pixel, row-major.  Returns the number of progress callbacks observed;
// *monotonic is set to 0 if the callback's completed count ever failed to increase by exactly one.
int pth_process_job(void *scene, void *camera, int width, int height, int min_samples, int max_samples, float epsilon, int worker_count, float *out,
                    int *total_tiles, int *monotonic) {
    auto *box = static_cast<SceneBox *>(scene);
    auto *cam = static_cast<Camera *>(camera);
    RenderOptions options{width, height, min_samples, max_samples, epsilon};
    FrameRenderJob job{*cam, *box->scene, options};

    int calls = 0;
    int last = 0;
    int total = 0;
    bool ok = true;
    const auto start = std::chrono::steady_clock::now();
    double first_callback = -1.0;
    double middle_callback = -1.0;
    auto callback = [&](int completed, int tiles) {
        calls++;
        ok = ok && (completed == last + 1);
        last = completed;
        total = tiles;
        const double now = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        if(first_callback < 0.0) {
            first_callback = now;
        }
        if(middle_callback < 0.0 && 2 * completed >= tiles) {
            middle_callback = now;
        }
    };
    Image<> image = processJob(job, callback, worker_count);
    const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    g_last_job_timeline[0] = elapsed;
    g_last_job_timeline[1] = first_callback;
    g_last_job_timeline[2] = middle_callback;
    if(image.size() > 0) {
        std::memcpy(out, image.data(), sizeof(float) * 4 * image.size());
    }
    if(total_tiles != nullptr) {
        *total_tiles = total;
    }
    if(monotonic != nullptr) {
        *monotonic = ok ? 1 : 0;
    }
    return calls;
}

// seconds: [0] duration of the last pth_process_job's processJob call, [1] when its first progress callback fired,
// [2] when the callback for half of the tiles fired (both measured from the start of the call; -1 = never)
void pth_last_job_timeline(double *out3) {
    out3[0] = g_last_job_timeline[0];
    out3[1] = g_last_job_timeline[1];
    out3[2] = g_last_job_timeline[2];
}

// ---------------------------------------------------------------- post-processing (host-side API kept as is)

// mode: 0 = toneMap, 1 = gammaCorrect(gamma), 2 = postProcess; in place on 4 floats per pixel
void pth_post_process(int mode, int width, int height, float gamma, float *pixels) {
    Image<> image(width, height);
    std::memcpy(image.data(), pixels, sizeof(float) * 4 * image.size());
    switch(mode) {
        case 0:
            toneMap(image);
            break;
        case 1:
            gammaCorrect(image, gamma);
            break;
        default:
            postProcess(image);
            break;
    }
    std::memcpy(pixels, image.data(), sizeof(float) * 4 * image.size());
}

} // extern "C"

#ifdef PATHTRACE_B200
// b200 build only: the C-ABI handle of the device scene behind a C++ Scene, so that callers can mix the public C++
// API (scene construction) with direct C-ABI calls (device-resident rendering, statistics).
extern "C" void *pth_scene_device_handle(void *scene) {
    return static_cast<SceneBox *>(scene)->scene->deviceScene();
}
#endif

#ifdef PATHTRACE_B200
// b200 build only: switches the result-neutral query options of ptb::RenderControl (certified closest hits on the SAH
// hierarchy, any-hit shadow rays, zero-weight shadow rays skipped) for every later call of this process.
extern "C" void pth_set_fast_queries(int certified_closest, int any_hit_shadows, int skip_null_shadows) {
    ptb::RenderControl &control = ptb::renderControl();
    control.certified_closest = certified_closest != 0;
    control.any_hit_shadows = any_hit_shadows != 0;
    control.skip_null_shadows = skip_null_shadows != 0;
}

#include <ptb.h> // b200 build only: ptb_device_count

// b200 build only: the remaining knobs of ptb::RenderControl (negative = leave unchanged); returns the GPUs processJob
// will really use (devices capped by the devices present)
extern "C" int pth_set_devices(int devices) {
    ptb::RenderControl &control = ptb::renderControl();
    int present = 1;
    ptb_device_count(&present);
    if(devices > 0) {
        control.devices = devices;
    }
    return std::min(control.devices, std::max(present, 1));
}

extern "C" void pth_set_render_control(int max_depth, int relaxed_guard) {
    ptb::RenderControl &control = ptb::renderControl();
    if(max_depth >= 0) {
        control.max_depth = max_depth;
    }
    if(relaxed_guard >= 0) {
        control.relaxed_guard = relaxed_guard != 0;
    }
}

// b200 build only: which share of processJob's tile grid this process renders (ptb::RenderControl::shard_index / shard_count)
// and the job seed (0 = std::random_device per call).
extern "C" void pth_set_sharding(int shard_index, int shard_count, unsigned long long fixed_seed) {
    ptb::RenderControl &control = ptb::renderControl();
    control.shard_index = shard_index;
    control.shard_count = shard_count;
    control.fixed_seed = fixed_seed;
}
#endif

#ifdef PATHTRACE_B200
#include <PathTrace/image/image_io.h>
// b200 build only (the reference's image_io.cpp needs libpng, absent here): PNG encode -> decode round trip of an RGBA
// float image through io::writeRGBImage / io::readRGBImage.  Returns the encoded size in bytes, or -1 on failure.
extern "C" long pth_png_roundtrip(int width, int height, const float *pixels_in, float *pixels_out) {
    try {
        Image<> image(width, height);
        std::memcpy(image.data(), pixels_in, sizeof(float) * 4 * image.size());
        std::stringstream stream(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
        io::writeRGBImage(stream, image);
        const long bytes = static_cast<long>(stream.str().size());
        Image<> decoded = io::readRGBImage(stream);
        if(decoded.getWidth() != width || decoded.getHeight() != height) {
            return -1;
        }
        std::memcpy(pixels_out, decoded.data(), sizeof(float) * 4 * decoded.size());
        return bytes;
    }
    catch(const std::exception &) {
        return -1;
    }
}

// b200 build only: io::readRGBImage on raw bytes into out[capacity_pixels * 4]; returns 0 and the size on success, 1 for
// std::logic_error, 2 for anything else, 3 when the image does not fit
extern "C" int pth_png_decode(const unsigned char *bytes, long length, int *width, int *height, float *out, long capacity_pixels) {
    try {
        std::stringstream stream(std::string(reinterpret_cast<const char *>(bytes), static_cast<size_t>(length)), std::ios_base::in | std::ios_base::binary);
        Image<> decoded = io::readRGBImage(stream);
        *width = decoded.getWidth();
        *height = decoded.getHeight();
        if(static_cast<long>(decoded.size()) > capacity_pixels) {
            return 3;
        }
        std::memcpy(out, decoded.data(), sizeof(float) * 4 * decoded.size());
        return 0;
    }
    catch(const std::logic_error &) {
        return 1;
    }
    catch(...) {
        return 2;
    }
}

// b200 build only: io::readRGBImage on raw bytes; 0 = decoded, 1 = std::logic_error (the documented failure), 2 = anything else
extern "C" int pth_png_decode_status(const unsigned char *bytes, long length) {
    try {
        std::stringstream stream(std::string(reinterpret_cast<const char *>(bytes), static_cast<size_t>(length)), std::ios_base::in | std::ios_base::binary);
        io::readRGBImage(stream);
        return 0;
    }
    catch(const std::logic_error &) {
        return 1;
    }
    catch(...) {
        return 2;
    }
}
#endif
