// Scene: lowering of the virtual object graph into the C-ABI's POD tables and forwarding of queries to the GPU.
//
// The reference's Scene constructor (src/scene/scene.cpp:153-181) wraps objects into an AABB pointer tree, registers
// emissive objects and builds a CDF; here those steps happen inside ptb_scene_create on flat arrays, and this file
// only translates types:
//   Triangle / Sphere / NullObject                      -> ptb_prim            (insertion order = primitive index)
//   ConstantMaterialHandler{ConstantMaterial, BSDF}     -> ptb_material        (de-duplicated per handler object)
//   PointLightSource                                    -> ptb_point_light
// Anything else is a user subclass whose virtuals would have to run on the CPU; north_star forbids a CPU fallback, so
// construction fails with std::logic_error.
#include "device.h"

#include <PathTrace/scene/scene.h>
#include <PathTrace/worker.h>

#include <algorithm>
#include <memory>
#include <atomic>
#include <thread>
#include <cstdio>
#include <stdexcept>
#include <typeinfo>
#include <unordered_map>
#include <utility>

Scene::Scene(std::vector<std::unique_ptr<Object>> &&objects_in, std::vector<std::unique_ptr<LightSource>> &&lights_in) :
  objects(std::move(objects_in)), light_sources(std::move(lights_in)) {
    // (not a std::vector: its constructor would zero 88 bytes per object on one core before the workers overwrite them)
    std::unique_ptr<ptb_prim[]> prims(new ptb_prim[objects.size()]);
    std::vector<ptb_material> materials;
    std::unordered_map<const MaterialHandler *, uint32_t> material_index;

    // Pass 1, in parallel for large scenes: every object's geometry and its material handler (virtual calls and
    // dynamic_casts per object -- 0.2 s on one core for a million triangles).  Pass 2, serial and cheap: material indices
    // in order of first appearance, so that the tables do not depend on the thread count.
    const std::size_t n = objects.size();
    std::vector<const MaterialHandler *> handlers(n);
    std::atomic<std::size_t> first_bad{n};
    auto lower_range = [&](std::size_t begin, std::size_t end) {
        for(std::size_t i = begin; i < end; i++) {
            const Object &object = *objects[i];
            if(!ptb::host::lowerObject(object, prims[i])) {
                std::size_t seen = first_bad.load();
                while(i < seen && !first_bad.compare_exchange_weak(seen, i)) {
                }
                return;
            }
            handlers[i] = object.getMaterialHandler();
        }
    };
    const std::size_t workers = n >= (1U << 16) ? std::min<std::size_t>(std::max(1U, std::thread::hardware_concurrency()), 16) : 1;
    if(workers > 1) {
        std::vector<std::thread> pool;
        const std::size_t chunk = (n + workers - 1) / workers;
        for(std::size_t w = 0; w < workers; w++) {
            pool.emplace_back(lower_range, std::min(n, w * chunk), std::min(n, (w + 1) * chunk));
        }
        for(std::thread &t : pool) {
            t.join();
        }
    }
    else {
        lower_range(0, n);
    }
    if(first_bad.load() < n) {
        const std::size_t i = first_bad.load();
        const Object &object = *objects[i];
        throw std::logic_error(std::string("PathTrace (B200): object ") + std::to_string(i) + " has type " + typeid(object).name() +
                               ", which the GPU scene cannot represent (supported: Triangle, Sphere, NullObject)");
    }

    const MaterialHandler *last_handler = nullptr;
    uint32_t last_index = 0;
    for(std::size_t i = 0; i < n; i++) {
        const MaterialHandler *handler = handlers[i];
        if(handler == last_handler && i > 0) {
            prims[i].material = last_index;
            continue;
        }
        auto found = material_index.find(handler);
        if(found == material_index.end()) {
            const auto *constant = dynamic_cast<const ConstantMaterialHandler *>(handler);
            const Material *material = constant != nullptr ? constant->getMaterial(vec3<float>{}) : nullptr;
            const BSDF *bsdf = constant != nullptr ? constant->getBSDF(vec3<float>{}) : nullptr;
            ptb_material pod{};
            if(material == nullptr || bsdf == nullptr || dynamic_cast<const ConstantMaterial *>(material) == nullptr ||
               !ptb::host::lowerMaterial(*material, *bsdf, vec3<float>{}, pod)) {
                throw std::logic_error(std::string("PathTrace (B200): object ") + std::to_string(i) +
                                       " uses a material handler, material or BSDF the GPU scene cannot represent (supported: "
                                       "ConstantMaterialHandler with ConstantMaterial and LambertianBRDF, GlassBDF or MirrorBRDF)");
            }
            found = material_index.emplace(handler, static_cast<uint32_t>(materials.size())).first;
            materials.push_back(pod);
        }
        prims[i].material = found->second;
        last_handler = handler;
        last_index = found->second;
    }

    std::vector<ptb_point_light> lights(light_sources.size());
    for(std::size_t i = 0; i < light_sources.size(); i++) {
        const auto *point = dynamic_cast<const PointLightSource *>(light_sources[i].get());
        if(point == nullptr) {
            throw std::logic_error(std::string("PathTrace (B200): light ") + std::to_string(i) +
                                   " is not a PointLightSource; user light subclasses cannot run on the GPU");
        }
        const auto pos = point->getPosition();
        const auto rgba = point->getSpectrum(Ray{}).getColor();
        for(int k = 0; k < 3; k++) {
            lights[i].pos[k] = pos[k];
        }
        for(int k = 0; k < 4; k++) {
            lights[i].rgba[k] = rgba[k];
        }
    }
    if(materials.empty()) {
        materials.push_back(ptb_material{{1, 1, 1, 1}, {0, 0, 0, 0}, 1.0F, PTB_BSDF_LAMBERT, 0, 0});
    }

    ptb_scene_desc desc{};
    desc.prims = prims.get();
    desc.n_prims = n;
    desc.materials = materials.data();
    desc.n_materials = static_cast<uint32_t>(materials.size());
    desc.lights = lights.data();
    desc.n_lights = static_cast<uint32_t>(lights.size());
    desc.bvh_mode = PTB_BVH_REFERENCE;
    ptb::host::check(ptb_scene_create(ptb::host::defaultContext(), &desc, &device_scene), "Scene upload");
}

Scene::~Scene() {
    for(ptb_scene *replica : replicas) {
        ptb_scene_destroy(replica);
    }
    ptb_scene_destroy(device_scene);
}

Scene::Scene(Scene &&other) noexcept :
  objects(std::move(other.objects)), light_sources(std::move(other.light_sources)), device_scene(std::exchange(other.device_scene, nullptr)),
  replicas(std::move(other.replicas)) {
    other.replicas.clear();
}

Scene &Scene::operator=(Scene &&other) noexcept {
    if(this != &other) {
        for(ptb_scene *replica : replicas) {
            ptb_scene_destroy(replica);
        }
        ptb_scene_destroy(device_scene);
        objects = std::move(other.objects);
        light_sources = std::move(other.light_sources);
        device_scene = std::exchange(other.device_scene, nullptr);
        replicas = std::move(other.replicas);
        other.replicas.clear();
    }
    return *this;
}

std::vector<ptb_scene *> Scene::deviceScenes(int count) const {
    std::vector<ptb_scene *> all{device_scene};
    int home = 0;
    ptb::host::check(ptb_context_device(ptb::host::defaultContext(), &home), "context device");
    int available = 0;
    ptb::host::check(ptb_device_count(&available), "device count");
    count = std::min(count, available);
    // replica k lives on the k-th device after the home device (wrapping around)
    for(int k = 1; k < count; k++) {
        if(static_cast<std::size_t>(k) > replicas.size()) {
            ptb_scene *copy = nullptr;
            ptb::host::check(ptb_scene_clone(device_scene, ptb::host::contextFor((home + k) % available), &copy), "scene replica");
            replicas.push_back(copy);
        }
        all.push_back(replicas[static_cast<std::size_t>(k) - 1]);
    }
    return all;
}

void Scene::getIntersections(const Ray *rays, std::size_t count, float *t_out, const Object **objects_out) const noexcept {
    if(count == 0) {
        return;
    }
    static_assert(sizeof(Ray) == 6 * sizeof(float), "Ray must be six packed floats");
    std::vector<int32_t> prim(count);
    const int status = ptb_intersect(device_scene, reinterpret_cast<const float *>(rays), count, t_out, prim.data(),
                                     ptb::renderControl().certified_closest ? PTB_FLAG_CERTIFIED_CLOSEST : 0U, nullptr);
    if(!ptb::host::ok(status, "Scene::getIntersections")) {
        for(std::size_t i = 0; i < count; i++) {
            t_out[i] = -1.0F;
            if(objects_out != nullptr) {
                objects_out[i] = nullptr;
            }
        }
        return;
    }
    if(objects_out != nullptr) {
        for(std::size_t i = 0; i < count; i++) {
            objects_out[i] = prim[i] >= 0 ? objects[static_cast<std::size_t>(prim[i])].get() : nullptr;
        }
    }
}

std::tuple<float, const Object *> Scene::getIntersection(const Ray &ray) const noexcept {
    float t = -1.0F;
    const Object *object = nullptr;
    getIntersections(&ray, 1, &t, &object);
    return std::make_tuple(t, object);
}

std::vector<std::tuple<vec3<float>, Spectrum, float>> Scene::sampleLights(vec3<float> pos, vec3<float> /*n*/, RandomEngine &re) const noexcept {
    std::vector<std::tuple<vec3<float>, Spectrum, float>> lights;
    ptb_scene_info info{};
    if(ptb_scene_get_info(device_scene, &info) != PTB_OK) {
        return lights;
    }
    const uint32_t capacity = info.n_lights + info.object_sample_count;
    std::vector<float> raw(8 * static_cast<std::size_t>(capacity) + 8);
    const float p[3] = {pos[0], pos[1], pos[2]};
    uint64_t state = re.state();
    uint32_t produced = 0;
    if(!ptb::host::ok(ptb_sample_lights(device_scene, p, &state, capacity, raw.data(), &produced), "Scene::sampleLights")) {
        return lights;
    }
    re.setState(state);
    lights.reserve(produced);
    for(uint32_t i = 0; i < produced && i < capacity; i++) {
        const float *o = raw.data() + 8 * static_cast<std::size_t>(i);
        lights.emplace_back(vec3<float>{o[0], o[1], o[2]}, Spectrum(Color<float>(o[3], o[4], o[5], o[6])), o[7]);
    }
    return lights;
}
