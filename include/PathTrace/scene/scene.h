// The scene container (API of the reference's include/PathTrace/scene/scene.h).
//
// Construction lowers the virtual object graph into the POD tables of the C-ABI (ptb_scene_desc), builds the
// reference-topology BVH and uploads everything to the GPU once; all queries run there.
#ifndef PATHTRACE_SCENE_H
#define PATHTRACE_SCENE_H

#include <PathTrace/base.h>
#include <PathTrace/scene/bounding_box.h>
#include <PathTrace/scene/light.h>
#include <PathTrace/scene/object.h>

#include <cstddef>
#include <memory>
#include <tuple>
#include <utility>
#include <vector>

struct ptb_scene;

class Scene {
  private:
    std::vector<std::unique_ptr<Object>> objects;
    std::vector<std::unique_ptr<LightSource>> light_sources;
    ptb_scene *device_scene = nullptr;
    //! copies of the device scene on further GPUs of this process (B200 extension, see deviceScenes())
    mutable std::vector<ptb_scene *> replicas;

  public:
    /**
     * Takes ownership of the objects and lights and uploads the scene.
     *
     * @throw std::logic_error if an object, material handler, material, BSDF or light is not one of the library's own
     *  concrete types (user subclasses cannot run on the GPU and there is no CPU fallback), or std::runtime_error if
     *  no CUDA device is available
     */
    Scene(std::vector<std::unique_ptr<Object>> &&objects, std::vector<std::unique_ptr<LightSource>> &&light_sources);
    ~Scene();

    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    Scene(Scene &&other) noexcept;
    Scene &operator=(Scene &&other) noexcept;

    //! closest hit: (distance or negative, object or nullptr)
    std::tuple<float, const Object *> getIntersection(const Ray &ray) const noexcept;

    //! samples explicit lights and emissive geometry as seen from pos: (position, spectrum, probability density)
    std::vector<std::tuple<vec3<float>, Spectrum, float>> sampleLights(vec3<float> pos, vec3<float> n, RandomEngine &re) const noexcept;

    // ---- B200 extensions

    //! closest hit for a batch of rays in one launch; objects_out may be nullptr
    void getIntersections(const Ray *rays, std::size_t count, float *t_out, const Object **objects_out) const noexcept;

    //! the device-resident scene (C-ABI handle, owned by this Scene)
    ptb_scene *deviceScene() const noexcept { return device_scene; }

    //! the device scene followed by copies of it on up to count - 1 further GPUs of this process (created on first use by
    //! device-to-device copies, no BVH rebuild; owned by this Scene): what processJob renders on when
    //! ptb::RenderControl::devices > 1.  Not thread-safe against concurrent calls on the same Scene.
    std::vector<ptb_scene *> deviceScenes(int count) const;

    std::size_t objectCount() const noexcept { return objects.size(); }
};

#endif /* PATHTRACE_SCENE_H */
