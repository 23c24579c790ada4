// Scene geometry and its material binding (API of the reference's include/PathTrace/scene/object.h).
//
// Triangle, Sphere and NullObject are the primitives the device knows (ptb_prim); the per-ray virtuals below are
// answered by unit launches of the same device functions the traversal kernel uses.
#ifndef PATHTRACE_OBJECT_H
#define PATHTRACE_OBJECT_H

#include <PathTrace/base.h>
#include <PathTrace/scene/propagation.h>

#include <memory>
#include <tuple>

//! Supplies the Material and BSDF of an object
class MaterialHandler {
  public:
    virtual ~MaterialHandler() = default;

    //! typical material of the object, used to probe emissiveness; defaults to the library's white diffuse material
    virtual const Material *probeMaterial() const noexcept;
    virtual const Material *getMaterial(vec3<float> pos) const noexcept = 0;
    virtual const BSDF *getBSDF(vec3<float> pos) const noexcept = 0;
};

//! Position-independent material binding (the only kind that can be lowered to the device)
class ConstantMaterialHandler final : public MaterialHandler {
  private:
    std::shared_ptr<Material> material;
    std::shared_ptr<BSDF> bsdf;

  public:
    virtual ~ConstantMaterialHandler() = default;
    ConstantMaterialHandler(std::shared_ptr<Material> material, std::shared_ptr<BSDF> bsdf);

    const Material *probeMaterial() const noexcept override;
    const Material *getMaterial(vec3<float> pos) const noexcept override;
    const BSDF *getBSDF(vec3<float> pos) const noexcept override;
};

struct AABBArea;

class Object {
  private:
    std::shared_ptr<MaterialHandler> material_handler;

  public:
    virtual ~Object() = default;
    //! binds the library-wide default handler (white Lambertian)
    Object();
    Object(std::shared_ptr<MaterialHandler> material_handler) noexcept;

    //! distance to the first intersection along the ray, negative if none
    virtual float getIntersection(const Ray &ray) const noexcept = 0;
    //! unit surface normal at a surface point
    virtual vec3<float> getSurfaceNormal(vec3<float> pos) const noexcept = 0;

    const MaterialHandler *getMaterialHandler() const noexcept;
    void setMaterialHandler(std::shared_ptr<MaterialHandler> material_handler);

    virtual AABBArea getBoundingVolume() const noexcept = 0;
    //! front-face area; 0 by default
    virtual float getSurfaceArea() const noexcept;
    //! uniformly sampled surface point, its density, and whether back faces are culled
    virtual std::tuple<vec3<float>, float, bool> sampleSurface(RandomEngine &re) const noexcept;
};

class NullObject final : public Object {
  public:
    virtual ~NullObject() = default;
    NullObject() = default;

    float getIntersection(const Ray &ray) const noexcept override;
    vec3<float> getSurfaceNormal(vec3<float> pos) const noexcept override;
    AABBArea getBoundingVolume() const noexcept override;
    float getSurfaceArea() const noexcept override;
};

class Sphere final : public Object {
  private:
    vec3<float> origin;
    float radius;
    float radius2;

  public:
    virtual ~Sphere() = default;
    Sphere(vec3<float> origin, float radius);

    float getIntersection(const Ray &ray) const noexcept override;
    vec3<float> getSurfaceNormal(vec3<float> pos) const noexcept override;
    AABBArea getBoundingVolume() const noexcept override;
    float getSurfaceArea() const noexcept override;
    std::tuple<vec3<float>, float, bool> sampleSurface(RandomEngine &re) const noexcept override;

    // B200 extension: read access for scene lowering
    vec3<float> getOrigin() const noexcept { return origin; }
    float getRadius() const noexcept { return radius; }
};

class Triangle final : public Object {
  public:
    vec3<float> a;
    vec3<float> b;
    vec3<float> c;
    vec3<float> normal_a;
    vec3<float> normal_b;
    vec3<float> normal_c;

  private:
    bool cull_backface;

  public:
    virtual ~Triangle() noexcept = default;
    //! all three vertex normals start as the face normal of (a, b, c)
    Triangle(vec3<float> a, vec3<float> b, vec3<float> c, bool cull_backface = false);

    float getIntersection(const Ray &ray) const noexcept override;
    vec3<float> getSurfaceNormal(vec3<float> pos) const noexcept override;
    AABBArea getBoundingVolume() const noexcept override;
    float getSurfaceArea() const noexcept override;
    std::tuple<vec3<float>, float, bool> sampleSurface(RandomEngine &re) const noexcept override;

    // B200 extension: read access for scene lowering
    bool cullsBackface() const noexcept { return cull_backface; }
};

#endif /* PATHTRACE_OBJECT_H */
