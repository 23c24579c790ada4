// Objects and material handlers.  Construction, bounds and areas are host-side set-up; the per-ray virtuals
// (getIntersection, getSurfaceNormal, sampleSurface) are answered by unit launches of the device functions the
// traversal / shade kernels use (ptb_prim_intersect, ptb_prim_normal, ptb_prim_sample), so a direct call returns
// exactly what the render path computes.
#include "device.h"

#include <PathTrace/scene/bounding_box.h>
#include <PathTrace/scene/object.h>

#include <cmath>
#include <limits>
#include <utility>

namespace {

    const std::shared_ptr<MaterialHandler> &libraryDefaultHandler() {
        static const std::shared_ptr<Material> material = std::make_shared<ConstantMaterial>(Color<float>(1.0F, 1.0F, 1.0F, 1.0F));
        static const std::shared_ptr<BSDF> bsdf = std::make_shared<LambertianBRDF>();
        static const std::shared_ptr<MaterialHandler> handler = std::make_shared<ConstantMaterialHandler>(material, bsdf);
        return handler;
    }

    float unitIntersect(const Object &object, const Ray &ray) noexcept {
        ptb_prim prim;
        if(!ptb::host::lowerObject(object, prim)) {
            return -1.0F;
        }
        const float packed[6] = {ray.origin[0], ray.origin[1], ray.origin[2], ray.dir[0], ray.dir[1], ray.dir[2]};
        float t = -1.0F;
        try {
            ptb::host::ok(ptb_prim_intersect(ptb::host::defaultContext(), &prim, 1, packed, &t), "Object::getIntersection");
        }
        catch(const std::exception &e) {
            std::fprintf(stderr, "%s\n", e.what());
        }
        return t;
    }

    vec3<float> unitNormal(const Object &object, vec3<float> pos) noexcept {
        ptb_prim prim;
        float n[3] = {0.0F, 1.0F, 0.0F};
        if(!ptb::host::lowerObject(object, prim)) {
            return {n[0], n[1], n[2]};
        }
        const float p[3] = {pos[0], pos[1], pos[2]};
        try {
            ptb::host::ok(ptb_prim_normal(ptb::host::defaultContext(), &prim, 1, p, n), "Object::getSurfaceNormal");
        }
        catch(const std::exception &e) {
            std::fprintf(stderr, "%s\n", e.what());
        }
        return {n[0], n[1], n[2]};
    }

    std::tuple<vec3<float>, float, bool> unitSample(const Object &object, RandomEngine &re) noexcept {
        ptb_prim prim;
        float out[5] = {0.0F, 0.0F, 0.0F, 0.0F, 0.0F};
        if(ptb::host::lowerObject(object, prim)) {
            uint64_t state = re.state();
            try {
                if(ptb::host::ok(ptb_prim_sample(ptb::host::defaultContext(), &prim, 1, &state, out), "Object::sampleSurface")) {
                    re.setState(state);
                }
            }
            catch(const std::exception &e) {
                std::fprintf(stderr, "%s\n", e.what());
            }
        }
        return std::make_tuple(vec3<float>{out[0], out[1], out[2]}, out[3], out[4] != 0.0F);
    }

}

const Material *MaterialHandler::probeMaterial() const noexcept {
    return libraryDefaultHandler()->getMaterial(vec3<float>{});
}

ConstantMaterialHandler::ConstantMaterialHandler(std::shared_ptr<Material> material, std::shared_ptr<BSDF> bsdf) :
  material(std::move(material)), bsdf(std::move(bsdf)) {}

const Material *ConstantMaterialHandler::probeMaterial() const noexcept {
    return material.get();
}

const Material *ConstantMaterialHandler::getMaterial(vec3<float> /*pos*/) const noexcept {
    return material.get();
}

const BSDF *ConstantMaterialHandler::getBSDF(vec3<float> /*pos*/) const noexcept {
    return bsdf.get();
}

Object::Object() : material_handler(libraryDefaultHandler()) {}

Object::Object(std::shared_ptr<MaterialHandler> material_handler) noexcept : material_handler(std::move(material_handler)) {}

const MaterialHandler *Object::getMaterialHandler() const noexcept {
    return material_handler.get();
}

void Object::setMaterialHandler(std::shared_ptr<MaterialHandler> handler) {
    material_handler = std::move(handler);
}

float Object::getSurfaceArea() const noexcept {
    return 0.0F;
}

std::tuple<vec3<float>, float, bool> Object::sampleSurface(RandomEngine & /*re*/) const noexcept {
    return std::make_tuple(vec3<float>{}, 0.0F, false);
}

// ---- NullObject

float NullObject::getIntersection(const Ray & /*ray*/) const noexcept {
    return -1.0F;
}

vec3<float> NullObject::getSurfaceNormal(vec3<float> /*pos*/) const noexcept {
    return {0.0F, 1.0F, 0.0F};
}

AABBArea NullObject::getBoundingVolume() const noexcept {
    return AABBArea{};
}

float NullObject::getSurfaceArea() const noexcept {
    return 0.0F;
}

// ---- Sphere

Sphere::Sphere(vec3<float> origin, float radius) : origin(origin), radius(radius), radius2(radius * radius) {}

float Sphere::getIntersection(const Ray &ray) const noexcept {
    return unitIntersect(*this, ray);
}

vec3<float> Sphere::getSurfaceNormal(vec3<float> pos) const noexcept {
    return unitNormal(*this, pos);
}

AABBArea Sphere::getBoundingVolume() const noexcept {
    const vec3<float> extent{radius, radius, radius};
    return {origin - extent, origin + extent};
}

float Sphere::getSurfaceArea() const noexcept {
    constexpr float pi = static_cast<float>(M_PI);
    return 4.0F * pi * radius2;
}

std::tuple<vec3<float>, float, bool> Sphere::sampleSurface(RandomEngine &re) const noexcept {
    return unitSample(*this, re);
}

// ---- Triangle

Triangle::Triangle(vec3<float> a, vec3<float> b, vec3<float> c, bool cull_backface) : a(a), b(b), c(c), cull_backface(cull_backface) {
    const vec3<float> face_normal = cross(b - a, c - a).normalize();
    normal_a = face_normal;
    normal_b = face_normal;
    normal_c = face_normal;
}

float Triangle::getIntersection(const Ray &ray) const noexcept {
    return unitIntersect(*this, ray);
}

vec3<float> Triangle::getSurfaceNormal(vec3<float> pos) const noexcept {
    return unitNormal(*this, pos);
}

AABBArea Triangle::getBoundingVolume() const noexcept {
    return {min(min(a, b), c), max(max(a, b), c)};
}

float Triangle::getSurfaceArea() const noexcept {
    return cross(b - a, c - a).getLength() / 2.0F;
}

std::tuple<vec3<float>, float, bool> Triangle::sampleSurface(RandomEngine &re) const noexcept {
    return unitSample(*this, re);
}
