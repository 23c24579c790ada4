#include "device.h"

#include <PathTrace/scene/propagation.h>

#include <cstdio>
#include <map>
#include <mutex>
#include <stdexcept>

namespace ptb::host {

    ptb_context *defaultContext() {
        static std::mutex mutex;
        static ptb_context *context = nullptr;
        std::lock_guard<std::mutex> lock(mutex);
        if(context == nullptr) {
            const int status = ptb_context_create(-1, &context);
            if(status != PTB_OK) {
                context = nullptr;
                throw std::runtime_error(std::string("PathTrace (B200): cannot create a device context: ") + ptb_last_error());
            }
        }
        return context;
    }

    ptb_context *contextFor(int device) {
        ptb_context *home = defaultContext();
        int home_device = 0;
        if(ptb_context_device(home, &home_device) == PTB_OK && home_device == device) {
            return home;
        }
        static std::mutex mutex;
        static std::map<int, ptb_context *> contexts;
        std::lock_guard<std::mutex> lock(mutex);
        auto found = contexts.find(device);
        if(found == contexts.end()) {
            ptb_context *context = nullptr;
            if(ptb_context_create(device, &context) != PTB_OK) {
                throw std::runtime_error(std::string("PathTrace (B200): cannot create a context on device ") + std::to_string(device) + ": " + ptb_last_error());
            }
            found = contexts.emplace(device, context).first;
        }
        return found->second;
    }

    void check(int status, const char *what) {
        if(status != PTB_OK) {
            throw std::runtime_error(std::string("PathTrace (B200): ") + what + " failed: " + ptb_last_error());
        }
    }

    bool ok(int status, const char *what) noexcept {
        if(status != PTB_OK) {
            std::fprintf(stderr, "PathTrace (B200): %s failed: %s\n", what, ptb_last_error());
            return false;
        }
        return true;
    }

    bool lowerObject(const Object &object, ptb_prim &out) noexcept {
        out = ptb_prim{};
        if(const auto *triangle = dynamic_cast<const Triangle *>(&object)) {
            out.kind = PTB_PRIM_TRIANGLE;
            out.cull_backface = triangle->cullsBackface() ? 1U : 0U;
            const vec3<float> *fields[6] = {&triangle->a, &triangle->b, &triangle->c, &triangle->normal_a, &triangle->normal_b, &triangle->normal_c};
            for(int f = 0; f < 6; f++) {
                for(int k = 0; k < 3; k++) {
                    out.p[3 * f + k] = (*fields[f])[k];
                }
            }
            return true;
        }
        if(const auto *sphere = dynamic_cast<const Sphere *>(&object)) {
            out.kind = PTB_PRIM_SPHERE;
            const auto origin = sphere->getOrigin();
            out.p[0] = origin[0];
            out.p[1] = origin[1];
            out.p[2] = origin[2];
            out.p[3] = sphere->getRadius();
            return true;
        }
        if(dynamic_cast<const NullObject *>(&object) != nullptr) {
            out.kind = PTB_PRIM_NULL;
            return true;
        }
        return false;
    }

    bool lowerMaterial(const Material &material, const BSDF &bsdf, vec3<float> pos, ptb_material &out) noexcept {
        out = ptb_material{};
        if(dynamic_cast<const LambertianBRDF *>(&bsdf) != nullptr) {
            out.bsdf = PTB_BSDF_LAMBERT;
        }
        else if(dynamic_cast<const GlassBDF *>(&bsdf) != nullptr) {
            out.bsdf = PTB_BSDF_GLASS;
        }
        else if(const auto *mirror = dynamic_cast<const MirrorBRDF *>(&bsdf)) {
            out.bsdf = PTB_BSDF_MIRROR;
            out.one_way = mirror->isOneWay() ? 1U : 0U;
        }
        else {
            return false;
        }
        const auto diffuse = material.getDiffuseColor(pos);
        const auto emission = material.getEmission(Ray{pos, vec3<float>{0.0F, 0.0F, 1.0F}}, pos).getColor();
        for(int k = 0; k < 4; k++) {
            out.diffuse[k] = diffuse[k];
            out.emission[k] = emission[k];
        }
        out.refractive_index = material.getRefractiveIndex(pos);
        return true;
    }

}
