// BSDF host objects.  The arithmetic lives on the device (cpupathtrace_b200/csrc/shading.cuh); a direct host call is
// a unit launch of those functions (ptb_bsdf_propagate / ptb_bsdf_spectrum).
#include "device.h"

#include <PathTrace/scene/propagation.h>

#include <cstdio>

namespace {

    std::tuple<Ray, float, float> unitPropagate(const BSDF &bsdf, Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                                const Material *material) noexcept {
        Ray out_ray{pos, ray.dir};
        float factor = 0.0F;
        float pd = 0.0F;
        ptb_material pod;
        if(material != nullptr && ptb::host::lowerMaterial(*material, bsdf, pos, pod)) {
            const float in[9] = {ray.dir[0], ray.dir[1], ray.dir[2], pos[0], pos[1], pos[2], normal[0], normal[1], normal[2]};
            float out[8] = {};
            uint64_t state = re.state();
            try {
                if(ptb::host::ok(ptb_bsdf_propagate(ptb::host::defaultContext(), &pod, epsilon, 1, in, &state, out), "BSDF::propagateRay")) {
                    re.setState(state);
                    out_ray = Ray{vec3<float>{out[0], out[1], out[2]}, vec3<float>{out[3], out[4], out[5]}};
                    factor = out[6];
                    pd = out[7];
                }
            }
            catch(const std::exception &e) {
                std::fprintf(stderr, "%s\n", e.what());
            }
        }
        return std::make_tuple(out_ray, factor, pd);
    }

    std::tuple<Spectrum, float, float> unitSpectrum(const BSDF &bsdf, Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light,
                                                    const Material *material, bool synthetic) noexcept {
        Spectrum spectrum;
        float shade = 0.0F;
        float pd = 0.0F;
        ptb_material pod;
        if(material != nullptr && ptb::host::lowerMaterial(*material, bsdf, pos, pod)) {
            const auto l = light.getColor();
            const float in[13] = {from_camera.dir[0], from_camera.dir[1], from_camera.dir[2], to_light.dir[0], to_light.dir[1], to_light.dir[2], normal[0],
                                  normal[1],          normal[2],          l[0],            l[1],            l[2],            l[3]};
            float out[6] = {};
            try {
                if(ptb::host::ok(ptb_bsdf_spectrum(ptb::host::defaultContext(), &pod, synthetic ? 1U : 0U, 1, in, out), "BSDF::getSpectrum")) {
                    spectrum = Spectrum(Color<float>(out[0], out[1], out[2], out[3]));
                    shade = out[4];
                    pd = out[5];
                }
            }
            catch(const std::exception &e) {
                std::fprintf(stderr, "%s\n", e.what());
            }
        }
        return std::make_tuple(spectrum, shade, pd);
    }

}

LambertianBRDF::LambertianBRDF() noexcept = default;

std::tuple<Ray, float, float> LambertianBRDF::propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                                           const Material *material) const noexcept {
    return unitPropagate(*this, ray, pos, normal, epsilon, re, material);
}

std::tuple<Spectrum, float, float> LambertianBRDF::getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                               const Material *material, bool synthetic) const noexcept {
    return unitSpectrum(*this, from_camera, to_light, pos, normal, light_spectrum, material, synthetic);
}

GlassBDF::GlassBDF() noexcept = default;

std::tuple<Ray, float, float> GlassBDF::propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                                     const Material *material) const noexcept {
    return unitPropagate(*this, ray, pos, normal, epsilon, re, material);
}

std::tuple<Spectrum, float, float> GlassBDF::getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                         const Material *material, bool synthetic) const noexcept {
    return unitSpectrum(*this, from_camera, to_light, pos, normal, light_spectrum, material, synthetic);
}

MirrorBRDF::MirrorBRDF(bool one_way) noexcept : one_way(one_way) {}

std::tuple<Ray, float, float> MirrorBRDF::propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                                       const Material *material) const noexcept {
    return unitPropagate(*this, ray, pos, normal, epsilon, re, material);
}

std::tuple<Spectrum, float, float> MirrorBRDF::getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                           const Material *material, bool synthetic) const noexcept {
    return unitSpectrum(*this, from_camera, to_light, pos, normal, light_spectrum, material, synthetic);
}
