// Scattering functions (API of the reference's include/PathTrace/scene/propagation.h).
//
// The three concrete BSDFs are evaluated on the GPU: inside the render path by the wavefront shade kernel, and for a
// direct host call of propagateRay / getSpectrum by a unit launch of the same device functions
// (ptb_bsdf_propagate / ptb_bsdf_spectrum), so that both give identical numbers.
#ifndef PATHTRACE_PROPAGATION_H
#define PATHTRACE_PROPAGATION_H

#include <PathTrace/base.h>
#include <PathTrace/scene/light.h>
#include <PathTrace/scene/material.h>

#include <array>
#include <memory>
#include <tuple>

class BSDF {
  public:
    virtual ~BSDF() = default;

    //! @return outgoing ray, radiance factor, probability density
    virtual std::tuple<Ray, float, float> propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                                       const Material *material) const noexcept = 0;

    //! @return incoming spectrum, shading factor, probability density of the ray pair
    virtual std::tuple<Spectrum, float, float> getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                           const Material *material, bool synthetic = false) const noexcept = 0;
};

//! ideal diffuse reflector, cosine-weighted sampling
class LambertianBRDF : public BSDF {
  public:
    virtual ~LambertianBRDF() = default;
    LambertianBRDF() noexcept;

    std::tuple<Ray, float, float> propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                               const Material *material) const noexcept override;
    std::tuple<Spectrum, float, float> getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                   const Material *material, bool synthetic = false) const noexcept override;
};

//! smooth dielectric: Fresnel-weighted choice between specular reflection and refraction
class GlassBDF : public BSDF {
  public:
    virtual ~GlassBDF() = default;
    GlassBDF() noexcept;

    std::tuple<Ray, float, float> propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                               const Material *material) const noexcept override;
    std::tuple<Spectrum, float, float> getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                   const Material *material, bool synthetic = false) const noexcept override;
};

//! perfect mirror, optionally transparent from the back
class MirrorBRDF : public BSDF {
  private:
    bool one_way;

  public:
    virtual ~MirrorBRDF() = default;
    MirrorBRDF(bool one_way = false) noexcept;

    std::tuple<Ray, float, float> propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                               const Material *material) const noexcept override;
    std::tuple<Spectrum, float, float> getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                   const Material *material, bool synthetic = false) const noexcept override;

    // B200 extension: read access for scene lowering
    bool isOneWay() const noexcept { return one_way; }
};

//! Declared for source compatibility only: the reference declares this template but never defines or uses it.
template<class... Ts>
class CombinedBSDF : public BSDF {
  private:
    std::tuple<Ts...> components;
    std::array<float, sizeof...(Ts)> weights;
    std::array<float, sizeof...(Ts)> probabilities;

  public:
    virtual ~CombinedBSDF() = default;
    CombinedBSDF(Ts... components, std::array<float, sizeof...(Ts)> weights) noexcept;

    std::tuple<Ray, float, float> propagateRay(Ray ray, vec3<float> pos, vec3<float> normal, float epsilon, RandomEngine &re,
                                               const Material *material) const noexcept override;
    std::tuple<Spectrum, float, float> getSpectrum(Ray from_camera, Ray to_light, vec3<float> pos, vec3<float> normal, Spectrum light_spectrum,
                                                   const Material *material, bool synthetic = false) const noexcept override;
};

#endif /* PATHTRACE_PROPAGATION_H */
