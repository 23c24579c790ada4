// Light spectra and explicit light sources (API of the reference's include/PathTrace/scene/light.h).
#ifndef PATHTRACE_LIGHT_H
#define PATHTRACE_LIGHT_H

#include <PathTrace/base.h>
#include <PathTrace/util/color.h>

#include <tuple>
#include <utility>

//! RGBA radiance / reflectance value with channel-wise arithmetic
class Spectrum {
  private:
    Color<float> color;

  public:
    Spectrum(Color<float> color = {0.0F, 0.0F, 0.0F, 0.0F}) noexcept : color(color) {}

    Color<float> getColor() const noexcept { return color; }

    Spectrum operator+(Spectrum other) const noexcept { return {Color<float>(color + other.color)}; }
    Spectrum operator*(Spectrum other) const noexcept { return {Color<float>(color * other.color)}; }
    Spectrum operator*(float factor) const noexcept { return {Color<float>(color * factor)}; }
    Spectrum operator/(float divisor) const noexcept { return {Color<float>(color / divisor)}; }
};

//! A light that can be sampled from a surface point.  Only PointLightSource can be lowered to the device;
//! scenes holding other subclasses are rejected when the Scene is constructed.
class LightSource {
  public:
    virtual ~LightSource() = default;

    //! @return sampled position on the light and its probability density
    virtual std::tuple<vec3<float>, float> importanceSample(vec3<float> pos) const noexcept = 0;

    //! @return spectrum emitted along a ray pointing at the light
    virtual Spectrum getSpectrum(Ray ray) const noexcept = 0;
};

//! Isotropic point light
class PointLightSource final : public LightSource {
  private:
    vec3<float> pos;
    Spectrum spectrum;

  public:
    virtual ~PointLightSource() = default;
    PointLightSource(vec3<float> pos, Spectrum spectrum) noexcept : pos(pos), spectrum(spectrum) {}

    std::tuple<vec3<float>, float> importanceSample(vec3<float> /*from*/) const noexcept override { return std::make_tuple(pos, 1.0F); }
    Spectrum getSpectrum(Ray /*ray*/) const noexcept override { return spectrum; }

    // B200 extension: read access for scene lowering
    vec3<float> getPosition() const noexcept { return pos; }
};

#endif /* PATHTRACE_LIGHT_H */
