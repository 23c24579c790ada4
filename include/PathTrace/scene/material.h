// Surface materials (API of the reference's include/PathTrace/scene/material.h).
// Only ConstantMaterial can be lowered to the device material table (ptb_material).
#ifndef PATHTRACE_MATERIAL_H
#define PATHTRACE_MATERIAL_H

#include <PathTrace/base.h>
#include <PathTrace/scene/light.h>
#include <PathTrace/util/color.h>
#include <PathTrace/util/vector.h>

class Material {
  public:
    virtual ~Material() = default;

    virtual Color<float> getDiffuseColor(vec3<float> pos) const noexcept = 0;
    //! defaults to white
    virtual Color<float> getSpecularColor(vec3<float> pos) const noexcept;
    //! defaults to 1
    virtual float getRefractiveIndex(vec3<float> pos) const noexcept;
    //! defaults to no emission
    virtual Spectrum getEmission(Ray ray, vec3<float> pos) const noexcept;
    //! heuristic whole-object emission used to register emissive geometry as light sources; defaults to none
    virtual Spectrum probeEmission() const noexcept;
};

//! Position-independent material
class ConstantMaterial final : public Material {
  private:
    Color<float> diffuse_color;
    float refractive_index;
    Spectrum emission;

  public:
    virtual ~ConstantMaterial() = default;
    ConstantMaterial(Color<float> diffuse_color = Color<float>(1.0F, 1.0F, 1.0F, 1.0F), float refractive_index = 1.0F, Spectrum emission = {}) noexcept;

    Color<float> getDiffuseColor(vec3<float> pos) const noexcept override;
    float getRefractiveIndex(vec3<float> pos) const noexcept override;
    Spectrum getEmission(Ray ray, vec3<float> pos) const noexcept override;
    Spectrum probeEmission() const noexcept override;
};

#endif // PATHTRACE_MATERIAL_H
