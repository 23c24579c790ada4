// Material defaults and ConstantMaterial (behaviour of the reference's src/scene/material.cpp).
#include <PathTrace/scene/material.h>

Color<float> Material::getSpecularColor(vec3<float> /*pos*/) const noexcept {
    return Color<float>(1.0F, 1.0F, 1.0F, 1.0F);
}

float Material::getRefractiveIndex(vec3<float> /*pos*/) const noexcept {
    return 1.0F;
}

Spectrum Material::getEmission(Ray /*ray*/, vec3<float> /*pos*/) const noexcept {
    return Spectrum();
}

Spectrum Material::probeEmission() const noexcept {
    return Spectrum();
}

ConstantMaterial::ConstantMaterial(Color<float> diffuse_color, float refractive_index, Spectrum emission) noexcept :
  diffuse_color(diffuse_color), refractive_index(refractive_index), emission(emission) {}

Color<float> ConstantMaterial::getDiffuseColor(vec3<float> /*pos*/) const noexcept {
    return diffuse_color;
}

float ConstantMaterial::getRefractiveIndex(vec3<float> /*pos*/) const noexcept {
    return refractive_index;
}

Spectrum ConstantMaterial::getEmission(Ray /*ray*/, vec3<float> /*pos*/) const noexcept {
    return emission;
}

Spectrum ConstantMaterial::probeEmission() const noexcept {
    return emission;
}
