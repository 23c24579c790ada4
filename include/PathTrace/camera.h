// Perspective camera with optional aperture and thin lens (API of the reference's include/PathTrace/camera.h).
// Primary rays are generated on the GPU by the wavefront `generate` kernel from the POD form of this class.
#ifndef PATHTRACE_CAMERA_H
#define PATHTRACE_CAMERA_H

#include <PathTrace/base.h>

#include <memory>
#include <tuple>

struct ptb_camera;

//! Uniformly sampleable aperture shape inside [-1, 1] x [-1, 1]
class ApertureSampler {
  public:
    virtual ~ApertureSampler() = default;
    virtual std::tuple<float, float> sampleAperture(RandomEngine &re) const noexcept = 0;
};

class CircularApertureSampler final : public ApertureSampler {
  public:
    virtual ~CircularApertureSampler() = default;
    std::tuple<float, float> sampleAperture(RandomEngine &re) const noexcept override;
};

class HexagonalApertureSampler final : public ApertureSampler {
  private:
    float horizontal_ratio;

  public:
    virtual ~HexagonalApertureSampler() = default;
    //! @param horizontal_ratio share of the hexagon's width taken by its flat top / bottom edges, clamped to [0, 1]
    HexagonalApertureSampler(float horizontal_ratio) noexcept;
    std::tuple<float, float> sampleAperture(RandomEngine &re) const noexcept override;

    // B200 extension: read access for lowering
    float getHorizontalRatio() const noexcept { return horizontal_ratio; }
};

class Camera {
  private:
    vec3<float> origin;
    vec3<float> forward;
    vec3<float> up;
    vec3<float> right;

    float aperture_width_half;
    float aperture_height_half;
    std::unique_ptr<ApertureSampler> aperture_sampler;

    float focal_plane_dist;

  public:
    //! pinhole camera
    Camera(vec3<float> origin, vec3<float> look_at, vec3<float> up, float focal_length, float height, float aspect_ratio) noexcept;

    //! camera with an aperture and, for focal_plane_dist > 0, a thin lens
    Camera(vec3<float> origin, vec3<float> look_at, vec3<float> up, float focal_length, float height, float aspect_ratio, float aperture_width,
           float aperture_height, std::unique_ptr<ApertureSampler> &&aperture_sampler, float focal_plane_dist = 0.0F) noexcept;

    //! ray through sensor coordinates (x, y) in [-1, 1]^2, jittered inside the pixel extent
    Ray shootRay(float x, float y, float pixel_width, float pixel_height, RandomEngine &re) const noexcept;

    // B200 extension: POD form for the C-ABI; false if the aperture sampler is a user subclass
    bool lower(ptb_camera &out) const noexcept;
};

#endif /* PATHTRACE_CAMERA_H */
