// Camera host object: the constructor derives the camera frame (host set-up arithmetic, ptb_camera_init); rays are
// generated on the device — by the wavefront generate kernel inside a render, by ptb_camera_shoot for a direct call.
#include "device.h"

#include <PathTrace/camera.h>

#include <algorithm>
#include <cstdio>
#include <utility>

namespace {

    std::tuple<float, float> unitAperture(uint32_t kind, float ratio, RandomEngine &re) noexcept {
        float out[2] = {0.0F, 0.0F};
        uint64_t state = re.state();
        try {
            if(ptb::host::ok(ptb_aperture_sample(ptb::host::defaultContext(), kind, ratio, 1, &state, out), "ApertureSampler::sampleAperture")) {
                re.setState(state);
            }
        }
        catch(const std::exception &e) {
            std::fprintf(stderr, "%s\n", e.what());
        }
        return std::make_tuple(out[0], out[1]);
    }

}

std::tuple<float, float> CircularApertureSampler::sampleAperture(RandomEngine &re) const noexcept {
    return unitAperture(PTB_APERTURE_CIRCULAR, 0.0F, re);
}

HexagonalApertureSampler::HexagonalApertureSampler(float horizontal_ratio) noexcept : horizontal_ratio(std::min(std::max(horizontal_ratio, 0.0F), 1.0F)) {}

std::tuple<float, float> HexagonalApertureSampler::sampleAperture(RandomEngine &re) const noexcept {
    return unitAperture(PTB_APERTURE_HEXAGONAL, horizontal_ratio, re);
}

Camera::Camera(vec3<float> origin, vec3<float> look_at, vec3<float> up, float focal_length, float height, float aspect_ratio) noexcept :
  Camera(origin, look_at, up, focal_length, height, aspect_ratio, 0.0F, 0.0F, nullptr, 0.0F) {}

Camera::Camera(vec3<float> origin_in, vec3<float> look_at, vec3<float> up_in, float focal_length, float height, float aspect_ratio, float aperture_width,
               float aperture_height, std::unique_ptr<ApertureSampler> &&sampler, float focal_plane) noexcept {
    ptb_camera pod{};
    const float o[3] = {origin_in[0], origin_in[1], origin_in[2]};
    const float l[3] = {look_at[0], look_at[1], look_at[2]};
    const float u[3] = {up_in[0], up_in[1], up_in[2]};
    ptb_camera_init(&pod, o, l, u, focal_length, height, aspect_ratio, aperture_width, aperture_height, PTB_APERTURE_NONE, 0.0F, focal_plane);

    origin = vec3<float>{pod.origin[0], pod.origin[1], pod.origin[2]};
    forward = vec3<float>{pod.forward[0], pod.forward[1], pod.forward[2]};
    up = vec3<float>{pod.up[0], pod.up[1], pod.up[2]};
    right = vec3<float>{pod.right[0], pod.right[1], pod.right[2]};
    aperture_width_half = pod.aperture_width_half;
    aperture_height_half = pod.aperture_height_half;
    aperture_sampler = std::move(sampler);
    focal_plane_dist = focal_plane;
}

bool Camera::lower(ptb_camera &out) const noexcept {
    out = ptb_camera{};
    for(int k = 0; k < 3; k++) {
        out.origin[k] = origin[k];
        out.forward[k] = forward[k];
        out.up[k] = up[k];
        out.right[k] = right[k];
    }
    out.aperture_width_half = aperture_width_half;
    out.aperture_height_half = aperture_height_half;
    out.focal_plane_dist = focal_plane_dist;
    out.aperture_kind = PTB_APERTURE_NONE;
    if(aperture_sampler) {
        if(dynamic_cast<const CircularApertureSampler *>(aperture_sampler.get()) != nullptr) {
            out.aperture_kind = PTB_APERTURE_CIRCULAR;
        }
        else if(const auto *hexagon = dynamic_cast<const HexagonalApertureSampler *>(aperture_sampler.get())) {
            out.aperture_kind = PTB_APERTURE_HEXAGONAL;
            out.hexagon_horizontal_ratio = hexagon->getHorizontalRatio();
        }
        else {
            return false;
        }
    }
    return true;
}

Ray Camera::shootRay(float x, float y, float pixel_width, float pixel_height, RandomEngine &re) const noexcept {
    Ray ray{origin, forward.normalize()};
    ptb_camera pod;
    if(!lower(pod)) {
        std::fprintf(stderr, "PathTrace (B200): Camera::shootRay: user ApertureSampler subclasses cannot run on the GPU\n");
        return ray;
    }
    const float xy[2] = {x, y};
    float out[6] = {};
    uint64_t state = re.state();
    try {
        if(ptb::host::ok(ptb_camera_shoot(ptb::host::defaultContext(), &pod, 1, xy, pixel_width, pixel_height, &state, out), "Camera::shootRay")) {
            re.setState(state);
            ray = Ray{vec3<float>{out[0], out[1], out[2]}, vec3<float>{out[3], out[4], out[5]}};
        }
    }
    catch(const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
    }
    return ray;
}
