// Host-side scalar set-up arithmetic that must round like the reference's x86 build (compiled by g++ with
// -ffp-contract=off): camera frame, primitive areas, emissive registration and its cumulative distribution.
// None of this is on the per-ray path; it runs once per camera / scene.
#include "host_math.h"

#include <algorithm>
#include <cmath>

namespace ptb {

    namespace {

        struct H3 {
            float x, y, z;
        };

        inline H3 sub(H3 a, H3 b) {
            return {a.x - b.x, a.y - b.y, a.z - b.z};
        }

        inline H3 scale(H3 a, float s) {
            return {a.x * s, a.y * s, a.z * s};
        }

        inline H3 crossH(H3 a, H3 b) {
            return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
        }

        inline float lengthH(H3 a) {
            float l2 = 0.0F;
            l2 += a.x * a.x;
            l2 += a.y * a.y;
            l2 += a.z * a.z;
            return std::sqrt(l2);
        }

        inline H3 normalizeH(H3 a) {
            const float inv = 1.0F / lengthH(a);
            return scale(a, inv);
        }

        inline H3 h3(const float *p) {
            return {p[0], p[1], p[2]};
        }

        inline void put(float *out, H3 v) {
            out[0] = v.x;
            out[1] = v.y;
            out[2] = v.z;
        }

    }

    // Camera::Camera (reference src/camera.cpp:54-76)
    void cameraInit(ptb_camera *out, const float origin[3], const float look_at[3], const float up[3], float focal_length, float height, float aspect_ratio,
                    float aperture_width, float aperture_height, uint32_t aperture_kind, float hexagon_horizontal_ratio, float focal_plane_dist) {
        const H3 o = h3(origin);
        const H3 forward_dir = normalizeH(sub(h3(look_at), o));
        const H3 forward = scale(forward_dir, focal_length);
        const H3 up_dir = normalizeH(h3(up));
        const float height_half = height / 2.0F;
        const H3 up_scaled = scale(up_dir, height_half);
        const H3 right_dir = normalizeH(crossH(forward, up_scaled));
        const float width_half = height_half * aspect_ratio;
        const H3 right = scale(right_dir, width_half);

        put(out->origin, o);
        put(out->forward, forward);
        put(out->up, up_scaled);
        put(out->right, right);
        out->aperture_width_half = aperture_width / 2.0F;
        out->aperture_height_half = aperture_height / 2.0F;
        out->aperture_kind = aperture_kind;
        // HexagonalApertureSampler's constructor clamps the ratio to [0, 1] (camera.cpp:21-23)
        out->hexagon_horizontal_ratio = std::min(std::max(hexagon_horizontal_ratio, 0.0F), 1.0F);
        out->focal_plane_dist = focal_plane_dist;
    }

    // Object::getSurfaceArea: Triangle (object.cpp:188-190), Sphere (:95-99), NullObject (:64-66)
    float primSurfaceArea(const ptb_prim &prim) {
        if(prim.kind == PTB_PRIM_TRIANGLE) {
            const H3 a = h3(prim.p);
            const H3 b = h3(prim.p + 3);
            const H3 c = h3(prim.p + 6);
            return lengthH(crossH(sub(b, a), sub(c, a))) / 2.0F;
        }
        if(prim.kind == PTB_PRIM_SPHERE) {
            constexpr float pi = static_cast<float>(M_PI);
            const float radius2 = prim.p[3] * prim.p[3];
            return 4.0F * pi * radius2;
        }
        return 0.0F;
    }

    // Object::sampleSurface's density: Triangle 1 / area (object.cpp:203-204), Sphere 1 / (4 pi r^2) (:113)
    float primSampleDensity(const ptb_prim &prim) {
        if(prim.kind == PTB_PRIM_TRIANGLE) {
            return 1.0F / primSurfaceArea(prim);
        }
        if(prim.kind == PTB_PRIM_SPHERE) {
            constexpr float pi = static_cast<float>(M_PI);
            const float radius2 = prim.p[3] * prim.p[3];
            return 1.0F / (4.0F * pi * radius2);
        }
        return 0.0F;
    }

    // Scene::registerEmissiveObjects + the prefix sum / normalisation of Scene::Scene (scene.cpp:165-208),
    // visiting primitives in leaf (slot) order.
    EmissiveTable buildEmissiveTable(const ptb_prim *prims, const ptb_material *materials, const uint32_t *slot_to_prim, uint64_t n_prims) {
        EmissiveTable table;
        for(uint64_t slot = 0; slot < n_prims; slot++) {
            const ptb_prim &prim = prims[slot_to_prim[slot]];
            const float *e = materials[prim.material].emission;
            const float emissive_power = (e[0] + e[1] + e[2]) * e[3];
            if(emissive_power <= 0.0F) {
                continue;
            }
            const float object_probability = emissive_power * primSurfaceArea(prim);
            if(object_probability <= 0.0F) {
                continue;
            }
            table.slots.push_back(static_cast<uint32_t>(slot));
            table.cdf.push_back(object_probability);
        }

        float cumulative = 0.0F;
        for(float &p : table.cdf) {
            const float probability = p;
            p += cumulative;
            cumulative += probability;
        }
        for(float &p : table.cdf) {
            p /= cumulative;
        }

        // scene.cpp:226 (the count depends only on the scene, so it is computed once here)
        const int emissive_object_count = static_cast<int>(table.slots.size());
        table.object_sample_count = static_cast<uint32_t>(std::min(2 + static_cast<int>(std::log10(emissive_object_count + 1)), emissive_object_count));
        return table;
    }

}
