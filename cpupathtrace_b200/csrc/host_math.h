#ifndef PTB_HOST_MATH_H
#define PTB_HOST_MATH_H

#include <cstdint>
#include <vector>

#include "../../include/ptb.h"

namespace ptb {

    void cameraInit(ptb_camera *out, const float origin[3], const float look_at[3], const float up[3], float focal_length, float height, float aspect_ratio,
                    float aperture_width, float aperture_height, uint32_t aperture_kind, float hexagon_horizontal_ratio, float focal_plane_dist);

    float primSurfaceArea(const ptb_prim &prim);
    float primSampleDensity(const ptb_prim &prim);

    struct EmissiveTable {
        std::vector<uint32_t> slots; // leaf slots of the emissive primitives, registration order
        std::vector<float> cdf;      // normalised cumulative selection probabilities
        uint32_t object_sample_count = 0;
    };

    EmissiveTable buildEmissiveTable(const ptb_prim *prims, const ptb_material *materials, const uint32_t *slot_to_prim, uint64_t n_prims);

}

#endif
