// Device-resident scene: the flattened form of the reference's pointer structures.
//
// HBM layout (every array 16-byte aligned, read with 128-bit loads through the read-only path):
//
//   nodes  : 4 x float4 per inner record (64 B), DFS preorder              <- AABB tree (scene/bounding_box.h:22-68)
//   occ_nodes: same record format, a second (SAH) hierarchy over the same primitives used by any-hit queries only
//   geom   : 4 x float4 per leaf slot  (64 B, last lane padding), leaf order <- Triangle / Sphere geometry (scene/object.h)
//              triangle: (a, flags) (b - a, 0) (c - a, 0) (pad)  flags = kind | cull << 2
//              sphere  : (origin, flags) (radius, radius^2, 0, 0) (0) (pad)
//   shade  : 3 x float4 per leaf slot  (48 B)                              <- shading normals + material index
//              (normal_a, material) (normal_b, 0) (normal_c, 0)
//   mats   : 3 x float4 per material   (48 B): diffuse, emission, (ior, bsdf, one_way, 0)
//   lights : 2 x float4 per point light: (pos, 0) (rgba)
//   emis   : 3 x float4 per emissive primitive, registration order (scene.cpp:183-208): triangle (a,slot)(b,p)(c,0);
//            sphere (origin, slot)(radius, radius^2, 0, p)(0); p = sampling density 1/area   [the un-differenced
//            vertices are needed by Triangle::sampleSurface, object.cpp:192-207]
//   cdf    : one float per emissive primitive (scene.cpp:169-180)
//
// One primitive per leaf, as in the reference; slot = position in left-to-right leaf order, which is also the
// order registerEmissiveObjects visits leaves in.
#ifndef PTB_DEVICE_SCENE_CUH
#define PTB_DEVICE_SCENE_CUH

#include "device_math.cuh"

namespace ptb {

    constexpr uint32_t kGeomLanes = 4U; // 3 used + 1 pad: 64-byte records so that (lane0, lane1) is one 256-bit load
    constexpr uint32_t kKindMask = 3U;
    constexpr uint32_t kCullBit = 4U;

    struct DeviceScene {
        const float4 *nodes;
        const float4 *occ_nodes; // occlusion hierarchy for any-hit queries (same record format, shares `geom`); may be null
        const float4 *geom;
        const float4 *shade;
        const float4 *mats;
        const float4 *lights;
        const float4 *emis;
        const float *cdf;
        const uint32_t *slot_to_prim;
        uint32_t n_prims;
        uint32_t n_lights;
        uint32_t n_emissive;
        uint32_t object_sample_count;
        int32_t root_ref;
        int32_t occ_root_ref;
        float root_lo[3];
        float root_hi[3];
    };

    struct VisitCounters {
        unsigned long long inner;
        unsigned long long leaf;
        unsigned long long suspect; // certified walk: hits reported more than 2^-8 in front of the primitive's own leaf box (see traverse.cuh)
    };

}

#endif
