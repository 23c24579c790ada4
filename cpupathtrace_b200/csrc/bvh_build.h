// Host-side BVH construction and flattening for the device scene.
//
// PTB_BVH_REFERENCE reproduces the topology of the reference's impl::constructBVH
// (reference src/scene/scene.cpp:12-102) decision for decision, because bit-exact closest-hit parity needs the same
// visiting order and the same tie behaviour ("Traversal contract").  Only the topology is
// shared; the representation is not: the reference keeps a pointer tree of 56-byte AABB objects owning one heap
// primitive each, this builder works on index spans and writes the flat device layout directly.
#ifndef PTB_BVH_BUILD_H
#define PTB_BVH_BUILD_H

#include <cstdint>
#include <vector>

#include "../../include/ptb.h"

namespace ptb {

    // 64-byte inner record, four 16-byte lanes (one LDG.128 each).
    //   lane0 = (L.lo.x, L.lo.y, L.lo.z, L.hi.x)
    //   lane1 = (L.hi.y, L.hi.z, R.lo.x, R.lo.y)
    //   lane2 = (R.lo.z, R.hi.x, R.hi.y, R.hi.z)
    //   lane3 = (left ref, right ref, leaf-count of the subtree, preorder index of the parent)  [int32 bit patterns]
    // child ref >= 0: index of an inner record; child ref < 0: ~slot of a primitive (one primitive per leaf).
    struct alignas(64) NodeRecord {
        float lane[3][4];
        int32_t left;
        int32_t right;
        int32_t leaf_count;
        int32_t parent;
    };
    static_assert(sizeof(NodeRecord) == 64, "inner record must be 64 bytes");

    struct FlatBvh {
        std::vector<NodeRecord> nodes;    // n_prims - 1 records (0 when n_prims <= 1), DFS preorder
        std::vector<uint32_t> slot_to_prim; // leaf order (left-to-right DFS) -> index in ptb_scene_desc.prims
        int32_t root_ref = -1;            // inner index 0, or ~0 for a single primitive; meaningless when empty
        uint32_t depth = 0;               // deepest leaf, root = 1
        float root_low[3] = {0, 0, 0};
        float root_high[3] = {0, 0, 0};
    };

    // Bounding volume of one primitive exactly as the reference computes it
    // (Triangle::getBoundingVolume object.cpp:184-186, Sphere::getBoundingVolume :90-93, NullObject :60-62).
    void primBounds(const ptb_prim &prim, float low[3], float high[3]);

    // Builds the reference topology over `prims` and flattens it.  `threads` <= 0 picks hardware concurrency.
    FlatBvh buildReferenceBvh(const ptb_prim *prims, uint64_t n_prims, int threads);

    // A second hierarchy over the same primitives for ANY-HIT (shadow) queries only.  Whether a segment is occluded
    // does not depend on the order in which primitives are tested, so this tree is free to be a better one than the
    // reference's median split: binned surface-area heuristic (16 bins per axis, all three axes), one primitive per
    // leaf, same 64-byte record.  Leaf references are ~slot with slot = position of the primitive in the REFERENCE
    // tree's leaf order (`prim_to_slot`), so both trees share one geometry array.  `slot_to_prim` of the result is unused.
    FlatBvh buildOcclusionBvh(const ptb_prim *prims, uint64_t n_prims, const uint32_t *prim_to_slot, int threads);

}

#endif
