// Axis-aligned boxes (API of the reference's include/PathTrace/scene/bounding_box.h).
//
// The GPU scene keeps its hierarchy as flat 64-byte records (cpupathtrace_b200/csrc/bvh_build.h); this class remains
// for source compatibility: a node that owns either one object or two children, with the reference's slab test
// answered by the device function the traversal kernel uses.
#ifndef PATHTRACE_BOUNDING_BOX_H
#define PATHTRACE_BOUNDING_BOX_H

#include <PathTrace/base.h>
#include <PathTrace/scene/object.h>

#include <memory>

struct AABBArea {
    vec3<float> low;
    vec3<float> high;
};

class AABB {
  public:
    AABBArea area;

    std::unique_ptr<AABB> left;
    std::unique_ptr<AABB> right;
    std::unique_ptr<Object> child;

    bool leaf;

    //! leaf holding a NullObject
    AABB();
    AABB(AABB &&other) noexcept;
    AABB &operator=(AABB &&other) noexcept;

    //! inner node bounding both children
    AABB(AABB &&left, AABB &&right);

    //! leaf holding `child`
    AABB(AABBArea area, std::unique_ptr<Object> &&child) noexcept;

    //! entry distance of the ray into the box; 0 when the origin is inside; negative on a miss
    float getIntersection(const Ray &ray) const noexcept;
};

#endif /* PATHTRACE_BOUNDING_BOX_H */
