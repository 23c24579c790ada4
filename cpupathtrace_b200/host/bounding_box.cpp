// AABB node type kept for source compatibility; the slab test runs on the device (ptb_aabb_intersect).
#include "device.h"

#include <PathTrace/scene/bounding_box.h>

#include <cstdio>
#include <utility>

AABB::AABB() : area{}, child(std::make_unique<NullObject>()), leaf(true) {}

AABB::AABB(AABB &&other) noexcept :
  area(other.area), left(std::move(other.left)), right(std::move(other.right)), child(std::move(other.child)), leaf(other.leaf) {}

AABB &AABB::operator=(AABB &&other) noexcept {
    if(this != &other) {
        area = other.area;
        left = std::move(other.left);
        right = std::move(other.right);
        child = std::move(other.child);
        leaf = other.leaf;
    }
    return *this;
}

AABB::AABB(AABB &&left_node, AABB &&right_node) : leaf(false) {
    area = AABBArea{min(left_node.area.low, right_node.area.low), max(left_node.area.high, right_node.area.high)};
    left = std::make_unique<AABB>(std::move(left_node));
    right = std::make_unique<AABB>(std::move(right_node));
}

AABB::AABB(AABBArea area, std::unique_ptr<Object> &&child) noexcept : area(area), child(std::move(child)), leaf(true) {}

float AABB::getIntersection(const Ray &ray) const noexcept {
    const float low[3] = {area.low[0], area.low[1], area.low[2]};
    const float high[3] = {area.high[0], area.high[1], area.high[2]};
    const float packed[6] = {ray.origin[0], ray.origin[1], ray.origin[2], ray.dir[0], ray.dir[1], ray.dir[2]};
    float t = -1.0F;
    try {
        ptb::host::ok(ptb_aabb_intersect(ptb::host::defaultContext(), low, high, 1, packed, &t), "AABB::getIntersection");
    }
    catch(const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
    }
    return t;
}
