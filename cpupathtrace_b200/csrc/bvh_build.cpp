// See bvh_build.h.  Compiled by g++ with -ffp-contract=off: the split decisions below compare float sums that must
// round exactly like the reference's x86 build (no FMA contraction).
#include "bvh_build.h"

#include <algorithm>
#include <cmath>
#include <future>
#include <limits>
#include <thread>

namespace ptb {

    namespace {

        inline float lesser(float a, float b) {
            return b < a ? b : a;
        }

        inline float greater(float a, float b) {
            return a < b ? b : a;
        }

        struct Box {
            float lo[3];
            float hi[3];
        };

        struct BuildContext {
            const Box *boxes;        // per primitive
            uint32_t *order;         // primitive indices, node [begin, end) owns order[begin..end)
            uint32_t *scratch_index; // same extent, scratch for the stable partition
            float *scratch_value;    // same extent, scratch for the median selection
            NodeRecord *nodes;
            uint32_t *slot_to_prim;
            int spawn_levels;        // tasks are forked for the first few levels only
        };

        struct SubtreeResult {
            Box box;
            int32_t ref;
            uint32_t depth;
        };

        // Split rule of the reference (scene.cpp:23-94), restated over an index span:
        //   1. per axis, the value at rank n/2-1 of the boxes' lower corners is the cut-off;
        //   2. per axis, the two groups {low <= cut-off} / {low > cut-off} are bounded and their surface areas summed;
        //   3. the axis with the smallest sum wins, the first axis on ties;
        //   4. stable partition by that rule, then the tail of the left group moves (last first) to the end of the
        //      right group until left <= 2 * right (or left == 1).
        // Returns the size of the left group; order[begin..end) is rearranged in place.
        size_t splitSpan(const BuildContext &ctx, size_t begin, size_t end) {
            const size_t n = end - begin;
            uint32_t *span = ctx.order + begin;
            float *values = ctx.scratch_value + begin;

            float cut[3];
            for(int axis = 0; axis < 3; axis++) {
                for(size_t k = 0; k < n; k++) {
                    values[k] = ctx.boxes[span[k]].lo[axis];
                }
                float *rank = values + (static_cast<int>(n) / 2 - 1);
                std::nth_element(values, rank, values + n);
                cut[axis] = *rank;
            }

            constexpr float inf = std::numeric_limits<float>::infinity();
            int best_axis = 0;
            float best_area = 0.0F;
            for(int axis = 0; axis < 3; axis++) {
                Box group[2];
                for(auto &g : group) {
                    for(int c = 0; c < 3; c++) {
                        g.lo[c] = inf;
                        g.hi[c] = -inf;
                    }
                }
                for(size_t k = 0; k < n; k++) {
                    const Box &b = ctx.boxes[span[k]];
                    Box &g = group[b.lo[axis] <= cut[axis] ? 0 : 1];
                    for(int c = 0; c < 3; c++) {
                        g.lo[c] = lesser(g.lo[c], b.lo[c]);
                        g.hi[c] = greater(g.hi[c], b.hi[c]);
                    }
                }
                float area = 0.0F;
                for(const auto &g : group) {
                    const float dx = g.hi[0] - g.lo[0];
                    const float dy = g.hi[1] - g.lo[1];
                    const float dz = g.hi[2] - g.lo[2];
                    area += 2.0F * (dx * dy + dy * dz + dx * dz);
                }
                if(axis == 0 || area < best_area) {
                    best_area = area;
                    best_axis = axis;
                }
            }

            // stable partition through the scratch span
            uint32_t *scratch = ctx.scratch_index + begin;
            size_t n_left = 0;
            size_t n_right = 0;
            for(size_t k = 0; k < n; k++) {
                if(ctx.boxes[span[k]].lo[best_axis] <= cut[best_axis]) {
                    n_left++;
                }
            }
            {
                size_t li = 0;
                size_t ri = n_left;
                for(size_t k = 0; k < n; k++) {
                    if(ctx.boxes[span[k]].lo[best_axis] <= cut[best_axis]) {
                        scratch[li++] = span[k];
                    }
                    else {
                        scratch[ri++] = span[k];
                    }
                }
                n_right = n - n_left;
            }

            // rebalance: the left tail is appended to the right group in reverse order
            size_t keep_left = n_left;
            size_t grown_right = n_right;
            while(keep_left > 1 && keep_left > 2 * grown_right) {
                keep_left--;
                grown_right++;
            }

            size_t w = 0;
            for(size_t k = 0; k < keep_left; k++) {
                span[w++] = scratch[k];
            }
            for(size_t k = 0; k < n_right; k++) {
                span[w++] = scratch[n_left + k];
            }
            for(size_t k = n_left; k > keep_left; k--) {
                span[w++] = scratch[k - 1];
            }

            return keep_left;
        }

        // A subtree over n primitives owns inner records [node, node + n - 1) and leaf slots [slot, slot + n):
        // both are known before recursing, which makes sibling subtrees independent tasks.
        SubtreeResult buildSubtree(const BuildContext &ctx, size_t begin, size_t end, int32_t node, int32_t parent, uint32_t slot, int level) {
            const size_t n = end - begin;
            if(n == 1) {
                const uint32_t prim = ctx.order[begin];
                ctx.slot_to_prim[slot] = prim;
                return {ctx.boxes[prim], ~static_cast<int32_t>(slot), 1U};
            }

            const size_t n_left = splitSpan(ctx, begin, end);
            const size_t mid = begin + n_left;

            const int32_t left_node = node + 1;
            const int32_t right_node = node + static_cast<int32_t>(n_left);

            SubtreeResult left;
            SubtreeResult right;
            if(level < ctx.spawn_levels && n > 4096) {
                auto future = std::async(std::launch::async, [&]() { return buildSubtree(ctx, begin, mid, left_node, node, slot, level + 1); });
                right = buildSubtree(ctx, mid, end, right_node, node, slot + static_cast<uint32_t>(n_left), level + 1);
                left = future.get();
            }
            else {
                left = buildSubtree(ctx, begin, mid, left_node, node, slot, level + 1);
                right = buildSubtree(ctx, mid, end, right_node, node, slot + static_cast<uint32_t>(n_left), level + 1);
            }

            NodeRecord &record = ctx.nodes[node];
            record.lane[0][0] = left.box.lo[0];
            record.lane[0][1] = left.box.lo[1];
            record.lane[0][2] = left.box.lo[2];
            record.lane[0][3] = left.box.hi[0];
            record.lane[1][0] = left.box.hi[1];
            record.lane[1][1] = left.box.hi[2];
            record.lane[1][2] = right.box.lo[0];
            record.lane[1][3] = right.box.lo[1];
            record.lane[2][0] = right.box.lo[2];
            record.lane[2][1] = right.box.hi[0];
            record.lane[2][2] = right.box.hi[1];
            record.lane[2][3] = right.box.hi[2];
            record.left = left.ref;
            record.right = right.ref;
            record.leaf_count = static_cast<int32_t>(n);
            record.parent = parent;

            // AABB(AABB&&, AABB&&) -> impl::combineAreas (bounding_box.cpp:8-12, 21-27)
            SubtreeResult result;
            for(int c = 0; c < 3; c++) {
                result.box.lo[c] = lesser(left.box.lo[c], right.box.lo[c]);
                result.box.hi[c] = greater(left.box.hi[c], right.box.hi[c]);
            }
            result.ref = node;
            result.depth = 1U + std::max(left.depth, right.depth);
            return result;
        }

    }

    namespace {

        constexpr int kBins = 16;

        struct SahContext {
            const Box *boxes;
            const float *centroids; // 3 per primitive
            uint32_t *order;
            const uint32_t *prim_to_slot;
            NodeRecord *nodes;
            int spawn_levels;
        };

        inline void grow(Box &b, const Box &o) {
            for(int c = 0; c < 3; c++) {
                b.lo[c] = lesser(b.lo[c], o.lo[c]);
                b.hi[c] = greater(b.hi[c], o.hi[c]);
            }
        }

        inline Box emptyBox() {
            constexpr float inf = std::numeric_limits<float>::infinity();
            return Box{{inf, inf, inf}, {-inf, -inf, -inf}};
        }

        inline float halfArea(const Box &b) {
            const float dx = b.hi[0] - b.lo[0];
            const float dy = b.hi[1] - b.lo[1];
            const float dz = b.hi[2] - b.lo[2];
            return dx * dy + dy * dz + dx * dz;
        }

        SubtreeResult buildSah(const SahContext &ctx, size_t begin, size_t end, int32_t node, int32_t parent, int level) {
            const size_t n = end - begin;
            uint32_t *span = ctx.order + begin;
            if(n == 1) {
                const uint32_t prim = span[0];
                return {ctx.boxes[prim], ~static_cast<int32_t>(ctx.prim_to_slot[prim]), 1U};
            }

            // centroid bounds decide the binning
            float clo[3] = {std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity()};
            float chi[3] = {-clo[0], -clo[1], -clo[2]};
            for(size_t k = 0; k < n; k++) {
                const float *c = ctx.centroids + 3 * static_cast<size_t>(span[k]);
                for(int a = 0; a < 3; a++) {
                    clo[a] = lesser(clo[a], c[a]);
                    chi[a] = greater(chi[a], c[a]);
                }
            }

            int best_axis = -1;
            int best_split = 0;
            float best_cost = std::numeric_limits<float>::infinity();
            for(int axis = 0; axis < 3; axis++) {
                const float extent = chi[axis] - clo[axis];
                if(!(extent > 0.0F)) {
                    continue;
                }
                const float scale = static_cast<float>(kBins) / extent;
                Box bin_box[kBins];
                size_t bin_count[kBins] = {};
                for(auto &b : bin_box) {
                    b = emptyBox();
                }
                for(size_t k = 0; k < n; k++) {
                    const uint32_t prim = span[k];
                    int b = static_cast<int>((ctx.centroids[3 * static_cast<size_t>(prim) + axis] - clo[axis]) * scale);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    bin_count[b]++;
                    grow(bin_box[b], ctx.boxes[prim]);
                }
                // sweep: cost of splitting after bin s
                float right_area[kBins];
                size_t right_count[kBins];
                Box acc = emptyBox();
                size_t cnt = 0;
                for(int b = kBins - 1; b > 0; b--) {
                    grow(acc, bin_box[b]);
                    cnt += bin_count[b];
                    right_area[b] = halfArea(acc);
                    right_count[b] = cnt;
                }
                acc = emptyBox();
                cnt = 0;
                for(int b = 0; b < kBins - 1; b++) {
                    grow(acc, bin_box[b]);
                    cnt += bin_count[b];
                    if(cnt == 0 || right_count[b + 1] == 0) {
                        continue;
                    }
                    const float cost = halfArea(acc) * static_cast<float>(cnt) + right_area[b + 1] * static_cast<float>(right_count[b + 1]);
                    if(cost < best_cost) {
                        best_cost = cost;
                        best_axis = axis;
                        best_split = b + 1;
                    }
                }
            }

            size_t n_left = 0;
            if(best_axis >= 0) {
                const float scale = static_cast<float>(kBins) / (chi[best_axis] - clo[best_axis]);
                size_t i = 0;
                size_t j = n;
                while(i < j) {
                    const uint32_t prim = span[i];
                    int b = static_cast<int>((ctx.centroids[3 * static_cast<size_t>(prim) + best_axis] - clo[best_axis]) * scale);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    if(b < best_split) {
                        i++;
                    }
                    else {
                        j--;
                        std::swap(span[i], span[j]);
                    }
                }
                n_left = i;
            }
            if(n_left == 0 || n_left == n) {
                n_left = n / 2; // coincident centroids: split by position
            }

            const size_t mid = begin + n_left;
            const int32_t left_node = node + 1;
            const int32_t right_node = node + static_cast<int32_t>(n_left);
            SubtreeResult left;
            SubtreeResult right;
            if(level < ctx.spawn_levels && n > 4096) {
                auto future = std::async(std::launch::async, [&]() { return buildSah(ctx, begin, mid, left_node, node, level + 1); });
                right = buildSah(ctx, mid, end, right_node, node, level + 1);
                left = future.get();
            }
            else {
                left = buildSah(ctx, begin, mid, left_node, node, level + 1);
                right = buildSah(ctx, mid, end, right_node, node, level + 1);
            }

            NodeRecord &record = ctx.nodes[node];
            record.lane[0][0] = left.box.lo[0];
            record.lane[0][1] = left.box.lo[1];
            record.lane[0][2] = left.box.lo[2];
            record.lane[0][3] = left.box.hi[0];
            record.lane[1][0] = left.box.hi[1];
            record.lane[1][1] = left.box.hi[2];
            record.lane[1][2] = right.box.lo[0];
            record.lane[1][3] = right.box.lo[1];
            record.lane[2][0] = right.box.lo[2];
            record.lane[2][1] = right.box.hi[0];
            record.lane[2][2] = right.box.hi[1];
            record.lane[2][3] = right.box.hi[2];
            record.left = left.ref;
            record.right = right.ref;
            record.leaf_count = static_cast<int32_t>(n);
            record.parent = parent;

            SubtreeResult result;
            result.box = left.box;
            grow(result.box, right.box);
            result.ref = node;
            result.depth = 1U + std::max(left.depth, right.depth);
            return result;
        }

    }

    void primBounds(const ptb_prim &prim, float low[3], float high[3]) {
        switch(prim.kind) {
            case PTB_PRIM_TRIANGLE:
                for(int c = 0; c < 3; c++) {
                    const float a = prim.p[c];
                    const float b = prim.p[3 + c];
                    const float cc = prim.p[6 + c];
                    low[c] = lesser(lesser(a, b), cc);
                    high[c] = greater(greater(a, b), cc);
                }
                break;
            case PTB_PRIM_SPHERE:
                for(int c = 0; c < 3; c++) {
                    low[c] = prim.p[c] - prim.p[3];
                    high[c] = prim.p[c] + prim.p[3];
                }
                break;
            default:
                for(int c = 0; c < 3; c++) {
                    low[c] = 0.0F;
                    high[c] = 0.0F;
                }
                break;
        }
    }

    FlatBvh buildReferenceBvh(const ptb_prim *prims, uint64_t n_prims, int threads) {
        FlatBvh flat;
        if(n_prims == 0) {
            return flat;
        }

        std::vector<Box> boxes(n_prims);
        for(uint64_t i = 0; i < n_prims; i++) {
            primBounds(prims[i], boxes[i].lo, boxes[i].hi);
        }

        std::vector<uint32_t> order(n_prims);
        for(uint64_t i = 0; i < n_prims; i++) {
            order[i] = static_cast<uint32_t>(i);
        }
        std::vector<uint32_t> scratch_index(n_prims);
        std::vector<float> scratch_value(n_prims);

        flat.nodes.resize(n_prims - 1);
        flat.slot_to_prim.resize(n_prims);

        if(threads <= 0) {
            threads = static_cast<int>(std::thread::hardware_concurrency());
        }
        int spawn_levels = 0;
        while((1 << spawn_levels) < std::max(threads, 1) && spawn_levels < 6) {
            spawn_levels++;
        }

        BuildContext ctx{boxes.data(), order.data(), scratch_index.data(), scratch_value.data(), flat.nodes.data(), flat.slot_to_prim.data(), spawn_levels};
        SubtreeResult root = buildSubtree(ctx, 0, n_prims, 0, -1, 0U, 0);

        flat.root_ref = root.ref;
        flat.depth = root.depth;
        for(int c = 0; c < 3; c++) {
            flat.root_low[c] = root.box.lo[c];
            flat.root_high[c] = root.box.hi[c];
        }
        return flat;
    }

}

namespace ptb {

    FlatBvh buildOcclusionBvh(const ptb_prim *prims, uint64_t n_prims, const uint32_t *prim_to_slot, int threads) {
        FlatBvh flat;
        if(n_prims == 0) {
            return flat;
        }
        std::vector<Box> boxes(n_prims);
        std::vector<float> centroids(3 * n_prims);
        std::vector<uint32_t> order(n_prims);
        for(uint64_t i = 0; i < n_prims; i++) {
            primBounds(prims[i], boxes[i].lo, boxes[i].hi);
            for(int c = 0; c < 3; c++) {
                centroids[3 * i + c] = 0.5F * (boxes[i].lo[c] + boxes[i].hi[c]);
            }
            order[i] = static_cast<uint32_t>(i);
        }
        flat.nodes.resize(n_prims - 1);

        if(threads <= 0) {
            threads = static_cast<int>(std::thread::hardware_concurrency());
        }
        int spawn_levels = 0;
        while((1 << spawn_levels) < std::max(threads, 1) && spawn_levels < 6) {
            spawn_levels++;
        }
        SahContext ctx{boxes.data(), centroids.data(), order.data(), prim_to_slot, flat.nodes.data(), spawn_levels};
        const SubtreeResult root = buildSah(ctx, 0, n_prims, 0, -1, 0);
        flat.root_ref = root.ref;
        flat.depth = root.depth;
        for(int c = 0; c < 3; c++) {
            flat.root_low[c] = root.box.lo[c];
            flat.root_high[c] = root.box.hi[c];
        }
        return flat;
    }

}
