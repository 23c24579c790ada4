// Scene setup on the device: both hierarchies of a scene are built by level-synchronous top-down splitting.
//
//   * the reference-topology tree: impl::constructBVH (reference src/scene/scene.cpp:12-102) decision for decision --
//     per node and axis the cut-off is the value at rank n/2 - 1 of the boxes' lower corners, the two groups
//     {low <= cut-off} / {low > cut-off} are bounded and their surface areas summed, the axis with the smallest sum wins
//     (the first on ties), the node's primitives are partitioned stably and the tail of the left group moves, last
//     first, to the end of the right group until left <= 2 * right.  The host builder (bvh_build.cpp) restates this as
//     a recursion over index spans; here every level of the tree is one pass over all primitives:
//       - three index lists sorted once by the lower corner in x / y / z make the rank-(n/2 - 1) value of every
//         node a single load; a stable partition keeps each node's part of the lists sorted, level after level;
//       - group bounds are min / max reductions (order-independent, so atomics on order-preserving integer keys give the
//         same bits as the reference's sequential loop), aggregated per warp;
//       - the primary list carries the reference's own primitive order, the stable partition and the reversed tail.
//     A node over n primitives at preorder index i owns records [i, i + n - 1) and leaf slots [begin, begin + n): the
//     children's indices are i + 1 and i + n_left, so the records land in the host builder's DFS preorder without a
//     renumbering pass, and a primitive's final position in the primary list IS its leaf slot.
//
//   * the query tree (any-hit and certified closest-hit walks, traverse.cuh): results cannot depend on its shape, so it
//     is built for traversal cost -- full-sweep surface-area heuristic: per node and axis, prefix and suffix boxes along
//     the centroid-sorted lists (one segmented scan over all six directions), cost(k) = A(prefix_k) (k + 1) +
//     A(suffix_k+1) (n - k - 1), minimum over the three axes and all positions, ties to the most balanced split.
//
// Boxes of both trees are fitted bottom-up by lbvh.cuh's fitKernel (the second thread to arrive at a node merges).
#ifndef PTB_GPU_BUILD_CUH
#define PTB_GPU_BUILD_CUH

#include <cub/cub.cuh>

#include <cstdint>

#include "../../include/ptb.h"
#include "device_scene.cuh"

namespace ptb {

    constexpr int kBuildBlock = 256;

    // order-preserving map float -> uint32 (total order, -0 < +0, infinities included)
    __host__ __device__ __forceinline__ uint32_t orderedKey(float f) {
#ifdef __CUDA_ARCH__
        const uint32_t u = __float_as_uint(f);
#else
        uint32_t u;
        memcpy(&u, &f, sizeof(u));
#endif
        return (u & 0x80000000U) != 0U ? ~u : (u | 0x80000000U);
    }

    __device__ __forceinline__ float orderedValue(uint32_t k) {
        return __uint_as_float((k & 0x80000000U) != 0U ? (k & 0x7FFFFFFFU) : ~k);
    }

    // Object::getBoundingVolume exactly as the reference computes it (Triangle: object.cpp:184-186, Sphere: :90-93,
    // NullObject: :60-62) -- the device twin of primBounds() in bvh_build.cpp.  boxes: 6 floats per primitive.
    __global__ void __launch_bounds__(kBuildBlock) primBoundsKernel(const ptb_prim *__restrict__ prims, uint32_t n, float *__restrict__ boxes) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i >= n) {
            return;
        }
        const ptb_prim &prim = prims[i];
        float lo[3];
        float hi[3];
        if(prim.kind == PTB_PRIM_TRIANGLE) {
            for(int c = 0; c < 3; c++) {
                const float a = prim.p[c];
                const float b = prim.p[3 + c];
                const float cc = prim.p[6 + c];
                const float ab_lo = b < a ? b : a;
                const float ab_hi = a < b ? b : a;
                lo[c] = cc < ab_lo ? cc : ab_lo;
                hi[c] = ab_hi < cc ? cc : ab_hi;
            }
        }
        else if(prim.kind == PTB_PRIM_SPHERE) {
            for(int c = 0; c < 3; c++) {
                lo[c] = prim.p[c] - prim.p[3];
                hi[c] = prim.p[c] + prim.p[3];
            }
        }
        else {
            for(int c = 0; c < 3; c++) {
                lo[c] = 0.0F;
                hi[c] = 0.0F;
            }
        }
        float *b = boxes + 6 * static_cast<size_t>(i);
        for(int c = 0; c < 3; c++) {
            b[c] = lo[c];
            b[3 + c] = hi[c];
        }
    }

    // sort keys of the three axis lists: lower corner (reference tree) or box centre (query tree) of primitive i
    __global__ void __launch_bounds__(kBuildBlock) axisKeysKernel(const float *__restrict__ boxes, uint32_t n, int centroids, float *__restrict__ keys, uint32_t *__restrict__ ids) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i >= n) {
            return;
        }
        const float *b = boxes + 6 * static_cast<size_t>(i);
        for(int a = 0; a < 3; a++) {
            keys[static_cast<size_t>(a) * n + i] = centroids != 0 ? 0.5F * (b[a] + b[3 + a]) : b[a];
        }
        ids[i] = i;
    }

    // The state of a level-synchronous build over n primitives ("ids": caller's primitive numbers for the reference tree,
    // leaf slots for the query tree).  A node is identified by its preorder index; position p of every list belongs to
    // node node_of_pos[p] (-1: the primitive at p has become a leaf).
    struct BuildState {
        uint32_t n;
        const float *boxes;          // 6 floats per id
        uint32_t *list[4];           // [0..2]: ids sorted by the axis key within every node's range; [3]: the primary order (reference tree)
        uint32_t *list_next[4];
        int32_t *node_of_pos;
        int32_t *node_of_pos_next;
        uint32_t *seg_begin;         // per node
        uint32_t *seg_count;
        int2 *children;              // per node: ref >= 0 inner node, < 0 ~position of the leaf in the primary list
        int32_t *node_parent;
        int32_t *leaf_parent;        // per position of the primary list
        uint32_t *side;              // per id: 1 = goes to the left child of its node at this level
        uint32_t *flags;             // 3 x (n + 1): partition predicate along a list; [n] = 0 so that an exclusive sum ends with the total
        uint32_t *ranks;             // 3 x (n + 1): exclusive sums of flags
        uint32_t *any_active;        // set when this level created a node with two or more primitives
        // reference split
        float *cut;                  // 3 per node
        uint32_t *group_keys;        // 36 per node: [axis][group][lo xyz, hi xyz] as ordered keys
        // sweep split
        uint32_t *best_cost;         // per node: ordered bits of the smallest cost
        unsigned long long *best_key; // per node: (imbalance, axis, k) of the split chosen among the positions with that cost
    };

    __device__ __forceinline__ void loadBox6(const float *boxes, uint32_t id, float lo[3], float hi[3]) {
        const float2 *p = reinterpret_cast<const float2 *>(boxes + 6 * static_cast<size_t>(id));
        const float2 a = __ldg(p);
        const float2 b = __ldg(p + 1);
        const float2 c = __ldg(p + 2);
        lo[0] = a.x;
        lo[1] = a.y;
        lo[2] = b.x;
        hi[0] = b.y;
        hi[1] = c.x;
        hi[2] = c.y;
    }

    // what the thread that places an element into a child does for the tree: the first element of a child of two or more
    // primitives creates the child node, a child of one primitive becomes a leaf
    __device__ __forceinline__ void placeInChild(const BuildState &s, int32_t node, uint32_t begin, uint32_t count, uint32_t n_left, bool left, uint32_t newpos) {
        const uint32_t child_count = left ? n_left : count - n_left;
        const uint32_t child_begin = left ? begin : begin + n_left;
        int32_t *slot_in_parent = left ? &s.children[node].x : &s.children[node].y;
        if(child_count >= 2U) {
            const int32_t child = left ? node + 1 : node + static_cast<int32_t>(n_left);
            s.node_of_pos_next[newpos] = child;
            if(newpos == child_begin) {
                s.seg_begin[child] = child_begin;
                s.seg_count[child] = child_count;
                s.node_parent[child] = node;
                *slot_in_parent = child;
                *s.any_active = 1U;
            }
        }
        else {
            s.node_of_pos_next[newpos] = -1;
            s.leaf_parent[newpos] = node;
            *slot_in_parent = ~static_cast<int32_t>(newpos);
        }
    }

    // ---------------------------------------------------------------------------------------------- reference split

    // scene.cpp:23-37: the cut-off of every active node in each dimension; also resets the node's group bounds
    __global__ void __launch_bounds__(kBuildBlock) refCutKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        if(p >= s.n) {
            return;
        }
        const int32_t node = s.node_of_pos[p];
        if(node < 0 || s.seg_begin[node] != p) {
            return;
        }
        const uint32_t count = s.seg_count[node];
        const uint32_t rank = p + static_cast<uint32_t>(static_cast<int>(count) / 2 - 1);
        for(int a = 0; a < 3; a++) {
            s.cut[3 * static_cast<size_t>(node) + a] = s.boxes[6 * static_cast<size_t>(s.list[a][rank]) + a];
        }
        uint32_t *g = s.group_keys + 36 * static_cast<size_t>(node);
        const uint32_t lo_init = orderedKey(__int_as_float(0x7F800000));  // +inf
        const uint32_t hi_init = orderedKey(__int_as_float(0xFF800000)); // -inf
        for(int k = 0; k < 6; k++) {
            for(int c = 0; c < 3; c++) {
                g[6 * k + c] = lo_init;
                g[6 * k + 3 + c] = hi_init;
            }
        }
    }

    // scene.cpp:39-55: bounds of the two groups of every node for each candidate dimension
    __global__ void __launch_bounds__(kBuildBlock) refGroupKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int32_t node = p < s.n ? s.node_of_pos[p] : -1;
        const uint32_t active = __ballot_sync(__activemask(), node >= 0);
        if(node < 0) {
            return;
        }
        float lo[3];
        float hi[3];
        loadBox6(s.boxes, s.list[3][p], lo, hi);
        uint32_t lo_key[3];
        uint32_t hi_key[3];
        for(int c = 0; c < 3; c++) {
            lo_key[c] = orderedKey(lo[c]);
            hi_key[c] = orderedKey(hi[c]);
        }
        uint32_t *g = s.group_keys + 36 * static_cast<size_t>(node);
        const float *cut = s.cut + 3 * static_cast<size_t>(node);
        const uint32_t peers = __match_any_sync(active, node);
        const bool uniform = peers == active;
        const uint32_t leader = static_cast<uint32_t>(__ffs(static_cast<int>(active))) - 1U;
        const uint32_t lane = threadIdx.x & 31U;
        for(int a = 0; a < 3; a++) {
            const int group = lo[a] <= cut[a] ? 0 : 1;
            if(uniform) {
                // the whole warp feeds one node (every level above the last few): one atomic per warp and bound
                for(int k = 0; k < 2; k++) {
                    for(int c = 0; c < 3; c++) {
                        const uint32_t mn = __reduce_min_sync(active, group == k ? lo_key[c] : 0xFFFFFFFFU);
                        const uint32_t mx = __reduce_max_sync(active, group == k ? hi_key[c] : 0U);
                        if(lane == leader) {
                            if(mn != 0xFFFFFFFFU) {
                                atomicMin(&g[12 * a + 6 * k + c], mn);
                            }
                            if(mx != 0U) {
                                atomicMax(&g[12 * a + 6 * k + 3 + c], mx);
                            }
                        }
                    }
                }
            }
            else {
                for(int c = 0; c < 3; c++) {
                    atomicMin(&g[12 * a + 6 * group + c], lo_key[c]);
                    atomicMax(&g[12 * a + 6 * group + 3 + c], hi_key[c]);
                }
            }
        }
    }

    // scene.cpp:57-73 (summed surface areas, first smallest dimension) and the partition predicate of scene.cpp:81-88
    __global__ void __launch_bounds__(kBuildBlock) refChooseKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        if(p > s.n) {
            return;
        }
        if(p == s.n) {
            s.flags[p] = 0U;
            return;
        }
        const int32_t node = s.node_of_pos[p];
        if(node < 0) {
            s.flags[p] = 0U;
            return;
        }
        const uint32_t *g = s.group_keys + 36 * static_cast<size_t>(node);
        int best_axis = 0;
        float best_area = 0.0F;
        for(int a = 0; a < 3; a++) {
            float area = 0.0F;
            for(int k = 0; k < 2; k++) {
                const uint32_t *b = g + 12 * a + 6 * k;
                const float dx = orderedValue(b[3]) - orderedValue(b[0]);
                const float dy = orderedValue(b[4]) - orderedValue(b[1]);
                const float dz = orderedValue(b[5]) - orderedValue(b[2]);
                area += 2.0F * (dx * dy + dy * dz + dx * dz);
            }
            if(a == 0 || area < best_area) {
                best_area = area;
                best_axis = a;
            }
        }
        const float low = s.boxes[6 * static_cast<size_t>(s.list[3][p]) + best_axis];
        s.flags[p] = low <= s.cut[3 * static_cast<size_t>(node) + best_axis] ? 1U : 0U;
    }

    // scene.cpp:75-94 on the primary list: stable partition, then the left tail moves (last first) behind the right group;
    // creates the children.  ranks = exclusive sums of flags over the primary list.
    __global__ void __launch_bounds__(kBuildBlock) refPartitionKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        if(p >= s.n) {
            return;
        }
        const uint32_t id = s.list[3][p];
        const int32_t node = s.node_of_pos[p];
        if(node < 0) {
            s.list_next[3][p] = id;
            s.node_of_pos_next[p] = -1;
            return;
        }
        const uint32_t begin = s.seg_begin[node];
        const uint32_t count = s.seg_count[node];
        const uint32_t n_left = s.ranks[begin + count] - s.ranks[begin];
        const uint32_t n_right = count - n_left;
        // while(left > 1 && left > 2 * right) { move one }: the smallest m with left - m <= 2 (right + m), at most left - 1
        uint32_t moved = 0U;
        if(n_left > 1U && n_left > 2U * n_right) {
            moved = (n_left - 2U * n_right + 2U) / 3U;
            moved = moved < n_left - 1U ? moved : n_left - 1U;
        }
        const uint32_t keep = n_left - moved;
        const bool in_left = s.flags[p] != 0U;
        const uint32_t left_before = s.ranks[p] - s.ranks[begin];
        uint32_t newpos;
        bool final_left;
        if(in_left) {
            if(left_before < keep) {
                newpos = begin + left_before;
                final_left = true;
            }
            else {
                newpos = begin + keep + n_right + (n_left - 1U - left_before);
                final_left = false;
            }
        }
        else {
            newpos = begin + keep + ((p - begin) - left_before);
            final_left = false;
        }
        s.list_next[3][newpos] = id;
        s.side[id] = final_left ? 1U : 0U;
        placeInChild(s, node, begin, count, keep, final_left, newpos);
    }

    // ---------------------------------------------------------------------------------------------- list partition

    // flags of the three axis lists from the sides decided on the primary list: blockIdx.y = axis
    __global__ void __launch_bounds__(kBuildBlock) sideFlagsKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int a = static_cast<int>(blockIdx.y);
        if(p > s.n) {
            return;
        }
        uint32_t *flags = s.flags + static_cast<size_t>(a) * (s.n + 1U);
        if(p == s.n) {
            flags[p] = 0U;
            return;
        }
        flags[p] = s.node_of_pos[p] >= 0 ? s.side[s.list[a][p]] : 0U;
    }

    // stable partition of the axis lists by those flags (ranks = their exclusive sums): blockIdx.y = axis.  With
    // bookkeep != 0 the pass over list 0 also creates the children (query tree: list 0 is the primary list).
    __global__ void __launch_bounds__(kBuildBlock) listPartitionKernel(BuildState s, int bookkeep) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int a = static_cast<int>(blockIdx.y);
        if(p >= s.n) {
            return;
        }
        const uint32_t id = s.list[a][p];
        const int32_t node = s.node_of_pos[p];
        if(node < 0) {
            s.list_next[a][p] = id;
            if(bookkeep != 0 && a == 0) {
                s.node_of_pos_next[p] = -1;
            }
            return;
        }
        const uint32_t *flags = s.flags + static_cast<size_t>(a) * (s.n + 1U);
        const uint32_t *ranks = s.ranks + static_cast<size_t>(a) * (s.n + 1U);
        const uint32_t begin = s.seg_begin[node];
        const uint32_t count = s.seg_count[node];
        const uint32_t n_left = ranks[begin + count] - ranks[begin];
        const uint32_t left_before = ranks[p] - ranks[begin];
        const bool left = flags[p] != 0U;
        const uint32_t newpos = left ? begin + left_before : begin + n_left + ((p - begin) - left_before);
        s.list_next[a][newpos] = id;
        if(bookkeep != 0 && a == 0) {
            placeInChild(s, node, begin, count, n_left, left, newpos);
        }
    }

    // ---------------------------------------------------------------------------------------------- sweep split

    struct SweepBox {
        float lo[3];
        float hi[3];
        uint32_t head; // 1: a segment starts here
    };

    struct SweepMerge {
        __host__ __device__ __forceinline__ SweepBox operator()(const SweepBox &a, const SweepBox &b) const {
            if(b.head != 0U) {
                return b;
            }
            SweepBox r;
            for(int c = 0; c < 3; c++) {
                r.lo[c] = fminf(a.lo[c], b.lo[c]);
                r.hi[c] = fmaxf(a.hi[c], b.hi[c]);
            }
            r.head = a.head;
            return r;
        }
    };

    // Six sequences of n boxes, back to back: axis a forward at [2a n, 2a n + n) and backward (positions reversed) at
    // [(2a + 1) n, (2a + 2) n); every node's range is a segment, every finished position a segment of its own.
    // blockIdx.y = axis.  Also resets the nodes' best cost (axis 0 pass).
    __global__ void __launch_bounds__(kBuildBlock) sweepFillKernel(BuildState s, SweepBox *__restrict__ seq) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int a = static_cast<int>(blockIdx.y);
        if(p >= s.n) {
            return;
        }
        const int32_t node = s.node_of_pos[p];
        SweepBox b;
        loadBox6(s.boxes, s.list[a][p], b.lo, b.hi);
        bool first = true;
        bool last = true;
        if(node >= 0) {
            const uint32_t begin = s.seg_begin[node];
            first = p == begin;
            last = p == begin + s.seg_count[node] - 1U;
            if(first && a == 0) {
                s.best_cost[node] = 0xFFFFFFFFU;
                s.best_key[node] = ~0ULL;
            }
        }
        b.head = first ? 1U : 0U;
        seq[static_cast<size_t>(2 * a) * s.n + p] = b;
        b.head = last ? 1U : 0U;
        seq[static_cast<size_t>(2 * a + 1) * s.n + (s.n - 1U - p)] = b;
    }

    __device__ __forceinline__ float sweepHalfArea(const SweepBox &b) {
        const float dx = b.hi[0] - b.lo[0];
        const float dy = b.hi[1] - b.lo[1];
        const float dz = b.hi[2] - b.lo[2];
        return dx * dy + dy * dz + dx * dz;
    }

    // cost of splitting a node after position p of axis list a; the node's smallest cost by atomicMin on ordered bits
    __global__ void __launch_bounds__(kBuildBlock) sweepCostKernel(BuildState s, const SweepBox *__restrict__ seq, float *__restrict__ costs) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int a = static_cast<int>(blockIdx.y);
        const int32_t node = p < s.n ? s.node_of_pos[p] : -1;
        bool candidate = false;
        float cost = 0.0F;
        if(node >= 0) {
            const uint32_t begin = s.seg_begin[node];
            const uint32_t count = s.seg_count[node];
            if(p + 1U < begin + count) {
                const uint32_t k = p - begin;
                const SweepBox left = seq[static_cast<size_t>(2 * a) * s.n + p];
                const SweepBox right = seq[static_cast<size_t>(2 * a + 1) * s.n + (s.n - 2U - p)];
                cost = sweepHalfArea(left) * static_cast<float>(k + 1U) + sweepHalfArea(right) * static_cast<float>(count - k - 1U);
                if(!(cost >= 0.0F)) {
                    cost = __int_as_float(0x7F800000); // NaN from overflowing extents: worst
                }
                candidate = true;
                costs[static_cast<size_t>(a) * s.n + p] = cost;
            }
        }
        const uint32_t voters = __ballot_sync(__activemask(), candidate);
        if(!candidate) {
            return;
        }
        const uint32_t key = orderedKey(cost);
        const uint32_t peers = __match_any_sync(voters, node);
        if(peers == voters) {
            const uint32_t mn = __reduce_min_sync(voters, key);
            if((threadIdx.x & 31U) == static_cast<uint32_t>(__ffs(static_cast<int>(voters))) - 1U) {
                atomicMin(&s.best_cost[node], mn);
            }
        }
        else {
            atomicMin(&s.best_cost[node], key);
        }
    }

    // among the positions that reach the node's smallest cost: the most balanced split, then the lowest axis and position
    __global__ void __launch_bounds__(kBuildBlock) sweepPickKernel(BuildState s, const float *__restrict__ costs) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int a = static_cast<int>(blockIdx.y);
        if(p >= s.n) {
            return;
        }
        const int32_t node = s.node_of_pos[p];
        if(node < 0) {
            return;
        }
        const uint32_t begin = s.seg_begin[node];
        const uint32_t count = s.seg_count[node];
        if(p + 1U >= begin + count) {
            return;
        }
        if(orderedKey(costs[static_cast<size_t>(a) * s.n + p]) != s.best_cost[node]) {
            return;
        }
        const uint32_t k = p - begin;
        const uint32_t twice_left = 2U * (k + 1U);
        const uint32_t imbalance = twice_left > count ? twice_left - count : count - twice_left;
        const unsigned long long key = (static_cast<unsigned long long>(imbalance) << 33) | (static_cast<unsigned long long>(a) << 31) | static_cast<unsigned long long>(k);
        atomicMin(&s.best_key[node], key);
    }

    // sides from the chosen split: the first k + 1 primitives of the node's range in the chosen axis list go left
    __global__ void __launch_bounds__(kBuildBlock) sweepSideKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        const int a = static_cast<int>(blockIdx.y);
        if(p >= s.n) {
            return;
        }
        const int32_t node = s.node_of_pos[p];
        if(node < 0) {
            return;
        }
        const unsigned long long key = s.best_key[node];
        if(static_cast<int>((key >> 31) & 3ULL) != a) {
            return;
        }
        const uint32_t k = static_cast<uint32_t>(key & 0x7FFFFFFFULL);
        s.side[s.list[a][p]] = (p - s.seg_begin[node]) <= k ? 1U : 0U;
    }

    // ---------------------------------------------------------------------------------------------- after the build

    __global__ void __launch_bounds__(kBuildBlock) buildInitKernel(BuildState s) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        if(p >= s.n) {
            return;
        }
        s.node_of_pos[p] = s.n >= 2U ? 0 : -1;
        if(p == 0U && s.n >= 2U) {
            s.seg_begin[0] = 0U;
            s.seg_count[0] = s.n;
            s.node_parent[0] = -1;
        }
    }

    __global__ void __launch_bounds__(kBuildBlock) iotaKernel(uint32_t *__restrict__ out, uint32_t n) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i < n) {
            out[i] = i;
        }
    }

    __global__ void __launch_bounds__(kBuildBlock) invertKernel(const uint32_t *__restrict__ slot_to_prim, uint32_t n, uint32_t *__restrict__ prim_to_slot) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if(i < n) {
            prim_to_slot[slot_to_prim[i]] = i;
        }
    }

    // The geometry and shading records of every leaf slot from the caller's primitives (the flattening loop of
    // ptb_scene_create on the host), plus the boxes in slot order for the query-tree build.
    __global__ void __launch_bounds__(kBuildBlock) packSlotsKernel(const ptb_prim *__restrict__ prims, const uint32_t *__restrict__ slot_to_prim, const float *__restrict__ boxes_by_prim,
                                                                    uint32_t n, float4 *__restrict__ geom, float4 *__restrict__ shade, float *__restrict__ boxes_by_slot) {
        const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
        if(slot >= n) {
            return;
        }
        const uint32_t index = slot_to_prim[slot];
        const ptb_prim &prim = prims[index];
        const float *p = prim.p;
        uint32_t flags = prim.kind & kKindMask;
        if(prim.kind == PTB_PRIM_TRIANGLE && prim.cull_backface != 0U) {
            flags |= kCullBit;
        }
        float4 *g = geom + kGeomLanes * static_cast<size_t>(slot);
        float4 *sh = shade + 3 * static_cast<size_t>(slot);
        const float4 zero = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
        float4 g0 = zero;
        float4 g1 = zero;
        float4 g2 = zero;
        float4 s0 = zero;
        float4 s1 = zero;
        float4 s2 = zero;
        if(prim.kind == PTB_PRIM_TRIANGLE) {
            // edges are differenced once, exactly as Triangle::getIntersection does per call (object.cpp:149-150)
            g0 = make_float4(p[0], p[1], p[2], 0.0F);
            g1 = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0F);
            g2 = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0F);
            s0 = make_float4(p[9], p[10], p[11], 0.0F);
            s1 = make_float4(p[12], p[13], p[14], 0.0F);
            s2 = make_float4(p[15], p[16], p[17], 0.0F);
        }
        else if(prim.kind == PTB_PRIM_SPHERE) {
            g0 = make_float4(p[0], p[1], p[2], 0.0F);
            g1 = make_float4(p[3], p[3] * p[3], 0.0F, 0.0F);
        }
        g0.w = __uint_as_float(flags);
        s0.w = __uint_as_float(prim.material);
        g[0] = g0;
        g[1] = g1;
        g[2] = g2;
        g[3] = zero;
        sh[0] = s0;
        sh[1] = s1;
        sh[2] = s2;
        for(int c = 0; c < 6; c++) {
            boxes_by_slot[6 * static_cast<size_t>(slot) + c] = boxes_by_prim[6 * static_cast<size_t>(index) + c];
        }
    }

}

#endif
