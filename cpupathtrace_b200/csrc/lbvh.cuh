// Device-side construction of the query hierarchy.
//
// The scene keeps two hierarchies over one geometry array: the reference-topology tree, which
// defines the answer of a closest-hit query and is built on the host exactly like impl::constructBVH
// (reference src/scene/scene.cpp:12-102), and a second tree that any-hit queries and certified closest-hit queries
// walk.  Results do not depend on the shape of that second tree -- visibility is order-independent and closest hits
// carry a certificate (traverse.cuh) -- so it is free to be built by whatever is fastest.  This file builds it on the
// GPU as a linear BVH:
//
//   1. 63-bit Morton code of every primitive's box centre relative to the scene box            mortonKernel
//   2. radix sort of (code, slot) pairs                                                        cub::DeviceRadixSort
//   3. one inner node per adjacent pair of sorted leaves, ranges found by binary search on the
//      common-prefix length with the index as tie-break (Karras 2012)                          hierarchyKernel
//   4. bottom-up box fit: the second thread to arrive at a node owns it, merges its children's
//      boxes and writes the 64-byte record the traversal kernels read                           fitKernel
//
// Leaves hold one primitive and refer to it as ~slot, slot = position in the reference tree's leaf order, like the
// host-built tree (bvh_build.h).  Primitive boxes come from the host (primBounds, the reference's exact
// getBoundingVolume arithmetic) because a leaf's box must be bit-identical in both trees.
#ifndef PTB_LBVH_CUH
#define PTB_LBVH_CUH

#include <cub/cub.cuh>

#include <cstdint>

namespace ptb {

    struct LbvhWorkspace {
        const float *boxes;        // 6 floats per slot: lo.xyz, hi.xyz
        uint64_t *keys;            // sorted Morton codes
        uint32_t *slots;           // sorted position -> slot (null: the position is the slot)
        const uint32_t *box_ids;   // sorted position -> index of the primitive's box in `boxes` (null: the slot)
        int32_t *leaf_parent;      // sorted position -> inner node
        int32_t *node_parent;      // inner node -> inner node, -1 for the root
        int2 *children;            // inner node -> (left ref, right ref); ref >= 0 inner node, < 0 ~sorted position
        uint32_t *arrivals;        // inner node -> threads that have reached it in the fit pass
        float *node_box;           // 6 floats per inner node
        uint32_t *node_height;     // inner node -> inner levels below and including it
        uint32_t *node_count;      // inner node -> primitives below it
        float4 *records;           // n - 1 records, 4 lanes each (NodeRecord layout)
        uint32_t n;
        float root_lo[3];
        float root_hi[3];
    };

    __device__ __forceinline__ uint64_t spreadBits21(uint32_t v) {
        uint64_t x = v & 0x1FFFFFULL;
        x = (x | (x << 32)) & 0x001F00000000FFFFULL;
        x = (x | (x << 16)) & 0x001F0000FF0000FFULL;
        x = (x | (x << 8)) & 0x100F00F00F00F00FULL;
        x = (x | (x << 4)) & 0x10C30C30C30C30C3ULL;
        x = (x | (x << 2)) & 0x1249249249249249ULL;
        return x;
    }

    __global__ void __launch_bounds__(256) mortonKernel(LbvhWorkspace w, uint64_t *__restrict__ keys_out, uint32_t *__restrict__ slots_out) {
        const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
        if(slot >= w.n) {
            return;
        }
        const float *b = w.boxes + 6 * static_cast<size_t>(slot);
        uint32_t q[3];
        for(int c = 0; c < 3; c++) {
            const float extent = w.root_hi[c] - w.root_lo[c];
            const float centre = 0.5F * (b[c] + b[3 + c]);
            float u = extent > 0.0F ? (centre - w.root_lo[c]) / extent : 0.0F;
            u = fminf(fmaxf(u, 0.0F), 1.0F);
            q[c] = min(static_cast<uint32_t>(u * 2097152.0F), 2097151U);
        }
        keys_out[slot] = (spreadBits21(q[0]) << 2) | (spreadBits21(q[1]) << 1) | spreadBits21(q[2]);
        slots_out[slot] = slot;
    }

    // length of the common prefix of the keys at sorted positions i and j; equal keys continue with the positions
    // themselves, which makes all keys distinct (Karras 2012, section 4); -1 outside the array
    __device__ __forceinline__ int commonPrefix(const uint64_t *keys, int n, int i, int j) {
        if(j < 0 || j >= n) {
            return -1;
        }
        const uint64_t a = keys[i];
        const uint64_t b = keys[j];
        if(a != b) {
            return __clzll(static_cast<long long>(a ^ b));
        }
        return 64 + __clz(i ^ j);
    }

    __global__ void __launch_bounds__(256) hierarchyKernel(LbvhWorkspace w) {
        const int i = static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x);
        const int n = static_cast<int>(w.n);
        if(i >= n - 1) {
            return;
        }
        const uint64_t *keys = w.keys;
        // direction of the node's range and a bound on its length
        const int d = commonPrefix(keys, n, i, i + 1) - commonPrefix(keys, n, i, i - 1) >= 0 ? 1 : -1;
        const int delta_min = commonPrefix(keys, n, i, i - d);
        int l_max = 2;
        while(commonPrefix(keys, n, i, i + l_max * d) > delta_min) {
            l_max *= 2;
        }
        int l = 0;
        for(int t = l_max / 2; t >= 1; t /= 2) {
            if(commonPrefix(keys, n, i, i + (l + t) * d) > delta_min) {
                l += t;
            }
        }
        const int j = i + l * d;
        // split position: the last position that shares more than the node's own prefix with i
        const int delta_node = commonPrefix(keys, n, i, j);
        int s = 0;
        int t = l;
        do {
            t = (t + 1) / 2;
            if(commonPrefix(keys, n, i, i + (s + t) * d) > delta_node) {
                s += t;
            }
        } while(t > 1);
        const int gamma = i + s * d + min(d, 0);
        const int lo = min(i, j);
        const int hi = max(i, j);
        const int left = lo == gamma ? ~gamma : gamma;
        const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
        w.children[i] = make_int2(left, right);
        if(left >= 0) {
            w.node_parent[left] = i;
        }
        else {
            w.leaf_parent[gamma] = i;
        }
        if(right >= 0) {
            w.node_parent[right] = i;
        }
        else {
            w.leaf_parent[gamma + 1] = i;
        }
        if(i == 0) {
            w.node_parent[0] = -1;
        }
    }

    struct FitBox {
        float lo[3];
        float hi[3];
    };

    __device__ __forceinline__ FitBox loadBox(const float *p) {
        FitBox b;
        for(int c = 0; c < 3; c++) {
            b.lo[c] = p[c];
            b.hi[c] = p[3 + c];
        }
        return b;
    }

    // results of another thread (possibly on another SM): read through L2, never from this SM's L1
    __device__ __forceinline__ FitBox loadBoxCoherent(const float *p) {
        FitBox b;
        for(int c = 0; c < 3; c++) {
            b.lo[c] = __ldcg(p + c);
            b.hi[c] = __ldcg(p + 3 + c);
        }
        return b;
    }

    // One thread per leaf climbs towards the root; at every inner node the first arrival stops, the second one finds
    // both children finished (the fence before the counter makes their results visible) and completes the node.
    __global__ void __launch_bounds__(256) fitKernel(LbvhWorkspace w, uint32_t *__restrict__ height_out) {
        const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
        if(p >= w.n) {
            return;
        }
        int node = w.leaf_parent[p];
        while(node >= 0) {
            __threadfence();
            if(atomicAdd(&w.arrivals[node], 1U) == 0U) {
                return;
            }
            __threadfence();
            const int2 ch = w.children[node];
            FitBox lb;
            FitBox rb;
            uint32_t l_height = 0U;
            uint32_t r_height = 0U;
            uint32_t l_count = 1U;
            uint32_t r_count = 1U;
            int32_t l_ref;
            int32_t r_ref;
            if(ch.x >= 0) {
                lb = loadBoxCoherent(w.node_box + 6 * static_cast<size_t>(ch.x));
                l_height = __ldcg(w.node_height + ch.x);
                l_count = __ldcg(w.node_count + ch.x);
                l_ref = ch.x;
            }
            else {
                const uint32_t slot = w.slots != nullptr ? w.slots[~ch.x] : static_cast<uint32_t>(~ch.x);
                lb = loadBox(w.boxes + 6 * static_cast<size_t>(w.box_ids != nullptr ? w.box_ids[~ch.x] : slot));
                l_ref = ~static_cast<int32_t>(slot);
            }
            if(ch.y >= 0) {
                rb = loadBoxCoherent(w.node_box + 6 * static_cast<size_t>(ch.y));
                r_height = __ldcg(w.node_height + ch.y);
                r_count = __ldcg(w.node_count + ch.y);
                r_ref = ch.y;
            }
            else {
                const uint32_t slot = w.slots != nullptr ? w.slots[~ch.y] : static_cast<uint32_t>(~ch.y);
                rb = loadBox(w.boxes + 6 * static_cast<size_t>(w.box_ids != nullptr ? w.box_ids[~ch.y] : slot));
                r_ref = ~static_cast<int32_t>(slot);
            }
            float *nb = w.node_box + 6 * static_cast<size_t>(node);
            for(int c = 0; c < 3; c++) {
                // impl::combineAreas (bounding_box.cpp:8-12, 21-27) with std::min / std::max's choice between equal values,
                // like the host builder: the parity tree's records come out bit for bit the same from either builder
                nb[c] = rb.lo[c] < lb.lo[c] ? rb.lo[c] : lb.lo[c];
                nb[3 + c] = lb.hi[c] < rb.hi[c] ? rb.hi[c] : lb.hi[c];
            }
            const uint32_t height = 1U + max(l_height, r_height);
            w.node_height[node] = height;
            w.node_count[node] = l_count + r_count;
            const int parent = w.node_parent[node];
            float4 *rec = w.records + 4 * static_cast<size_t>(node);
            rec[0] = make_float4(lb.lo[0], lb.lo[1], lb.lo[2], lb.hi[0]);
            rec[1] = make_float4(lb.hi[1], lb.hi[2], rb.lo[0], rb.lo[1]);
            rec[2] = make_float4(rb.lo[2], rb.hi[0], rb.hi[1], rb.hi[2]);
            rec[3] = make_float4(__int_as_float(l_ref), __int_as_float(r_ref), __int_as_float(static_cast<int>(l_count + r_count)), __int_as_float(parent));
            if(parent < 0) {
                *height_out = height;
            }
            node = parent;
        }
    }

}

#endif
