// BVH traversal + primitive tests on the flattened scene.
//
// Result-identical restatement of
//   Scene::getIntersection          reference src/scene/scene.cpp:210-220
//   impl::getChildIntersection      reference src/scene/scene.cpp:104-150   (recursive, near child first)
//   AABB::getIntersection           reference src/scene/bounding_box.cpp:38-73
//   Triangle::getIntersection       reference src/scene/object.cpp:146-182  (Moeller-Trumbore, |det| <= 1e-6 rejected)
//   Sphere::getIntersection         reference src/scene/object.cpp:72-84    (near root only)
// as an iterative loop with a short per-thread stack of deferred far children:
//   - both child boxes of an inner record are tested; the child with the smaller entry distance is "close",
//     the RIGHT child on ties (scene.cpp:122-123 uses `left_t < right_t`);
//   - a child is entered iff 0 <= entry < best_t (strict); a deferred far child is re-tested against the current
//     best_t when popped (the recursion's `far_t < t_max` after `t_max = min(t_max, close_t)`, scene.cpp:130-138);
//   - a leaf hit replaces the best iff t >= 0 and (no best yet or t <= best_t): the later-visited primitive wins
//     ties, as the recursion's `close_t < far_hit_t ? close : far` does (scene.cpp:141-146).
#ifndef PTB_TRAVERSE_CUH
#define PTB_TRAVERSE_CUH

#include "cert_guard.h"
#include "device_scene.cuh"

namespace ptb {

    constexpr int kStackCapacity = 64; // deferred far siblings only; the host asserts bvh depth <= capacity

    // Optionally the bottom SMEM levels of every thread's stack live in shared memory, [level][thread] so that the 32
    // lanes of a warp never collide on a bank whatever their individual depths are; deeper entries spill to local
    // memory.  It is a template parameter of warpTrace because it pays in one regime and not in the other (round 2):
    //   * wavefront kernels on the bench scene (233 MiB scene served by L2 / L1, kernel bound by instruction issue):
    //     0 levels 7.77, 8 levels 7.52, 12 levels 7.49 Grays/s closest -- the local stack's sectors are L1 hits that the
    //     kernel hides, the extra predicated instructions and the smaller L1 are not: PTB_SMEM_STACK = 0;
    //   * batch queries on the 4 / 16 Mi-triangle soups (2-4 GB scene, kernel bound by memory latency, L1 hit rate 12 %, the
    //     local stack's 2.3 G sectors per launch go to L2): closest hits with 8 levels 562 -> 611 and 288 -> 306 Mrays/s
    //     (16 levels: half the speed), any-hit with 16 levels 566 -> 591 Mrays/s: kQuerySmemClosest / kQuerySmemAnyHit.
#ifndef PTB_SMEM_STACK
#define PTB_SMEM_STACK 0
#endif
    constexpr int kSmemStackLevels = PTB_SMEM_STACK;
    constexpr int kQuerySmemClosest = 8;
    constexpr int kQuerySmemAnyHit = 16;
    constexpr int kTraceBlock = 128; // threads per CTA of every kernel that calls warpTrace (== kBlock in kernels.cuh)

    struct RayInv {
        V3 o;
        V3 d;
        V3 inv; // 1/d, or FLT_MAX where d == 0 (bounding_box.cpp:43-45)
    };

    PTB_DEV float slabReciprocal(float d) {
        return fabsf(d) > 0.0F ? 1.0F / d : kFloatMax;
    }

    PTB_DEV RayInv makeRay(V3 o, V3 d) {
        RayInv r;
        r.o = o;
        r.d = d;
        r.inv = mk3(slabReciprocal(d.x), slabReciprocal(d.y), slabReciprocal(d.z));
        return r;
    }

    // entry distance, 0 if the origin is inside, -1 on a miss (bounding_box.cpp:47-72)
    PTB_DEV float slab(const RayInv &r, float lox, float loy, float loz, float hix, float hiy, float hiz) {
        const float t1 = (lox - r.o.x) * r.inv.x;
        const float t2 = (hix - r.o.x) * r.inv.x;
        const float t3 = (loy - r.o.y) * r.inv.y;
        const float t4 = (hiy - r.o.y) * r.inv.y;
        const float t5 = (loz - r.o.z) * r.inv.z;
        const float t6 = (hiz - r.o.z) * r.inv.z;
        const float t_min = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
        const float t_max = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
        if(t_max < 0.0F || t_min > t_max) {
            return -1.0F;
        }
        return t_min < 0.0F ? 0.0F : t_min;
    }

    // object.cpp:146-182 with ab = b - a and ac = c - a taken from the geometry lanes
    PTB_DEV float hitTriangle(const RayInv &r, V3 a, V3 ab, V3 ac, bool cull) {
        const V3 pvec = cross(r.d, ac);
        const float det = dot(ab, pvec);
        if(cull) {
            if(det <= 1E-6F) {
                return -1.0F;
            }
        }
        else {
            if(fabsf(det) <= 1E-6F) {
                return -1.0F;
            }
        }
        const float inv_det = 1.0F / det;
        const V3 tvec = r.o - a;
        const float u = dot(tvec, pvec) * inv_det;
        if(u < 0.0F || u > 1.0F) {
            return -1.0F;
        }
        const V3 qvec = cross(tvec, ab);
        const float v = dot(r.d, qvec) * inv_det;
        if(v < 0.0F || u + v > 1.0F) {
            return -1.0F;
        }
        return dot(ac, qvec) * inv_det;
    }

    // object.cpp:72-84
    PTB_DEV float hitSphere(const RayInv &r, V3 origin, float radius2) {
        const V3 co = r.o - origin;
        const float d = dot(r.d, co);
        const float discriminant = d * d - length2(co) + radius2;
        if(discriminant >= 0.0F) {
            return -(d + sqrtf(discriminant));
        }
        return -1.0F;
    }

    PTB_DEV float hitSlot(const DeviceScene &s, const RayInv &r, uint32_t slot) {
        const float4 *g = s.geom + kGeomLanes * static_cast<size_t>(slot);
        float4 g0;
        float4 g1;
        ld256(g, g0, g1);
        const uint32_t flags = __float_as_uint(g0.w);
        const uint32_t kind = flags & kKindMask;
        if(kind == PTB_PRIM_TRIANGLE) {
            const float4 g2 = __ldg(g + 2);
            return hitTriangle(r, mk3(g0.x, g0.y, g0.z), mk3(g1.x, g1.y, g1.z), mk3(g2.x, g2.y, g2.z), (flags & kCullBit) != 0U);
        }
        if(kind == PTB_PRIM_SPHERE) {
            return hitSphere(r, mk3(g0.x, g0.y, g0.z), g1.y);
        }
        return -1.0F; // NullObject (object.cpp:52-54)
    }

    struct Hit {
        float t;      // < 0: miss
        int32_t slot; // -1: none
    };

    // ------------------------------------------------------------------------------------------------ warp scheduling
    //
    // Persistent-warp traversal with warp-level votes.  Every lane runs the per-ray algorithm described at the top of
    // this file on its own ray (same node order, same pruning, same tie rules); what the votes decide is WHEN a lane
    // takes its next step, so that the 32 lanes of a warp execute the same kind of step together:
    //
    //   MODE = kTraceClosest  : closest hit on the reference-topology tree, `limit` ignored.
    //   MODE = kTraceAnyHit   : the ray is finished at the first primitive with 0 <= t < limit.
    //   MODE = kTraceCertified: closest hit on the SAH hierarchy plus a certificate that the reference walk returns the
    //                           same primitive (see "certified closest hit" below); commit() receives certain = false
    //                           for the rays that must be re-traced on the reference tree.
    //
    //   * a lane that has finished its ray does not wait for the slowest ray of a 32-ray batch: finished lanes are
    //     refilled from the device-side queue cursor as soon as kRefillVote of them are idle (one atomic per refill,
    //     claimed by ballot/popc/shfl);
    //   * a lane that has arrived at a leaf parks until vote.leaf (default 12) lanes are parked (or no lane has inner work left),
    //     then the parked lanes run the primitive test together; inner-node steps run for all unparked lanes.
    //
    // Measured with ncu before this change (profiles/r01_ncu_trace_baseline.md): 6.3 (closest) and 3.4 (shadow)
    // active threads per issued instruction with issue slots 75 % busy, i.e. the kernels were bound by SIMT divergence,
    // not by memory.
    struct VoteParams {
        int refill;     // idle lanes that trigger a refill
        int leaf;       // parked lanes that trigger the primitive tests
        int leaf_burst; // consecutive leaves one lane may test per primitive-test phase
        int inner_burst; // inner-node steps between two votes of the descend loop
    };

    enum LaneStatus : uint32_t { kLaneIdle = 0U, kLaneInner = 1U, kLaneLeaf = 2U };

    enum TraceMode : int { kTraceClosest = 0, kTraceAnyHit = 1, kTraceCertified = 2 };

    // ---- certified closest hit
    //
    // The reference's result depends on its tree only through ORDER: a primitive P (leaf-box entry e_P, own distance
    // t_P) is tested iff e_P < best_t at the moment its leaf is reached (its ancestors' entries are <= e_P and were
    // compared with an earlier, larger best_t), and replaces the best iff 0 <= t_P <= best_t.  Hence, with
    // Q = argmin t_P over the primitives whose leaf box is hit and t_P >= 0:
    //     if every other such primitive R has  t_R > t_Q  and  t_R > e_Q,
    // then in ANY visiting order Q is tested (best_t is a minimum of some t_R, or FLT_MAX) and wins strictly, so the
    // reference returns (t_Q, Q) whatever its tree looks like.  The SAH walk finds Q visiting every subtree whose entry
    // is <= t_Q (1 + 2^-7) and checks the condition on everything it tested; primitives it never reached have
    // e_R > t_Q (1 + 2^-7), and the certificate additionally asks e_Q <= t_Q (1 + 2^-9), so such an R could only matter if
    // its own test returned t_R < e_R (1 - 2^-8), i.e. if Moeller-Trumbore put the hit more than 0.4 % in front of the
    // triangle's own bounding box (only possible for |det| within rounding of the 1e-6 rejection threshold on very
    // large triangles; the parity tests compare both modes ray for ray).  Rays without a certificate -- exact ties on
    // shared edges and vertices, near-ties within an ulp or two of a box face -- are handed back (certain = false)
    // and re-traced on the reference tree, so ties keep the reference's later-visited-wins outcome.
    constexpr float kCertifiedPruneSlack = 1.0078125F;   // 1 + 2^-7
    constexpr float kCertifiedEntrySlack = 1.001953125F; // 1 + 2^-9
    constexpr float kCertifiedSuspectFactor = 0.99609375F; // 1 - 2^-8

    // The guard of the certified walk (cert_guard.h): true when the ray must be traced by the reference-order walk because
    // some large triangle or nearby sphere could report a distance that is rounding noise.
    PTB_DEV bool guardFlagsRay(const ptb_guard::CertGuard &g, const RayInv &r) {
        const V3 o = r.o;
        const V3 d = r.d;
        for(uint32_t j = 0; j < g.n_planes; j++) {
            const ptb_guard::GuardPlane &pl = g.planes[j];
            const float hd = ((pl.nx * o.x + pl.ny * o.y) + pl.nz * o.z) - pl.h;
            const float cd = (pl.nx * d.x + pl.ny * d.y) + pl.nz * d.z;
            const float ahd = fabsf(hd);
            const float dx = o.x - 0.5F * (pl.lox + pl.hix);
            const float dy = o.y - 0.5F * (pl.loy + pl.hiy);
            const float dz = o.z - 0.5F * (pl.loz + pl.hiz);
            const float band = pl.k * (sqrtf((dx * dx + dy * dy) + dz * dz) + pl.r);
            if(fabsf(cd) < pl.cone || ahd <= pl.w + band) {
                // near the plane or grazing it: now the exact question -- does the ray enter the box of the triangles behind
                // the plane (the union of their leaf boxes, same slab arithmetic as the reference's), and if so, closer
                // than band / |n.d|?  A ray that misses this box is never tested against those triangles by the reference
                // either (a ray LEAVING a wall it just bounced off misses the wall's flat box).
                const float entry = slab(r, pl.lox, pl.loy, pl.loz, pl.hix, pl.hiy, pl.hiz);
                if(entry >= 0.0F && (fabsf(cd) < pl.cone || entry * fabsf(cd) < band)) {
                    return true;
                }
            }
        }
        for(uint32_t j = 0; j < g.n_spheres; j++) {
            const float cx = o.x - g.spheres[j][0];
            const float cy = o.y - g.spheres[j][1];
            const float cz = o.z - g.spheres[j][2];
            const float radius = g.spheres[j][3];
            const float r2 = radius * radius;
            const float reach = fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fabsf(cz));
            // a ray that starts inside the sphere's box always reaches its leaf (entry 0) and is tested exactly
            if(reach > radius) {
                if(reach <= 1.01F * radius) {
                    return true;
                }
                const float co2 = (cx * cx + cy * cy) + cz * cz;
                if(co2 < 9.0F * r2) {
                    const float dd = (d.x * cx + d.y * cy) + d.z * cz;
                    const float disc = dd * dd - co2 + r2;
                    if(fabsf(disc) < r2 * 0.0625F) {
                        return true;
                    }
                }
            }
        }
        return false;
    }

    // Hit test of one box in "visit" form: hit <=> the reference's slab result is >= 0, entry = that result.
    // (bounding_box.cpp:61-72: -1 iff t_max < 0 or t_min > t_max; otherwise max(t_min, 0).)
    PTB_DEV bool slabVisit(const RayInv &r, float lox, float loy, float loz, float hix, float hiy, float hiz, float best_t, float &entry) {
        const float t1 = (lox - r.o.x) * r.inv.x;
        const float t2 = (hix - r.o.x) * r.inv.x;
        const float t3 = (loy - r.o.y) * r.inv.y;
        const float t4 = (hiy - r.o.y) * r.inv.y;
        const float t5 = (loz - r.o.z) * r.inv.z;
        const float t6 = (hiz - r.o.z) * r.inv.z;
        const float t_min = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
        const float t_max = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
        entry = fmaxf(t_min, 0.0F);
        // child entered iff 0 <= slab result < best_t (scene.cpp:126, 137)
        return t_max >= 0.0F && t_min <= t_max && entry < best_t;
    }

    // ticket = fetch(k, o, d, limit) loads work item k of the queue and returns what commit() needs to find the ray's
    // destination again (the path index, shadow slot or ray number the queue entry held -- kept in the register that would
    // otherwise hold k, so that commit does not re-read the queue: that dependent load was 5-10 % of the trace kernels'
    // instructions in round 1); commit(ticket, hit, certain) stores the result.  `cursor` is a zero-initialised device
    // counter shared by all warps of the launch; `count` the number of rays.
    //
    // Warp collectives and convergence.  `__ballot_sync(0xFFFFFFFF, ...)` is only as good as the barrier in front of the
    // VOTE instruction, and ptxas (CUDA 12.9, sm_100a) emits that barrier only where its convergence analysis thinks
    // the warp may be split; elsewhere it trusts BSYNC.RECONVERGENT, drops an explicit __syncwarp() as redundant, and
    // does the same for a member mask passed at run time.  On B200 the lanes demonstrably do not always arrive
    // together: a variant of this loop was caught executing its refill vote with part of the warp (64 times in one
    // 60-launch render); the halves then disagreed on `exhausted`, one half left and the other waited for an all-idle
    // ballot that could no longer come.  The loop therefore does not ASSUME full-warp convergence anywhere: every
    // group of collectives runs over `present = __activemask()`, the lanes that really are executing together, and
    // every decision taken from a vote is a decision of that group -- it refills, parks, tests and terminates as a
    // sub-warp of its own.  `exhausted` is re-agreed at every refill vote (groups can merge again), and a full-mask
    // __syncwarp() at the top of the outer loop invites split groups to merge where the compiler keeps it.
    // one allocation per kernel, however many warpTrace instantiations the kernel contains
    template<int SMEM>
    PTB_DEV uint2 *sharedStackBase() {
        if constexpr(SMEM > 0) {
            __shared__ uint2 shared_stack[SMEM * kTraceBlock];
            return shared_stack;
        }
        else {
            return nullptr;
        }
    }

    // `guard` (certified mode only, may be null = relaxed): rays it flags are committed as uncertain without a walk, and a
    // certified hit closer than guard->tau_safe loses its certificate.
    template<int MODE, bool COUNT, int SMEM = kSmemStackLevels, typename Fetch, typename Commit>
    PTB_DEV void warpTrace(const DeviceScene &s, VoteParams vote, uint32_t *cursor, uint32_t count, Fetch fetch, Commit commit, VisitCounters *counters,
                           const ptb_guard::CertGuard *guard = nullptr) {
        constexpr bool ANY_HIT = MODE == kTraceAnyHit;
        constexpr bool CERTIFIED = MODE == kTraceCertified;
        const int kRefillVote = vote.refill;
        const int kLeafVote = vote.leaf;
        const uint32_t lane = threadIdx.x & 31U;
        const uint32_t lanes_below = (1U << lane) - 1U;

        uint32_t status = kLaneIdle;
        uint32_t k = 0U;
        RayInv r;
        r.o = mk3(0.0F, 0.0F, 0.0F);
        r.d = mk3(0.0F, 0.0F, 1.0F);
        r.inv = r.d;
        float limit = 0.0F;
        float best_t = 0.0F;
        float prune_t = 0.0F;    // certified mode: what subtree entries are compared with (best_t, widened)
        float leaf_entry = 0.0F; // certified mode: entry distance of the leaf box the lane is parked at
        float rival_t = 0.0F;    // certified mode: max(t, leaf entry) of the best; another hit at or below it voids the certificate
        bool certain = true;
        Hit hit;
        hit.t = -1.0F;
        hit.slot = -1;
        int32_t node = 0;
        int sp = 0;
        // (node ref, entry distance bits) of deferred far children
        uint2 local_stack[kStackCapacity - SMEM];
        uint2 *const my_shared_stack = sharedStackBase<SMEM>() + threadIdx.x;
        auto stackStore = [&](int at, uint2 e) {
            if(at < SMEM) {
                my_shared_stack[at * kTraceBlock] = e;
            }
            else {
                local_stack[at - SMEM] = e;
            }
        };
        auto stackLoad = [&](int at) -> uint2 {
            if(at < SMEM) {
                return my_shared_stack[at * kTraceBlock];
            }
            return local_stack[at - SMEM];
        };
        bool exhausted = count == 0U;
        unsigned long long n_inner = 0;
        unsigned long long n_leaf = 0;
        unsigned long long n_suspect = 0;

        // next deferred far child that still beats the best distance, or the ray is finished
        auto advance = [&]() {
            while(sp > 0) {
                sp--;
                const uint2 e = stackLoad(sp);
                // any-hit: the bound never shrinks, so an entry that passed `entry < limit` when deferred still passes
                if(ANY_HIT || __uint_as_float(e.y) < (CERTIFIED ? prune_t : best_t)) {
                    node = static_cast<int32_t>(e.x);
                    if(CERTIFIED) {
                        leaf_entry = __uint_as_float(e.y);
                    }
                    status = node >= 0 ? kLaneInner : kLaneLeaf;
                    return;
                }
            }
            if(CERTIFIED && guard != nullptr && hit.slot >= 0 && !(hit.t >= guard->tau_safe)) {
                certain = false;
            }
            commit(k, hit, certain);
            status = kLaneIdle;
        };

        // one inner-node step of a lane whose `node` is an inner record
        auto innerStep = [&]() {
            const float4 *rec = s.nodes + 4 * static_cast<size_t>(node);
            float4 n0;
            float4 n1;
            float4 n2;
            float4 n3;
            ld256(rec, n0, n1);
            ld256(rec + 2, n2, n3);
            if(COUNT) {
                n_inner++;
            }
            float lt;
            float rt;
            const float bound = CERTIFIED ? prune_t : best_t;
            const bool visit_left = slabVisit(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, bound, lt);
            const bool visit_right = slabVisit(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, bound, rt);
            const int32_t left = __float_as_int(n3.x);
            const int32_t right = __float_as_int(n3.y);
            if(visit_left && visit_right) {
                // nearer child first, the right one on ties (scene.cpp:122-123); the other is deferred
                const bool left_first = lt < rt;
                // (an L2 prefetch of the deferred sibling was measured and rejected: -9 % on the bench scene, -10 % on
                // the 16 Mi-triangle soup)
                stackStore(sp, make_uint2(static_cast<uint32_t>(left_first ? right : left), __float_as_uint(left_first ? rt : lt)));
                sp++;
                node = left_first ? left : right;
                if(CERTIFIED) {
                    leaf_entry = left_first ? lt : rt;
                }
                status = node >= 0 ? kLaneInner : kLaneLeaf;
            }
            else if(visit_left || visit_right) {
                node = visit_left ? left : right;
                if(CERTIFIED) {
                    leaf_entry = visit_left ? lt : rt;
                }
                status = node >= 0 ? kLaneInner : kLaneLeaf;
            }
            else {
                advance();
            }
        };

        for(;;) {
            // ---- (A) refill idle lanes once enough of them wait
            __syncwarp();
            uint32_t present = __activemask();
            exhausted = __any_sync(present, exhausted);
            const uint32_t idle_mask = __ballot_sync(present, status == kLaneIdle);
            if(idle_mask == present && exhausted) {
                break;
            }
            if(!exhausted && (__popc(idle_mask) >= kRefillVote || idle_mask == present)) {
                const uint32_t wanted = static_cast<uint32_t>(__popc(idle_mask));
                const uint32_t leader = static_cast<uint32_t>(__ffs(static_cast<int>(present))) - 1U;
                uint32_t base = 0U;
                if(lane == leader) {
                    base = atomicAdd(cursor, wanted);
                }
                base = __shfl_sync(present, base, static_cast<int>(leader));
                if(base + wanted >= count) {
                    exhausted = true;
                }
                if(status == kLaneIdle) {
                    const uint32_t mine = base + static_cast<uint32_t>(__popc(idle_mask & lanes_below));
                    if(mine < count) {
                        V3 o;
                        V3 d;
                        k = fetch(mine, o, d, limit);
                        r = makeRay(o, d);
                        hit.t = -1.0F;
                        hit.slot = -1;
                        sp = 0;
                        best_t = ANY_HIT ? limit : kFloatMax;
                        prune_t = best_t;
                        certain = true;
                        rival_t = 0.0F;
                        if(s.n_prims == 0U) {
                            commit(k, hit, true);
                        }
                        else if(CERTIFIED && guard != nullptr && guardFlagsRay(*guard, r)) {
                            commit(k, hit, false);
                        }
                        else {
                            const float root_t = slab(r, s.root_lo[0], s.root_lo[1], s.root_lo[2], s.root_hi[0], s.root_hi[1], s.root_hi[2]);
                            if(!(root_t >= 0.0F)) {
                                hit.t = root_t;
                                commit(k, hit, true);
                            }
                            else {
                                node = s.root_ref;
                                leaf_entry = root_t;
                                status = node >= 0 ? kLaneInner : kLaneLeaf;
                            }
                        }
                    }
                }
            }

            // ---- (B) descend: inner-node steps until enough lanes are parked at a leaf or waiting for a refill
            for(;;) {
                present = __activemask();
                const uint32_t inner_mask = __ballot_sync(present, status == kLaneInner);
                if(inner_mask == 0U) {
                    break;
                }
                // up to vote.inner_burst steps between two votes: a lane that leaves the Inner state early sits the rest of the
                // burst out, which costs less than voting after every step
                for(int step = 0; step < vote.inner_burst; step++) {
                    if(status == kLaneInner) {
                        innerStep();
                    }
                }
                present = __activemask();
                const uint32_t parked = __ballot_sync(present, status == kLaneLeaf);
                if(__popc(parked) >= kLeafVote) {
                    break;
                }
                if(!exhausted) {
                    const uint32_t waiting = __ballot_sync(present, status == kLaneIdle);
                    if(__popc(waiting) >= kRefillVote) {
                        break;
                    }
                }
            }

            // ---- (C) primitive tests for every parked lane; a lane whose next deferred sibling is a leaf as well (the
            // common case at the bottom of a one-primitive-per-leaf tree) tests it in the same phase instead of parking again
            for(int burst = 0; burst < vote.leaf_burst && status == kLaneLeaf; burst++) {
                const uint32_t slot = static_cast<uint32_t>(~node);
                if(COUNT) {
                    n_leaf++;
                }
                const float t = hitSlot(s, r, slot);
                if(ANY_HIT) {
                    if(t >= 0.0F && t < limit) {
                        hit.t = t;
                        hit.slot = static_cast<int32_t>(slot);
                        commit(k, hit, true);
                        status = kLaneIdle;
                    }
                    else {
                        advance();
                    }
                }
                else if(CERTIFIED) {
                    // The one assumption behind the certificate: a primitive the walk never reaches could matter only if its
                    // test put the hit more than 2^-8 in front of its own leaf box.  Every primitive that IS tested is
                    // checked for exactly that, in every build: a ray that meets such a primitive is in the numerically
                    // degenerate regime (grazing a large triangle, |det| near the 1e-6 rejection threshold); its walk is
                    // abandoned on the spot and the ray handed back to the reference-order walk.  Counting builds count them.
                    if(t >= 0.0F && t < leaf_entry * kCertifiedSuspectFactor) {
                        if(COUNT) {
                            n_suspect++;
                        }
                        commit(k, hit, false);
                        status = kLaneIdle;
                    }
                    else {
                        if(t >= 0.0F) {
                            if(t < best_t) {
                                // every primitive tested so far has t >= the old best; they clear the new rival bound iff it does
                                const float rival = fmaxf(t, leaf_entry);
                                certain = best_t > rival && leaf_entry <= t * kCertifiedEntrySlack && t > 0.0F;
                                rival_t = rival;
                                best_t = t;
                                prune_t = t * kCertifiedPruneSlack;
                                hit.t = t;
                                hit.slot = static_cast<int32_t>(slot);
                            }
                            else if(t <= rival_t) {
                                certain = false;
                            }
                        }
                        advance();
                    }
                }
                else {
                    if(t >= 0.0F && (hit.slot < 0 || t <= best_t)) {
                        best_t = t;
                        hit.t = t;
                        hit.slot = static_cast<int32_t>(slot);
                    }
                    advance();
                }
            }
        }

        if(COUNT && counters != nullptr) {
            atomicAdd(&counters->inner, n_inner);
            atomicAdd(&counters->leaf, n_leaf);
            if(n_suspect != 0ULL) {
                atomicAdd(&counters->suspect, n_suspect);
            }
        }
    }

}

#endif
