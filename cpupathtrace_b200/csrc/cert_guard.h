// Guard table of the certified closest-hit walk (see "certified closest hit" in traverse.cuh).
//
// The certificate is exact about every primitive the walk TESTS.  About a primitive R it never reaches (leaf-box entry
// e_R > t_Q (1 + 2^-7)) it assumes that R's own intersection routine cannot report a distance below e_R (1 - 2^-8).
// That is a statement about fp32 rounding in Triangle / Sphere::getIntersection (reference src/scene/object.cpp:72-84,
// 146-182) and it fails for large triangles met at grazing incidence (|det| near the 1e-6 rejection threshold: the
// reported distance is cancellation noise) and for spheres grazed from nearby; tests/stress_cases.py builds such scenes.
//
// Error model (eps = 2^-24, safety factor K = 2, s = 2^-7):
//   triangle  |t_computed - t_true| <= 9 eps K |ab||ac| (|o - a| + 2 t_true) / |det|      (only |det| > 1e-6 reports a hit)
//   sphere    |t_computed - t_true| <= 3 eps K |co| + min(D / (2 sqrt(disc)), sqrt(D)),   D = 4 eps K max(|co|^2, r^2)
// A triangle whose worst case (|det| = 1e-6) keeps the relative part below s/12 is SAFE from every direction; what remains
// is an absolute term dmax * diameter, covered by demanding t_Q >= tau_safe.  Every other triangle contributes a PLANE
// (coplanar ones share one) with the union of the leaf boxes of the triangles behind it: a ray is handed to the
// reference-order walk without a certified walk when it enters that box (same slab arithmetic as the reference's) AND
// either is nearly parallel to the plane (|n.d| < cone) or enters closer than band / |n.d|, band = k (|o - box centre| +
// box half diagonal); a cheap plane-distance test (|n.o - h| <= w + band) comes first.  A ray that misses the box is
// never tested against those triangles by the reference either.  Spheres flag rays that start just outside their box
// (within 1 % of the radius; from inside the box the leaf is always reached and tested exactly) or graze
// (|disc| < r^2 / 16) from less than 3 r.  A scene with more distinct planes or spheres than the table
// holds, or with a sliver whose cone would exceed 0.25, cannot be certified: `certifiable` = 0 and guarded queries walk
// the reference tree.
//
// Guarded (exact) is the default of PTB_FLAG_CERTIFIED_CLOSEST; PTB_FLAG_CERTIFIED_RELAXED skips the guard (production
// renders with the counter-based generator, whose results are not comparable ray by ray with the reference anyway).
#ifndef PTB_CERT_GUARD_H
#define PTB_CERT_GUARD_H

#include "../../include/ptb.h"

#include <stdint.h>

namespace ptb_guard {

    constexpr int kGuardPlanes = 24;
    constexpr int kGuardSpheres = 8;
    constexpr float kGuardSlack = 0.0078125F; // s = 2^-7: prune slack 1 + s, entry slack 1 + s/4, suspect factor 1 - s/2

    struct GuardPlane {
        float nx, ny, nz, h;       // unit normal, offset n.x = h
        float lox, loy, loz, w;    // box of the triangles behind this plane; w = half thickness of their boxes across the plane
        float hix, hiy, hiz, cone; // |n.d| < cone: grazing
        float k, r, pad0, pad1;    // band = k (|o - box centre| + r), r = half diagonal of the box
    };

    struct CertGuard {
        uint32_t n_planes;
        uint32_t n_spheres;
        float tau_safe;       // a certified hit needs t >= tau_safe
        uint32_t certifiable; // 0: the scene does not fit the table
        GuardPlane planes[kGuardPlanes];
        float spheres[kGuardSpheres][4]; // centre, radius
    };

    // host: builds the table from the scene's primitives (cert_guard.cpp)
    void buildCertGuard(const ptb_prim *prims, uint64_t n_prims, CertGuard *out);

}

#endif
