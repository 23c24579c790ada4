// Host-side construction of the certified walk's guard table (cert_guard.h).
#include "cert_guard.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace ptb_guard {

    namespace {

        constexpr double kEps = 5.9604644775390625e-8; // 2^-24
        constexpr double kSafety = 2.0;
        constexpr double kMargin = 1.0001; // fp32 rounding of the table entries themselves

        struct Vec {
            double x, y, z;
        };

        Vec sub(Vec a, Vec b) {
            return {a.x - b.x, a.y - b.y, a.z - b.z};
        }

        double dot(Vec a, Vec b) {
            return a.x * b.x + a.y * b.y + a.z * b.z;
        }

        double norm(Vec a) {
            return std::sqrt(dot(a, a));
        }

    }

    void buildCertGuard(const ptb_prim *prims, uint64_t n_prims, CertGuard *out) {
        CertGuard &g = *out;
        std::memset(&g, 0, sizeof(g));
        g.certifiable = 1;
        const double slack = kGuardSlack;
        double tau_safe = 0.0;

        // Pass 1 (parallel for large scenes): almost every triangle of a mesh is "small" -- its worst relative distance
        // error is far below the slack -- and contributes nothing but a candidate for tau_safe, a maximum.  Everything else
        // (spheres, large triangles, degenerate ones) is collected and goes through the sequential pass 2 in index order,
        // where planes are merged.
        std::vector<uint64_t> special;
        {
            const unsigned workers = n_prims >= (1U << 16) ? std::min(16U, std::max(1U, std::thread::hardware_concurrency())) : 1U;
            std::vector<std::vector<uint64_t>> found(workers);
            std::vector<double> tau(workers, 0.0);
            auto scan = [&](unsigned w) {
                const uint64_t begin = n_prims * w / workers;
                const uint64_t end = n_prims * (w + 1) / workers;
                for(uint64_t i = begin; i < end; i++) {
                    const ptb_prim &p = prims[i];
                    if(p.kind != PTB_PRIM_TRIANGLE) {
                        if(p.kind == PTB_PRIM_SPHERE) {
                            found[w].push_back(i);
                        }
                        continue;
                    }
                    const Vec a{p.p[0], p.p[1], p.p[2]};
                    const Vec b{p.p[3], p.p[4], p.p[5]};
                    const Vec c{p.p[6], p.p[7], p.p[8]};
                    const Vec ab = sub(b, a);
                    const Vec ac = sub(c, a);
                    const Vec n{ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
                    const double area2 = norm(n);
                    const double m = norm(ab) * norm(ac);
                    if(!(area2 > 0.0) || !(m > 0.0)) {
                        continue; // degenerate: never hit
                    }
                    const double worst_relative = 9.0 * kEps * kSafety * m / 1e-6;
                    if(worst_relative <= slack / 12.0) {
                        const double diameter = std::max(norm(ab), std::max(norm(ac), norm(sub(c, b))));
                        tau[w] = std::max(tau[w], worst_relative * diameter / (slack / 4.0));
                    }
                    else {
                        found[w].push_back(i);
                    }
                }
            };
            if(workers > 1U) {
                std::vector<std::thread> pool;
                for(unsigned w = 0; w < workers; w++) {
                    pool.emplace_back(scan, w);
                }
                for(std::thread &t : pool) {
                    t.join();
                }
            }
            else {
                scan(0U);
            }
            for(unsigned w = 0; w < workers; w++) {
                tau_safe = std::max(tau_safe, tau[w]);
                special.insert(special.end(), found[w].begin(), found[w].end());
            }
        }

        for(size_t at = 0; at < special.size() && g.certifiable != 0U; at++) {
            const ptb_prim &p = prims[special[at]];
            if(p.kind == PTB_PRIM_SPHERE) {
                if(g.n_spheres == static_cast<uint32_t>(kGuardSpheres)) {
                    g.certifiable = 0;
                    break;
                }
                for(int c = 0; c < 4; c++) {
                    g.spheres[g.n_spheres][c] = p.p[c];
                }
                g.n_spheres++;
                continue;
            }
            if(p.kind != PTB_PRIM_TRIANGLE) {
                continue;
            }
            const Vec a{p.p[0], p.p[1], p.p[2]};
            const Vec b{p.p[3], p.p[4], p.p[5]};
            const Vec c{p.p[6], p.p[7], p.p[8]};
            const Vec ab = sub(b, a);
            const Vec ac = sub(c, a);
            const Vec n{ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
            const double area2 = norm(n);
            const double m = norm(ab) * norm(ac);
            const double diameter = std::max(norm(ab), std::max(norm(ac), norm(sub(c, b))));
            if(!(area2 > 0.0) || !(m > 0.0)) {
                continue; // degenerate: det is 0 for every ray, the triangle is never hit
            }
            const double worst_relative = 9.0 * kEps * kSafety * m / 1e-6;
            if(worst_relative <= slack / 12.0) {
                tau_safe = std::max(tau_safe, worst_relative * diameter / (slack / 4.0));
                continue;
            }
            const double relative_at_normal_incidence = 9.0 * kEps * kSafety * m / area2;
            const double cone = 12.0 * relative_at_normal_incidence / slack;
            if(cone > 0.25) {
                g.certifiable = 0; // a sliver: its distances are noise from almost every direction
                break;
            }
            const Vec unit{n.x / area2, n.y / area2, n.z / area2};
            const double h = dot(unit, a);
            double w = 0.0;
            for(int corner = 0; corner < 8; corner++) {
                // the fp32 box of the triangle (object.cpp:184-186): min / max of the vertices per axis
                const Vec q{(corner & 1) ? std::max({a.x, b.x, c.x}) : std::min({a.x, b.x, c.x}), (corner & 2) ? std::max({a.y, b.y, c.y}) : std::min({a.y, b.y, c.y}),
                            (corner & 4) ? std::max({a.z, b.z, c.z}) : std::min({a.z, b.z, c.z})};
                w = std::max(w, std::fabs(dot(unit, q) - h));
            }
            const Vec lo{std::min({a.x, b.x, c.x}), std::min({a.y, b.y, c.y}), std::min({a.z, b.z, c.z})};
            const Vec hi{std::max({a.x, b.x, c.x}), std::max({a.y, b.y, c.y}), std::max({a.z, b.z, c.z})};
            const double k = 4.0 * relative_at_normal_incidence / slack;

            bool merged = false;
            for(uint32_t j = 0; j < g.n_planes && !merged; j++) {
                GuardPlane &pl = g.planes[j];
                const double along = unit.x * pl.nx + unit.y * pl.ny + unit.z * pl.nz;
                const double hj = along >= 0.0 ? h : -h;
                if(std::fabs(std::fabs(along) - 1.0) < 1e-8 && std::fabs(hj - pl.h) <= 1e-6 * (1.0 + std::fabs(hj))) {
                    pl.lox = std::min(pl.lox, static_cast<float>(lo.x));
                    pl.loy = std::min(pl.loy, static_cast<float>(lo.y));
                    pl.loz = std::min(pl.loz, static_cast<float>(lo.z));
                    pl.hix = std::max(pl.hix, static_cast<float>(hi.x));
                    pl.hiy = std::max(pl.hiy, static_cast<float>(hi.y));
                    pl.hiz = std::max(pl.hiz, static_cast<float>(hi.z));
                    pl.w = static_cast<float>(std::max<double>(pl.w, w * kMargin + 1e-7 * (1.0 + std::fabs(h))));
                    pl.cone = static_cast<float>(std::max<double>(pl.cone, cone * kMargin));
                    pl.k = static_cast<float>(std::max<double>(pl.k, k * kMargin));
                    merged = true;
                }
            }
            if(merged) {
                continue;
            }
            if(g.n_planes == static_cast<uint32_t>(kGuardPlanes)) {
                g.certifiable = 0;
                break;
            }
            GuardPlane &pl = g.planes[g.n_planes++];
            pl.nx = static_cast<float>(unit.x);
            pl.ny = static_cast<float>(unit.y);
            pl.nz = static_cast<float>(unit.z);
            pl.h = static_cast<float>(h);
            pl.lox = static_cast<float>(lo.x);
            pl.loy = static_cast<float>(lo.y);
            pl.loz = static_cast<float>(lo.z);
            pl.hix = static_cast<float>(hi.x);
            pl.hiy = static_cast<float>(hi.y);
            pl.hiz = static_cast<float>(hi.z);
            pl.w = static_cast<float>(w * kMargin + 1e-7 * (1.0 + std::fabs(h)));
            pl.cone = static_cast<float>(cone * kMargin);
            pl.k = static_cast<float>(k * kMargin);
            pl.r = 0.0F;
            pl.pad0 = pl.pad1 = 0.0F;
        }
        for(uint32_t j = 0; j < g.n_planes; j++) {
            GuardPlane &pl = g.planes[j];
            const double dx = static_cast<double>(pl.hix) - pl.lox;
            const double dy = static_cast<double>(pl.hiy) - pl.loy;
            const double dz = static_cast<double>(pl.hiz) - pl.loz;
            pl.r = static_cast<float>(0.5 * std::sqrt(dx * dx + dy * dy + dz * dz) * kMargin);
        }
        g.tau_safe = static_cast<float>(tau_safe * kMargin);
    }

}
