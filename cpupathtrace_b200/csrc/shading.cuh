// Path-vertex arithmetic: camera rays, surface normals, light sampling, BSDFs and the per-vertex step of
// impl::getSample, restated for the device in the reference's evaluation order.
//
//   Camera::shootRay, aperture samplers      reference src/camera.cpp:7-49, 78-113
//   Triangle/Sphere::getSurfaceNormal        reference src/scene/object.cpp:126-144, 86-88
//   Triangle/Sphere::sampleSurface           reference src/scene/object.cpp:192-207, 101-116
//   Scene::sampleLights                      reference src/scene/scene.cpp:222-289
//   importanceSampleCosine / localToGlobal   reference src/scene/propagation.cpp:11-62
//   getFresnelReflectance                    reference src/scene/propagation.cpp:64-83
//   Lambertian / Glass / Mirror              reference src/scene/propagation.cpp:89-217
//   impl::getSample (one loop iteration)     reference src/worker.cpp:44-138
//
// libm: sqrt and division are IEEE-exact; sin, cos, acos and pow(x, 0.5) are the bit-exact restatements of glibc's
// routines in glibc_libm.cuh; pow(x, 1) is x (glibc's powf returns x exactly).  With identical random numbers the
// radiance of a sample is therefore expected to be bit-identical to the reference's (DESIGN.md "Numerics").
#ifndef PTB_SHADING_CUH
#define PTB_SHADING_CUH

#include "../../include/ptb.h"
#include "device_scene.cuh"
#include "glibc_libm.cuh"
#include "rng.cuh"

namespace ptb {

    // ------------------------------------------------------------------------------------------------ camera

    template<typename RNG>
    PTB_DEV void sampleAperture(const ptb_camera &c, RNG &rng, float &sx, float &sy) {
        if(c.aperture_kind == PTB_APERTURE_CIRCULAR) {
            float u_r;
            float u_theta;
            rng.canonicalPair(u_r, u_theta);
            const float r = sqrtf(u_r);
            const float theta = kTwoPi * u_theta;
            sx = r * pathCosf(theta);
            sy = r * pathSinf(theta);
            return;
        }
        // hexagonal: rejection sample the upper-right quadrant, then two fair coin flips for the signs
        const float ratio = c.hexagon_horizontal_ratio;
        float x;
        float y;
        bool inside;
        do {
            x = rng.uniform01();
            y = rng.uniform01();
            const float relative_x = x - ratio;
            inside = (relative_x <= 0.0F) || (relative_x / (1.0F - ratio)) >= y;
        } while(!inside);
        if(rng.bernoulli(0.5)) {
            x = -x;
        }
        if(rng.bernoulli(0.5)) {
            y = -y;
        }
        sx = x;
        sy = y;
    }

    template<typename RNG>
    PTB_DEV void shootRay(const ptb_camera &c, float x, float y, float pixel_width, float pixel_height, RNG &rng, V3 &ray_o, V3 &ray_d) {
        // uniform_real_distribution<float>(a, b): u * (b - a) + a, x first (camera.cpp:79-83)
        float u_x;
        float u_y;
        rng.canonicalPair(u_x, u_y);
        const float offset_x = u_x * (pixel_width / 2.0F - -pixel_width / 2.0F) + -pixel_width / 2.0F;
        const float offset_y = u_y * (pixel_height / 2.0F - -pixel_height / 2.0F) + -pixel_height / 2.0F;
        const float sensor_x = x + offset_x;
        const float sensor_y = y + offset_y;

        const V3 origin = mk3(c.origin[0], c.origin[1], c.origin[2]);
        const V3 forward = mk3(c.forward[0], c.forward[1], c.forward[2]);
        const V3 up = mk3(c.up[0], c.up[1], c.up[2]);
        const V3 right = mk3(c.right[0], c.right[1], c.right[2]);

        const V3 sensor_pos = ((origin - forward) - up * sensor_y) - right * sensor_x;

        float aperture_offset_x = 0.0F;
        float aperture_offset_y = 0.0F;
        if(c.aperture_kind != PTB_APERTURE_NONE) {
            float sx;
            float sy;
            sampleAperture(c, rng, sx, sy);
            aperture_offset_x = sx * c.aperture_width_half;
            aperture_offset_y = sy * c.aperture_height_half;
        }
        // quirk kept: the x offset moves along `up`, the y offset along `right` (camera.cpp:99)
        ray_o = (origin + up * aperture_offset_x) + right * aperture_offset_y;

        if(c.focal_plane_dist > 0.0F) {
            const V3 base_dir = normalize(origin - sensor_pos);
            const V3 ray_target = origin + base_dir * (c.focal_plane_dist / dot(forward, base_dir));
            ray_d = normalize(ray_target - ray_o);
        }
        else {
            ray_d = normalize(ray_o - sensor_pos);
        }
    }

    // processItem's pixel -> camera-space mapping (worker.cpp:168-170)
    PTB_DEV void pixelToCamera(int px, int py, int width, int height, float &x_camera, float &y_camera) {
        x_camera = 2.0F * ((static_cast<float>(px) + 0.5F) / static_cast<float>(width) - 0.5F);
        y_camera = 2.0F * ((static_cast<float>(py) + 0.5F) / static_cast<float>(height) - 0.5F);
        y_camera = -y_camera;
    }

    // ------------------------------------------------------------------------------------------------ surfaces

    struct Material {
        V4 diffuse;
        V4 emission;
        float ior;
        uint32_t bsdf;
        bool one_way;
    };

    PTB_DEV Material loadMaterial(const DeviceScene &s, uint32_t index) {
        const float4 *m = s.mats + 3 * static_cast<size_t>(index);
        Material out;
        out.diffuse = ld4(m);
        out.emission = ld4(m + 1);
        const float4 misc = __ldg(m + 2);
        out.ior = misc.x;
        out.bsdf = __float_as_uint(misc.y);
        out.one_way = __float_as_uint(misc.z) != 0U;
        return out;
    }

    // Object::getSurfaceNormal for the primitive in `slot`; also returns its material index
    PTB_DEV V3 surfaceNormal(const DeviceScene &s, uint32_t slot, V3 pos, uint32_t &material) {
        const float4 *g = s.geom + kGeomLanes * static_cast<size_t>(slot);
        const float4 *sh = s.shade + 3 * static_cast<size_t>(slot);
        const float4 g0 = __ldg(g);
        const float4 s0 = __ldg(sh);
        material = __float_as_uint(s0.w);
        const uint32_t kind = __float_as_uint(g0.w) & kKindMask;
        if(kind == PTB_PRIM_SPHERE) {
            return normalize(pos - mk3(g0.x, g0.y, g0.z));
        }
        if(kind != PTB_PRIM_TRIANGLE) {
            return mk3(0.0F, 1.0F, 0.0F); // NullObject::getSurfaceNormal (object.cpp:56-58)
        }
        const float4 g1 = __ldg(g + 1);
        const float4 g2 = __ldg(g + 2);
        const float4 s1 = __ldg(sh + 1);
        const float4 s2 = __ldg(sh + 2);
        const V3 ab = mk3(g1.x, g1.y, g1.z);
        const V3 ac = mk3(g2.x, g2.y, g2.z);
        const V3 ap = pos - mk3(g0.x, g0.y, g0.z);
        const float d00 = dot(ab, ab);
        const float d01 = dot(ab, ac);
        const float d11 = dot(ac, ac);
        const float d20 = dot(ap, ab);
        const float d21 = dot(ap, ac);
        const float inv_d = 1.0F / (d00 * d11 - d01 * d01);
        const float v = (d11 * d20 - d01 * d21) * inv_d;
        const float w = (d00 * d21 - d01 * d20) * inv_d;
        const float u = 1.0F - v - w;
        const V3 na = mk3(s0.x, s0.y, s0.z);
        const V3 nb = mk3(s1.x, s1.y, s1.z);
        const V3 nc = mk3(s2.x, s2.y, s2.z);
        return normalize((na * u + nb * v) + nc * w);
    }

    // ------------------------------------------------------------------------------------------------ lights

    // Object::sampleSurface for a primitive given as its three un-differenced lanes
    // (triangle: a, b, c; sphere: origin / (radius, radius^2)): uniformly sampled point, density, cull flag.
    template<typename RNG>
    PTB_DEV void samplePrimSurface(float4 e0, float4 e1, float4 e2, uint32_t flags, RNG &rng, V3 &surface_pos, float &surface_p, bool &surface_cull) {
        if((flags & kKindMask) == PTB_PRIM_TRIANGLE) {
            const V3 a = mk3(e0.x, e0.y, e0.z);
            const V3 b = mk3(e1.x, e1.y, e1.z);
            const V3 c = mk3(e2.x, e2.y, e2.z);
            float r1;
            float r2;
            rng.canonicalPair(r1, r2);
            // correctly rounded in every build: the point's last bits take part in the light's self-occlusion coin (device_math.cuh)
            const float rr1 = __fsqrt_rn(r1);
            surface_pos = exactAdd(exactAdd(exactScale(a, __fsub_rn(1.0F, rr1)), exactScale(b, __fmul_rn(rr1, __fsub_rn(1.0F, r2)))), exactScale(c, __fmul_rn(rr1, r2)));
            // p = 1 / (|cross(b - a, c - a)| / 2), evaluated once on the host with the same operations (host_math.cpp)
            surface_p = e1.w;
            surface_cull = (flags & kCullBit) != 0U;
        }
        else if((flags & kKindMask) == PTB_PRIM_SPHERE) {
            const V3 origin = mk3(e0.x, e0.y, e0.z);
            const float radius = e1.x;
            float u_theta;
            float u_phi;
            rng.canonicalPair(u_theta, u_phi);
            const float theta = kTwoPi * u_theta;
            const float phi = pathAcosf(1.0F - 2.0F * u_phi);
            const float x = pathSinf(phi) * pathCosf(theta);
            const float y = pathSinf(phi) * pathSinf(theta);
            const float z = pathCosf(phi);
            surface_pos = origin + mk3(x, y, z) * radius;
            surface_p = e1.w; // 1 / (4 pi r^2), evaluated once on the host
            surface_cull = false;
        }
        else {
            // Object::sampleSurface default (object.cpp:48-50)
            surface_pos = mk3(0.0F, 0.0F, 0.0F);
            surface_p = 0.0F;
            surface_cull = false;
        }
    }

    struct LightSample {
        V3 pos;
        V4 spectrum;
        float pd;
    };

    // Scene::sampleLights.  `emit(LightSample)` is called once per produced sample, in the reference's order:
    // explicit lights first, then the emissive-object samples that survive the rejection tests.  All engine draws of
    // a rejected sample are consumed before the rejection (scene.cpp:239-277).
    template<typename RNG, typename Emit>
    PTB_DEV void sampleLights(const DeviceScene &s, V3 pos, RNG &rng, Emit emit) {
        for(uint32_t i = 0; i < s.n_lights; i++) {
            LightSample ls;
            const float4 p = __ldg(s.lights + 2 * i);
            ls.pos = mk3(p.x, p.y, p.z);
            ls.spectrum = ld4(s.lights + 2 * i + 1);
            ls.pd = 1.0F; // PointLightSource::importanceSample (light.cpp:35-37)
            emit(ls);
        }

        const uint32_t count = s.object_sample_count;
        for(uint32_t i = 0; i < count; i++) {
            const float r = rng.uniform01();

            // std::lower_bound over the cumulative probabilities
            uint32_t lo = 0;
            uint32_t n = s.n_emissive;
            while(n > 0U) {
                const uint32_t half = n >> 1;
                if(__ldg(s.cdf + lo + half) < r) {
                    lo += half + 1U;
                    n -= half + 1U;
                }
                else {
                    n = half;
                }
            }
            const uint32_t index = lo < s.n_emissive ? lo : s.n_emissive - 1U;

            float selection_p = __ldg(s.cdf + index);
            if(index > 0U) {
                selection_p -= __ldg(s.cdf + index - 1U);
            }
            selection_p *= static_cast<float>(count);

            const float4 *e = s.emis + 3 * static_cast<size_t>(index);
            const float4 e0 = __ldg(e);
            const float4 e1 = __ldg(e + 1);
            const uint32_t slot = __float_as_uint(e0.w);
            const float4 g0 = __ldg(s.geom + kGeomLanes * static_cast<size_t>(slot));
            const uint32_t flags = __float_as_uint(g0.w);

            V3 surface_pos;
            float surface_p;
            bool surface_cull;
            samplePrimSurface(e0, e1, __ldg(e + 2), flags, rng, surface_pos, surface_p, surface_cull);

            uint32_t material_index;
            const V3 surface_n = surfaceNormal(s, slot, surface_pos, material_index);

            const V3 to_light = surface_pos - pos;
            const V3 dir = normalize(to_light);
            const float abs_dot = fabsf(dot(-dir, surface_n));
            if(!(abs_dot > 0.0F)) {
                continue;
            }
            if(!(length2(to_light) > 0.0F)) {
                continue;
            }
            if(surface_cull) {
                if(!(dot(dir, surface_n) < 0.0F)) {
                    continue;
                }
            }
            const float conversion_factor = length2(to_light) / abs_dot;

            LightSample ls;
            ls.pos = surface_pos;
            ls.spectrum = ld4(s.mats + 3 * static_cast<size_t>(material_index) + 1);
            ls.pd = selection_p * surface_p * conversion_factor;
            emit(ls);
        }
    }

    // ------------------------------------------------------------------------------------------------ BSDFs

    PTB_DEV V3 localToGlobal(V3 vec, V3 n) {
        V3 d;
        if(fabsf(n.x) > 0.0F) {
            if(fabsf(n.y) > 0.0F) {
                d = mk3(0.0F, -n.x, n.y);
            }
            else {
                d = mk3(0.0F, -n.x, n.z);
            }
        }
        else {
            if(fabsf(n.y) > 0.0F) {
                d = mk3(-n.y, n.z, 0.0F);
            }
            else {
                d = mk3(1.0F, 0.0F, 0.0F);
            }
        }
        d = normalize(d);
        const V3 b1 = normalize(cross(d, n));
        const V3 b2 = normalize(cross(b1, n));
        const V3 vx = mk3(b1.x, b2.x, n.x);
        const V3 vy = mk3(b1.y, b2.y, n.y);
        const V3 vz = mk3(b1.z, b2.z, n.z);
        return mk3(dot(vx, vec), dot(vy, vec), dot(vz, vec));
    }

    PTB_DEV void fresnelReflectance(float ray_dot, float ri_leaving, float ri_entering, float &reflectance, float &cos_theta_t) {
        const float sin_theta_i = sqrtf(stdmax(1.0F - ray_dot * ray_dot, 0.0F));
        const float sin_theta_t = ri_leaving / ri_entering * sin_theta_i;
        if(sin_theta_t >= 1.0F) {
            reflectance = 1.0F;
            cos_theta_t = 0.0F;
            return;
        }
        cos_theta_t = sqrtf(stdmax(1.0F - sin_theta_t * sin_theta_t, 0.0F));
        const float r_parallel = ((ri_entering * ray_dot) - (ri_leaving * cos_theta_t)) / ((ri_entering * ray_dot) + (ri_leaving * cos_theta_t));
        const float r_perpendicular = ((ri_leaving * ray_dot) - (ri_entering * cos_theta_t)) / ((ri_leaving * ray_dot) + (ri_entering * cos_theta_t));
        reflectance = (r_parallel * r_parallel + r_perpendicular * r_perpendicular) / 2.0F;
    }

    // BSDF::propagateRay: next ray, radiance factor, probability density
    template<typename RNG>
    PTB_DEV void propagateRay(const Material &m, V3 ray_d, V3 pos, V3 normal, float epsilon, RNG &rng, V3 &out_o, V3 &out_d, float &factor, float &pd) {
        if(m.bsdf == PTB_BSDF_LAMBERT) {
            // importanceSampleCosine(dist(re), dist(re), 1.0F): g++ evaluates the second argument first, so the first
            // draw is r2 (the cos-theta variate) and the second is r1 (the azimuth)   [SURVEY.md App. B]
            float r2;
            float r1;
            rng.canonicalPair(r2, r1);
            const float fac = sqrtf(1.0F - r2);         // pow(r2, 2/(e+1)) with e = 1: powf(x, 1) == x
            const float cos_theta = pathPowfHalf(r2);  // pow(r2, 1/(e+1))
            const float angle = kTwoPi * r1;
            const V3 local = mk3(fac * pathCosf(angle), fac * pathSinf(angle), cos_theta);
            const float p = 2.0F * cos_theta / kTwoPi; // (e+1) * pow(cos_theta, e) / (2 pi)
            const V3 dir = localToGlobal(local, normal);
            out_o = pos + dir * epsilon;
            out_d = dir;
            factor = 1.0F;
            pd = p;
            return;
        }
        if(m.bsdf == PTB_BSDF_GLASS) {
            const float ray_dot = -dot(ray_d, normal);
            const float ri_leaving = ray_dot >= 0.0F ? 1.0F : m.ior;
            const float ri_entering = ray_dot >= 0.0F ? m.ior : 1.0F;
            float rat;
            float cos_theta_t;
            fresnelReflectance(fabsf(ray_dot), ri_leaving, ri_entering, rat, cos_theta_t);
            const float side = ray_dot < 0.0F ? -1.0F : 1.0F;
            if(rng.bernoulli(static_cast<double>(rat))) {
                const V3 dir = reflect(ray_d, normal * side);
                out_o = pos + dir * epsilon;
                out_d = dir;
                factor = rat;
                pd = rat;
            }
            else {
                const float ri_ratio = ri_leaving / ri_entering;
                V3 dir = ray_d * ri_ratio + (normal * (ri_ratio * fabsf(ray_dot) - cos_theta_t)) * side;
                dir = normalize(dir);
                const float ri_fac = (ri_entering * ri_entering) / (ri_leaving * ri_leaving);
                out_o = pos + dir * epsilon;
                out_d = dir;
                factor = ri_fac * (1.0F - rat);
                pd = 1.0F - rat;
            }
            return;
        }
        // mirror
        const bool unaligned = dot(ray_d, normal) > 0.0F;
        if(m.one_way && unaligned) {
            out_o = pos + ray_d * epsilon;
            out_d = ray_d;
            factor = 1.0F;
            pd = 1.0F;
            return;
        }
        V3 normal_dir = normal;
        if(!m.one_way && unaligned) {
            normal_dir = normal_dir * -1.0F;
        }
        const V3 dir = reflect(ray_d, normal_dir);
        out_o = pos + dir * epsilon;
        out_d = dir;
        factor = 1.0F;
        pd = 1.0F;
    }

    // BSDF::getSpectrum: incoming spectrum, shading factor, probability density
    PTB_DEV void bsdfSpectrum(const Material &m, V3 from_d, V3 to_d, V3 normal, V4 light, bool synthetic, V4 &spectrum, float &shade, float &pd) {
        if(m.bsdf == PTB_BSDF_LAMBERT) {
            shade = stdmax(dot(normal, to_d), 0.0F) / kPi;
            spectrum = m.diffuse * light;
            pd = 1.0F;
            return;
        }
        shade = 1.0F;
        pd = synthetic ? 0.0F : 1.0F;
        if(m.bsdf == PTB_BSDF_GLASS) {
            // reflection side: specular colour (white, material.cpp:15-17); transmission side: diffuse colour
            spectrum = dot(from_d, to_d) <= 0.0F ? light : light * m.diffuse;
            return;
        }
        spectrum = light; // mirror: specular colour is white whether or not the one-way test passes
    }

    // ------------------------------------------------------------------------------------------------ path vertex

    template<typename RNG>
    struct PathRegs {
        V3 ray_o;
        V3 ray_d;
        V4 throughput;   // sample_spectrum
        V4 radiance;     // out_spectrum
        double divisor;  // sample_divisor
        double bounce_pd; // sample_bounce_pd
        float contribution_unweighted;
        int path_length;
        RNG rng;
    };

    template<typename RNG>
    PTB_DEV void initPath(PathRegs<RNG> &p) {
        p.throughput = V4{1.0F, 1.0F, 1.0F, 1.0F};
        p.radiance = V4{0.0F, 0.0F, 0.0F, 0.0F};
        p.divisor = 1.0;
        p.bounce_pd = 1.0;
        p.contribution_unweighted = 1.0F;
        p.path_length = 0;
    }

    struct ShadowCandidate {
        V3 o;
        V3 d;
        float limit;       // occluded iff a hit with 0 <= t < limit exists  (worker.cpp:86: |to_light| - epsilon)
        V4 contribution;   // added to the radiance iff unoccluded
    };

    // One iteration of the getSample loop for a path whose ray hit `slot` at distance t (worker.cpp:50-138).
    // `shadow(candidate, is_null)` is invoked per next-event-estimation sample, in light order; is_null marks the
    // samples that cannot change the radiance whatever their visibility: those whose BSDF returns pd 0 for synthetic
    // rays (Glass, Mirror: the reference traces their shadow ray and discards the result, worker.cpp:84-92), and those
    // whose contribution is +-0 in all three channels (a Lambertian surface facing away from the light has shading
    // factor max(n.l, 0) = 0, propagation.cpp:104-108; adding +-0 to the non-negative radiance sum leaves every bit of
    // it unchanged).  They carry geometry but no weight.
    // Returns true when the path continues with the new ray in p.
    template<typename RNG, typename Shadow>
    PTB_DEV bool shadeVertex(const DeviceScene &s, float epsilon, int max_depth, PathRegs<RNG> &p, float t, uint32_t slot, Shadow shadow) {
        p.path_length++;

        const V3 pos = exactAdd(p.ray_o, exactScale(p.ray_d, t));
        uint32_t material_index;
        const V3 n = surfaceNormal(s, slot, pos, material_index);
        const Material m = loadMaterial(s, material_index);

        // a non-emissive surface adds exactly +0 to every channel (the divisor is finite and non-zero here), so the four
        // IEEE divisions are skipped for it
        if(m.emission.x != 0.0F || m.emission.y != 0.0F || m.emission.z != 0.0F) {
            p.radiance = p.radiance + (p.throughput * m.emission) / static_cast<float>(p.divisor * p.bounce_pd);
        }

        const float contribution = ((p.throughput.x + p.throughput.y) + p.throughput.z) / 3.0F;
        const float bounce_probability = p.path_length <= 4 ? 1.0F : 0.1F + 0.1F * stdmin(p.contribution_unweighted * contribution, 1.0F);

        p.rng.startBounce(p.path_length);
        const bool do_bounce = p.rng.uniform01() < bounce_probability;

        sampleLights(s, pos, p.rng, [&](const LightSample &ls) {
            // shadow-ray geometry: correctly rounded in every build (device_math.cuh "correctly rounded, never contracted")
            const V3 to_light = exactSub(ls.pos, pos);
            const V3 light_dir = exactNormalize(to_light);
            V4 base;
            float shading_factor;
            float shadow_ray_pd;
            bsdfSpectrum(m, p.ray_d, light_dir, n, ls.spectrum, true, base, shading_factor, shadow_ray_pd);
            ShadowCandidate c;
            c.o = exactAdd(pos, exactScale(light_dir, epsilon));
            c.d = light_dir;
            c.limit = __fsub_rn(exactLength(to_light), epsilon);
            if(shadow_ray_pd > 0.0F) {
                const V4 combined = (base * shading_factor) * p.throughput;
                c.contribution = combined / static_cast<float>(p.divisor * p.bounce_pd * ls.pd * shadow_ray_pd);
                shadow(c, c.contribution.x == 0.0F && c.contribution.y == 0.0F && c.contribution.z == 0.0F);
            }
            else {
                c.contribution = V4{0.0F, 0.0F, 0.0F, 0.0F};
                shadow(c, true);
            }
        });

        if(!do_bounce) {
            return false;
        }
        p.bounce_pd *= bounce_probability;
        if(p.bounce_pd <= 1E-20) {
            return false;
        }
        if(max_depth > 0 && p.path_length >= max_depth) {
            return false;
        }

        V3 next_o;
        V3 next_d;
        float ray_factor;
        float ray_pd;
        propagateRay(m, p.ray_d, pos, n, epsilon, p.rng, next_o, next_d, ray_factor, ray_pd);
        p.divisor *= ray_pd;
        p.divisor /= ray_factor;
        p.contribution_unweighted *= ray_factor;

        V4 shaded;
        float shading_factor;
        float shading_pd;
        bsdfSpectrum(m, p.ray_d, next_d, n, p.throughput, false, shaded, shading_factor, shading_pd);
        p.divisor *= shading_pd;
        p.divisor /= shading_factor;
        p.contribution_unweighted *= shading_factor;
        p.throughput = shaded;

        if(p.divisor <= 1E-20) {
            return false;
        }
        p.ray_o = next_o;
        p.ray_d = next_d;
        return true;
    }

}

#endif
