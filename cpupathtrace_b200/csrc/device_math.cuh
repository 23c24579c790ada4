// Small fp32 vector helpers for the device code.
//
// Numerics contract (DESIGN.md "Numerics"): this translation unit is compiled with `-fmad=false`, IEEE division and
// square root (nvcc defaults -prec-div=true -prec-sqrt=true), so that every +,-,*,/ and sqrt below rounds exactly like
// the reference's x86 build without FMA contraction.  Sums are written in the association order of the reference's
// loops (util/vector.h:138-147, 196-205: ((x*x + y*y) + z*z)).
#ifndef PTB_DEVICE_MATH_CUH
#define PTB_DEVICE_MATH_CUH

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define PTB_DEV __device__ __forceinline__

namespace ptb {

    // Path-pool, queue and per-sample traffic is streamed: every array is read or written once per kernel and is far larger
    // than the 126 MB L2, while the scene's nodes and primitives (233 MiB on the bench scene) are re-read by every ray.
    // PTB_STREAM_HINTS=1 marks the former evict-first (ld.global.cs / st.global.cs) so that it does not push the scene out
    // of L2 between two traversal launches; PTB_STREAM_HINTS=2 does so in the traversal kernels only (sld_t / sst_t).
#ifndef PTB_STREAM_HINTS
#define PTB_STREAM_HINTS 2 // measured on the bench scene: 1 -> trace -2 % but shade + accumulate +13 % (what shade writes is re-read at once by the
                           // shadow trace and accumulate); 2 -> closest trace 316.4 -> 309.6 ms, shadow 166.7 -> 164.4 ms per frame, the rest unchanged
#endif
    template<typename T>
    __device__ __forceinline__ T sld_t(const T *p) {
#if PTB_STREAM_HINTS
        return __ldcs(p);
#else
        return *p;
#endif
    }
    template<typename T>
    __device__ __forceinline__ void sst_t(T *p, T v) {
#if PTB_STREAM_HINTS
        __stcs(p, v);
#else
        *p = v;
#endif
    }
    template<typename T>
    __device__ __forceinline__ T sld(const T *p) {
#if PTB_STREAM_HINTS == 1
        return __ldcs(p);
#else
        return *p;
#endif
    }
    template<typename T>
    __device__ __forceinline__ void sst(T *p, T v) {
#if PTB_STREAM_HINTS == 1
        __stcs(p, v);
#else
        *p = v;
#endif
    }
#if PTB_STREAM_HINTS == 1
    template<>
    __device__ __forceinline__ unsigned long sld<unsigned long>(const unsigned long *p) {
        return static_cast<unsigned long>(__ldcs(reinterpret_cast<const unsigned long long *>(p)));
    }
    template<>
    __device__ __forceinline__ void sst<unsigned long>(unsigned long *p, unsigned long v) {
        __stcs(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v));
    }
#endif


    constexpr float kFloatMax = 3.402823466e+38F;
    constexpr float kPi = 3.14159274101257324219F;    // static_cast<float>(M_PI)
    constexpr float kTwoPi = 6.28318548202514648438F; // 2.0F * kPi (exact doubling)

    struct V3 {
        float x, y, z;
    };

    struct V4 {
        float x, y, z, w;
    };

    PTB_DEV V3 mk3(float x, float y, float z) {
        return V3{x, y, z};
    }

    PTB_DEV V3 operator+(V3 a, V3 b) {
        return V3{a.x + b.x, a.y + b.y, a.z + b.z};
    }

    PTB_DEV V3 operator-(V3 a, V3 b) {
        return V3{a.x - b.x, a.y - b.y, a.z - b.z};
    }

    PTB_DEV V3 operator-(V3 a) {
        return V3{-a.x, -a.y, -a.z};
    }

    PTB_DEV V3 operator*(V3 a, float s) {
        return V3{a.x * s, a.y * s, a.z * s};
    }

    PTB_DEV float dot(V3 a, V3 b) {
        return (a.x * b.x + a.y * b.y) + a.z * b.z;
    }

    PTB_DEV V3 cross(V3 a, V3 b) {
        return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    }

    PTB_DEV float length2(V3 a) {
        return (a.x * a.x + a.y * a.y) + a.z * a.z;
    }

    PTB_DEV float length(V3 a) {
        return sqrtf(length2(a));
    }

    // rt_vector::normalize (util/vector.h:161-167): multiply by the reciprocal of the length
    PTB_DEV V3 normalize(V3 a) {
        const float inv = 1.0F / length(a);
        return a * inv;
    }

    // reflect (util/vector.h:250-255): v - n * 2 * d, evaluated as ((n * 2) * d)
    PTB_DEV V3 reflect(V3 v, V3 n) {
        const float d = dot(v, n);
        return v - (n * 2.0F) * d;
    }

    // ---- correctly rounded, never contracted: the same IEEE operations in every build.
    //
    // The production-math build of the shade kernels (ptb_fast.cu: FMA contraction, approximate division / square root)
    // may evaluate BSDFs and weights any way it likes, but one chain of the reference's arithmetic decides a coin that
    // the image's brightness depends on and must round identically: a next-event sample lies ON the emissive triangle,
    // the shadow ray starts at pos + dir * eps and the reference calls the light unoccluded iff the closest hit --
    // the light's own surface, at |to_light| - eps up to rounding -- has t >= |to_light| - eps (worker.cpp:80-86).
    // Whether that holds is decided by the last bits of pos, of the sampled point, of the direction and of the limit.
    // With contracted or approximate arithmetic the coin lands differently often and the image comes out 3-4 % brighter
    // than the reference's (measured, tests/test_parity_gpu.py::test_bench_scene_image_rmse_within_noise).  Hence: hit
    // position, sampled light position, shadow-ray direction, origin and limit use these helpers in every build.
    PTB_DEV V3 exactAdd(V3 a, V3 b) {
        return V3{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)};
    }

    PTB_DEV V3 exactSub(V3 a, V3 b) {
        return V3{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)};
    }

    PTB_DEV V3 exactScale(V3 a, float s) {
        return V3{__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)};
    }

    PTB_DEV float exactLength2(V3 a) {
        return __fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z));
    }

    PTB_DEV float exactLength(V3 a) {
        return __fsqrt_rn(exactLength2(a));
    }

    PTB_DEV V3 exactNormalize(V3 a) {
        return exactScale(a, __fdiv_rn(1.0F, exactLength(a)));
    }

    PTB_DEV V4 operator+(V4 a, V4 b) {
        return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
    }

    PTB_DEV V4 operator*(V4 a, V4 b) {
        return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
    }

    PTB_DEV V4 operator*(V4 a, float s) {
        return V4{a.x * s, a.y * s, a.z * s, a.w * s};
    }

    PTB_DEV V4 operator/(V4 a, float s) {
        return V4{a.x / s, a.y / s, a.z / s, a.w / s};
    }

    PTB_DEV V4 ld4(const float4 *p) {
        const float4 v = __ldg(p);
        return V4{v.x, v.y, v.z, v.w};
    }

    // One 256-bit read-only load (LDG.E.256, sm_100+): 32 bytes, 32-byte aligned, two float4 lanes per request.
    // Node records are fetched with two of these instead of four 128-bit loads, which halves the L1 wavefronts per
    // node visit (profiles/r01_ncu_trace_vote.md: l1tex__data_pipe_lsu_wavefronts was the top limiter at 71 %).
    // (eviction priorities on these loads -- ld.global.nc.L2::evict_last, with and without L1::evict_last -- were measured in
    // round 2: no effect on the trace kernels, 309.6 ms of closest-hit trace per frame either way)
    PTB_DEV void ld256(const float4 *p, float4 &a, float4 &b) {
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(p));
    }

    // the same for read-write data (path pool): plain global load, cached in L2 only
    PTB_DEV void ld256cg(const float4 *p, float4 &a, float4 &b) {
        asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(p)
                     : "memory");
    }

    PTB_DEV float4 f4(V4 v) {
        return make_float4(v.x, v.y, v.z, v.w);
    }

    PTB_DEV V4 v4(float4 v) {
        return V4{v.x, v.y, v.z, v.w};
    }

    // std::min / std::max select semantics (first argument wins ties and NaN comparisons)
    PTB_DEV float stdmin(float a, float b) {
        return b < a ? b : a;
    }

    PTB_DEV float stdmax(float a, float b) {
        return a < b ? b : a;
    }

}

#endif
