// Bit-exact device restatement of the four glibc 2.39 libm routines the reference's hot path calls through <cmath>:
//
//   sinf, cosf   sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h, s_sincosf_data.c   (ARM optimized-routines)
//   powf         sysdeps/ieee754/flt-32/e_powf.c, e_powf_log2_data.c, e_exp2f_data.c         (ARM optimized-routines)
//   acosf        sysdeps/ieee754/flt-32/e_acosf.c                                            (fdlibm, fp32 arithmetic)
//
// glibc is a toolchain dependency of the reference (std::sin/std::cos/std::pow/std::acos on float), not vendored in
// /root/reference; pinned version: Ubuntu GLIBC 2.39-0ubuntu8.5, the libm both this container and the GPU box run.
// The reference's callers: propagation.cpp:12-18 (pow, cos, sin), camera.cpp:15-16 (cos, sin), object.cpp:107-110
// (acos, sin, cos).  CUDA's own sinf/cosf/powf/acosf are accurate to 1-2 ulp but not IDENTICAL to glibc's, and this
// scene amplifies about half of all 1-ulp differences past the 1e-4 radiance tolerance (DESIGN.md "Numerics"), so the
// published algorithms are restated here operation for operation:
//   * sinf/cosf/powf evaluate in fp64 (polynomials + table look-ups) and round once to fp32.  glibc selects an
//     FMA-contracted build of the same source on FMA-capable hosts; fused and unfused evaluation round to the same
//     fp32 result on every input tested (oracle/libm_check.c: 2e8 arguments each), fma() is used here.
//   * acosf evaluates in fp32 WITHOUT contraction (glibc ships no FMA variant of it); this file is compiled with
//     -fmad=false.
// Only the argument ranges the path can produce take the exact route: |x| < 120 for sinf/cosf (angles are in
// [0, 2 pi]), x in [0, 1) with exponent 0.5 for powf.  The polynomial and table constants are those of the published
// algorithms; oracle/libm_check.c verifies this restatement against the system libm on the CPU.
#ifndef PTB_GLIBC_LIBM_CUH
#define PTB_GLIBC_LIBM_CUH

#include "device_math.cuh"

namespace ptb {

    // ---- sinf / cosf

    struct SinCosTable {
        double sign[4];
        double hpi_inv; // 2 / pi * 2^24
        double hpi;     // pi / 2
        double c0, c1, c2, c3, c4;
        double s1, s2, s3;
    };

    __device__ const SinCosTable kSinCosTable[2] = {
        {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10,
         0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
        {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10,
         -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};

    PTB_DEV uint32_t absTop12(float x) {
        return (__float_as_uint(x) >> 20) & 0x7ffU;
    }

    // sinf_poly of s_sincosf.h: sine polynomial for even quadrants, cosine polynomial for odd ones
    PTB_DEV float sinCosPoly(double x, double x2, const SinCosTable &p, int n) {
        if((n & 1) == 0) {
            const double x3 = x * x2;
            const double s1 = fma(x2, p.s3, p.s2);
            const double x7 = x3 * x2;
            const double s = fma(x3, p.s1, x);
            return static_cast<float>(fma(x7, s1, s));
        }
        const double x4 = x2 * x2;
        const double c2 = fma(x2, p.c4, p.c3);
        const double c1 = fma(x2, p.c1, p.c0);
        const double x6 = x4 * x2;
        const double c = fma(x4, p.c2, c1);
        return static_cast<float>(fma(x6, c2, c));
    }

    // Arguments outside the ranges the render path can produce go to CUDA's own routines through real calls, so that
    // the (never taken) fallbacks are not if-converted into the hot path.
    __device__ __noinline__ float libmFallbackSinCos(float y, int which) {
        return which == 0 ? sinf(y) : cosf(y);
    }

    __device__ __noinline__ float libmFallbackPow(float x, float y) {
        return powf(x, y);
    }

    // which = 0: sinf(y), which = 1: cosf(y)
    PTB_DEV float glibcSinCos(float y, int which) {
        const uint32_t top = absTop12(y);
        double x = static_cast<double>(y);
        if(top < absTop12(0x1.921FB6p-1F)) { // |y| < pi/4
            if(top < absTop12(0x1p-12F)) {
                return which == 0 ? y : 1.0F;
            }
            return sinCosPoly(x, x * x, kSinCosTable[0], which);
        }
        if(top < absTop12(120.0F)) {
            const double r = x * kSinCosTable[0].hpi_inv;
            const int n = (static_cast<int>(r) + 0x800000) >> 24;
            x = fma(-static_cast<double>(n), kSinCosTable[0].hpi, x);
            const double s = kSinCosTable[0].sign[n & 3];
            const SinCosTable &p = kSinCosTable[(n & 2) ? 1 : 0];
            return sinCosPoly(x * s, x * x, p, which == 0 ? n : (n ^ 1));
        }
        return libmFallbackSinCos(y, which); // |y| >= 120: outside the path's argument range
    }

    PTB_DEV float glibcSinf(float y) {
        return glibcSinCos(y, 0);
    }

    PTB_DEV float glibcCosf(float y) {
        return glibcSinCos(y, 1);
    }

    // ---- powf(x, 0.5F) for finite x >= 0

#define PTB_TABLE __device__ const
#include "glibc_libm_tables.h"
#undef PTB_TABLE

    // powf(x, y) for finite x >= 0 and finite y != 0 (e_powf.c: log2_inline, exp2_inline and the zero / subnormal /
    // overflow / underflow branches that this argument range can reach).  Callers: propagation.cpp:14 (y = 0.5) and
    // post_processing.cpp:170 (y = 1 / gamma - 1).
    PTB_DEV float glibcPowfPositive(float x, float y) {
        uint32_t ix = __float_as_uint(x);
        if(ix - 0x00800000U >= 0x7f800000U - 0x00800000U) {
            if(ix >= 0x7f800000U) {
                return libmFallbackPow(x, y); // negative, inf or NaN: outside the supported argument range
            }
            if(ix == 0U) {
                // x == +0: x^y = 0 for y > 0, +inf for y < 0 (1 / (x * x))
                return (__float_as_uint(y) & 0x80000000U) != 0U ? CUDART_INF_F : 0.0F;
            }
            // subnormal: normalise
            ix = __float_as_uint(x * 0x1p23F);
            ix &= 0x7fffffffU;
            ix -= 23U << 23;
        }
        // log2_inline
        const uint32_t tmp = ix - 0x3f330000U;
        const int i = static_cast<int>((tmp >> (23 - 4)) % 16U);
        const uint32_t top = tmp & 0xff800000U;
        const uint32_t iz = ix - top;
        const int k = static_cast<int>(top) >> 23;
        const double invc = kPowfLog2Tab[2 * i];
        const double logc = kPowfLog2Tab[2 * i + 1];
        const double z = static_cast<double>(__uint_as_float(iz));
        const double r = fma(z, invc, -1.0);
        const double y0 = logc + static_cast<double>(k);
        const double r2 = r * r;
        double lg = fma(kPowfLog2Poly[0], r, kPowfLog2Poly[1]);
        const double p = fma(kPowfLog2Poly[2], r, kPowfLog2Poly[3]);
        const double r4 = r2 * r2;
        double q = fma(kPowfLog2Poly[4], r, y0);
        q = fma(p, r2, q);
        lg = fma(lg, r4, q);
        const double ylogx = static_cast<double>(y) * lg;
        // |y * log2(x)| >= 126: overflow / underflow (e_powf.c)
        if(((static_cast<unsigned long long>(__double_as_longlong(ylogx)) >> 47) & 0xffffULL) >=
           (static_cast<unsigned long long>(__double_as_longlong(126.0)) >> 47)) {
            if(ylogx > 0x1.fffffffd1d571p+6) {
                return CUDART_INF_F; // __math_oflowf
            }
            if(ylogx <= -150.0) {
                return 0.0F; // __math_uflowf
            }
        }
        // exp2_inline
        double kd = ylogx + kExp2fShiftScaled;
        const unsigned long long ki = static_cast<unsigned long long>(__double_as_longlong(kd));
        kd -= kExp2fShiftScaled;
        const double rr = ylogx - kd;
        unsigned long long t = kExp2fTab[ki % 32ULL];
        t += ki << (52 - 5);
        const double s = __longlong_as_double(static_cast<long long>(t));
        const double zz = fma(kExp2fPoly0, rr, kExp2fPoly1);
        const double rr2 = rr * rr;
        double out = fma(kExp2fPoly2, rr, 1.0);
        out = fma(zz, rr2, out);
        out = out * s;
        return static_cast<float>(out);
    }

    PTB_DEV float glibcPowfHalf(float x) {
        return glibcPowfPositive(x, 0.5F);
    }

    // ---- acosf (fdlibm, fp32, no contraction)

    PTB_DEV float glibcAcosf(float x) {
        const float one = 1.0000000000e+00F, pi = 3.1415925026e+00F, pio2_hi = 1.5707962513e+00F, pio2_lo = 7.5497894159e-08F;
        const float pS0 = 1.6666667163e-01F, pS1 = -3.2556581497e-01F, pS2 = 2.0121252537e-01F, pS3 = -4.0055535734e-02F, pS4 = 7.9153501429e-04F,
                    pS5 = 3.4793309169e-05F;
        const float qS1 = -2.4033949375e+00F, qS2 = 2.0209457874e+00F, qS3 = -6.8828397989e-01F, qS4 = 7.7038154006e-02F;
        const int hx = __float_as_int(x);
        const int ix = hx & 0x7fffffff;
        if(ix == 0x3f800000) {
            return hx > 0 ? 0.0F : pi + 2.0F * pio2_lo;
        }
        if(ix > 0x3f800000) {
            return (x - x) / (x - x);
        }
        if(ix < 0x3f000000) { // |x| < 0.5
            if(ix <= 0x32800000) {
                return pio2_hi + pio2_lo;
            }
            const float z = x * x;
            const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
            const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
            const float r = p / q;
            return pio2_hi - (x - (pio2_lo - x * r));
        }
        if(hx < 0) { // x < -0.5
            const float z = (one + x) * 0.5F;
            const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
            const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
            const float s = sqrtf(z);
            const float r = p / q;
            const float w = r * s - pio2_lo;
            return pi - 2.0F * (s + w);
        }
        const float z = (one - x) * 0.5F; // x > 0.5
        const float s = sqrtf(z);
        const float df = __int_as_float(__float_as_int(s) & 0xfffff000);
        const float c = (z - df * df) / (s + df);
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        const float w = r * s + c;
        return 2.0F * (df + w);
    }

    // ---- what the path-vertex code calls
    //
    // Exact build (validation, and any TU compiled without PTB_FAST_MATH): the glibc restatements above.
    // Production build (ptb_fast.cu, -use_fast_math, FMA contraction on; CounterRng kernels only): the SFU approximations
    // (MUFU.SIN/COS after a range-reduction multiply, absolute error < 2^-21 on [0, 2 pi]), sqrt for pow(x, 1/2) and CUDA's
    // acosf.  The production generator's paths are not comparable sample by sample with the reference anyway (different
    // random numbers); what the north star asks of them is an image within the Monte Carlo noise bound, which
    // tests/test_parity_gpu.py::test_bench_scene_image_rmse_within_noise checks with these kernels.
#if defined(PTB_FAST_MATH)
    PTB_DEV float pathSinf(float x) {
        return __sinf(x);
    }
    PTB_DEV float pathCosf(float x) {
        return __cosf(x);
    }
    PTB_DEV float pathAcosf(float x) {
        return acosf(x);
    }
    PTB_DEV float pathPowfHalf(float x) {
        return sqrtf(x);
    }
#else
    PTB_DEV float pathSinf(float x) {
        return glibcSinf(x);
    }
    PTB_DEV float pathCosf(float x) {
        return glibcCosf(x);
    }
    PTB_DEV float pathAcosf(float x) {
        return glibcAcosf(x);
    }
    PTB_DEV float pathPowfHalf(float x) {
        return glibcPowfHalf(x);
    }
#endif

}

#endif
