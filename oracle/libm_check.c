/*
 * libm_check.c — CPU check of the glibc libm restatement used by the device code.  TEST INFRASTRUCTURE ONLY.
 *
 * cpupathtrace_b200/csrc/glibc_libm.cuh restates glibc 2.39's sinf, cosf, powf(x, 0.5) and acosf operation for
 * operation so that the CUDA path rounds like the reference's <cmath> calls.  This file holds the same restatement in
 * plain C — same constants (the tables come from the shared header glibc_libm_tables.h), same operation order — and
 * compares it with the system libm on pseudo-random arguments drawn the way the path draws them:
 *   angle = 2 pi u   (sinf, cosf: camera.cpp:13, propagation.cpp:16, object.cpp:106)
 *   u                (powf(u, 0.5): propagation.cpp:14)
 *   1 - 2 u          (acosf: object.cpp:107)
 * with u = float(x) / 2^32 from a 32-bit generator.  PTO_FMA selects fused evaluation of the fp64 polynomials (what
 * glibc's *_fma ifunc variants do); both builds must agree with libm.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define PTB_TABLE static const
#include "../cpupathtrace_b200/csrc/glibc_libm_tables.h"

#ifdef PTO_FMA
#define FMA(a, b, c) fma((a), (b), (c))
#else
#define FMA(a, b, c) ((a) * (b) + (c))
#endif

static uint32_t asuint(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static float asfloat(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static uint64_t asuint64(double f) {
    uint64_t u;
    memcpy(&u, &f, 8);
    return u;
}
static double asdouble(uint64_t u) {
    double f;
    memcpy(&f, &u, 8);
    return f;
}
static uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

typedef struct {
    double sign[4];
    double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
} sincos_t;

static const sincos_t TBL[2] = {
  {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10,
   0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
  {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10,
   -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};

static float sincos_poly(double x, double x2, const sincos_t *p, int n) {
    if((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = FMA(x2, p->s3, p->s2);
        double x7 = x3 * x2;
        double s = FMA(x3, p->s1, x);
        return (float)FMA(x7, s1, s);
    }
    double x4 = x2 * x2;
    double c2 = FMA(x2, p->c4, p->c3);
    double c1 = FMA(x2, p->c1, p->c0);
    double x6 = x4 * x2;
    double c = FMA(x4, p->c2, c1);
    return (float)FMA(x6, c2, c);
}

static float restated_sincos(float y, int which) {
    double x = y;
    const sincos_t *p = &TBL[0];
    if(abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if(abstop12(y) < abstop12(0x1p-12f)) return which == 0 ? y : 1.0f;
        return sincos_poly(x, x * x, p, which);
    }
    double r = x * p->hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = FMA(-(double)n, p->hpi, x);
    double s = p->sign[n & 3];
    if(n & 2) p = &TBL[1];
    return sincos_poly(x * s, x * x, p, which == 0 ? n : (n ^ 1));
}

static float restated_powf(float x, float y_exp) {
    uint32_t ix = asuint(x);
    if(ix == 0) return (asuint(y_exp) & 0x80000000u) ? INFINITY : 0.0f;
    if(ix < 0x00800000) {
        ix = asuint(x * 0x1p23f);
        ix &= 0x7fffffff;
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000;
    int i = (tmp >> (23 - 4)) % 16;
    uint32_t top = tmp & 0xff800000;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = kPowfLog2Tab[2 * i], logc = kPowfLog2Tab[2 * i + 1];
    double z = (double)asfloat(iz);
    double r = FMA(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double y = FMA(kPowfLog2Poly[0], r, kPowfLog2Poly[1]);
    double p = FMA(kPowfLog2Poly[2], r, kPowfLog2Poly[3]);
    double r4 = r2 * r2;
    double q = FMA(kPowfLog2Poly[4], r, y0);
    q = FMA(p, r2, q);
    y = FMA(y, r4, q);
    double xd = (double)y_exp * y;
    if(((asuint64(xd) >> 47) & 0xffff) >= (asuint64(126.0) >> 47)) {
        if(xd > 0x1.fffffffd1d571p+6) return INFINITY;
        if(xd <= -150.0) return 0.0f;
    }
    double kd = xd + kExp2fShiftScaled;
    uint64_t ki = asuint64(kd);
    kd -= kExp2fShiftScaled;
    double rr = xd - kd;
    uint64_t t = kExp2fTab[ki % 32];
    t += ki << (52 - 5);
    double s = asdouble(t);
    double zz = FMA(kExp2fPoly0, rr, kExp2fPoly1);
    double rr2 = rr * rr;
    double out = FMA(kExp2fPoly2, rr, 1.0);
    out = FMA(zz, rr2, out);
    return (float)(out * s);
}

/* fp32 throughout, compiled with -ffp-contract=off */
static float restated_acosf(float x) {
    const float one = 1.0000000000e+00f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f,
                pS5 = 3.4793309169e-05f;
    const float qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    int32_t hx = (int32_t)asuint(x), ix = hx & 0x7fffffff;
    if(ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if(ix > 0x3f800000) return (x - x) / (x - x);
    if(ix < 0x3f000000) {
        if(ix <= 0x32800000) return pio2_hi + pio2_lo;
        float z = x * x;
        float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        float r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if(hx < 0) {
        float z = (one + x) * 0.5f;
        float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        float s = sqrtf(z);
        float r = p / q;
        float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    float z = (one - x) * 0.5f;
    float s = sqrtf(z);
    float df = asfloat(asuint(s) & 0xfffff000);
    float c = (z - df * df) / (s + df);
    float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    float r = p / q;
    float w = r * s + c;
    return 2.0f * (df + w);
}

/* mismatches[0..3] = sinf, cosf, powf(.,0.5), acosf disagreements with the system libm over n arguments */
void pto_libm_check(uint64_t n, uint64_t seed, uint64_t mismatches[4]) {
    uint64_t s = seed ? seed : 88172645463325252ULL;
    mismatches[0] = mismatches[1] = mismatches[2] = mismatches[3] = 0;
    for(uint64_t i = 0; i < n; i++) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        float u = (float)(uint32_t)(s >> 32) / 4294967296.0f;
        if(u >= 1.0f) u = 0.99999994f;
        volatile float vu = u;
        volatile float angle = 6.28318548202514648438f * vu;
        volatile float arg = 1.0f - 2.0f * vu;
        if(sinf(angle) != restated_sincos(angle, 0)) mismatches[0]++;
        if(cosf(angle) != restated_sincos(angle, 1)) mismatches[1]++;
        if(powf(vu, 0.5f) != restated_powf(vu, 0.5f)) mismatches[2]++;
        /* post_processing.cpp:170: pow(brightness, 1 / gamma - 1) over a wide range of brightness and gamma */
        {
            volatile float bx = ldexpf(vu + 1e-9f, (int)((s >> 8) % 61) - 40);
            volatile float by = ((s >> 20) & 1) ? (1.0f / 1.8f - 1.0f) : (float)(int)((s >> 24) % 4001 - 2000) / 500.0f;
            if(by != 0.0f && powf(bx, by) != restated_powf(bx, by)) mismatches[2]++;
        }
        float a = acosf(arg), b = restated_acosf(arg);
        if(a != b) mismatches[3]++;
        /* the sphere sampler also takes sin/cos of phi = acos(.) in [0, pi] */
        volatile float phi = a;
        if(sinf(phi) != restated_sincos(phi, 0)) mismatches[0]++;
        if(cosf(phi) != restated_sincos(phi, 1)) mismatches[1]++;
    }
}
