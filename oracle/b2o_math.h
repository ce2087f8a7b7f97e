/*
 * b2o_math.h — TEST INFRASTRUCTURE (part of the CPU oracle; never linked into the product).
 *
 * Scalar fp32 math with a FIXED operation order so that the oracle is reproducible bit for
 * bit on any IEEE-754 host (compiled with -ffp-contract=off): only + - * / sqrt and
 * explicitly written fmaf() are used; the few transcendentals are polynomial restatements
 * built from the same primitives.
 *
 * The operation order follows what the reference's Enoki expressions expand to on the
 * scalar_rgb variant:
 *   dot(a,b)      = (ax*bx + ay*by) + az*bz     (_mm_dp_ps, ext/enoki/include/enoki/array_sse42.h:2162)
 *   cross(a,b)    = fmsub(a.yzx, b.zxy, a.zxy*b.yzx)     (ext/enoki/include/enoki/array_router.h:684)
 *   normalize(v)  = v * rsqrt(dot(v,v))                  (array_router.h:665-667)
 *   Array / scalar = Array * rcp(scalar)                 (array_router.h:271-284)
 * The reference evaluates rcp/rsqrt with vrcp14ss/vrsqrt14ss + 1 Newton step
 * (ext/enoki/include/enoki/array_fallbacks.h:104-140), which is accurate to ~1 ulp but not
 * correctly rounded; the oracle uses the correctly rounded 1/x and 1/sqrt(x). This is one of
 * the documented sources of the oracle-vs-reference noise floor (DESIGN.md "Float parity").
 */
#ifndef B2O_MATH_H
#define B2O_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

#define B2O_PI        3.14159265358979323846f
#define B2O_INV_PI    0.31830988618379067154f
#define B2O_INV_SQRT_PI 0.56418958354775628695f
/* include/mitsuba/core/math.h:36-38 */
#define B2O_RAY_EPS    (1500.0f * 5.9604644775390625e-08f)  /* 1500 * 2^-24 (Epsilon<float> = 2^-24) */
#define B2O_SHADOW_EPS (B2O_RAY_EPS * 10.0f)
#define B2O_ONE_MINUS_EPS 0.999999940395355225f                   /* math::OneMinusEpsilon<float>, math.h:31-33 */

static inline float b2o_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
#if defined(B2O_REF_RCP) && defined(__AVX512F__) && defined(__FMA__)
/* ANALYSIS build only (oracle/Makefile: libb2oracle_refrcp.so; scripts/refrcp_analysis.py; never what the CUDA path is compared
 * with): rcp / rsqrt by the very instruction sequences the reference executes on an AVX-512 host — vrcp14ss / vrsqrt14ss, one
 * Newton-Raphson step, vfixupimmss (ext/enoki/include/enoki/array_fallbacks.h:32-58, 112-140) — to MEASURE how much of the
 * oracle-vs-reference noise floor (which closest-hit wins between coplanar triangles) comes from the last-ulp
 * difference between these and the correctly rounded 1/x, 1/sqrt(x) that the oracle and the CUDA kernels use. */
#include <immintrin.h>
static inline float b2o_rcp(float x) {
    __m128 v = _mm_set_ss(x), r = _mm_rcp14_ss(v, v);
    __m128 t0 = _mm_add_ss(r, r), t1 = _mm_mul_ss(r, v);
    r = _mm_fnmadd_ss(r, t1, t0);
    r = _mm_fixupimm_ss(r, v, _mm_set1_epi32(0x0087A622), 0);
    return _mm_cvtss_f32(r);
}
static inline float b2o_rsqrt(float x) {
    __m128 v = _mm_set_ss(x), r = _mm_rsqrt14_ss(v, v);
    const __m128 c0 = _mm_set_ss(0.5f), c1 = _mm_set_ss(3.0f);
    __m128 t0 = _mm_mul_ss(r, c0), t1 = _mm_mul_ss(r, v);
    r = _mm_mul_ss(_mm_fnmadd_ss(t1, r, c1), t0);
    r = _mm_fixupimm_ss(r, v, _mm_set1_epi32(0x0383A622), 0);
    return _mm_cvtss_f32(r);
}
#else
static inline float b2o_rcp(float x) { return 1.0f / x; }
static inline float b2o_rsqrt(float x) { return 1.0f / sqrtf(x); }
#endif
static inline float b2o_sqr(float x) { return x * x; }
static inline float b2o_safe_sqrt(float x) { return sqrtf(x > 0.0f ? x : 0.0f); } /* sqrt(max(x,0)) */
static inline float b2o_min(float a, float b) { return a < b ? a : b; }   /* minss semantics */
static inline float b2o_max(float a, float b) { return a > b ? a : b; }

static inline uint32_t b2o_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float b2o_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
/* enoki::sign = copysign(1, x); mulsign(a, b) = a ^ signbit(b); mulsign_neg(a,b) = a ^ ~signbit(b) */
static inline float b2o_sign(float x) { return b2o_u2f(0x3f800000u | (b2o_f2u(x) & 0x80000000u)); }
static inline float b2o_mulsign(float a, float b) { return b2o_u2f(b2o_f2u(a) ^ (b2o_f2u(b) & 0x80000000u)); }
static inline float b2o_mulsign_neg(float a, float b) { return b2o_u2f(b2o_f2u(a) ^ (~b2o_f2u(b) & 0x80000000u)); }

static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(b2o_fma(a.y, b.z, -(a.z * b.y)),
                   b2o_fma(a.z, b.x, -(a.x * b.z)),
                   b2o_fma(a.x, b.y, -(a.y * b.x)));
}
static inline v3 v3_normalize(v3 a) { return v3_scale(a, b2o_rsqrt(v3_dot(a, a))); }
static inline float v3_hmax_abs(v3 a) { return b2o_max(b2o_max(fabsf(a.x), fabsf(a.y)), fabsf(a.z)); }
static inline float v3_hmax(v3 a) { return b2o_max(b2o_max(a.x, a.y), a.z); }
/* fnmadd(a, s, c) = c - a*s with one rounding, per component */
static inline v3 v3_fnmadd(v3 a, float s, v3 c) {
    return v3_make(b2o_fma(-a.x, s, c.x), b2o_fma(-a.y, s, c.y), b2o_fma(-a.z, s, c.z));
}

/* ---- sincos: Cody-Waite reduction by pi/2 + minimax polynomials on [-pi/4, pi/4] ----------
 * Stands in for glibc sinf/cosf (which the reference reaches through enoki::sincos ->
 * std::sin/std::cos for scalars, ext/enoki/include/enoki/array_math.h). Accurate to <= 2 ulp
 * for |x| < 1e4; uses only fmaf, + and x, so the CUDA twin (csrc/b2_math.cuh) is bit-identical. */
static inline void b2o_sincos_poly(float x, float *s_out, float *c_out) {
    float q = rintf(x * 0.63661977236758134308f);          /* x * 2/pi, round-half-even */
    int n = (int) q;
    float r = b2o_fma(q, -1.5707962512969970703125f, x);    /* pi/2 split in 3 parts */
    r = b2o_fma(q, -7.54978995489188e-08f, r);
    r = b2o_fma(q, -5.390302953e-15f, r);
    float r2 = r * r;
    /* sin(r) ~ r + r^3 * P(r^2) */
    float ps = b2o_fma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = b2o_fma(ps, r2, -1.6666654611e-1f);
    float sn = b2o_fma(ps * r2, r, r);
    /* cos(r) ~ 1 - r^2/2 + r^4 * Q(r^2) */
    float pc = b2o_fma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = b2o_fma(pc, r2, 4.166664568298827e-2f);
    float cs = b2o_fma(pc * r2, r2, b2o_fma(r2, -0.5f, 1.0f));
    float s, c;
    if (n & 1) { s = cs; c = sn; } else { s = sn; c = cs; }
    if (n & 2) s = -s;
    if ((n + 1) & 2) c = -c;
    *s_out = s; *c_out = c;
}

/* ---- expf / logf / erf / erfinv twins (polynomial; used by Beckmann + gaussian filter) ---- */
static inline float b2o_exp_poly(float x) {
    /* exp(x) = 2^k * exp(r), r in [-ln2/2, ln2/2] */
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) return INFINITY;
    float q = rintf(x * 1.44269504088896341f);
    float r = b2o_fma(q, -0.693145751953125f, x);
    r = b2o_fma(q, -1.428606765330187045e-06f, r);
    float p = 1.9875691500e-4f;
    p = b2o_fma(p, r, 1.3981999507e-3f);
    p = b2o_fma(p, r, 8.3334519073e-3f);
    p = b2o_fma(p, r, 4.1665795894e-2f);
    p = b2o_fma(p, r, 1.6666665459e-1f);
    p = b2o_fma(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = b2o_fma(p, r2, r) + 1.0f;
    int k = (int) q;
    /* scale by 2^k in two steps to stay in range */
    int k1 = k / 2, k2 = k - k1;
    e = e * b2o_u2f((uint32_t) (k1 + 127) << 23);
    e = e * b2o_u2f((uint32_t) (k2 + 127) << 23);
    return e;
}

static inline float b2o_log_poly(float x) {
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return -INFINITY;
    if (!(x < INFINITY)) return x;
    int e = 0;
    uint32_t u = b2o_f2u(x);
    if (u < 0x00800000u) { x = x * 8388608.0f; u = b2o_f2u(x); e = -23; }
    e += (int) (u >> 23) - 127;
    float m = b2o_u2f((u & 0x007fffffu) | 0x3f800000u);     /* [1,2) */
    if (m > 1.41421356237f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    float z = f * f;
    float p = 7.0376836292e-2f;
    p = b2o_fma(p, f, -1.1514610310e-1f);
    p = b2o_fma(p, f, 1.1676998740e-1f);
    p = b2o_fma(p, f, -1.2420140846e-1f);
    p = b2o_fma(p, f, 1.4249322787e-1f);
    p = b2o_fma(p, f, -1.6668057665e-1f);
    p = b2o_fma(p, f, 2.0000714765e-1f);
    p = b2o_fma(p, f, -2.4999993993e-1f);
    p = b2o_fma(p, f, 3.3333331174e-1f);
    float y = (p * f) * z;
    float fe = (float) e;
    y = b2o_fma(fe, -2.12194440e-4f, y);
    y = b2o_fma(z, -0.5f, y);
    float r = f + y;
    r = b2o_fma(fe, 0.693359375f, r);
    return r;
}

/* ANALYSIS build only (libb2oracle_reflibm.so, scripts/refrcp_analysis.py): glibc's sinf / cosf / expf / logf / atan2f — what the
 * reference reaches for scalars — instead of the polynomial twins the oracle shares with the CUDA kernels. */
#ifdef B2O_REF_LIBM
static inline void b2o_sincos(float x, float *s_out, float *c_out) { *s_out = sinf(x); *c_out = cosf(x); }
static inline float b2o_exp(float x) { return expf(x); }
static inline float b2o_log(float x) { return logf(x); }
#else
static inline void b2o_sincos(float x, float *s_out, float *c_out) { b2o_sincos_poly(x, s_out, c_out); }
static inline float b2o_exp(float x) { return b2o_exp_poly(x); }
static inline float b2o_log(float x) { return b2o_log_poly(x); }
#endif

/* erf: Abramowitz-Stegun-free, uses the same structure as Enoki's erf
 * (ext/enoki/include/enoki/array_math.h erf(): polynomial for |x|<1, exp-based tail) */
static inline float b2o_erf(float x) {
    float ax = fabsf(x);
    float r;
    if (ax < 1.0f) {
        float x2 = x * x;
        float p = -5.61802298761904239654541015625e-4f;
        p = b2o_fma(p, x2, 4.91381669417023658752441406250e-3f);
        p = b2o_fma(p, x2, -2.67075151205062866210937500000e-2f);
        p = b2o_fma(p, x2, 1.12799220979213714599609375000e-1f);
        p = b2o_fma(p, x2, -3.76123636960983276367187500000e-1f);
        p = b2o_fma(p, x2, 1.12837922573089599609375000000e+0f);
        r = x * p;
    } else {
        /* Abramowitz-Stegun 7.1.26 complementary form for the tail */
        float t = 1.0f / b2o_fma(0.3275911f, ax, 1.0f);
        float y = 1.061405429f;
        y = b2o_fma(y, t, -1.453152027f);
        y = b2o_fma(y, t, 1.421413741f);
        y = b2o_fma(y, t, -0.284496736f);
        y = b2o_fma(y, t, 0.254829592f);
        r = 1.0f - (y * t) * b2o_exp(-(ax * ax));
        r = b2o_mulsign(r, x);
    }
    return r;
}

/* erfinv: Giles' single-precision approximation ("Approximating the erfinv function", 2010) */
static inline float b2o_erfinv(float x) {
    float w = -b2o_log((1.0f - x) * (1.0f + x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = b2o_fma(p, w, 3.43273939e-07f);
        p = b2o_fma(p, w, -3.5233877e-06f);
        p = b2o_fma(p, w, -4.39150654e-06f);
        p = b2o_fma(p, w, 0.00021858087f);
        p = b2o_fma(p, w, -0.00125372503f);
        p = b2o_fma(p, w, -0.00417768164f);
        p = b2o_fma(p, w, 0.246640727f);
        p = b2o_fma(p, w, 1.50140941f);
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = b2o_fma(p, w, 0.000100950558f);
        p = b2o_fma(p, w, 0.00134934322f);
        p = b2o_fma(p, w, -0.00367342844f);
        p = b2o_fma(p, w, 0.00573950773f);
        p = b2o_fma(p, w, -0.0076224613f);
        p = b2o_fma(p, w, 0.00943887047f);
        p = b2o_fma(p, w, 1.00167406f);
        p = b2o_fma(p, w, 2.83297682f);
    }
    return p * x;
}

/* ---- acos / atan2 twins (cephes-style polynomials; used by the environment map) -------------- */
static inline float b2o_asin_poly(float x) {       /* |x| <= 0.5 */
    float z = x * x;
    float p = 4.2163199048e-2f;
    p = b2o_fma(p, z, 2.4181311049e-2f);
    p = b2o_fma(p, z, 4.5470025998e-2f);
    p = b2o_fma(p, z, 7.4953002686e-2f);
    p = b2o_fma(p, z, 1.6666752422e-1f);
    return b2o_fma(p * z, x, x);
}
static inline float b2o_safe_acos(float x) {       /* acos(clamp(x, -1, 1)) */
    x = b2o_min(b2o_max(x, -1.f), 1.f);
    float ax = fabsf(x);
    if (ax <= 0.5f) return 1.57079632679489661923f - b2o_asin_poly(x);
    float z = 0.5f * (1.f - ax);
    float s = sqrtf(z);
    float r = 2.f * b2o_asin_poly(s);
    return x > 0.f ? r : 3.14159265358979323846f - r;
}
static inline float b2o_asin(float x) {            /* |x| <= 1 */
    float ax = fabsf(x);
    if (ax <= 0.5f) return b2o_asin_poly(x);
    float r = 1.57079632679489661923f - 2.f * b2o_asin_poly(sqrtf(0.5f * (1.f - ax)));
    return b2o_mulsign(r, x);
}
static inline float b2o_atan2_poly(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = b2o_max(ax, ay), mn = b2o_min(ax, ay);
    float a = mx == 0.f ? 0.f : mn / mx;           /* in [0, 1] */
    float base = 0.f;
    if (a > 0.4142135623730950f) { a = (a - 1.f) / (a + 1.f); base = 0.78539816339744830962f; }
    float z = a * a;
    float p = 8.05374449538e-2f;
    p = b2o_fma(p, z, -1.38776856032e-1f);
    p = b2o_fma(p, z, 1.99777106478e-1f);
    p = b2o_fma(p, z, -3.33329491539e-1f);
    float r = base + b2o_fma(p * z, a, a);
    if (ay > ax) r = 1.57079632679489661923f - r;
    if (x < 0.f) r = 3.14159265358979323846f - r;
    return b2o_mulsign(r, y);
}

#ifdef B2O_REF_LIBM
static inline float b2o_atan2(float y, float x) { return atan2f(y, x); }
#else
static inline float b2o_atan2(float y, float x) { return b2o_atan2_poly(y, x); }
#endif

#endif /* B2O_MATH_H */
