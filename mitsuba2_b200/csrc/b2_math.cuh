// b2_math.cuh — fp32 math of the wavefront kernels with a FIXED operation order.
//
// The whole library is compiled with -fmad=false: the compiler never contracts a*b+c, every
// fused multiply-add is written explicitly (__fmaf_rn) at exactly the places the reference's
// Enoki expressions use fmadd/fmsub/fnmadd; division and sqrt are the IEEE-rounded ones
// (nvcc defaults -prec-div=true -prec-sqrt=true). The transcendentals are polynomial
// evaluations built from the same primitives (no libdevice sinf/cosf/expf/logf, whose results
// are not specified bit for bit), so a CPU implementation of the same formulas reproduces the
// kernels' results exactly.
//
// Operation order follows the scalar_rgb expansion of the reference's Enoki expressions:
//   dot    = (ax*bx + ay*by) + az*bz        ext/enoki/include/enoki/array_sse42.h:2162 (_mm_dp_ps)
//   cross  = fmsub(a.yzx, b.zxy, a.zxy*b.yzx)   ext/enoki/include/enoki/array_router.h:684
//   normalize(v) = v * rsqrt(dot(v,v))      array_router.h:665-667
//   Array / scalar = Array * rcp(scalar)    array_router.h:271-284
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define B2_D __device__ __forceinline__

#define B2_PI          3.14159265358979323846f
#define B2_INV_PI      0.31830988618379067154f
#define B2_INV_SQRT_PI 0.56418958354775628695f
#define B2_EPS         5.9604644775390625e-08f            /* math::Epsilon<float> = 2^-24 */
#define B2_RAY_EPS     (1500.0f * 5.9604644775390625e-08f)  /* include/mitsuba/core/math.h:36 */
#define B2_SHADOW_EPS  (B2_RAY_EPS * 10.0f)                /* math.h:38 */
#define B2_ONE_MINUS_EPS 0.999999940395355225f              /* math::OneMinusEpsilon<float>, math.h:31-33 */
#define B2_INF         __int_as_float(0x7f800000)

B2_D float b2_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// rcp.rn.f32 is the correctly rounded reciprocal: bit-identical to the IEEE division 1.0f / x that the
// CPU twin evaluates (correct rounding is unique), at roughly half the instruction count of div.rn.f32
B2_D float b2_rcp(float x) { return __frcp_rn(x); }
B2_D float b2_rsqrt(float x) { return __frcp_rn(__fsqrt_rn(x)); }
B2_D float b2_sqr(float x) { return x * x; }
B2_D float b2_safe_sqrt(float x) { return sqrtf(x > 0.0f ? x : 0.0f); }
B2_D float b2_min(float a, float b) { return a < b ? a : b; }
B2_D float b2_max(float a, float b) { return a > b ? a : b; }
B2_D float b2_sign(float x) { return __uint_as_float(0x3f800000u | (__float_as_uint(x) & 0x80000000u)); }
B2_D float b2_mulsign(float a, float b) { return __uint_as_float(__float_as_uint(a) ^ (__float_as_uint(b) & 0x80000000u)); }
B2_D float b2_mulsign_neg(float a, float b) { return __uint_as_float(__float_as_uint(a) ^ (~__float_as_uint(b) & 0x80000000u)); }

B2_D float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
B2_D float3 v3(float4 a) { return make_float3(a.x, a.y, a.z); }
B2_D float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2_D float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2_D float3 operator*(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
B2_D float3 operator*(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
B2_D float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
B2_D float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
B2_D float3 cross(float3 a, float3 b) {
    return v3(b2_fma(a.y, b.z, -(a.z * b.y)), b2_fma(a.z, b.x, -(a.x * b.z)), b2_fma(a.x, b.y, -(a.y * b.x)));
}
B2_D float3 normalize(float3 a) { return a * b2_rsqrt(dot(a, a)); }
B2_D float hmax_abs(float3 a) { return b2_max(b2_max(fabsf(a.x), fabsf(a.y)), fabsf(a.z)); }
B2_D float hmax(float3 a) { return b2_max(b2_max(a.x, a.y), a.z); }
B2_D float3 fnmadd(float3 a, float s, float3 c) {      // c - a*s, one rounding per component
    return v3(b2_fma(-a.x, s, c.x), b2_fma(-a.y, s, c.y), b2_fma(-a.z, s, c.z));
}
B2_D bool any_nonzero(float3 a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f; }

// sincos: Cody-Waite reduction by pi/2 (3 constants) + minimax polynomials on [-pi/4, pi/4].
// <= 2 ulp for |x| < 1e4; stands in for the libm sinf/cosf the reference's scalar sincos uses.
B2_D void b2_sincos(float x, float *s_out, float *c_out) {
    float q = rintf(x * 0.63661977236758134308f);
    int n = (int) q;
    float r = b2_fma(q, -1.5707962512969970703125f, x);
    r = b2_fma(q, -7.54978995489188e-08f, r);
    r = b2_fma(q, -5.390302953e-15f, r);
    float r2 = r * r;
    float ps = b2_fma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = b2_fma(ps, r2, -1.6666654611e-1f);
    float sn = b2_fma(ps * r2, r, r);
    float pc = b2_fma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = b2_fma(pc, r2, 4.166664568298827e-2f);
    float cs = b2_fma(pc * r2, r2, b2_fma(r2, -0.5f, 1.0f));
    float s, c;
    if (n & 1) { s = cs; c = sn; } else { s = sn; c = cs; }
    if (n & 2) s = -s;
    if ((n + 1) & 2) c = -c;
    *s_out = s; *c_out = c;
}

B2_D float b2_exp(float x) {
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) return B2_INF;
    float q = rintf(x * 1.44269504088896341f);
    float r = b2_fma(q, -0.693145751953125f, x);
    r = b2_fma(q, -1.428606765330187045e-06f, r);
    float p = 1.9875691500e-4f;
    p = b2_fma(p, r, 1.3981999507e-3f);
    p = b2_fma(p, r, 8.3334519073e-3f);
    p = b2_fma(p, r, 4.1665795894e-2f);
    p = b2_fma(p, r, 1.6666665459e-1f);
    p = b2_fma(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = b2_fma(p, r2, r) + 1.0f;
    int k = (int) q;
    int k1 = k / 2, k2 = k - k1;
    e = e * __uint_as_float((uint32_t) (k1 + 127) << 23);
    e = e * __uint_as_float((uint32_t) (k2 + 127) << 23);
    return e;
}

B2_D float b2_log(float x) {
    if (x < 0.0f) return __int_as_float(0x7fc00000);
    if (x == 0.0f) return -B2_INF;
    if (!(x < B2_INF)) return x;
    int e = 0;
    uint32_t u = __float_as_uint(x);
    if (u < 0x00800000u) { x = x * 8388608.0f; u = __float_as_uint(x); e = -23; }
    e += (int) (u >> 23) - 127;
    float m = __uint_as_float((u & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356237f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    float z = f * f;
    float p = 7.0376836292e-2f;
    p = b2_fma(p, f, -1.1514610310e-1f);
    p = b2_fma(p, f, 1.1676998740e-1f);
    p = b2_fma(p, f, -1.2420140846e-1f);
    p = b2_fma(p, f, 1.4249322787e-1f);
    p = b2_fma(p, f, -1.6668057665e-1f);
    p = b2_fma(p, f, 2.0000714765e-1f);
    p = b2_fma(p, f, -2.4999993993e-1f);
    p = b2_fma(p, f, 3.3333331174e-1f);
    float y = (p * f) * z;
    float fe = (float) e;
    y = b2_fma(fe, -2.12194440e-4f, y);
    y = b2_fma(z, -0.5f, y);
    float r = f + y;
    r = b2_fma(fe, 0.693359375f, r);
    return r;
}

B2_D float b2_erf(float x) {
    float ax = fabsf(x);
    float r;
    if (ax < 1.0f) {
        float x2 = x * x;
        float p = -5.61802298761904239654541015625e-4f;
        p = b2_fma(p, x2, 4.91381669417023658752441406250e-3f);
        p = b2_fma(p, x2, -2.67075151205062866210937500000e-2f);
        p = b2_fma(p, x2, 1.12799220979213714599609375000e-1f);
        p = b2_fma(p, x2, -3.76123636960983276367187500000e-1f);
        p = b2_fma(p, x2, 1.12837922573089599609375000000e+0f);
        r = x * p;
    } else {
        float t = 1.0f / b2_fma(0.3275911f, ax, 1.0f);
        float y = 1.061405429f;
        y = b2_fma(y, t, -1.453152027f);
        y = b2_fma(y, t, 1.421413741f);
        y = b2_fma(y, t, -0.284496736f);
        y = b2_fma(y, t, 0.254829592f);
        r = 1.0f - (y * t) * b2_exp(-(ax * ax));
        r = b2_mulsign(r, x);
    }
    return r;
}

B2_D float b2_erfinv(float x) {
    float w = -b2_log((1.0f - x) * (1.0f + x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = b2_fma(p, w, 3.43273939e-07f);
        p = b2_fma(p, w, -3.5233877e-06f);
        p = b2_fma(p, w, -4.39150654e-06f);
        p = b2_fma(p, w, 0.00021858087f);
        p = b2_fma(p, w, -0.00125372503f);
        p = b2_fma(p, w, -0.00417768164f);
        p = b2_fma(p, w, 0.246640727f);
        p = b2_fma(p, w, 1.50140941f);
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = b2_fma(p, w, 0.000100950558f);
        p = b2_fma(p, w, 0.00134934322f);
        p = b2_fma(p, w, -0.00367342844f);
        p = b2_fma(p, w, 0.00573950773f);
        p = b2_fma(p, w, -0.0076224613f);
        p = b2_fma(p, w, 0.00943887047f);
        p = b2_fma(p, w, 1.00167406f);
        p = b2_fma(p, w, 2.83297682f);
    }
    return p * x;
}


// ---- acos / atan2 (cephes-style polynomials; used by the environment map) ----------------------
B2_D float b2_asin_poly(float x) {       // |x| <= 0.5
    float z = x * x;
    float p = 4.2163199048e-2f;
    p = b2_fma(p, z, 2.4181311049e-2f);
    p = b2_fma(p, z, 4.5470025998e-2f);
    p = b2_fma(p, z, 7.4953002686e-2f);
    p = b2_fma(p, z, 1.6666752422e-1f);
    return b2_fma(p * z, x, x);
}
B2_D float b2_safe_acos(float x) {
    x = b2_min(b2_max(x, -1.f), 1.f);
    float ax = fabsf(x);
    if (ax <= 0.5f) return 1.57079632679489661923f - b2_asin_poly(x);
    float z = 0.5f * (1.f - ax);
    float s = sqrtf(z);
    float r = 2.f * b2_asin_poly(s);
    return x > 0.f ? r : 3.14159265358979323846f - r;
}
B2_D float b2_asin(float x) {            // |x| <= 1
    float ax = fabsf(x);
    if (ax <= 0.5f) return b2_asin_poly(x);
    float r = 1.57079632679489661923f - 2.f * b2_asin_poly(sqrtf(0.5f * (1.f - ax)));
    return b2_mulsign(r, x);
}
B2_D float b2_atan2(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = b2_max(ax, ay), mn = b2_min(ax, ay);
    float a = mx == 0.f ? 0.f : mn / mx;
    float base = 0.f;
    if (a > 0.4142135623730950f) { a = (a - 1.f) / (a + 1.f); base = 0.78539816339744830962f; }
    float z = a * a;
    float p = 8.05374449538e-2f;
    p = b2_fma(p, z, -1.38776856032e-1f);
    p = b2_fma(p, z, 1.99777106478e-1f);
    p = b2_fma(p, z, -3.33329491539e-1f);
    float r = base + b2_fma(p * z, a, a);
    if (ay > ax) r = 1.57079632679489661923f - r;
    if (x < 0.f) r = 3.14159265358979323846f - r;
    return b2_mulsign(r, y);
}

// ---- PCG32 (ext/enoki/include/enoki/random.h:62-119) ----------------------------------------
#define B2_PCG32_DEFAULT_STREAM 0xda3e39cb94b95bdbULL
#define B2_PCG32_MULT           0x5851f42d4c957f2dULL
#define B2_PCG32_INC            ((B2_PCG32_DEFAULT_STREAM << 1u) | 1u)

B2_D uint32_t pcg32_next_uint32(uint64_t &state) {
    uint64_t oldstate = state;
    state = oldstate * B2_PCG32_MULT + B2_PCG32_INC;
    uint32_t xorshifted = (uint32_t) (((oldstate >> 18u) ^ oldstate) >> 27u);
    uint32_t rot = (uint32_t) (oldstate >> 59u);
    return __funnelshift_r(xorshifted, xorshifted, rot);
}
B2_D uint64_t pcg32_seed(uint64_t initstate) {          // seed(initstate, PCG32_DEFAULT_STREAM)
    uint64_t state = 0u;
    pcg32_next_uint32(state);
    state += initstate;
    pcg32_next_uint32(state);
    return state;
}
B2_D float pcg32_next_float(uint64_t &state) {
    return __uint_as_float((pcg32_next_uint32(state) >> 9) | 0x3f800000u) - 1.0f;
}
