// refmath.h -- float transcendental functions that reproduce GNU libc 2.39's results bit for bit.
//
// pbrt-v2 calls the host libm's sinf/cosf/expf/powf/atan2f/acosf on its hot path
// (core/montecarlo.cpp:346-347,419-420, core/reflection.cpp:47,349-359, core/reflection.h:429,
// core/geometry.h:640-648, shapes/sphere.cpp:89,110, shapes/disk.cpp:76).  None of them is
// correctly rounded in glibc, and CUDA's are different again, so the only way to make hit/miss and
// Russian-roulette decisions (and with them the RNG stream) identical is to restate glibc's own
// published algorithms:
//   sinf, cosf   sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h, sincosf_data.c
//                (ARM optimized-routines; polynomial in double after a 2/pi reduction)
//   expf, powf   sysdeps/ieee754/flt-32/e_expf.c, e_powf.c, e_exp2f_data.c, e_powf_log2_data.c
//                (ARM optimized-routines; 32-entry 2^(i/32) table, 16-entry log2 table)
//   atanf, atan2f, acosf   sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c, e_acosf.c
//                (Sun fdlibm float versions: float arithmetic, rational / odd polynomials)
// The numeric tables were read out of /lib/x86_64-linux-gnu/libm.so.6 of this image (glibc 2.39)
// and the implementations are verified against that libm exhaustively / on 10^9 random arguments
// by tests/test_refmath.py (this header also compiles as plain C++ for that purpose).
// Arithmetic must not be contracted: nvcc -fmad=false / g++ -ffp-contract=off.
// Domain: the special-case branches pbrt cannot reach (NaN/inf arguments, x < 0 for powf, |x| >= 120
// for sinf/cosf) are reduced to the documented glibc results where cheap and otherwise noted.
#ifndef PBRT_B200_REFMATH_H
#define PBRT_B200_REFMATH_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RM_FN __host__ __device__ inline
#else
#define RM_FN static inline
#endif

namespace refmath {

RM_FN uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RM_FN float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
RM_FN uint64_t asuint64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
RM_FN double asdouble(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
RM_FN uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

// ------------------------------------------------------------------ sinf / cosf
// sincosf_data.c: {sign[4], hpi_inv * 2^24, hpi, c0..c4, s1..s3}; table 1 = cosine terms negated
RM_FN double sc_poly_sin(double x, double x2) {   // sinf_poly, (n & 1) == 0
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    double x3 = x * x2;
    double t = s2 + x2 * s3;
    double x7 = x3 * x2;
    double s = x + x3 * s1;
    return s + x7 * t;
}
RM_FN double sc_poly_cos(double x2, int neg) {    // sinf_poly, (n & 1) == 1; neg selects __sincosf_table[1]
    double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
           c4 = 0x1.99343027bf8c3p-16;
    if (neg) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; c4 = -c4; }
    double x4 = x2 * x2;
    double t2 = c3 + x2 * c4;
    double t1 = c0 + x2 * c1;
    double x6 = x4 * x2;
    double c = t1 + x4 * c2;
    return c + x6 * t2;
}
RM_FN double sc_reduce_fast(double x, int *np) {   // sincosf.h reduce_fast, !TOINT_INTRINSICS
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    double r = x * hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return x - n * hpi;
}
// |y| < 120 (pbrt's arguments are within [0, 2*pi]); larger arguments are not restated
RM_FN float ref_sinf(float y) {
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {   // |y| < pi/4
        double s = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return (float)sc_poly_sin(x, s);
    }
    int n;
    x = sc_reduce_fast(x, &n);
    const double sign = (n & 3) == 0 || (n & 3) == 3 ? 1.0 : -1.0;   // sign[] = {1, -1, -1, 1}
    const int neg = (n & 2) != 0;
    if ((n & 1) == 0) return (float)sc_poly_sin(x * sign, x * x);
    return (float)sc_poly_cos(x * x, neg);
}
RM_FN float ref_cosf(float y) {
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return (float)sc_poly_cos(x2, 0);            // sinf_poly(x, x2, p, 1)
    }
    int n;
    x = sc_reduce_fast(x, &n);
    const double sign = (n & 3) == 0 || (n & 3) == 3 ? 1.0 : -1.0;
    const int neg = (n & 2) != 0;
    // s_cosf.c: sinf_poly(x * s, x * x, p, n ^ 1)
    if (((n ^ 1) & 1) == 0) return (float)sc_poly_sin(x * sign, x * x);
    return (float)sc_poly_cos(x * x, neg);
}

// ------------------------------------------------------------------ expf / powf
// (plain __device__ globals, not __constant__: lanes index them divergently)
#ifdef __CUDACC__
__device__
#endif
static const uint64_t kExp2fTab[32] = {   // e_exp2f_data.c: asuint64(2^(i/32)) - (i << 47)
    0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f, 0x3fef9301d0125b51, 0x3fef72b83c7d517b,
    0x3fef54873168b9aa, 0x3fef387a6e756238, 0x3fef1e9df51fdee1, 0x3fef06fe0a31b715, 0x3feef1a7373aa9cb,
    0x3feedea64c123422, 0x3feece086061892d, 0x3feebfdad5362a27, 0x3feeb42b569d4f82, 0x3feeab07dd485429,
    0x3feea47eb03a5585, 0x3feea09e667f3bcd, 0x3fee9f75e8ec5f74, 0x3feea11473eb0187, 0x3feea589994cce13,
    0x3feeace5422aa0db, 0x3feeb737b0cdc5e5, 0x3feec49182a3f090, 0x3feed503b23e255d, 0x3feee89f995ad3ad,
    0x3feeff76f2fb5e47, 0x3fef199bdd85529c, 0x3fef3720dcef9069, 0x3fef5818dcfba487, 0x3fef7c97337b9b5f,
    0x3fefa4afa2a490da, 0x3fefd0765b6e4540};
#ifdef __CUDACC__
__device__
#endif
static const double kPowfLog2Tab[32] = {   // e_powf_log2_data.c: {invc, logc} x 16
    0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
    0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
    0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
    0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1.0000000000000p+0, 0x0.0p+0,
    0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3,
    0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
    0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2};

#if defined(__CUDA_ARCH__)
#define RM_EXP2TAB(i) __ldg(reinterpret_cast<const unsigned long long *>(kExp2fTab) + (i))
#define RM_LOG2TAB(i) __ldg(kPowfLog2Tab + (i))
#else
RM_FN uint64_t rm_exp2tab_host(int i) {
    const uint64_t t[32] = {
        0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f, 0x3fef9301d0125b51, 0x3fef72b83c7d517b,
        0x3fef54873168b9aa, 0x3fef387a6e756238, 0x3fef1e9df51fdee1, 0x3fef06fe0a31b715, 0x3feef1a7373aa9cb,
        0x3feedea64c123422, 0x3feece086061892d, 0x3feebfdad5362a27, 0x3feeb42b569d4f82, 0x3feeab07dd485429,
        0x3feea47eb03a5585, 0x3feea09e667f3bcd, 0x3fee9f75e8ec5f74, 0x3feea11473eb0187, 0x3feea589994cce13,
        0x3feeace5422aa0db, 0x3feeb737b0cdc5e5, 0x3feec49182a3f090, 0x3feed503b23e255d, 0x3feee89f995ad3ad,
        0x3feeff76f2fb5e47, 0x3fef199bdd85529c, 0x3fef3720dcef9069, 0x3fef5818dcfba487, 0x3fef7c97337b9b5f,
        0x3fefa4afa2a490da, 0x3fefd0765b6e4540};
    return t[i];
}
RM_FN double rm_log2tab_host(int i) {
    const double t[32] = {
        0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
        0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
        0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3,
        0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
        0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1.0000000000000p+0, 0x0.0p+0,
        0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3,
        0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
        0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2};
    return t[i];
}
#define RM_EXP2TAB(i) rm_exp2tab_host(i)
#define RM_LOG2TAB(i) rm_log2tab_host(i)
#endif

RM_FN float rm_uflowf() { volatile float y = 0x1p-95f; return y * y; }          // __math_uflowf(0)
RM_FN float rm_may_uflowf() { volatile float y = 0x1.4p-75f; return y * y; }    // __math_may_uflowf(0)
RM_FN float rm_oflowf() { volatile float y = 0x1p97f; return y * y; }           // __math_oflowf(0)

RM_FN float ref_expf(float x) {   // e_expf.c
    const double InvLn2N = 0x1.71547652b82fep+5;   // N/ln2, N = 32
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;   // poly_scaled
    double xd = (double)x;
    uint32_t abstop = (asuint(x) >> 20) & 0x7ff;
    if (abstop >= ((asuint(88.0f) >> 20) & 0x7ff)) {
        if (asuint(x) == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return x + x;
        if (x > 0x1.62e42ep6f) return rm_oflowf();
        if (x < -0x1.9fe368p6f) return rm_uflowf();
        if (x < -0x1.9d1d9ep6f) return rm_may_uflowf();
    }
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    uint64_t ki = asuint64(kd);
    kd -= SHIFT;
    double r = z - kd;
    uint64_t t = RM_EXP2TAB(ki % 32);
    t += ki << (52 - 5);
    double s = asdouble(t);
    z = C0 * r + C1;
    double r2 = r * r;
    double y = C2 * r + 1;
    y = z * r2 + y;
    y = y * s;
    return (float)y;
}

// x > 0 finite or x == 0; y finite and non-zero (pbrt: x in [0, 1], y in (0, 1e4])
RM_FN float ref_powf(float x, float y) {   // e_powf.c
    const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
                 A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
    uint32_t ix = asuint(x), iy = asuint(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (2 * ix == 0) {   // x == +-0: x2 = x * x; y > 0 -> x2, y < 0 -> 1 / x2
            float x2 = x * x;
            return (iy & 0x80000000u) ? 1 / x2 : x2;
        }
        if (ix < 0x00800000u) {   // subnormal x
            ix = asuint(x * 0x1p23f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    // log2_inline
    const uint32_t OFF = 0x3f330000u;
    uint32_t tmp = ix - OFF;
    int i = (tmp >> (23 - 4)) % 16;
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = RM_LOG2TAB(2 * i), logc = RM_LOG2TAB(2 * i + 1);
    double z = (double)asfloat(iz);
    double r = z * invc - 1;
    double y0 = logc + (double)k;
    double r2 = r * r;
    double yy = A0 * r + A1;
    double p = A2 * r + A3;
    double r4 = r2 * r2;
    double q = A4 * r + y0;
    q = p * r2 + q;
    double logx = yy * r4 + q;
    double ylogx = y * logx;
    if ((asuint64(ylogx) >> 47 & 0xffff) >= asuint64(126.0) >> 47) {
        if (ylogx > 0x1.fffffffd1d571p+6) return rm_oflowf();
        if (ylogx <= -150.0) return rm_uflowf();
        if (ylogx < -149.0) return rm_may_uflowf();
    }
    // exp2_inline, !TOINT_INTRINSICS: poly, shift_scaled
    const double SHIFT = 0x1.8p+47;
    const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    double kd = ylogx + SHIFT;
    uint64_t ki = asuint64(kd);
    kd -= SHIFT;
    r = ylogx - kd;
    uint64_t t = RM_EXP2TAB(ki % 32);
    t += ki << (52 - 5);
    double s = asdouble(t);
    z = C0 * r + C1;
    r2 = r * r;
    double v = C2 * r + 1;
    v = z * r2 + v;
    v = v * s;
    return (float)v;
}

// ------------------------------------------------------------------ atanf / atan2f / acosf (fdlibm, float)
RM_FN float rm_fabsf(float x) { return asfloat(asuint(x) & 0x7fffffffu); }

RM_FN float ref_atanf(float x) {   // s_atanf.c
    // atanhi[id], atanlo[id] as selects: an indexed local array is re-initialised in local memory on every call
#define RM_ATANHI(id) ((id) == 0 ? 4.6364760399e-01f : (id) == 1 ? 7.8539812565e-01f : (id) == 2 ? 9.8279368877e-01f : 1.5707962513e+00f)
#define RM_ATANLO(id) ((id) == 0 ? 5.0121582440e-09f : (id) == 1 ? 3.7748947079e-08f : (id) == 2 ? 3.4473217170e-08f : 7.5497894159e-08f)
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    float w, s1, s2, z;
    int32_t ix, hx, id;
    hx = (int32_t)asuint(x);
    ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) {   // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return RM_ATANHI(3) + RM_ATANLO(3);
        else return -RM_ATANHI(3) - RM_ATANLO(3);
    }
    if (ix < 0x3ee00000) {   // |x| < 0.4375
        if (ix < 0x31000000) return x;   // |x| < 2^-29
        id = -1;
    } else {
        x = rm_fabsf(x);
        if (ix < 0x3f980000) {   // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = ((float)2.0 * x - 1.0f) / ((float)2.0 + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - (float)1.5) / (1.0f + (float)1.5 * x); }
            else { id = 3; x = -(float)1.0 / x; }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    z = RM_ATANHI(id) - ((x * (s1 + s2) - RM_ATANLO(id)) - x);
    return (hx < 0) ? -z : z;
}

RM_FN float ref_atan2f(float y, float x) {   // e_atan2f.c
    const float tiny = 1.0e-30f, zero = 0.0f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t k, m, hx, hy, ix, iy;
    hx = (int32_t)asuint(x); ix = hx & 0x7fffffff;
    hy = (int32_t)asuint(y); iy = hy & 0x7fffffff;
    if ((ix > 0x7f800000) || (iy > 0x7f800000)) return x + y;
    if (hx == 0x3f800000) return ref_atanf(y);
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
        case 0:
        case 1: return y;
        case 2: return pi + tiny;
        case 3: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return (float)3.0 * pi_o_4 + tiny;
            case 3: return (float)-3.0 * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
            case 0: return zero;
            case 1: return -zero;
            case 2: return pi + tiny;
            case 3: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + (float)0.5 * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = ref_atanf(rm_fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return asfloat(asuint(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

RM_FN float rm_sqrtf(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}

RM_FN float ref_acosf(float x) {   // e_acosf.c
    const float one = 1.0000000000e+00f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
                pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
                pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
                qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    float z, p, q, r, w, s, c, df;
    int32_t hx, ix;
    hx = (int32_t)asuint(x);
    ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) {
        if (hx > 0) return 0.0f;
        return pi + (float)2.0 * pio2_lo;
    } else if (ix > 0x3f800000) {
        return (x - x) / (x - x);
    }
    if (ix < 0x3f000000) {   // |x| < 0.5
        if (ix <= 0x32800000) return pio2_hi + pio2_lo;   // |x| <= 2^-26
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) {   // x < -0.5
        z = (one + x) * (float)0.5;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = rm_sqrtf(z);
        r = p / q;
        w = r * s - pio2_lo;
        return pi - (float)2.0 * (s + w);
    } else {   // x > 0.5
        z = (one - x) * (float)0.5;
        s = rm_sqrtf(z);
        df = asfloat(asuint(s) & 0xfffff000u);
        c = (z - df * df) / (s + df);
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        w = r * s + c;
        return (float)2.0 * (df + w);
    }
}

}  // namespace refmath
#endif
