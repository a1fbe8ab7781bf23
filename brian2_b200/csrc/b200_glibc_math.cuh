// b200_glibc_math.cuh -- exp / expm1 / log / pow / tanh / sinh / cosh / sin / cos with the arithmetic of
// the host's glibc, for device code.
//
// Why: the oracle of this path is the reference's cpp_standalone build, whose exp()/expm1()/log()/pow()
// are glibc's.  CUDA's functions differ from them by <= 1 ulp in a few per cent of the arguments,
// which is enough to leave Hodgkin-Huxley state variables only "within rtol" of the oracle.  With
// `prefs.devices.b200.libm = 'glibc'` (-DB200_GLIBC_MATH) the device runs the functions below
// instead: every floating-point operation -- including which a*b+c are fused -- is the one the
// x86-64 FMA variants of glibc 2.39 execute (`__exp_fma`, `__log_fma`, `__pow_fma`: the table-driven
// algorithms of sysdeps/ieee754/dbl-64/e_exp.c, e_log.c, e_pow.c; `__expm1_fma`: the fdlibm algorithm of
// s_expm1.c; further down tanh/sinh/cosh and sin/cos), so the results are bit-identical for every
// argument (sin/cos: |x| < 1.05e8).  The four lookup tables and the
// polynomial coefficients come from the host's libm itself (b200_libm_tables.h, written into the
// project by brian2_b200/libm_tables.py).  tests/cuda/glibc_math_test.cpp compiles this header for
// the host and compares it with the real functions over >= 10^7 arguments per function.
//
// All arithmetic goes through B200G_* so that neither nvcc (-fmad) nor g++ (-ffp-contract) can
// change the sequence.  errno and floating-point exception flags are not modelled.
#pragma once
#include <stdint.h>
#include "b200_libm_tables.h"

#if defined(__CUDACC__)
#define B200G_FN static __device__ __forceinline__
#define B200G_SLOW static __device__ __noinline__
#define B200G_TABLE static __device__ const
#define B200G_FMA(a, b, c) __fma_rn((a), (b), (c))
#define B200G_MUL(a, b) __dmul_rn((a), (b))
#define B200G_ADD(a, b) __dadd_rn((a), (b))
#define B200G_SUB(a, b) __dsub_rn((a), (b))
#define B200G_DIV(a, b) __ddiv_rn((a), (b))
#define B200G_BITS(x) ((uint64_t)__double_as_longlong(x))
#define B200G_DBL(u) __longlong_as_double((long long)(u))
#else   // host build of the test harness (g++ -ffp-contract=off)
#include <math.h>
#include <string.h>
#define B200G_FN static inline
#define B200G_SLOW static __attribute__((noinline))
#define B200G_TABLE static const
#define B200G_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define B200G_MUL(a, b) ((a) * (b))
#define B200G_ADD(a, b) ((a) + (b))
#define B200G_SUB(a, b) ((a) - (b))
#define B200G_DIV(a, b) ((a) / (b))
static inline uint64_t b200g_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double b200g_dbl(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define B200G_BITS(x) b200g_bits(x)
#define B200G_DBL(u) b200g_dbl(u)
#endif

namespace b200g {

B200G_TABLE uint64_t kExpTab[256] = {B200_LIBM_EXP_TAB};        // {error term, value - (i << 45)} x 128
B200G_TABLE uint64_t kPowLogTab[384] = {B200_LIBM_POW_TAB};     // {1/c, log(c) high, log(c) low} x 128
B200G_TABLE uint64_t kLogTab[256] = {B200_LIBM_LOG_TAB};        // {1/c, log(c)} x 128
B200G_TABLE uint64_t kSinCosTab[440] = {B200_LIBM_SINCOS_TAB};  // {sin, sin low, cos, cos low}(k/128) x 110

static const uint64_t kInfBits = 0x7ff0000000000000ull;
static const uint64_t kOneBits = 0x3ff0000000000000ull;
static const uint64_t kSignBit = 0x8000000000000000ull;

// Results whose scale 2^(k/128) is not a normal double (|x| >= 512, or the pow analogue):
// e_exp.c `specialcase`.  `sign` only matters for pow (a negative base with an odd exponent).
B200G_SLOW double exp_special(double tmp, uint64_t sbits, uint64_t ki, bool pow_variant) {
    if ((ki & 0x80000000ull) == 0) {                       // k > 0: the result may overflow
        sbits -= 1009ull << 52;
        const double scale = B200G_DBL(sbits);
        return B200G_MUL(0x1p1009, B200G_FMA(scale, tmp, scale));
    }
    sbits += 1022ull << 52;                                // k < 0: the result may be subnormal
    const double scale = B200G_DBL(sbits);
    const double st = B200G_MUL(scale, tmp);
    double y = B200G_ADD(scale, st);
    const double ay = pow_variant ? B200G_DBL(B200G_BITS(y) & ~kSignBit) : y;
    if (ay < 1.0) {
        // round once: the double rounding of y followed by the final scaling is avoided by
        // computing the rounding in the right place relative to the subnormal range
        const double one = (pow_variant && y < 0.0) ? -1.0 : 1.0;
        double lo = B200G_ADD(B200G_SUB(scale, y), st);
        const double hi = B200G_ADD(one, y);
        lo = B200G_ADD(B200G_ADD(B200G_SUB(one, hi), y), lo);
        y = B200G_SUB(B200G_ADD(hi, lo), one);
        if (y == 0.0) y = B200G_DBL(sbits & kSignBit);   // keep the sign of a zero result
    }
    return B200G_MUL(0x1p-1022, y);
}

// sysdeps/ieee754/dbl-64/e_exp.c (__exp), FMA variant.
B200G_FN double exp(double x) {
    const uint64_t ix = B200G_BITS(x);
    uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) {                        // |x| < 2^-54 or |x| >= 512 or nan
        if (abstop - 0x3c9u >= 0x80000000u) return B200G_ADD(1.0, x);
        if (abstop >= 0x409u) {                            // |x| >= 1024
            if (ix == (kInfBits | kSignBit)) return 0.0;
            if (abstop >= 0x7ffu) return B200G_ADD(1.0, x);
            return (ix >> 63) ? 0.0 : B200G_DBL(kInfBits);
        }
        abstop = 0;
    }
    // x = k ln2/128 + r; the product with 128/ln2 is fused into the rounding-by-shift addition
    const double ks = B200G_FMA(x, B200_LIBM_EXP_INVLN2N, B200_LIBM_EXP_SHIFT);
    const uint64_t ki = B200G_BITS(ks);
    const double kd = B200G_SUB(ks, B200_LIBM_EXP_SHIFT);
    double r = B200G_FMA(kd, B200_LIBM_EXP_NEGLN2HIN, x);
    r = B200G_FMA(kd, B200_LIBM_EXP_NEGLN2LON, r);
    const uint32_t idx = 2u * ((uint32_t)ki & 127u);
    const double tail = B200G_DBL(kExpTab[idx]);
    const uint64_t sbits = kExpTab[idx + 1] + (ki << 45);
    const double r2 = B200G_MUL(r, r);
    const double p23 = B200G_FMA(r, B200_LIBM_EXP_C3, B200_LIBM_EXP_C2);
    const double p45 = B200G_FMA(r, B200_LIBM_EXP_C5, B200_LIBM_EXP_C4);
    double tmp = B200G_FMA(p23, r2, B200G_ADD(r, tail));
    tmp = B200G_FMA(B200G_MUL(r2, r2), p45, tmp);
    if (abstop == 0) return exp_special(tmp, sbits, ki, false);
    const double scale = B200G_DBL(sbits);
    return B200G_FMA(scale, tmp, scale);
}

// sysdeps/ieee754/dbl-64/s_expm1.c (__expm1, fdlibm), FMA variant.
B200G_FN double expm1(double x) {
    const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33;
    const double invln2 = 0x1.71547652b82fep+0, o_threshold = 0x1.62e42fefa39efp+9;
    const double Q1 = -0x1.11111111110f4p-5, Q2 = 0x1.a01a019fe5585p-10, Q3 = -0x1.4ce199eaadbb7p-14;
    const double Q4 = 0x1.0cfca86e65239p-18, Q5 = -0x1.afdb76e09c32dp-23;
    const uint64_t ix = B200G_BITS(x);
    const uint32_t hx = (uint32_t)(ix >> 32) & 0x7fffffffu;
    const bool neg = (ix >> 63) != 0;
    double hi, lo, c = 0.0;
    int k;
    if (hx >= 0x4043687au) {                               // |x| >= 56 ln2
        if (hx >= 0x40862e42u) {                           // |x| >= 709.78
            if (hx >= 0x7ff00000u) {
                if ((ix & 0x000fffffffffffffull) != 0) return B200G_ADD(x, x);   // nan
                return neg ? -1.0 : x;                                           // +-inf
            }
            if (x > o_threshold) return B200G_DBL(kInfBits);
        }
        if (neg) return B200G_SUB(1e-300, 1.0);
    }
    if (hx > 0x3fd62e42u) {                                // |x| > 0.5 ln2
        if (hx < 0x3ff0a2b2u) {                            // |x| < 1.5 ln2
            if (!neg) { hi = B200G_SUB(x, ln2_hi); lo = ln2_lo; k = 1; }
            else { hi = B200G_ADD(x, ln2_hi); lo = -ln2_lo; k = -1; }
        } else {
            k = (int)B200G_ADD(neg ? -0.5 : 0.5, B200G_MUL(x, invln2));
            const double t = (double)k;
            hi = B200G_FMA(-t, ln2_hi, x);
            lo = B200G_MUL(t, ln2_lo);
        }
        x = B200G_SUB(hi, lo);
        c = B200G_SUB(B200G_SUB(hi, x), lo);
    } else if (hx < 0x3c900000u) {                         // |x| < 2^-54
        return x;
    } else {
        k = 0;
    }
    const double hfx = B200G_MUL(x, 0.5);
    const double hxs = B200G_MUL(x, hfx);
    const double R2 = B200G_FMA(hxs, Q3, Q2);
    const double R3 = B200G_FMA(hxs, Q5, Q4);
    const double h2 = B200G_MUL(hxs, hxs);
    const double R1 = B200G_FMA(hxs, Q1, 1.0);
    const double h4 = B200G_MUL(h2, h2);
    const double r1 = B200G_FMA(h4, R3, B200G_FMA(h2, R2, R1));
    const double t = B200G_FMA(-r1, hfx, 3.0);
    double e = B200G_MUL(B200G_DIV(B200G_SUB(r1, t), B200G_FMA(-x, t, 6.0)), hxs);
    if (k == 0) return B200G_SUB(x, B200G_FMA(e, x, -hxs));
    e = B200G_FMA(B200G_SUB(e, c), x, -c);
    e = B200G_SUB(e, hxs);
    if (k == -1) return B200G_FMA(0.5, B200G_SUB(x, e), -0.5);
    if (k == 1) {
        if (x < -0.25) return B200G_MUL(B200G_SUB(e, B200G_ADD(x, 0.5)), -2.0);
        return B200G_FMA(2.0, B200G_SUB(x, e), 1.0);
    }
    if (k <= -2 || k > 56) {                               // exp(x) - 1 with exp(x) from the series
        const double y = B200G_SUB(1.0, B200G_SUB(e, x));
        const uint64_t u = B200G_BITS(y) + ((uint64_t)(uint32_t)(k << 20) << 32);
        return B200G_SUB(B200G_DBL(u), 1.0);
    }
    double y;
    if (k < 20) {
        const double tk = B200G_DBL((uint64_t)(0x3ff00000u - (0x200000u >> k)) << 32);   // 1 - 2^-k
        y = B200G_SUB(tk, B200G_SUB(e, x));
    } else {
        const double tk = B200G_DBL((uint64_t)((uint32_t)(0x3ff - k) << 20) << 32);       // 2^-k
        y = B200G_ADD(B200G_SUB(x, B200G_ADD(e, tk)), 1.0);
    }
    // add k to the exponent (32-bit arithmetic on the high word, as the original does)
    const uint64_t yb = B200G_BITS(y);
    const uint32_t high = (uint32_t)(yb >> 32) + (uint32_t)(k << 20);
    return B200G_DBL(((uint64_t)high << 32) | (yb & 0xffffffffull));
}

// sysdeps/ieee754/dbl-64/e_log.c (__log), FMA variant.
B200G_SLOW double log_near_one(double x) {                 // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (B200G_BITS(x) == kOneBits) return 0.0;
    const double r = B200G_SUB(x, 1.0);
    const double r2 = B200G_MUL(r, r);
    const double r3 = B200G_MUL(r, r2);
    const double a = B200G_FMA(r2, B200_LIBM_LOG_B3, B200G_FMA(r, B200_LIBM_LOG_B2, B200_LIBM_LOG_B1));
    const double b = B200G_FMA(r2, B200_LIBM_LOG_B6, B200G_FMA(r, B200_LIBM_LOG_B5, B200_LIBM_LOG_B4));
    double c = B200G_FMA(r2, B200_LIBM_LOG_B9, B200G_FMA(r, B200_LIBM_LOG_B8, B200_LIBM_LOG_B7));
    c = B200G_FMA(r3, B200_LIBM_LOG_B10, c);
    const double e = B200G_FMA(B200G_FMA(c, r3, b), r3, a);
    // r*r*B0 with the high part of r (27 bits, so rhi*rhi is exact), both roundings fused away
    const double t = B200G_FMA(r, 0x1p27, r);
    const double rhi = B200G_FMA(-0x1p27, r, t);
    const double rlo = B200G_SUB(r, rhi);
    const double rhi2 = B200G_MUL(rhi, rhi);
    const double hi = B200G_FMA(rhi2, B200_LIBM_LOG_B0, r);
    double lo = B200G_FMA(rhi2, B200_LIBM_LOG_B0, B200G_SUB(r, hi));
    lo = B200G_FMA(B200G_MUL(B200_LIBM_LOG_B0, rlo), B200G_ADD(r, rhi), lo);
    return B200G_ADD(hi, B200G_FMA(e, r3, lo));
}
B200G_FN double log(double x) {
    uint64_t ix = B200G_BITS(x);
    const uint32_t top = (uint32_t)(ix >> 48);
    if (ix - 0x3fee000000000000ull < 0x0003090000000000ull) return log_near_one(x);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {              // x < 2^-1022, inf or nan
        if (ix * 2 == 0) return B200G_DBL(kInfBits | kSignBit);
        if (ix == kInfBits) return x;
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return B200G_DBL(0x7ff8000000000000ull | kSignBit);
        ix = B200G_BITS(B200G_MUL(x, 0x1p52));             // subnormal: normalise
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
    const int k = (int)((int64_t)tmp >> 52);
    const double z = B200G_DBL(ix - (tmp & 0xfff0000000000000ull));
    const double invc = B200G_DBL(kLogTab[2 * i]);
    const double logc = B200G_DBL(kLogTab[2 * i + 1]);
    const double r = B200G_FMA(z, invc, -1.0);
    const double kd = (double)k;
    const double w = B200G_FMA(kd, B200_LIBM_LOG_LN2HI, logc);
    const double hi = B200G_ADD(w, r);
    const double lo = B200G_FMA(kd, B200_LIBM_LOG_LN2LO, B200G_ADD(B200G_SUB(w, hi), r));
    const double r2 = B200G_MUL(r, r);
    const double p12 = B200G_FMA(r, B200_LIBM_LOG_A2, B200_LIBM_LOG_A1);
    const double p34 = B200G_FMA(r, B200_LIBM_LOG_A4, B200_LIBM_LOG_A3);
    const double q = B200G_FMA(p34, r2, p12);
    const double y = B200G_FMA(B200G_MUL(r, r2), q, B200G_FMA(r2, B200_LIBM_LOG_A0, lo));
    return B200G_ADD(y, hi);
}

// ---- functions glibc builds on the ones above (fdlibm; compiled without fused operations) ------
// sysdeps/ieee754/dbl-64/s_tanh.c
B200G_FN double tanh(double x) {
    const uint64_t ix64 = B200G_BITS(x);
    const uint32_t ix = (uint32_t)(ix64 >> 32) & 0x7fffffffu;
    const bool neg = (ix64 >> 63) != 0;
    if (ix >= 0x7ff00000u)                                 // inf: +-1, nan: nan
        return neg ? B200G_SUB(B200G_DIV(1.0, x), 1.0) : B200G_ADD(B200G_DIV(1.0, x), 1.0);
    double z;
    if (ix < 0x40360000u) {                                // |x| < 22
        if ((ix64 << 1) == 0) return x;
        if (ix < 0x3c800000u) return B200G_MUL(x, B200G_ADD(1.0, x));            // |x| < 2^-55
        const double ax = B200G_DBL(ix64 & ~kSignBit);
        if (ix >= 0x3ff00000u) {                           // |x| >= 1
            const double t = expm1(B200G_ADD(ax, ax));
            z = B200G_SUB(1.0, B200G_DIV(2.0, B200G_ADD(t, 2.0)));
        } else {
            const double t = expm1(B200G_MUL(ax, -2.0));
            z = B200G_DIV(-t, B200G_ADD(t, 2.0));
        }
    } else {
        z = B200G_SUB(1.0, 1e-300);
    }
    return neg ? -z : z;
}
// sysdeps/ieee754/dbl-64/e_sinh.c
B200G_FN double sinh(double x) {
    const uint64_t ix64 = B200G_BITS(x);
    const uint32_t ix = (uint32_t)(ix64 >> 32) & 0x7fffffffu, lx = (uint32_t)ix64;
    if (ix >= 0x7ff00000u) return B200G_ADD(x, x);
    const double h = (ix64 >> 63) ? -0.5 : 0.5;
    const double ax = B200G_DBL(ix64 & ~kSignBit);
    if (ix < 0x40360000u) {                                // |x| < 22
        if (ix < 0x3e300000u) return x;                    // |x| < 2^-28
        const double t = expm1(ax);
        if (ix < 0x3ff00000u)
            return B200G_MUL(h, B200G_SUB(B200G_MUL(2.0, t), B200G_DIV(B200G_MUL(t, t), B200G_ADD(t, 1.0))));
        return B200G_MUL(h, B200G_ADD(t, B200G_DIV(t, B200G_ADD(t, 1.0))));
    }
    if (ix < 0x40862e42u) return B200G_MUL(h, exp(ax));    // |x| < log(DBL_MAX)
    if (ix < 0x408633ceu || (ix == 0x408633ceu && lx <= 0x8fb9f87du)) {
        const double w = exp(B200G_MUL(0.5, ax));
        return B200G_MUL(B200G_MUL(h, w), w);
    }
    return B200G_MUL(x, 1.0e307);                          // overflow
}
// sysdeps/ieee754/dbl-64/e_cosh.c
B200G_FN double cosh(double x) {
    const uint64_t ix64 = B200G_BITS(x);
    const uint32_t ix = (uint32_t)(ix64 >> 32) & 0x7fffffffu, lx = (uint32_t)ix64;
    const double ax = B200G_DBL(ix64 & ~kSignBit);
    if (ix < 0x40360000u) {                                // |x| < 22
        if (ix < 0x3fd62e43u) {                            // |x| < 0.5 ln2
            if (ix < 0x3c800000u) return 1.0;
            const double t = expm1(ax);
            const double w = B200G_ADD(1.0, t);
            return B200G_ADD(1.0, B200G_DIV(B200G_MUL(t, t), B200G_ADD(w, w)));
        }
        const double t = exp(ax);
        return B200G_ADD(B200G_MUL(0.5, t), B200G_DIV(0.5, t));
    }
    if (ix < 0x40862e42u) return B200G_MUL(0.5, exp(ax));
    if (ix < 0x408633ceu || (ix == 0x408633ceu && lx <= 0x8fb9f87du)) {
        const double w = exp(B200G_MUL(0.5, ax));
        return B200G_MUL(B200G_MUL(0.5, w), w);
    }
    if (ix >= 0x7ff00000u) return B200G_MUL(x, x);
    return B200G_DBL(kInfBits);                            // overflow
}

// ---- sin / cos: sysdeps/ieee754/dbl-64/s_sin.c (__sin, __cos; IBM Accurate Mathematical Library),
// FMA variants.  |x| < 105414350 is restated; beyond that glibc reduces the argument with a
// 1200-bit 2/pi (branred.c), which is not: CUDA's sin/cos take those arguments (<= 1-2 ulp).
B200G_FN double sc_copysign(double mag, double sgn) {
    return B200G_DBL((B200G_BITS(mag) & ~kSignBit) | (B200G_BITS(sgn) & kSignBit));
}
// sin(a + dx) for |a| < ~0.86 (do_sin): odd Taylor polynomial below 0.126, else
// sin(x0 + x) = sin x0 cos x + cos x0 sin x with x0 = a rounded to 1/128 from the table
B200G_FN double sc_do_sin(double a, double dx) {
    const double aa = B200G_DBL(B200G_BITS(a) & ~kSignBit);
    if (aa < 0.126) {
        const double xx = B200G_MUL(a, a);
        double p = B200G_FMA(xx, -0x1.addffc2fcdf59p-26, 0x1.71de27b9a7ed9p-19);
        p = B200G_FMA(xx, p, -0x1.a01a019db08b8p-13);
        p = B200G_FMA(xx, p, 0x1.1111111110ecep-7);
        p = B200G_FMA(xx, p, -0x1.5555555555555p-3);
        const double t = B200G_FMA(xx, B200G_FMA(p, a, -B200G_MUL(dx, 0.5)), dx);
        return B200G_ADD(a, t);
    }
    if (!(a > 0.0)) dx = -dx;
    const double u = B200G_ADD(aa, 0x1.8p45);
    const double x = B200G_SUB(aa, B200G_SUB(u, 0x1.8p45));
    const uint32_t k = (uint32_t)B200G_BITS(u) << 2;
    const double xx = B200G_MUL(x, x);
    const double p = B200G_FMA(xx, 0x1.11110e829872fp-7, -0x1.5555555555515p-3);
    const double s = B200G_ADD(x, B200G_FMA(B200G_MUL(x, xx), p, dx));
    const double c0 = B200G_FMA(xx, B200G_FMA(xx, 0x1.6c16bedd9e239p-10, -0x1.5555555555535p-5), 0.5);
    const double c = B200G_FMA(dx, x, B200G_MUL(xx, c0));
    const double sn = B200G_DBL(kSinCosTab[k]), ssn = B200G_DBL(kSinCosTab[k + 1]);
    const double cs = B200G_DBL(kSinCosTab[k + 2]), ccs = B200G_DBL(kSinCosTab[k + 3]);
    const double cor = B200G_FMA(s, cs, B200G_FMA(-c, sn, B200G_FMA(s, ccs, ssn)));
    return sc_copysign(B200G_ADD(sn, cor), a);
}
// cos(a + dx) for |a| < ~0.86 (do_cos)
B200G_FN double sc_do_cos(double a, double dx) {
    const double aa = B200G_DBL(B200G_BITS(a) & ~kSignBit);
    if (a < 0.0) dx = -dx;
    const double u = B200G_ADD(aa, 0x1.8p45);
    const double x = B200G_ADD(B200G_SUB(aa, B200G_SUB(u, 0x1.8p45)), dx);
    const uint32_t k = (uint32_t)B200G_BITS(u) << 2;
    const double xx = B200G_MUL(x, x);
    const double p = B200G_FMA(xx, 0x1.11110e829872fp-7, -0x1.5555555555515p-3);
    const double s = B200G_FMA(B200G_MUL(x, xx), p, x);
    const double c0 = B200G_FMA(xx, B200G_FMA(xx, 0x1.6c16bedd9e239p-10, -0x1.5555555555535p-5), 0.5);
    const double c = B200G_MUL(xx, c0);
    const double sn = B200G_DBL(kSinCosTab[k]), ssn = B200G_DBL(kSinCosTab[k + 1]);
    const double cs = B200G_DBL(kSinCosTab[k + 2]), ccs = B200G_DBL(kSinCosTab[k + 3]);
    const double cor = B200G_FMA(-s, sn, B200G_FMA(-c, cs, B200G_FMA(-s, ssn, ccs)));
    return B200G_ADD(cs, cor);
}
// x = n pi/2 + (a + da), |a| <= pi/4, for |x| < 105414350 (reduce_sincos: pi/2 in four pieces)
B200G_FN int sc_reduce(double x, double& a, double& da) {
    const double t = B200G_FMA(x, 0x1.45f306dc9c883p-1, 0x1.8p52);
    const double xn = B200G_SUB(t, 0x1.8p52);
    const int n = (int)((uint32_t)B200G_BITS(t) & 3u);
    const double y = B200G_FMA(-xn, -0x1.dde973c000000p-27, B200G_FMA(-xn, 0x1.921fb58000000p+0, x));
    const double pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;
    const double t2 = B200G_FMA(-xn, pp3, y);
    double db = B200G_FMA(-pp3, xn, B200G_SUB(y, t2));
    const double b = B200G_FMA(-xn, pp4, t2);
    db = B200G_ADD(db, B200G_FMA(-xn, pp4, B200G_SUB(t2, b)));
    a = b;
    da = db;
    return n;
}
B200G_FN double sc_quadrant(double a, double da, int n) {
    const double r = (n & 1) ? sc_do_cos(a, da) : sc_do_sin(a, da);
    return (n & 2) ? -r : r;
}
B200G_FN double sin(double x) {
    const uint32_t k = (uint32_t)(B200G_BITS(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;                         // |x| < 2^-26
    if (k < 0x3feb6000u) return sc_do_sin(x, 0.0);         // |x| < 0.855469
    if (k < 0x400368fdu) {                                 // |x| < 2.426265: cos(pi/2 - |x|)
        const double t = B200G_SUB(0x1.921fb54442d18p+0, B200G_DBL(B200G_BITS(x) & ~kSignBit));
        return sc_copysign(sc_do_cos(t, 0x1.1a62633145c07p-54), x);
    }
    if (k < 0x419921fbu) {                                 // |x| < 105414350
        double a, da;
        const int n = sc_reduce(x, a, da);
        return sc_quadrant(a, da, n);
    }
    if (k < 0x7ff00000u) return ::sin(x);                  // not restated (see above)
    return B200G_DIV(x, x);                                // inf, nan
}
B200G_FN double cos(double x) {
    const uint32_t k = (uint32_t)(B200G_BITS(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e400000u) return 1.0;                       // |x| < 2^-27
    if (k < 0x3feb6000u) return sc_do_cos(x, 0.0);
    if (k < 0x400368fdu) {                                 // sin(pi/2 - |x|)
        const double y = B200G_SUB(0x1.921fb54442d18p+0, B200G_DBL(B200G_BITS(x) & ~kSignBit));
        const double a = B200G_ADD(y, 0x1.1a62633145c07p-54);
        const double da = B200G_ADD(B200G_SUB(y, a), 0x1.1a62633145c07p-54);
        return sc_do_sin(a, da);
    }
    if (k < 0x419921fbu) {
        double a, da;
        const int n = sc_reduce(x, a, da);
        return sc_quadrant(a, da, n + 1);
    }
    if (k < 0x7ff00000u) return ::cos(x);
    return B200G_DIV(x, x);
}

// ---- pow: sysdeps/ieee754/dbl-64/e_pow.c (__pow), FMA variant ---------------------------------
B200G_FN int pow_checkint(uint64_t iy) {                   // 0: not an integer, 1: odd, 2: even
    const int e = (int)(iy >> 52) & 0x7ff;
    if (e < 0x3ff) return 0;
    if (e > 0x3ff + 52) return 2;
    if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
    if (iy & (1ull << (0x3ff + 52 - e))) return 1;
    return 2;
}
B200G_FN bool pow_zeroinfnan(uint64_t i) { return 2 * i - 1 >= 2 * kInfBits - 1; }
B200G_FN bool pow_issignaling(uint64_t i) {
    return (i & 0x7ff8000000000000ull) == kInfBits && (i & 0x0007ffffffffffffull) != 0;
}

// Everything that is not "x normal and positive, 2^-65 <= |y| < 2^63": returns true when `out`
// is the result, otherwise ix / sign_bias are adjusted for the main path.
B200G_SLOW bool pow_special(double x, double y, uint64_t& ix, uint32_t& sign_bias, double& out) {
    const uint64_t iy = B200G_BITS(y);
    uint32_t topx = (uint32_t)(ix >> 52);
    const uint32_t topy = (uint32_t)(iy >> 52);
    if (pow_zeroinfnan(iy)) {
        if (2 * iy == 0) { out = pow_issignaling(ix) ? B200G_ADD(x, y) : 1.0; return true; }
        if (ix == kOneBits) { out = pow_issignaling(iy) ? B200G_ADD(x, y) : 1.0; return true; }
        if (2 * ix > 2 * kInfBits || 2 * iy > 2 * kInfBits) { out = B200G_ADD(x, y); return true; }
        if (2 * ix == 2 * kOneBits) { out = 1.0; return true; }
        if ((2 * ix < 2 * kOneBits) == !(iy >> 63)) { out = 0.0; return true; }
        out = B200G_MUL(y, y);
        return true;
    }
    if (pow_zeroinfnan(ix)) {
        double x2 = B200G_MUL(x, x);
        if ((ix >> 63) && pow_checkint(iy) == 1) x2 = -x2;
        out = (iy >> 63) ? B200G_DIV(1.0, x2) : x2;
        return true;
    }
    if (ix >> 63) {                                        // finite x < 0
        const int yint = pow_checkint(iy);
        if (yint == 0) { out = B200G_DBL(0x7ff8000000000000ull | kSignBit); return true; }   // (x-x)/(x-x)
        if (yint == 1) sign_bias = 0x800u << 7;
        ix &= ~kSignBit;
        topx &= 0x7ffu;
    }
    if ((topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) {     // |y| < 2^-65 or |y| >= 2^63
        if (ix == kOneBits) { out = 1.0; return true; }
        if ((topy & 0x7ffu) < 0x3beu) {
            out = ix > kOneBits ? B200G_ADD(1.0, y) : B200G_SUB(1.0, y);
            return true;
        }
        out = ((ix > kOneBits) == (topy < 0x800u)) ? B200G_DBL(kInfBits) : 0.0;
        return true;
    }
    if (topx == 0) {                                       // subnormal x: normalise
        ix = B200G_BITS(B200G_MUL(x, 0x1p52));
        ix &= ~kSignBit;
        ix -= 52ull << 52;
    }
    return false;
}

B200G_FN double pow(double x, double y) {
    uint64_t ix = B200G_BITS(x);
    const uint64_t iy = B200G_BITS(y);
    const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
    uint32_t sign_bias = 0;
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) {
        double out;
        if (pow_special(x, y, ix, sign_bias, out)) return out;
    }
    // ---- log(x) = k ln2 + log(c) + log1p(z/c - 1) as hi + lo (log_inline) ----
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
    const int k = (int)((int64_t)tmp >> 52);
    const double z = B200G_DBL(ix - (tmp & 0xfff0000000000000ull));
    const double kd = (double)k;
    const double invc = B200G_DBL(kPowLogTab[3 * i]);
    const double logc = B200G_DBL(kPowLogTab[3 * i + 1]);
    const double logctail = B200G_DBL(kPowLogTab[3 * i + 2]);
    const double r = B200G_FMA(z, invc, -1.0);             // exact
    const double t1 = B200G_FMA(kd, B200_LIBM_POW_LN2HI, logc);
    const double t2 = B200G_ADD(t1, r);
    const double lo1 = B200G_FMA(kd, B200_LIBM_POW_LN2LO, logctail);
    const double lo2 = B200G_ADD(B200G_SUB(t1, t2), r);
    const double ar = B200G_MUL(B200_LIBM_POW_A0, r);
    const double ar2 = B200G_MUL(r, ar);
    const double ar3 = B200G_MUL(r, ar2);
    const double hi0 = B200G_ADD(t2, ar2);
    const double lo3 = B200G_FMA(ar, r, -ar2);
    const double lo4 = B200G_ADD(B200G_SUB(t2, hi0), ar2);
    const double p12 = B200G_FMA(r, B200_LIBM_POW_A2, B200_LIBM_POW_A1);
    const double p34 = B200G_FMA(r, B200_LIBM_POW_A4, B200_LIBM_POW_A3);
    const double p56 = B200G_FMA(r, B200_LIBM_POW_A6, B200_LIBM_POW_A5);
    const double q = B200G_FMA(ar2, B200G_FMA(p56, ar2, p34), p12);
    const double losum = B200G_ADD(B200G_ADD(B200G_ADD(lo1, lo2), lo3), lo4);
    const double lo = B200G_FMA(ar3, q, losum);
    const double lhi = B200G_ADD(hi0, lo);
    const double llo = B200G_ADD(B200G_SUB(hi0, lhi), lo);
    // ---- y * log(x) as ehi + elo ----
    const double ehi = B200G_MUL(y, lhi);
    const double elo = B200G_FMA(y, llo, B200G_FMA(y, lhi, -ehi));
    // ---- exp(ehi + elo) (exp_inline) ----
    uint32_t abstop = (uint32_t)(B200G_BITS(ehi) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop - 0x3c9u >= 0x80000000u) {              // |y log x| < 2^-54
            const double one = B200G_ADD(1.0, ehi);
            return sign_bias ? -one : one;
        }
        if (abstop >= 0x409u) {                            // over/underflow
            const double big = (B200G_BITS(ehi) >> 63) ? 0.0 : B200G_DBL(kInfBits);
            return sign_bias ? -big : big;
        }
        abstop = 0;
    }
    const double ks = B200G_FMA(ehi, B200_LIBM_EXP_INVLN2N, B200_LIBM_EXP_SHIFT);
    const uint64_t ki = B200G_BITS(ks);
    const double kd2 = B200G_SUB(ks, B200_LIBM_EXP_SHIFT);
    double rr = B200G_FMA(kd2, B200_LIBM_EXP_NEGLN2HIN, ehi);
    rr = B200G_FMA(kd2, B200_LIBM_EXP_NEGLN2LON, rr);
    rr = B200G_ADD(elo, rr);
    const uint32_t idx = 2u * ((uint32_t)ki & 127u);
    const double tail = B200G_DBL(kExpTab[idx]);
    const uint64_t sbits = kExpTab[idx + 1] + ((ki + sign_bias) << 45);
    const double r2 = B200G_MUL(rr, rr);
    const double p23 = B200G_FMA(rr, B200_LIBM_EXP_C3, B200_LIBM_EXP_C2);
    const double p45 = B200G_FMA(rr, B200_LIBM_EXP_C5, B200_LIBM_EXP_C4);
    double tmp2 = B200G_FMA(p23, r2, B200G_ADD(rr, tail));
    tmp2 = B200G_FMA(B200G_MUL(r2, r2), p45, tmp2);
    if (abstop == 0) return exp_special(tmp2, sbits, ki, true);
    const double scale = B200G_DBL(sbits);
    return B200G_FMA(scale, tmp2, scale);
}

}  // namespace b200g
