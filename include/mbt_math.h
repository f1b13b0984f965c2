/*
 * mbt_math.h -- bit-reproducible exp / log / normal-quantile for float and double,
 * shared by the sm_100a kernels and the CPU oracle.  Plain C99 / CUDA C++.
 *
 * Why it exists: one env-step takes DISCRETE decisions from transcendental values --
 * a fill happens iff  u < exp(-kappa * depth)  (reference:
 * mbt_gym/stochastic_processes/fill_probability_models.py:33,58) -- so a 1-ulp
 * difference between glibc's and CUDA's `exp` can flip a fill and move a trajectory's
 * inventory by one unit.  To make "GPU == oracle" a BIT-EXACT statement rather than a
 * statistical one, both sides evaluate the functions below: only IEEE-754 correctly
 * rounded primitives (+, -, *, /, sqrt, fma, int<->float conversion) in a fixed order.
 * Build rules that keep that true: nvcc `-fmad=false` (no implicit contraction; every
 * fused multiply-add here is an explicit fma), no --use_fast_math (IEEE div/sqrt, no
 * FTZ); gcc `-ffp-contract=off`.
 *
 * Accuracy (measured in tests/test_primitives.py against libm / scipy):
 *   mbt_exp_f32, mbt_log_f32   <= 2 ulp ;  mbt_exp_f64, mbt_log_f64  <= 4 ulp
 *   mbt_normal_from_bits_f32   rel err <= 1e-6 ; _f64  rel err <= 1e-13
 */
#ifndef MBT_MATH_H
#define MBT_MATH_H

#include "mbt_rtc.h" /* <math.h>, <stdint.h>, <string.h> -- or their NVRTC stand-ins */

#include "mbt_philox.h" /* MBT_HD */

/* ------------------------------------------------------------------ bit casts */
MBT_HD uint32_t mbt_f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, sizeof u);
    return u;
#endif
}
MBT_HD float mbt_bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, sizeof f);
    return f;
#endif
}
MBT_HD uint64_t mbt_f64_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, sizeof u);
    return u;
#endif
}
MBT_HD double mbt_bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
#endif
}

/* ------------------------------------------------------------------ exp, float */
/* Cody-Waite reduction x = n*ln2 + r, |r| <= ln2/2, degree-5 minimax tail (the classic
 * Cephes single-precision coefficient set), scaled by 2^n through the exponent field. */
/* exp(x) * 2^k, k a small non-negative integer folded into the exponent add (free).  Branch-free: x is
 * clamped to [-87, 88 - 0.7k] (exp(-87) = 1.6e-38 is the smallest value returned, the upper clamp keeps the
 * scaled result finite; NaN is treated as -87). */
MBT_HD float mbt_exp2k_f32(float x, int k) {
    x = fminf(fmaxf(x, -87.0f), 88.0f - 0.7f * (float)k); /* upper clamp keeps exp(x)*2^k finite */
    const float magic = 12582912.0f; /* 1.5 * 2^23: adding it rounds to nearest integer */
    float t = fmaf(x, 1.44269504088896341f, magic);
    float n = t - magic;
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r) + 1.0f;
    /* n is an integer in [-126, 127]: the low bits of t hold it (t = 1.5*2^23 + n exactly) */
    int32_t ni = (int32_t)(mbt_f32_bits(t) & 0x007FFFFFu) - 0x00400000;
    return mbt_bits_f32(mbt_f32_bits(y) + ((uint32_t)(ni + k) << 23));
}
MBT_HD float mbt_exp_f32(float x) { return mbt_exp2k_f32(x, 0); }

/* ------------------------------------------------------------------ log, float */
/* x = m * 2^e, m in [sqrt(1/2), sqrt(2)); log(m) by the Cephes degree-8 polynomial in
 * f = m - 1.  Defined for finite x > 0 (the only use is on products in (0, 1]); x <= 0
 * returns -inf / NaN like libm. */
MBT_HD float mbt_log_f32(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;
    uint32_t b = mbt_f32_bits(x);
    int32_t e = (int32_t)(b >> 23) - 126;
    if ((b >> 23) == 0u) { /* subnormal: renormalise exactly */
        b = mbt_f32_bits(x * 8388608.0f);
        e = (int32_t)(b >> 23) - 126 - 23;
    }
    float m = mbt_bits_f32((b & 0x007FFFFFu) | 0x3F000000u); /* [0.5, 1) */
    float f;
    if (m < 0.707106781186547524f) {
        e -= 1;
        f = (m + m) - 1.0f;
    } else {
        f = m - 1.0f;
    }
    float z = f * f;
    float p = 7.0376836292e-2f;
    p = fmaf(p, f, -1.1514610310e-1f);
    p = fmaf(p, f, 1.1676998740e-1f);
    p = fmaf(p, f, -1.2420140846e-1f);
    p = fmaf(p, f, 1.4249322787e-1f);
    p = fmaf(p, f, -1.6668057665e-1f);
    p = fmaf(p, f, 2.0000714765e-1f);
    p = fmaf(p, f, -2.4999993993e-1f);
    p = fmaf(p, f, 3.3333331174e-1f);
    float fe = (float)e;
    float y = (p * f) * z;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    float res = f + y;
    return fmaf(fe, 0.693359375f, res);
}

/* log(x) for x known to be a positive NORMAL float (no zero / negative / subnormal handling): the argument of the
 * normal quantile below is always in [2^-32, 1]. Same arithmetic as mbt_log_f32 on that domain. */
MBT_HD float mbt_log_pn_f32(float x) {
    uint32_t b = mbt_f32_bits(x);
    int32_t e = (int32_t)(b >> 23) - 126;
    float m = mbt_bits_f32((b & 0x007FFFFFu) | 0x3F000000u); /* [0.5, 1) */
    float f;
    if (m < 0.707106781186547524f) {
        e -= 1;
        f = (m + m) - 1.0f;
    } else {
        f = m - 1.0f;
    }
    float z = f * f;
    float p = 7.0376836292e-2f;
    p = fmaf(p, f, -1.1514610310e-1f);
    p = fmaf(p, f, 1.1676998740e-1f);
    p = fmaf(p, f, -1.2420140846e-1f);
    p = fmaf(p, f, 1.4249322787e-1f);
    p = fmaf(p, f, -1.6668057665e-1f);
    p = fmaf(p, f, 2.0000714765e-1f);
    p = fmaf(p, f, -2.4999993993e-1f);
    p = fmaf(p, f, 3.3333331174e-1f);
    float fe = (float)e;
    float y = (p * f) * z;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    float res = f + y;
    return fmaf(fe, 0.693359375f, res);
}

/* ------------------------------------------------------------------ exp, double */
/* Polynomial coefficients live in a table: on the device a __constant__ array, so every DFMA takes its
 * coefficient straight from the constant bank (a 64-bit literal would cost two UMOVs each). */
#if defined(__CUDACC__)
#define MBT_TABLE(name, n, ...)                         \
    static const double name##_h[n] = {__VA_ARGS__};    \
    static __constant__ double name##_d[n] = {__VA_ARGS__};
#else
#define MBT_TABLE(name, n, ...) static const double name##_h[n] = {__VA_ARGS__};
#endif
#if defined(__CUDA_ARCH__)
#define MBT_T(name) name##_d
#else
#define MBT_T(name) name##_h
#endif

/* Taylor coefficients 1/13! ... 1/2! : truncation < 5e-18 on |r| <= ln2/2 */
MBT_TABLE(mbt_exp64_c, 12, 1.6059043836821614599e-10, 2.0876756987868098979e-9, 2.5052108385441718775e-8,
          2.7557319223985890653e-7, 2.7557319223985890653e-6, 2.4801587301587301587e-5, 1.9841269841269841270e-4,
          1.3888888888888888889e-3, 8.3333333333333333333e-3, 4.1666666666666666667e-2, 1.6666666666666666667e-1, 0.5)

MBT_HD double mbt_exp2k_f64(double x, int k) {
    x = fmin(fmax(x, -700.0), 700.0 - 0.7 * (double)k); /* branch-free clamp; NaN is treated as -700 */
    const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
    double t = fma(x, 1.4426950408889634074, magic);
    double n = t - magic;
    double r = fma(n, -6.93147180369123816490e-01, x); /* ln2 high part (fdlibm split) */
    r = fma(n, -1.90821492927058770002e-10, r);        /* ln2 low part               */
    double p = MBT_T(mbt_exp64_c)[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 1; j < 12; ++j) p = fma(p, r, MBT_T(mbt_exp64_c)[j]);
    double r2 = r * r;
    double y = fma(p, r2, r) + 1.0;
    /* n is an integer in [-1010, 1010]: the low word of t holds it in two's complement */
    int32_t ni = (int32_t)(uint32_t)mbt_f64_bits(t);
    return mbt_bits_f64(mbt_f64_bits(y) + ((uint64_t)(int64_t)(ni + k) << 52));
}
MBT_HD double mbt_exp_f64(double x) { return mbt_exp2k_f64(x, 0); }

/* ------------------------------------------------------------------ log, double */
/* log(m) = 2*atanh(s), s = f/(2+f), odd series to s^23 (|s| <= 0.1716 -> trunc < 2e-19). */
MBT_HD double mbt_log_f64(double x) {
    if (!(x > 0.0)) return (x == 0.0) ? -(double)INFINITY : (double)NAN;
    uint64_t b = mbt_f64_bits(x);
    int32_t e = (int32_t)(b >> 52) - 1022;
    if ((b >> 52) == 0u) {
        b = mbt_f64_bits(x * 4503599627370496.0);
        e = (int32_t)(b >> 52) - 1022 - 52;
    }
    double m = mbt_bits_f64((b & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull); /* [0.5,1) */
    double f;
    if (m < 0.70710678118654752440) {
        e -= 1;
        f = (m + m) - 1.0;
    } else {
        f = m - 1.0;
    }
    double s = f / (2.0 + f);
    double s2 = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, s2, 1.0 / 21.0);
    p = fma(p, s2, 1.0 / 19.0);
    p = fma(p, s2, 1.0 / 17.0);
    p = fma(p, s2, 1.0 / 15.0);
    p = fma(p, s2, 1.0 / 13.0);
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    /* log(m) = 2s + 2s*s2*p */
    double two_s = s + s;
    double lm = fma(two_s * s2, p, two_s);
    double fe = (double)e;
    double lo = fma(fe, 1.90821492927058770002e-10, lm);
    return fma(fe, 6.93147180369123816490e-01, lo);
}

/* ------------------------------------------------------------------ normal from 32 bits */
/*
 * Standard normal from 32 random bits by inversion: bit 31 is the sign, the low 31 bits m
 * give the half-normal probability level.  With mc = 2^31-1-m,
 *     t = (mc + 1/2) / 2^31  in (0,1)     (upper-tail mass; small t = far tail, converted
 *                                          from the integer so the tail keeps full precision)
 *     v = 1 - t,   w = -log(t * (2 - t)) = -log(1 - v^2),
 *     |z| = sqrt(2) * erfinv(v) = sqrt(2) * v * P(w)
 * P: Giles' piecewise polynomials ("Approximating the erfinv function", 2011) --
 * float: his published single-precision pair; double: three segments fitted by
 * tools/fit_normal_icdf.py.  Largest |z| = 6.3 (t = 2^-32).
 */
MBT_HD float mbt_normal_from_bits_f32(uint32_t bits) {
    uint32_t mc = 0x7FFFFFFFu - (bits & 0x7FFFFFFFu);
    float t = fmaf((float)mc, 4.656612873077392578125e-10f /* 2^-31 */, 2.3283064365386962890625e-10f /* 2^-32 */);
    float v = 1.0f - t;
    float w = -mbt_log_pn_f32(t * (2.0f - t));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = fmaf(p, w, 3.43273939e-07f);
        p = fmaf(p, w, -3.5233877e-06f);
        p = fmaf(p, w, -4.39150654e-06f);
        p = fmaf(p, w, 0.00021858087f);
        p = fmaf(p, w, -0.00125372503f);
        p = fmaf(p, w, -0.00417768164f);
        p = fmaf(p, w, 0.246640727f);
        p = fmaf(p, w, 1.50140941f);
    } else if (w >= 16.0f) { /* beyond Giles' single-precision range (t < 2^-24): own degree-4 fit,
                              * tools/fit_normal_icdf.py segment TAIL re-fitted at float accuracy */
        w = sqrtf(w) - 4.375f;
        p = 1.791975665e-04f;
        p = fmaf(p, w, -5.201602471e-04f);
        p = fmaf(p, w, 5.034005735e-04f);
        p = fmaf(p, w, 1.010129929e+00f);
        p = fmaf(p, w, 4.218480587e+00f);
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = fmaf(p, w, 0.000100950558f);
        p = fmaf(p, w, 0.00134934322f);
        p = fmaf(p, w, -0.00367342844f);
        p = fmaf(p, w, 0.00573950773f);
        p = fmaf(p, w, -0.0076224613f);
        p = fmaf(p, w, 0.00943887047f);
        p = fmaf(p, w, 1.00167406f);
        p = fmaf(p, w, 2.83297682f);
    }
    float z = (1.41421356237309504880f * v) * p;
    return (bits >> 31) ? -z : z;
}

MBT_HD double mbt_normal_from_bits_f64(uint32_t bits) {
    uint32_t mc = 0x7FFFFFFFu - (bits & 0x7FFFFFFFu);
    double t = ((double)mc + 0.5) * 4.656612873077392578125e-10; /* exact */
    double v = 1.0 - t;
    double w = -mbt_log_f64(t * (2.0 - t));
    double p;
    if (w < 6.25) {
        w = w - 3.125;
        p = -2.93378702136066883e-20;
        p = fma(p, w, -1.70618749633882542e-20);
        p = fma(p, w, 1.65276491700746045e-18);
        p = fma(p, w, 7.64220219799399885e-19);
        p = fma(p, w, -3.94324760557796949e-17);
        p = fma(p, w, -1.09079577435404181e-17);
        p = fma(p, w, 4.38503143106243651e-16);
        p = fma(p, w, 3.16888922080693809e-16);
        p = fma(p, w, 1.57760898062725170e-15);
        p = fma(p, w, -4.30211992333208736e-14);
        p = fma(p, w, -5.21610747729488213e-14);
        p = fma(p, w, 2.64690701492755575e-12);
        p = fma(p, w, -1.30878116424140578e-11);
        p = fma(p, w, -5.42012076052834127e-11);
        p = fma(p, w, 1.05149424965246830e-09);
        p = fma(p, w, -4.11252892034572366e-09);
        p = fma(p, w, -2.90708125408396845e-08);
        p = fma(p, w, 4.23478637261581625e-07);
        p = fma(p, w, -1.36546879478674747e-06);
        p = fma(p, w, -1.38825232602688395e-05);
        p = fma(p, w, 1.86734207843842592e-04);
        p = fma(p, w, -7.40702534198593764e-04);
        p = fma(p, w, -6.03367087139462018e-03);
        p = fma(p, w, 2.40158182425592087e-01);
        p = fma(p, w, 1.65365456268310140e+00);
    } else {
        double s = sqrt(w);
        if (s < 4.0) {
            s = s - 3.25;
            p = -7.62029190603530220e-07;
            p = fma(p, s, -7.16154295328572080e-08);
            p = fma(p, s, 2.09610158764196380e-06);
            p = fma(p, s, 2.77641204894734524e-07);
            p = fma(p, s, -2.71235730553805872e-06);
            p = fma(p, s, -1.86564814457334863e-07);
            p = fma(p, s, 3.06667208412655536e-06);
            p = fma(p, s, -3.89157826630587532e-06);
            p = fma(p, s, 2.32027017785594476e-06);
            p = fma(p, s, 1.24321426461071173e-05);
            p = fma(p, s, -4.71752686387452903e-05);
            p = fma(p, s, 6.82939857338359650e-05);
            p = fma(p, s, 2.40106755546017631e-05);
            p = fma(p, s, -3.55038642504593052e-04);
            p = fma(p, s, 9.53291036442542551e-04);
            p = fma(p, s, -1.68827548262202496e-03);
            p = fma(p, s, 2.49144202910259443e-03);
            p = fma(p, s, -3.75120850970072283e-03);
            p = fma(p, s, 5.37091455460582835e-03);
            p = fma(p, s, 1.00525896769417677e+00);
            p = fma(p, s, 3.08388561049221810e+00);
        } else {
            s = s - 4.375;
            p = -1.69490261523055272e-05;
            p = fma(p, s, -5.01780379509045184e-07);
            p = fma(p, s, 7.29149613974081338e-06);
            p = fma(p, s, -1.49946089970476237e-07);
            p = fma(p, s, -2.71249703167381524e-07);
            p = fma(p, s, -1.78229809009254970e-06);
            p = fma(p, s, 3.17653889324899587e-06);
            p = fma(p, s, -5.88398956522093252e-06);
            p = fma(p, s, 1.51372614712039134e-05);
            p = fma(p, s, -5.07338716845259286e-05);
            p = fma(p, s, 1.76456923611614401e-04);
            p = fma(p, s, -5.11090819314451855e-04);
            p = fma(p, s, 5.03497607799785864e-04);
            p = fma(p, s, 1.01012955238797497e+00);
            p = fma(p, s, 4.21848070765291716e+00);
        }
    }
    double z = (1.41421356237309504880 * v) * p;
    return (bits >> 31) ? -z : z;
}

/* 24-bit integer -> the uniform k * 2^-24 in [0,1), exactly, in either precision.  On the device the double
 * version avoids the slow I2F.F64 unit: OR the integer into the mantissa of 2^52 and subtract 2^52. */
MBT_HD float mbt_u24_to_unit_f32(uint32_t k) { return (float)k * 5.9604644775390625e-08f; }
MBT_HD double mbt_u24_to_unit_f64(uint32_t k) { return (double)k * 5.9604644775390625e-08; }
/* the integer itself as a real (exact): the kernels compare  k < p * 2^24  instead of  k * 2^-24 < p  -- the same
 * predicate, since scaling by a power of two is exact -- with the 2^24 folded into p for free */
MBT_HD float mbt_u24_to_real_f32(uint32_t k) { return (float)k; }
MBT_HD double mbt_u24_to_real_f64(uint32_t k) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)k) - 4503599627370496.0; /* no I2F.F64 */
#else
    return (double)k;
#endif
}

/*
 * The fill decision  unif < exp(x)  (fill_probability_models.py:33,57-58) on the draw contract's 24-bit uniform k * 2^-24,
 * i.e.  k < exp(x) * 2^24.  The float version is the definition in float arithmetic.  The double version returns exactly
 * what  (double)k < mbt_exp2k_f64(x, 24)  returns, but decides almost every draw with the FLOAT exponential: for
 * -16 <= x <= 0 the float estimate pf of exp(x) * 2^24 is within 1.3e-6 (relative) of the double value --
 *     |(float)x - x| <= 16 * 2^-24 = 9.6e-7 absolute, so exp changes by 9.6e-7 relative;  mbt_exp2k_f32 is accurate to
 *     2 ulp = 2.4e-7 (tests/test_primitives.py);  the double exponential itself to 4 ulp = 9e-16
 * -- so k below pf * (1 - 1.6e-5) or above pf * (1 + 1.6e-5) (a 12x margin) is on the same side of the double threshold;
 * only the ~3e-5 of the draws that land inside the band evaluate the double exponential (saves ~2 x 30 float64
 * instructions per env-step; the decisions, hence all results, are unchanged -- every float64 fixture still matches).
 */
MBT_HD int mbt_u24_below_exp_f32(uint32_t k, float x) { return (float)k < mbt_exp2k_f32(x, 24); }
MBT_HD int mbt_u24_below_exp_f64(uint32_t k, double x) {
    if (x <= 0.0 && x >= -16.0) {
        const float pf = mbt_exp2k_f32((float)x, 24);
        const float kf = (float)k; /* exact: k < 2^24 */
        const float band = pf * 1.6e-5f;
        if (kf < pf - band) return 1;
        if (kf > pf + band) return 0;
    }
    return mbt_u24_to_real_f64(k) < mbt_exp2k_f64(x, 24);
}

/* x^p for the inventory penalties (reference: RewardFunctions.py:60-67,100-107,133-137,
 * `q ** inventory_exponent`).  numpy evaluates `** 2.0` as a square and `** 1.0` as the
 * identity (fast_scalar_power), so those two cases are exact; any other exponent goes
 * through exp(p*log|x|) with numpy's sign rules (negative base, non-integer p -> NaN). */
MBT_HD double mbt_pow_f64(double x, double p) {
    if (p == 2.0) return x * x;
    if (p == 1.0) return x;
    if (p == 0.0) return 1.0;
    if (x == 0.0) return (p > 0.0) ? 0.0 : (double)INFINITY;
    double ax = x < 0.0 ? -x : x;
    double r = mbt_exp_f64(p * mbt_log_f64(ax));
    if (x < 0.0) {
        double ip = (double)(int64_t)p;
        if (ip != p) return (double)NAN;
        return (((int64_t)p) & 1) ? -r : r;
    }
    return r;
}
MBT_HD float mbt_pow_f32(float x, float p) {
    if (p == 2.0f) return x * x;
    if (p == 1.0f) return x;
    if (p == 0.0f) return 1.0f;
    if (x == 0.0f) return (p > 0.0f) ? 0.0f : INFINITY;
    float ax = x < 0.0f ? -x : x;
    float r = mbt_exp_f32(p * mbt_log_f32(ax));
    if (x < 0.0f) {
        float ip = (float)(int32_t)p;
        if (ip != p) return NAN;
        return (((int32_t)p) & 1) ? -r : r;
    }
    return r;
}

/* ------------------------------------------------------------------ division by a uniform divisor */
/*
 * a / b, correctly rounded, from the correctly rounded reciprocal y = RN(1/b) that the HOST formed once (the divisor is
 * uniform over the batch: an observation-space half-width, TradingEnvironment.py:112-118).  Markstein's sequence
 *     q0 = RN(a*y);  r0 = RN(a - b*q0);  q1 = RN(q0 + r0*y)      q1 is a faithful rounding of a/b (error O(u^2) before rounding)
 *     r1 =    a - b*q1 (exact);          q2 = RN(q1 + r1*y)      = RN(a/b)   [Markstein 1990: y = RN(1/b), q1 faithful]
 * five multiply-add class operations instead of the ~15-instruction IEEE division sequence with its reciprocal seed and
 * slow path -- the normalised observation row has one division per column.  Outside the range where no intermediate can
 * overflow or lose bits to underflow (zeros, subnormal-ish, huge, inf, NaN numerators; y == 0 marks a divisor the host
 * rejected) the plain division is used, so the result equals `a / b` bit for bit for EVERY input
 * (tests/test_primitives.py compares 10^8 random and adversarial pairs).
 */
MBT_HD double mbt_div_rcp_f64(double a, double b, double y) {
    const double aa = a < 0.0 ? -a : a;
    if (!(y != 0.0 && aa >= 3.0549363634996047e-151 /* 2^-500 */ && aa <= 3.2733906078961419e+150 /* 2^500 */)) return a / b;
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    return fma(r, y, q);
}
MBT_HD float mbt_div_rcp_f32(float a, float b, float y) {
    const float aa = a < 0.0f ? -a : a;
    if (!(y != 0.0f && aa >= 8.67361737988403547e-19f /* 2^-60 */ && aa <= 1.15292150460684698e+18f /* 2^60 */)) return a / b;
    float q = a * y;
    float r = fmaf(-b, q, a);
    q = fmaf(r, y, q);
    r = fmaf(-b, q, a);
    return fmaf(r, y, q);
}
/* the reciprocal the host passes along: RN(1/b) when b is in the range the sequence above is proven for, else 0 */
MBT_HD double mbt_rcp_for_div_f64(double b) {
    const double ab = b < 0.0 ? -b : b;
    return (ab >= 3.0549363634996047e-151 && ab <= 3.2733906078961419e+150) ? 1.0 / b : 0.0;
}
MBT_HD float mbt_rcp_for_div_f32(float b) {
    const float ab = b < 0.0f ? -b : b;
    return (ab >= 8.67361737988403547e-19f && ab <= 1.15292150460684698e+18f) ? 1.0f / b : 0.0f;
}

#endif /* MBT_MATH_H */
