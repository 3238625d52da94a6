// Quantisers: float/double sample -> integer code, bit-exact with the numpy
// ufunc chains of baseband/base/encoding.py:63-128, :147-158,
// baseband/mark5b/payload.py:86-106 and baseband/gsb/payload.py:45-53.
//
// numpy evaluates each ufunc in the dtype of the input array with the Python
// float constants cast to that dtype (NEP 50), and rounds after every ufunc.
// So every constant exists in a float and a double flavour, and every
// arithmetic step goes through add_rn/mul_rn (never contracted to FMA).
//
// 2-bit: clip(v, -1.5s, 1.5s) + 2s, floor_divide by s.  numpy's float
// floor_divide is fmod based and returns the exact floor of the real quotient,
// which for a in [0.5s, 3.5s] equals the number of thresholds s, 2s, 3s that
// are <= a.  s and 2s are exact in the working type; 3s is not, so t3 is the
// smallest representable value >= 3s (computed on the host in extended
// precision, make_quant_consts()).
#pragma once
#include <math.h>
#include "bb_common.cuh"

namespace bb {

template <typename T>
struct QuantConsts {
    T clip_lo, clip_hi, two_sigma;   // encoding.py:59-60
    T t1, t2, t3;                    // exact thresholds for floor_divide
    T x1, x2, x3;                    // the same thresholds referred to v
    T four_bit_scale;                // encoding.py:47  (2.95)
    T eight_bit_scale;               // encoding.py:49  (35.5)
};

inline float next_toward(float x, float to) { return nextafterf(x, to); }
inline double next_toward(double x, double to) { return nextafter(x, to); }

template <typename T>
inline QuantConsts<T> make_quant_consts() {
    const double sigma = 2.174564;               // TWO_BIT_1_SIGMA
    QuantConsts<T> c;
    c.clip_lo = (T)(-1.5 * sigma);
    c.clip_hi = (T)(1.5 * sigma);
    c.two_sigma = (T)(2 * sigma);
    T s = (T)sigma;
    c.t1 = s;
    c.t2 = (T)2 * s;                              // exact
    long double exact3 = 3.0L * (long double)s;   // exact in 64-bit mantissa
    T t3 = (T)exact3;
    if ((long double)t3 < exact3) t3 = next_toward(t3, (T)INFINITY);
    c.t3 = t3;
    // x_k = smallest v with fl(clip(v) + 2s) >= t_k.  Rounding is monotonic,
    // so (a >= t_k) <=> (v >= x_k) for every v, including +-inf and values
    // beyond the clip range (x_k lies strictly inside it); NaN compares false
    // both ways.  This folds clip + add into the comparison constants.
    const T tk[3] = {c.t1, c.t2, c.t3};
    T xk[3];
    for (int k = 0; k < 3; ++k) {
        const T two = c.two_sigma, t = tk[k];
        auto reaches = [&](T y) {
            volatile T a = y + two;          // one IEEE rounding in type T
            return a >= t;
        };
        // Bisection on the (monotonic) predicate between the clip bounds:
        // !reaches(lo), reaches(hi); stop when hi is the successor of lo.
        T lo = c.clip_lo, hi = c.clip_hi;
        for (;;) {
            T mid = lo + (hi - lo) / 2;
            if (!(mid > lo && mid < hi)) {   // interval no longer splits
                T nx = next_toward(lo, hi);
                if (nx == hi) break;
                mid = nx;
            }
            if (reaches(mid)) hi = mid; else lo = mid;
        }
        xk[k] = hi;
    }
    c.x1 = xk[0]; c.x2 = xk[1]; c.x3 = xk[2];
    c.four_bit_scale = (T)2.95;
    c.eight_bit_scale = (T)35.5;
    return c;
}

BB_HD float rint_t(float x) {
#if defined(__CUDA_ARCH__)
    return rintf(x);
#else
    return nearbyintf(x);
#endif
}
BB_HD double rint_t(double x) {
#if defined(__CUDA_ARCH__)
    return rint(x);
#else
    return nearbyint(x);
#endif
}

// numpy clip = minimum(maximum(v, lo), hi), NaN propagating.
template <typename T>
BB_HD T clip_nan(T v, T lo, T hi) {
    T a = v < lo ? lo : v;
    return a > hi ? hi : a;
}

template <typename T>
BB_HD uint32_t to_code(T x) {          // astype(uint8) of an in-range value
    return (x != x) ? 0u : (uint32_t)(int32_t)x;
}

// encoding.py:63-74
template <typename T>
BB_HD uint32_t quant1_offset(T v) { return v >= (T)0 ? 1u : 0u; }

// mark5b/payload.py:88-89 (np.signbit)
BB_HD uint32_t quant1_signbit(float v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__float_as_uint(v) >> 31;
#else
    return signbit(v) ? 1u : 0u;
#endif
}
BB_HD uint32_t quant1_signbit(double v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__double2hiint(v)) >> 31;
#else
    return signbit(v) ? 1u : 0u;
#endif
}

// encoding.py:77-102
template <typename T>
BB_HD uint32_t quant2_offset(T v, const QuantConsts<T> &c) {
    // == floor_divide(clip(v, -1.5s, 1.5s) + 2s, s); see make_quant_consts.
    return (uint32_t)(v >= c.x1) + (uint32_t)(v >= c.x2)
        + (uint32_t)(v >= c.x3);
}

// The literal ufunc chain, kept to cross-check the folded thresholds.
template <typename T>
BB_HD uint32_t quant2_offset_chain(T v, const QuantConsts<T> &c) {
    T a = add_rn(clip_nan(v, c.clip_lo, c.clip_hi), c.two_sigma);
    return (uint32_t)(a >= c.t1) + (uint32_t)(a >= c.t2)
        + (uint32_t)(a >= c.t3);
}

// encoding.py:105-128
template <typename T>
BB_HD uint32_t quant4_offset(T v, const QuantConsts<T> &c) {
    T x = add_rn(mul_rn(v, c.four_bit_scale), (T)8.5);
    return to_code(clip_nan(x, (T)0, (T)15));
}

// encoding.py:147-158
template <typename T>
BB_HD uint32_t quant8_offset(T v, const QuantConsts<T> &c);

// guppi/payload.py:17-18, dada/payload.py:17-18, gsb/payload.py:45-53
// = clip(rint(v), lo, hi) cast to int (rint = round half to even).  Clipping
// to the integer bounds first gives the same result (both steps are monotonic
// and the bounds are integers), and then the rounding is one addition: adding
// 1.5 * 2^23 (2^52 for double) to a value in [-128, 127] leaves the
// round-half-even integer in the low mantissa bits, in two's complement.
// This avoids the FRND + F2I pair, which both issue on the quarter-rate
// conversion pipe (the int8 encoders were limited by it).
BB_HD uint32_t magic_round_bits(float y) {
    const float t = add_rn(y, 12582912.0f);
#if defined(__CUDA_ARCH__)
    return __float_as_uint(t);
#else
    uint32_t u;
    __builtin_memcpy(&u, &t, 4);
    return u;
#endif
}
BB_HD uint32_t magic_round_bits(double y) {
    const double t = add_rn(y, 6755399441055744.0);
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(t);
#else
    unsigned long long u;
    __builtin_memcpy(&u, &t, 8);
    return (uint32_t)u;
#endif
}

template <typename T, int BPS>
BB_HD uint32_t quant_sint(T v) {
    const T lo = (T)(-(1 << (BPS - 1))), hi = (T)((1 << (BPS - 1)) - 1);
    const T y = clip_nan(v, lo, hi);
    const uint32_t bits = (y != y) ? 0u : magic_round_bits(y);
    return bits & ((1u << BPS) - 1u);
}

// encoding.py:147-158: clip(rint(v * 35.5 + 127.5), 0, 255), rounded the same
// way (clip to the integer bounds, then one magic addition).
template <typename T>
BB_HD uint32_t quant8_offset(T v, const QuantConsts<T> &c) {
    const T y = clip_nan(add_rn(mul_rn(v, c.eight_bit_scale), (T)127.5),
                         (T)0, (T)255);
    return (y != y) ? 0u : (magic_round_bits(y) & 0xffu);
}

enum { QUANT_OFFSET = 0, QUANT_MARK5B = 1, QUANT_SINT = 2 };

template <typename T, int BPS, int QUANT>
BB_HD uint32_t quantise(T v, const QuantConsts<T> &c) {
    if (QUANT == QUANT_SINT) {
        return quant_sint<T, (BPS >= 2 ? BPS : 2)>(v);
    } else if (BPS == 1) {
        return QUANT == QUANT_MARK5B ? quant1_signbit(v) : quant1_offset(v);
    } else if (BPS == 2) {
        uint32_t q = quant2_offset(v, c);
        if (QUANT == QUANT_MARK5B)            // swap codes 1 <-> 2
            q = (0xD8u >> (2u * q)) & 3u;
        return q;
    } else if (BPS == 4) {
        return quant4_offset(v, c);
    } else {
        return quant8_offset(v, c);
    }
}

}  // namespace bb
