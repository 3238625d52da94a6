// Shared helpers for the baseband_b200 kernels.
//
// The per-thread bodies of the streaming kernels are written as
// host/device inline functions over plain structs so that the index
// arithmetic can also be compiled by g++ for the CPU emulation used by
// tests/test_emulation.py (test infrastructure; the shipped library only
// contains the CUDA instantiations).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

namespace bb {

struct alignas(16) F4 { float x, y, z, w; };
struct alignas(16) D2 { double x, y; };
struct alignas(8) F2 { float x, y; };
struct alignas(16) U4 { uint32_t x, y, z, w; };

BB_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// PRMT: byte i of the result is byte (s >> 4i) & 7 of the pair {x, y}
// (selector nibbles 0..7 only: no sign replication).
BB_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, s);
#else
    const uint64_t xy = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i)
        r |= (uint32_t)((xy >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}

// Division of any 32-bit unsigned n by a runtime-constant d >= 1
// (Granlund & Montgomery round-up method): 4 integer instructions.
struct FastDiv {
    uint32_t d, mul, sh1, sh2;
    BB_HD uint32_t div(uint32_t n) const {
        uint32_t t = umulhi32(mul, n);
        return (t + ((n - t) >> sh1)) >> sh2;
    }
    BB_HD void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
        q = div(n);
        r = n - q * d;
    }
};

inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                 // ceil(log2 d)
    f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
    f.sh1 = l < 1 ? l : 1;
    f.sh2 = l > 1 ? l - 1 : 0;
    return f;
}

inline int ilog2_exact(uint32_t v) {            // -1 if not a power of two
    if (v == 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1u << l) < v) ++l;
    return l;
}

// Arithmetic that must round exactly once per numpy ufunc (no FMA fusion).
BB_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
BB_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
// Fused multiply-add, one rounding (explicitly wanted: not an accidental
// contraction, which the build disables with -fmad=false).
BB_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
BB_HD float uint_as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    __builtin_memcpy(&f, &u, 4);
    return f;
#endif
}
// Small signed integer (|i| < 2^22) -> float without the conversion pipe: i is
// added to the mantissa of 1.5 * 2^23, whose ulp is 1, and the offset removed
// by one exact subtraction.
BB_HD float small_int_to_float(int32_t i) {
    return add_rn(uint_as_float(0x4B400000u + (uint32_t)i), -12582912.0f);
}
BB_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
BB_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}

}  // namespace bb
