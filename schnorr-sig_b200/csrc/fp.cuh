// Goldilocks base field Fp, p = 2^64 - 2^32 + 1 (reference README.md:4; `cheetah::Fp`, call sites
// src/signature.rs:278-298).  Elements are canonical u64 values (the reference's raw limbs are
// canonical, not Montgomery -- SURVEY.md App. A).
//
// Device code is written on 32-bit halves: every 32x32->64 product is one IMAD.WIDE on the
// integer-multiply pipe; the special-form reduction (2^64 = 2^32-1, 2^96 = -1 mod p) is
// multiplier-free and runs on the ALU pipe.  Wide sums are kept as three 96-bit "column"
// accumulators (weights 2^0, 2^32, 2^64) so that each product costs one IMAD.WIDE with carry-out
// plus one carry-absorbing add, and only one reduction is paid per output coefficient.
//
// The same header compiles on the host (g++) with portable fallbacks for the PTX blocks; that
// build is used ONLY by tests/hostsim to unit-test the formulas without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SB_DEV __device__ __forceinline__
#define SB_DEV_NOINLINE __device__ __noinline__
#else
#define SB_DEV static inline __attribute__((always_inline))
#define SB_DEV_NOINLINE static __attribute__((noinline))
#endif

namespace sb {

typedef uint64_t fp_t;
static constexpr uint64_t FP_P = 0xffffffff00000001ULL;
static constexpr uint64_t FP_EPS = 0xffffffffULL;  // 2^64 mod p

SB_DEV fp_t fp_add(fp_t a, fp_t b) {
    uint64_t s = a + b;
    bool over = (s < a) | (s >= FP_P);
    return over ? s + FP_EPS : s;  // s - p == s + 2^32 - 1 (mod 2^64)
}
SB_DEV fp_t fp_sub(fp_t a, fp_t b) {
    uint64_t d = a - b;
    return (a < b) ? d - FP_EPS : d;  // d + p == d - (2^32 - 1) (mod 2^64)
}
SB_DEV fp_t fp_neg(fp_t a) { return a ? FP_P - a : 0; }
SB_DEV fp_t fp_dbl(fp_t a) { return fp_add(a, a); }

// 96-bit column accumulator
struct acc96 {
    uint32_t w0, w1, w2;
};

// acc += a * b   (a, b 32-bit; acc 96-bit)
SB_DEV void mac96(acc96& c, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c.w0), "+r"(c.w1), "+r"(c.w2)
        : "r"(a), "r"(b));
#else
    unsigned __int128 t = ((unsigned __int128)c.w2 << 64) | ((uint64_t)c.w1 << 32) | c.w0;
    t += (uint64_t)a * b;
    c.w0 = (uint32_t)t;
    c.w1 = (uint32_t)(t >> 32);
    c.w2 = (uint32_t)(t >> 64);
#endif
}

// Three columns of a sum of 64x64 products: value = c0 + c1*2^32 + c2*2^64
struct wide_acc {
    acc96 c0, c1, c2;
};
SB_DEV void wide_zero(wide_acc& w) {
    w.c0 = {0, 0, 0};
    w.c1 = {0, 0, 0};
    w.c2 = {0, 0, 0};
}
// w += a * b, a and b any 64-bit values
SB_DEV void wide_mac(wide_acc& w, uint64_t a, uint64_t b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    mac96(w.c0, a0, b0);
    mac96(w.c1, a0, b1);
    mac96(w.c1, a1, b0);
    mac96(w.c2, a1, b1);
}
// w += a * a (3 products; the cross term is doubled by adding it twice into the column)
SB_DEV void wide_mac_sqr(wide_acc& w, uint64_t a) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    mac96(w.c0, a0, a0);
    acc96 x = {0, 0, 0};
    mac96(x, a0, a1);
    // c1 += 2*x   (x < 2^64 so 2x < 2^65)
    uint64_t xl = ((uint64_t)x.w1 << 32) | x.w0;
    uint64_t c1l = ((uint64_t)w.c1.w1 << 32) | w.c1.w0;
    uint64_t s1 = c1l + xl;
    uint32_t k1 = s1 < xl;
    uint64_t s2 = s1 + xl;
    uint32_t k2 = s2 < xl;
    w.c1.w0 = (uint32_t)s2;
    w.c1.w1 = (uint32_t)(s2 >> 32);
    w.c1.w2 += k1 + k2;
    mac96(w.c2, a1, a1);
}
// double every column (used for the cross terms of a squaring); columns stay < 2^96 as long as
// the un-doubled sums are < 2^95, which holds for <= 2^30 accumulated products.
SB_DEV void wide_double(wide_acc& w) {
    w.c0.w2 = (w.c0.w2 << 1) | (w.c0.w1 >> 31);
    w.c0.w1 = (w.c0.w1 << 1) | (w.c0.w0 >> 31);
    w.c0.w0 <<= 1;
    w.c1.w2 = (w.c1.w2 << 1) | (w.c1.w1 >> 31);
    w.c1.w1 = (w.c1.w1 << 1) | (w.c1.w0 >> 31);
    w.c1.w0 <<= 1;
    w.c2.w2 = (w.c2.w2 << 1) | (w.c2.w1 >> 31);
    w.c2.w1 = (w.c2.w1 << 1) | (w.c2.w0 >> 31);
    w.c2.w0 <<= 1;
}

// x = x0 + x1*2^32 + x2*2^64 + x3*2^96 + x4*2^128  ->  canonical x mod p
// using 2^64 = 2^32-1, 2^96 = -1, 2^128 = -2^32 (mod p):
//   x = (x0 + x1*2^32) + x2*(2^32-1) - x3 - x4*2^32
// x4 must be < 2^32 (always: it holds at most ~8 bits here).
SB_DEV fp_t fp_reduce160(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4) {
    uint64_t lo = ((uint64_t)x1 << 32) | x0;
    // subtract x3 + x4*2^32 (a 64-bit value) -> may borrow once
    uint64_t sub = ((uint64_t)x4 << 32) | x3;
    uint64_t t = lo - sub;
    if (lo < sub) t -= FP_EPS;  // borrowed 2^64 == EPS too much; no second borrow: t >= 2^64 - sub' ...
    // the correction can itself wrap only if t < EPS after a borrow, i.e. lo - sub + 2^64 < 2^32-1,
    // impossible when sub < 2^64 - 2^32 + 1; sub's top word x4 is tiny so this always holds.
    uint64_t m = ((uint64_t)x2 << 32) - x2;  // x2 * (2^32 - 1) < 2^64
    uint64_t r = t + m;
    if (r < m) r += FP_EPS;  // overflowed 2^64 == EPS; r + EPS cannot overflow again (r <= 2^64 - 2^33)
    if (r >= FP_P) r -= FP_P;
    return r;
}

SB_DEV fp_t wide_reduce(const wide_acc& w) {
    uint32_t x0 = w.c0.w0, x1, x2, x3, x4;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %4, %5;\n\t"
        "addc.cc.u32 %1, %6, %7;\n\t"
        "addc.cc.u32 %2, %8, 0;\n\t"
        "addc.u32 %3, 0, 0;\n\t"
        "add.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.u32 %3, %3, %11;"
        : "=&r"(x1), "=&r"(x2), "=&r"(x3), "=&r"(x4)
        : "r"(w.c0.w1), "r"(w.c1.w0), "r"(w.c0.w2), "r"(w.c1.w1), "r"(w.c1.w2), "r"(w.c2.w0), "r"(w.c2.w1),
          "r"(w.c2.w2));
#else
    unsigned __int128 hi = (unsigned __int128)w.c0.w1 + w.c1.w0;                       // weight 2^32
    hi += ((unsigned __int128)w.c0.w2 + w.c1.w1 + w.c2.w0) << 32;                      // weight 2^64
    hi += ((unsigned __int128)w.c1.w2 + w.c2.w1) << 64;                                // weight 2^96
    hi += ((unsigned __int128)w.c2.w2) << 96;                                          // weight 2^128
    x1 = (uint32_t)hi;
    x2 = (uint32_t)(hi >> 32);
    x3 = (uint32_t)(hi >> 64);
    x4 = (uint32_t)(hi >> 96);
#endif
    return fp_reduce160(x0, x1, x2, x3, x4);
}

SB_DEV fp_t fp_mul(fp_t a, fp_t b) {
    wide_acc w;
    wide_zero(w);
    wide_mac(w, a, b);
    return wide_reduce(w);
}
SB_DEV fp_t fp_sqr(fp_t a) {
    wide_acc w;
    wide_zero(w);
    wide_mac_sqr(w, a);
    return wide_reduce(w);
}
// a * k for a small constant k < 2^32
SB_DEV fp_t fp_mul_small(fp_t a, uint32_t k) {
    uint64_t lo = (uint64_t)(uint32_t)a * k;
    uint64_t hi = (a >> 32) * k;  // weight 2^32
    uint64_t s = lo + (hi << 32);
    uint32_t c = s < lo;
    uint32_t x2 = (uint32_t)(hi >> 32) + c;  // weight 2^64, < 2^32
    return fp_reduce160((uint32_t)s, (uint32_t)(s >> 32), x2, 0, 0);
}
SB_DEV fp_t fp_mul7(fp_t a) { return fp_mul_small(a, 7); }

// a^(2^n)
SB_DEV fp_t fp_sqr_n(fp_t a, int n) {
    for (int i = 0; i < n; i++) a = fp_sqr(a);
    return a;
}
// a^-1 = a^(p-2), p-2 = 2^64 - 2^32 - 1 = (2^32-1)*2^32 + (2^32 - 1): 63 squarings + 7+... multiplications
SB_DEV fp_t fp_inv(fp_t a) {
    // t_k = a^(2^k - 1)
    fp_t t1 = a;
    fp_t t2 = fp_mul(fp_sqr(t1), t1);
    fp_t t4 = fp_mul(fp_sqr_n(t2, 2), t2);
    fp_t t8 = fp_mul(fp_sqr_n(t4, 4), t4);
    fp_t t16 = fp_mul(fp_sqr_n(t8, 8), t8);
    fp_t t32 = fp_mul(fp_sqr_n(t16, 16), t16);
    // p - 2 = (2^32 - 1) * 2^32 + (2^32 - 2) + 1 ... write p-2 = 0xffffffff_00000000 - 1 + ... :
    // p - 2 = 0xfffffffeffffffff = (2^31 - 1) << 33 | 0 << 32 | (2^32 - 1)
    fp_t t31 = fp_mul(fp_sqr_n(fp_mul(fp_sqr_n(fp_mul(fp_sqr_n(t16, 8), t8), 4), t4), 2), t2);  // 2^30-1
    t31 = fp_mul(fp_sqr(t31), a);                                                                // 2^31-1
    fp_t r = fp_sqr_n(t31, 33);   // (2^31-1) << 33
    r = fp_mul(r, t32);           // | (2^32 - 1)
    return r;
}

}  // namespace sb
