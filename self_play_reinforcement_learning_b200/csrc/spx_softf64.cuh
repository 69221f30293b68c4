// spx_softf64.cuh -- IEEE-754 binary64 arithmetic (round to nearest even) on the INTEGER pipe.
//
// Why: on B200 any FP64 instruction (DADD / DMUL / DFMA, and with them the DDIV / DSQRT sequences) that a warp executes while the
// SM's tensor pipe is running tcgen05.mma slows the MMAs down -- a single warp doing the search's fp64 PUCT arithmetic next to the
// network costs the SM pair 5-7 % of its MMA issue rate (measured: scripts/dbg_layer_trace_fused.py; integer
// arithmetic and L2 loads next to the MMAs cost nothing).  The fused tick kernel's shadow warp, which searches WHILE the network
// runs, therefore does its PUCT arithmetic with the functions below: same bits as __dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn
// (tests/test_softf64_gpu.py holds them against the hardware on 10^8 operand pairs; every parity test of the fused kernel runs
// through them), no FP64 instruction on the common path.  Anything unusual (subnormal or non-finite operands, results outside
// the normal range) takes the hardware instruction -- exact either way, and rare enough not to matter.
#pragma once
#include <cuda_runtime.h>

namespace spx {
namespace sf {

typedef unsigned long long u64;
constexpr u64 SIGN = 0x8000000000000000ULL, MASK52 = 0x000FFFFFFFFFFFFFULL, HID = 0x0010000000000000ULL;

// Bit casts the compiler cannot see through: with __double_as_longlong it recognises `(bits << 1) == 0` as `x != 0.0`, `bits & ~SIGN`
// as fabs(x) ... and emits DSETP / DADD for them, and it hoists the single-instruction hardware fallbacks above the branches that
// guard them -- FP64 instructions on the very path that must not have any (found in the SASS; ncu: smsp__inst_executed_pipe_fp64).
__device__ __forceinline__ u64 bits(double x) { u64 b; asm("mov.b64 %0, %1;" : "=l"(b) : "d"(x)); return b; }
__device__ __forceinline__ double dbl(u64 b) { double x; asm("mov.b64 %0, %1;" : "=d"(x) : "l"(b)); return x; }
// the rare cases (subnormal / non-finite operands or results) take the FP64 instruction, out of line
static __device__ __noinline__ double hw_mul(double a, double b) { return __dmul_rn(a, b); }
static __device__ __noinline__ double hw_add(double a, double b) { return __dadd_rn(a, b); }
static __device__ __noinline__ double hw_div_u32(double a, unsigned k) { return __ddiv_rn(a, (double)k); }
static __device__ __noinline__ double hw_sqrt_u32(unsigned m) { return __dsqrt_rn((double)m); }
static __device__ __noinline__ double hw_from_f32(float f) { return (double)f; }
__device__ __forceinline__ bool is_zero(u64 b) { return (b << 1) == 0; }
__device__ __forceinline__ int expo(u64 b) { return (int)(b >> 52) & 0x7FF; }
__device__ __forceinline__ bool normal(u64 b) { const int e = expo(b); return e != 0 && e != 0x7FF; }
__device__ __forceinline__ bool finite(u64 b) { return expo(b) != 0x7FF; }

// value = (m / 2^63) * 2^(e - 1023) with bit 63 of m set, `sticky` = something non-zero below bit 0 of m -> nearest even double
__device__ __forceinline__ double round_pack(u64 sign, int e, u64 m, bool sticky, bool& ok) {
    u64 mant = m >> 11;   // 53 bits
    const bool rnd = (m >> 10) & 1ULL;
    const bool st = ((m & 0x3FFULL) != 0) | sticky;
    if (rnd && (st || (mant & 1ULL))) {
        mant += 1;
        if (mant >> 53) { mant >>= 1; e += 1; }
    }
    ok = e >= 1 && e <= 2046;
    return dbl(sign | ((u64)(unsigned)e << 52) | (mant & MASK52));
}

__device__ __forceinline__ double mul(double a, double b) {
    const u64 x = bits(a), y = bits(b);
    if (normal(x) && normal(y)) {
        const u64 A = ((x & MASK52) | HID) << 11, B = ((y & MASK52) | HID) << 11;   // [2^63, 2^64)
        u64 hi = __umul64hi(A, B), lo = A * B;
        int e = expo(x) + expo(y) - 1023;
        if (hi >> 63) e += 1;
        else { hi = (hi << 1) | (lo >> 63); lo <<= 1; }
        bool ok;
        const double r = round_pack((x ^ y) & SIGN, e, hi, lo != 0, ok);
        if (ok) return r;
    } else if (finite(x) && finite(y) && (is_zero(x) || is_zero(y))) return dbl((x ^ y) & SIGN);
    return hw_mul(a, b);
}

__device__ __forceinline__ double add(double a, double b) {
    u64 x = bits(a), y = bits(b);
    if (normal(x) && normal(y)) {
        if ((x & ~SIGN) < (y & ~SIGN)) { const u64 t = x; x = y; y = t; }   // |x| >= |y|
        const int ex = expo(x), d = ex - expo(y);
        const u64 X = ((x & MASK52) | HID) << 10;                             // leading bit 62: room for the carry
        u64 Y = ((y & MASK52) | HID) << 10;
        bool sticky = false;
        if (d >= 64) { sticky = true; Y = 0; }
        else if (d > 0) { sticky = (Y << (64 - d)) != 0; Y >>= d; }
        u64 S;
        if (((x ^ y) >> 63) == 0) S = X + Y;
        else S = X - Y - (sticky ? 1ULL : 0ULL);                               // exact value = S + (a fraction in (0, 1)) when sticky
        if (S == 0) return dbl(0ULL);                                          // exact cancellation: +0 in round-to-nearest
        const int lz = __clzll((long long)S);
        S <<= lz;
        bool ok;
        const double r = round_pack(x & SIGN, ex + 1 - lz, S, sticky, ok);
        if (ok) return r;
        return hw_add(a, b);
    }
    if (finite(x) && finite(y)) {
        if (is_zero(x) && is_zero(y)) return dbl(x & y & SIGN);               // -0 only for (-0) + (-0)
        if (is_zero(x) && normal(y)) return b;
        if (is_zero(y) && normal(x)) return a;
    }
    return hw_add(a, b);
}

// a / k for an integer 1 <= k < 2^31 (the PUCT divisors are visit counts): == __ddiv_rn(a, (double)k)
__device__ __forceinline__ double div_u32(double a, unsigned k) {
    const u64 x = bits(a);
    if (is_zero(x) && k) return a;
    if (normal(x) && k >= 1u && k < 0x80000000u) {
        const u64 M = (x & MASK52) | HID;                      // [2^52, 2^53)
        const int j = 31 - __clz((int)k), s = j + 3;           // k in [2^j, 2^(j+1))
        const u64 kk = k, inv = 0xFFFFFFFFFFFFFFFFULL / kk;    // integer division (no FP64 inside)
        u64 qh = __umul64hi(M, inv), r = M - qh * kk;          // floor(M / k): the estimate is at most 2 short
        while (r >= kk) { qh += 1; r -= kk; }
        const u64 R = r << s;                                   // r < k < 2^(j+1): R < 2^(2j+4) <= 2^64
        u64 ql = __umul64hi(R, inv), r2 = R - ql * kk;
        while (r2 >= kk) { ql += 1; r2 -= kk; }
        u64 Q = (qh << s) | ql;                                 // floor(M 2^s / k), 55 or 56 bits
        const int lz = __clzll((long long)Q);
        Q <<= lz;
        bool ok;
        const double res = round_pack(x & SIGN, expo(x) - 52 - s + 63 - lz, Q, r2 != 0, ok);
        if (ok) return res;
    }
    return hw_div_u32(a, k);
}

// sqrt(m) for an integer m: a table of correctly rounded roots (filled with __dsqrt_rn when the engine is created)
constexpr unsigned SQRT_TABLE = 1u << 16;
__device__ __forceinline__ double sqrt_u32(unsigned m, const double* __restrict__ table) {
    if (m < SQRT_TABLE) return __ldg(table + m);
    return hw_sqrt_u32(m);
}

// (double)f, exact
__device__ __forceinline__ double from_f32(float f) {
    const unsigned b = __float_as_uint(f), e = (b >> 23) & 0xFFu, m = b & 0x7FFFFFu;
    if (e == 0u && m == 0u) return dbl((u64)(b >> 31) << 63);
    if (e != 0u && e != 0xFFu) return dbl(((u64)(b >> 31) << 63) | ((u64)(e + 896u) << 52) | ((u64)m << 29));
    return hw_from_f32(f);
}

// k * 2^-53 for k < 2^53 (the counter stream's uniform), exact
__device__ __forceinline__ double from_u53(u64 k) {
    if (k == 0) return dbl(0ULL);
    const int lz = __clzll((long long)k);   // >= 11
    return dbl(((u64)(unsigned)(1033 - lz) << 52) | ((k << (lz - 11)) & MASK52));
}

// a * 2^p, exact while the result stays normal
__device__ __forceinline__ double mul_pow2(double a, int p) {
    const u64 x = bits(a);
    if (is_zero(x)) return a;
    const int e = expo(x) + p;
    if (normal(x) && e >= 1 && e <= 2046) return dbl((x & ~(0x7FFULL << 52)) | ((u64)(unsigned)e << 52));
    return hw_mul(a, p == 2 ? 4.0 : 0.25);   // (only p = 2 and p = -2 are used)
}

__device__ __forceinline__ double neg_if(double a, bool neg) { return dbl(bits(a) ^ (neg ? SIGN : 0ULL)); }   // == a * (+1 | -1)

// total order of finite doubles and infinities as unsigned keys (-0 == +0); NaN does not occur
__device__ __forceinline__ u64 key(double a) {
    u64 b = bits(a);
    if (is_zero(b)) b = 0;
    return (b >> 63) ? ~b : (b | SIGN);
}
__device__ __forceinline__ bool gt(double a, double b) { return key(a) > key(b); }
__device__ __forceinline__ bool eq(double a, double b) { return key(a) == key(b); }

}  // namespace sf

// The arithmetic of the search, hardware (FP64 pipe) or integer pipe, behind one interface.
template <bool SOFT> struct FP;
template <> struct FP<false> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double div_int(double a, int k) { return __ddiv_rn(a, (double)k); }
    static __device__ __forceinline__ double sqrt_int(int m, const double*) { return __dsqrt_rn((double)m); }
    static __device__ __forceinline__ double from_f32(float f) { return (double)f; }
    static __device__ __forceinline__ double from_u53(unsigned long long k) { return (double)k * (1.0 / 9007199254740992.0); }
    static __device__ __forceinline__ double times4(double a) { return __dmul_rn(4.0, a); }
    static __device__ __forceinline__ double quarter(double a) { return __dmul_rn(a, 0.25); }
    static __device__ __forceinline__ double signed_by(double a, int player) { return __dmul_rn((double)player, a); }
    static __device__ __forceinline__ bool gt(double a, double b) { return a > b; }
    static __device__ __forceinline__ bool eq(double a, double b) { return a == b; }
};
template <> struct FP<true> {
    static __device__ __forceinline__ double mul(double a, double b) { return sf::mul(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return sf::add(a, b); }
    static __device__ __forceinline__ double div_int(double a, int k) { return sf::div_u32(a, (unsigned)k); }
    static __device__ __forceinline__ double sqrt_int(int m, const double* table) { return sf::sqrt_u32((unsigned)m, table); }
    static __device__ __forceinline__ double from_f32(float f) { return sf::from_f32(f); }
    static __device__ __forceinline__ double from_u53(unsigned long long k) { return sf::from_u53(k); }
    static __device__ __forceinline__ double times4(double a) { return sf::mul_pow2(a, 2); }
    static __device__ __forceinline__ double quarter(double a) { return sf::mul_pow2(a, -2); }
    static __device__ __forceinline__ double signed_by(double a, int player) { return sf::neg_if(a, player < 0); }
    static __device__ __forceinline__ bool gt(double a, double b) { return sf::gt(a, b); }
    static __device__ __forceinline__ bool eq(double a, double b) { return sf::eq(a, b); }
};

}  // namespace spx
