// fp.cuh — BLS12-381 base field, 12 x 32-bit little-endian limbs, Montgomery form (R = 2^384).
//
// Memory layout is byte-identical to the reference's blst_fp (6 x u64 LE limbs, Montgomery form;
// /root/reference/vendor/blst/bindings/blst.h:63, semantics vendor/blst/src/no_asm.h:29-82,
// :104-170), so SignatureSet buffers are consumed without any conversion.
//
// Device code: hand-written PTX carry chains.  The multiplier keeps two accumulators (even / odd
// limb alignment) so that every 32x32->64 partial product is a single IMAD.WIDE.U32 with carry
// (mad.lo.cc + madc.hi.cc on an aligned register pair), interleaved with the Montgomery reduction
// row by row: 288 IMAD.WIDE + 12 IMAD per multiplication.
//
// The same file compiles as plain C++ (no __CUDACC__) for tests/hostsim, where the PTX bodies are
// replaced by portable 64-bit arithmetic.  That build is TEST INFRASTRUCTURE: the shipped library
// contains device code only and has no CPU execution path.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BLS_FN __device__ __forceinline__
#define BLS_HD __host__ __device__ __forceinline__
#define BLS_NOINLINE __device__ __noinline__
#define BLS_TABLE __device__ __constant__ const
#else
#define BLS_FN static inline
#define BLS_HD static inline
#define BLS_NOINLINE static
#define BLS_TABLE static const
#endif

namespace bls {

// 16-byte alignment: every field element moves as three 128-bit accesses (local, shared, global).  All containing
// layouts keep their reference sizes and offsets (48 = 3 x 16; SignatureSet: pk @0, msg @96, sig @128).
struct alignas(16) fp { uint32_t l[12]; };

// p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
// (vendor/blst/src/consts.c:10-14).  A switch so that unrolled loops fold the limbs into immediates.
BLS_FN constexpr uint32_t P32(int i) {
    return i == 0 ? 0xffffaaabu : i == 1 ? 0xb9feffffu : i == 2 ? 0xb153ffffu : i == 3 ? 0x1eabfffeu
         : i == 4 ? 0xf6b0f624u : i == 5 ? 0x6730d2a0u : i == 6 ? 0xf38512bfu : i == 7 ? 0x64774b84u
         : i == 8 ? 0x434bacd7u : i == 9 ? 0x4b1ba7b6u : i == 10 ? 0x397fe69au : 0x1a0111eau;
}
// -p^-1 mod 2^32 (low half of vendor/blst/src/consts.h:12)
#define BLS_N0 0xfffcfffdu

BLS_FN void fp_set_zero(fp &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
}

BLS_FN bool fp_is_zero(const fp &a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= a.l[i];
    return t == 0;
}

BLS_FN bool fp_eq(const fp &a, const fp &b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}

// r = c ? a : b
BLS_FN void fp_select(fp &r, bool c, const fp &a, const fp &b) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = c ? a.l[i] : b.l[i];
}

#ifdef __CUDA_ARCH__
// ------------------------------------------------------------------------------------------
// PTX building blocks
#define BLS_R12(x) "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), \
                   "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11])

// In-place forms only: instruction i reads and writes operand i, so no early-clobber hazards.
// acc += b (12 limbs; the caller guarantees no carry out)
BLS_FN void add12(uint32_t *acc, const uint32_t *b) {
    asm("add.cc.u32 %0,%0,%12;\n\taddc.cc.u32 %1,%1,%13;\n\taddc.cc.u32 %2,%2,%14;\n\t"
        "addc.cc.u32 %3,%3,%15;\n\taddc.cc.u32 %4,%4,%16;\n\taddc.cc.u32 %5,%5,%17;\n\t"
        "addc.cc.u32 %6,%6,%18;\n\taddc.cc.u32 %7,%7,%19;\n\taddc.cc.u32 %8,%8,%20;\n\t"
        "addc.cc.u32 %9,%9,%21;\n\taddc.cc.u32 %10,%10,%22;\n\taddc.u32 %11,%11,%23;"
        : BLS_R12(acc)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]),
          "r"(b[9]), "r"(b[10]), "r"(b[11]));
}

// acc -= b; returns the borrow mask (0xffffffff when acc < b)
BLS_FN uint32_t sub12(uint32_t *acc, const uint32_t *b) {
    uint32_t bw;
    asm("sub.cc.u32 %0,%0,%13;\n\tsubc.cc.u32 %1,%1,%14;\n\tsubc.cc.u32 %2,%2,%15;\n\t"
        "subc.cc.u32 %3,%3,%16;\n\tsubc.cc.u32 %4,%4,%17;\n\tsubc.cc.u32 %5,%5,%18;\n\t"
        "subc.cc.u32 %6,%6,%19;\n\tsubc.cc.u32 %7,%7,%20;\n\tsubc.cc.u32 %8,%8,%21;\n\t"
        "subc.cc.u32 %9,%9,%22;\n\tsubc.cc.u32 %10,%10,%23;\n\tsubc.cc.u32 %11,%11,%24;\n\t"
        "subc.u32 %12,0,0;"
        : BLS_R12(acc), "=r"(bw)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]),
          "r"(b[9]), "r"(b[10]), "r"(b[11]));
    return bw;
}

// acc -= p (limbs as immediates); returns the borrow mask
BLS_FN uint32_t sub_p12(uint32_t *acc) {
    uint32_t bw;
    asm("sub.cc.u32 %0,%0,0xffffaaab;\n\tsubc.cc.u32 %1,%1,0xb9feffff;\n\tsubc.cc.u32 %2,%2,0xb153ffff;\n\t"
        "subc.cc.u32 %3,%3,0x1eabfffe;\n\tsubc.cc.u32 %4,%4,0xf6b0f624;\n\tsubc.cc.u32 %5,%5,0x6730d2a0;\n\t"
        "subc.cc.u32 %6,%6,0xf38512bf;\n\tsubc.cc.u32 %7,%7,0x64774b84;\n\tsubc.cc.u32 %8,%8,0x434bacd7;\n\t"
        "subc.cc.u32 %9,%9,0x4b1ba7b6;\n\tsubc.cc.u32 %10,%10,0x397fe69a;\n\tsubc.cc.u32 %11,%11,0x1a0111ea;\n\t"
        "subc.u32 %12,0,0;"
        : BLS_R12(acc), "=r"(bw));
    return bw;
}

// acc += p
BLS_FN void add_p12(uint32_t *acc) {
    asm("add.cc.u32 %0,%0,0xffffaaab;\n\taddc.cc.u32 %1,%1,0xb9feffff;\n\taddc.cc.u32 %2,%2,0xb153ffff;\n\t"
        "addc.cc.u32 %3,%3,0x1eabfffe;\n\taddc.cc.u32 %4,%4,0xf6b0f624;\n\taddc.cc.u32 %5,%5,0x6730d2a0;\n\t"
        "addc.cc.u32 %6,%6,0xf38512bf;\n\taddc.cc.u32 %7,%7,0x64774b84;\n\taddc.cc.u32 %8,%8,0x434bacd7;\n\t"
        "addc.cc.u32 %9,%9,0x4b1ba7b6;\n\taddc.cc.u32 %10,%10,0x397fe69a;\n\taddc.u32 %11,%11,0x1a0111ea;"
        : BLS_R12(acc));
}

// conditional final subtraction: r = t < p ? t : t - p
BLS_FN void reduce_once12(uint32_t *r, const uint32_t *t) {
    uint32_t u[12];
#pragma unroll
    for (int i = 0; i < 12; i++) u[i] = t[i];
    uint32_t bw = sub_p12(u);
#pragma unroll
    for (int i = 0; i < 12; i++) r[i] = bw ? t[i] : u[i];
}

// one 64-bit lane: (hi:lo) += a*b + carry-in, carry-out   (ptxas fuses the pair into IMAD.WIDE.U32.X)
#define BLS_LANE_C(lo, hi, a, b) \
    "madc.lo.cc.u32 " lo "," a "," b "," lo ";\n\tmadc.hi.cc.u32 " hi "," a "," b "," hi ";\n\t"
#define BLS_LANE_0(lo, hi, a, b) \
    "mad.lo.cc.u32 " lo "," a "," b "," lo ";\n\tmadc.hi.cc.u32 " hi "," a "," b "," hi ";\n\t"

// acc[0..11] += (x0 + x1*2^64 + ... + x5*2^320) * m ; acc[12] += carry.  13-limb accumulator.
BLS_FN void mad6_top(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4,
                     uint32_t x5, uint32_t m) {
    asm(BLS_LANE_0("%0", "%1", "%13", "%19") BLS_LANE_C("%2", "%3", "%14", "%19")
        BLS_LANE_C("%4", "%5", "%15", "%19") BLS_LANE_C("%6", "%7", "%16", "%19")
        BLS_LANE_C("%8", "%9", "%17", "%19") BLS_LANE_C("%10", "%11", "%18", "%19")
        "addc.u32 %12,%12,0;"
        : BLS_R12(acc), "+r"(acc[12])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4), "r"(x5), "r"(m));
}

// acc[0..11] += (x0 + x1*2^64 + ...) * m ; the carry out of the top lane is provably zero (see fp_mul)
BLS_FN void mad6(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4,
                 uint32_t x5, uint32_t m) {
    asm(BLS_LANE_0("%0", "%1", "%12", "%18") BLS_LANE_C("%2", "%3", "%13", "%18")
        BLS_LANE_C("%4", "%5", "%14", "%18") BLS_LANE_C("%6", "%7", "%15", "%18")
        BLS_LANE_C("%8", "%9", "%16", "%18") BLS_LANE_C("%10", "%11", "%17", "%18")
        : BLS_R12(acc)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4), "r"(x5), "r"(m));
}

// lo += s (carry into the chain), then acc[0..11] += x*m with that carry entering lane 0
BLS_FN void mad6_stray(uint32_t *acc, uint32_t &lo, uint32_t s, uint32_t x0, uint32_t x1, uint32_t x2,
                       uint32_t x3, uint32_t x4, uint32_t x5, uint32_t m) {
    asm("add.cc.u32 %12,%12,%13;\n\t"
        BLS_LANE_C("%0", "%1", "%14", "%20") BLS_LANE_C("%2", "%3", "%15", "%20")
        BLS_LANE_C("%4", "%5", "%16", "%20") BLS_LANE_C("%6", "%7", "%17", "%20")
        BLS_LANE_C("%8", "%9", "%18", "%20") BLS_LANE_C("%10", "%11", "%19", "%20")
        : BLS_R12(acc), "+r"(lo)
        : "r"(s), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4), "r"(x5), "r"(m));
}
#endif  // __CUDA_ARCH__

// r = a + b mod p
BLS_FN void fp_add(fp &r, const fp &a, const fp &b) {
#ifdef __CUDA_ARCH__
    uint32_t t[12];
#pragma unroll
    for (int i = 0; i < 12; i++) t[i] = a.l[i];
    add12(t, b.l);
    reduce_once12(r.l, t);
#else
    uint32_t t[12], u[12];
    uint64_t c = 0;
    for (int i = 0; i < 12; i++) { c += (uint64_t)a.l[i] + b.l[i]; t[i] = (uint32_t)c; c >>= 32; }
    int64_t bw = 0;
    for (int i = 0; i < 12; i++) { bw += (int64_t)t[i] - P32(i); u[i] = (uint32_t)bw; bw >>= 32; }
    for (int i = 0; i < 12; i++) r.l[i] = bw ? t[i] : u[i];
#endif
}

// r = a - b mod p
BLS_FN void fp_sub(fp &r, const fp &a, const fp &b) {
#ifdef __CUDA_ARCH__
    uint32_t t[12], u[12];
#pragma unroll
    for (int i = 0; i < 12; i++) t[i] = a.l[i];
    uint32_t bw = sub12(t, b.l);
#pragma unroll
    for (int i = 0; i < 12; i++) u[i] = t[i];
    add_p12(u);
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = bw ? u[i] : t[i];
#else
    uint32_t t[12];
    int64_t bw = 0;
    for (int i = 0; i < 12; i++) { bw += (int64_t)a.l[i] - b.l[i]; t[i] = (uint32_t)bw; bw >>= 32; }
    uint64_t c = 0;
    for (int i = 0; i < 12; i++) { c += (uint64_t)t[i] + (bw ? P32(i) : 0u); r.l[i] = (uint32_t)c; c >>= 32; }
#endif
}

// r = -a mod p  (0 stays 0)
BLS_FN void fp_neg(fp &r, const fp &a) {
    fp z;
    fp_set_zero(z);
    fp_sub(r, z, a);
}

BLS_FN void fp_cneg(fp &r, const fp &a, bool c) {
    fp n;
    fp_neg(n, a);
    fp_select(r, c, n, a);
}

BLS_FN void fp_dbl(fp &r, const fp &a) { fp_add(r, a, a); }

// Montgomery product r = a*b/R mod p, fully reduced.
//
// Running total T = E + 2^32*O.  E (13 limbs) takes the partial products whose weight is an even
// limb index, O (12 limbs) the odd ones, so each product lands on an aligned 64-bit register pair.
// After a row T is divisible by 2^32 and is shifted down by one limb: T/2^32 = O + E[1] + 2^32*(E>>64),
// i.e. the accumulators swap roles (new E = old O, new O = old E >> 64) and the stray limb E[1] is
// added into the new E[0]; its carry has weight 2^32 and enters the new O chain.
// Bounds: every partial sum is <= T + a*b_i + m*p < 2^414, hence O < 2^382 (no carry out of 12
// limbs) and E < 2^414 (13 limbs).
BLS_FN void fp_mul(fp &r, const fp &a, const fp &b) {
#ifdef __CUDA_ARCH__
    uint32_t E[13], O[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { E[i] = 0; O[i] = 0; }
    E[12] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint32_t bi = b.l[i];
        if (i == 0) {
            mad6(O, a.l[1], a.l[3], a.l[5], a.l[7], a.l[9], a.l[11], bi);
        } else {
            // shift: stray = E[1]; newE = O ; newO = E >> 64
            uint32_t s = E[1];
            uint32_t nO[12];
#pragma unroll
            for (int k = 0; k < 11; k++) nO[k] = E[k + 2];
            nO[11] = 0;
#pragma unroll
            for (int k = 0; k < 12; k++) E[k] = O[k];
            E[12] = 0;
#pragma unroll
            for (int k = 0; k < 12; k++) O[k] = nO[k];
            mad6_stray(O, E[0], s, a.l[1], a.l[3], a.l[5], a.l[7], a.l[9], a.l[11], bi);
        }
        mad6_top(E, a.l[0], a.l[2], a.l[4], a.l[6], a.l[8], a.l[10], bi);
        const uint32_t m = E[0] * BLS_N0;
        mad6(O, P32(1), P32(3), P32(5), P32(7), P32(9), P32(11), m);
        mad6_top(E, P32(0), P32(2), P32(4), P32(6), P32(8), P32(10), m);
    }
    // final shift: result = O + (E >> 32)
    add12(O, E + 1);
    reduce_once12(r.l, O);
#else
#ifdef BLS_COUNT_MULS
    extern unsigned long long g_fp_mul_count;
    g_fp_mul_count++;
#endif
    uint32_t t[14];
    for (int i = 0; i < 14; i++) t[i] = 0;
    for (int i = 0; i < 12; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 12; j++) {
            c += (uint64_t)a.l[j] * b.l[i] + t[j];
            t[j] = (uint32_t)c; c >>= 32;
        }
        c += t[12]; t[12] = (uint32_t)c; t[13] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * BLS_N0;
        c = ((uint64_t)m * P32(0) + t[0]) >> 32;
        for (int j = 1; j < 12; j++) {
            c += (uint64_t)m * P32(j) + t[j];
            t[j - 1] = (uint32_t)c; c >>= 32;
        }
        c += t[12]; t[11] = (uint32_t)c;
        t[12] = t[13] + (uint32_t)(c >> 32);
    }
    uint32_t u[12];
    int64_t bw = 0;
    for (int i = 0; i < 12; i++) { bw += (int64_t)t[i] - P32(i); u[i] = (uint32_t)bw; bw >>= 32; }
    for (int i = 0; i < 12; i++) r.l[i] = bw ? t[i] : u[i];
#endif
}

}  // namespace bls
#include "fp_gen.cuh"
namespace bls {

// Montgomery square.  Device: dedicated squaring (fp_gen.cuh: 78 product terms instead of 144, 234 IMADs instead of
// 300) unless BLS_NO_FP_SQR; the host build multiplies.
BLS_FN void fp_sqr(fp &r, const fp &a) {
#if defined(__CUDA_ARCH__) && !defined(BLS_NO_FP_SQR)
    uint32_t t[12];
    fp_sqr_ptx(t, a.l);
    reduce_once12(r.l, t);
#else
    fp_mul(r, a, a);
#endif
}

// out of Montgomery form: r = a/R mod p  (canonical integer limbs)
BLS_FN void fp_from_mont(fp &r, const fp &a) {
    fp one;
    fp_set_zero(one);
    one.l[0] = 1;
    fp_mul(r, a, one);
}

// parity of the canonical value (sgn0 for Fp; vendor/blst/src/no_asm.h:501-537)
BLS_FN uint32_t fp_parity(const fp &a) {
    fp t;
    fp_from_mont(t, a);
    return t.l[0] & 1;
}

// a > (p-1)/2 on the canonical value (the "sign" bit of the Zcash encoding)
BLS_FN bool fp_is_lexically_largest_canon(const fp &c) {
    // c > (p-1)/2  <=>  2c > p-1  <=> 2c >= p  (p odd)
    uint64_t carry = 0;
    uint32_t d[13];
    for (int i = 0; i < 12; i++) {
        carry += ((uint64_t)c.l[i] << 1);
        d[i] = (uint32_t)carry;
        carry >>= 32;
    }
    d[12] = (uint32_t)carry;
    int64_t bw = 0;
    for (int i = 0; i < 12; i++) { bw += (int64_t)d[i] - P32(i); bw >>= 32; }
    bw += d[12];
    return bw >= 0;
}

}  // namespace bls
