// fp2.cuh — Fp2 = Fp[u]/(u^2+1) (vendor/blst/src/fp12_tower.c:9, vect.h:54: (re, im) order),
// plus the noinline call boundary of the arithmetic: fp_mul/fp_sqr are force-inlined PTX,
// everything above Fp2 calls fp2_mul / fp2_sqr (and fp_mul_ni for Fp-only code) as real functions
// so that kernels stay a few thousand instructions instead of millions.
#pragma once
#include "fp.cuh"

namespace bls {

struct fp2 { fp c0, c1; };

// Fp-level noinline entry points (G1 arithmetic, exponentiation chains)
BLS_NOINLINE void fp_mul_ni(fp &r, const fp &a, const fp &b) {
    fp x = a, y = b, t;
    fp_mul(t, x, y);
    r = t;
}
BLS_NOINLINE void fp_sqr_ni(fp &r, const fp &a) {
    fp x = a, t;
    fp_sqr(t, x);
    r = t;
}

// Measured on B200 (profiles/): with fp_mul called (not inlined) inside fp2_mul/fp2_sqr the Fp2-heavy kernels fit 128
// registers -> 4 warps per scheduler, and the hash / line kernels run 4-11 % faster.  Level 0 inlines everything.
#if defined(__CUDACC__) && !defined(BLS_SMALL_CODE)
#define BLS_SMALL_CODE 1
#endif
#if defined(BLS_SMALL_CODE) && BLS_SMALL_CODE >= 2
// one copy of each Fp primitive in the instruction stream (I-cache footprint)
BLS_NOINLINE void fp_add_ni(fp &r, const fp &a, const fp &b) { fp t; fp_add(t, a, b); r = t; }
BLS_NOINLINE void fp_sub_ni(fp &r, const fp &a, const fp &b) { fp t; fp_sub(t, a, b); r = t; }
#define FP_ADD fp_add_ni
#define FP_SUB fp_sub_ni
#define FP_MUL fp_mul_ni
#elif defined(BLS_SMALL_CODE) && BLS_SMALL_CODE >= 1
#define FP_ADD fp_add
#define FP_SUB fp_sub
#define FP_MUL fp_mul_ni
#else
#define FP_ADD fp_add
#define FP_SUB fp_sub
#define FP_MUL fp_mul
#endif

BLS_FN void fp2_set_zero(fp2 &r) { fp_set_zero(r.c0); fp_set_zero(r.c1); }
BLS_FN bool fp2_is_zero(const fp2 &a) { return fp_is_zero(a.c0) & fp_is_zero(a.c1); }
BLS_FN bool fp2_eq(const fp2 &a, const fp2 &b) { return fp_eq(a.c0, b.c0) & fp_eq(a.c1, b.c1); }
BLS_FN void fp2_select(fp2 &r, bool c, const fp2 &a, const fp2 &b) {
    fp_select(r.c0, c, a.c0, b.c0);
    fp_select(r.c1, c, a.c1, b.c1);
}
BLS_FN void fp2_add(fp2 &r, const fp2 &a, const fp2 &b) { FP_ADD(r.c0, a.c0, b.c0); FP_ADD(r.c1, a.c1, b.c1); }
BLS_FN void fp2_sub(fp2 &r, const fp2 &a, const fp2 &b) { FP_SUB(r.c0, a.c0, b.c0); FP_SUB(r.c1, a.c1, b.c1); }
BLS_FN void fp2_neg(fp2 &r, const fp2 &a) { fp_neg(r.c0, a.c0); fp_neg(r.c1, a.c1); }
BLS_FN void fp2_dbl(fp2 &r, const fp2 &a) { FP_ADD(r.c0, a.c0, a.c0); FP_ADD(r.c1, a.c1, a.c1); }
BLS_FN void fp2_conj(fp2 &r, const fp2 &a) { r.c0 = a.c0; fp_neg(r.c1, a.c1); }
BLS_FN void fp2_cneg(fp2 &r, const fp2 &a, bool c) { fp_cneg(r.c0, a.c0, c); fp_cneg(r.c1, a.c1, c); }

// r = a * (1+u)
BLS_FN void fp2_mul_xi(fp2 &r, const fp2 &a) {
    fp t0, t1;
    FP_SUB(t0, a.c0, a.c1);
    FP_ADD(t1, a.c0, a.c1);
    r.c0 = t0;
    r.c1 = t1;
}

#ifdef __CUDA_ARCH__
// ---- double-width pieces for the lazily reduced Fp2 product ----
// T[0..23] = a * b as plain integers (a, b < 2^382): the even/odd accumulator rows of fp_mul without the reduction
// rows; the limb that fp_mul's reduction would clear is the finished output limb of the row.
BLS_FN void mul_wide12(uint32_t *T, const uint32_t *a, const uint32_t *b) {
    uint32_t E[13], O[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { E[i] = 0; O[i] = 0; }
    E[12] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint32_t bi = b[i];
        if (i == 0) {
            mad6(O, a[1], a[3], a[5], a[7], a[9], a[11], bi);
        } else {
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
            mad6_stray(O, E[0], s, a[1], a[3], a[5], a[7], a[9], a[11], bi);
        }
        mad6_top(E, a[0], a[2], a[4], a[6], a[8], a[10], bi);
        T[i] = E[0];
    }
    add12(O, E + 1);                                         // a b < 2^764: the high half fits twelve limbs
#pragma unroll
    for (int k = 0; k < 12; k++) T[12 + k] = O[k];
}
// r = T / R mod p, fully reduced, for T < p R:  (T_lo + M p) / R  [<= p]  +  T_hi  [< p], one conditional subtraction.
// Twelve reduction rows of fp_mul on the low half.
BLS_FN void redc24(uint32_t *r, const uint32_t *T) {
    uint32_t E[13], O[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { E[i] = T[i]; O[i] = 0; }
    E[12] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t m;
        if (i == 0) {
            m = E[0] * BLS_N0;
            mad6(O, P32(1), P32(3), P32(5), P32(7), P32(9), P32(11), m);
        } else {
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
            m = (E[0] + s) * BLS_N0;
            mad6_stray(O, E[0], s, P32(1), P32(3), P32(5), P32(7), P32(9), P32(11), m);
        }
        mad6_top(E, P32(0), P32(2), P32(4), P32(6), P32(8), P32(10), m);
    }
    add12(O, E + 1);
    add12(O, T + 12);                                        // < 2p < 2^382
    reduce_once12(r, O);
}
// acc -= b with a borrow (mask) coming in; returns the borrow mask going out
BLS_FN uint32_t sub12_b(uint32_t *acc, const uint32_t *b, uint32_t bw_in) {
    uint32_t bw;
    asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t,0,%25;\n\tsubc.cc.u32 %0,%0,%13;\n\tsubc.cc.u32 %1,%1,%14;\n\tsubc.cc.u32 %2,%2,%15;\n\t"
        "subc.cc.u32 %3,%3,%16;\n\tsubc.cc.u32 %4,%4,%17;\n\tsubc.cc.u32 %5,%5,%18;\n\t"
        "subc.cc.u32 %6,%6,%19;\n\tsubc.cc.u32 %7,%7,%20;\n\tsubc.cc.u32 %8,%8,%21;\n\t"
        "subc.cc.u32 %9,%9,%22;\n\tsubc.cc.u32 %10,%10,%23;\n\tsubc.cc.u32 %11,%11,%24;\n\t"
        "subc.u32 %12,0,0;\n\t}"
        : BLS_R12(acc), "=r"(bw)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]),
          "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(bw_in));
    return bw;
}
BLS_FN uint32_t sub24(uint32_t *acc, const uint32_t *b) {
    const uint32_t bw = sub12(acc, b);
    return sub12_b(acc + 12, b + 12, bw);
}
#endif

// Karatsuba with three reduced products
BLS_NOINLINE void fp2_mul_r3(fp2 &r, const fp2 &a, const fp2 &b) {
    fp s0, s1, t0, t1, t2;
    FP_ADD(s0, a.c0, a.c1);
    FP_ADD(s1, b.c0, b.c1);
    FP_MUL(t2, s0, s1);
    FP_MUL(t0, a.c0, b.c0);
    FP_MUL(t1, a.c1, b.c1);
    FP_SUB(t2, t2, t0);
    FP_SUB(t2, t2, t1);
    FP_SUB(t0, t0, t1);
    r.c0 = t0;
    r.c1 = t2;
}

#ifndef BLS_FP2_LAZY
#define BLS_FP2_LAZY 0
#endif

// Karatsuba: 3 Fp multiplications.  Operands are read from memory right where they are used and the
// result is written last (r may alias a or b), which keeps the live register set near one multiplication.
// -DBLS_FP2_LAZY=1 (measured in round 2 and NOT adopted): the three products stay double-width, the Karatsuba
// subtractions run on 24 limbs and only the two results are reduced — 3 x 144 + 2 x 156 = 744 multiply-adds instead of
// 3 x 300 = 900, same canonical output (parity suite green).  Timed alone at 131 072 sets: the G2 MSM 4.00 -> 3.57 ms,
// but the hash kernel 24.0 -> 25.5 ms and the line kernel 10.7 -> 12.6 ms: one 1 100-instruction body with ~100 live
// registers loses more (spills around it, 17 KB of a 32 KB instruction cache) than the 17 % fewer multiply-adds win.
BLS_NOINLINE void fp2_mul(fp2 &r, const fp2 &a, const fp2 &b) {
#if defined(__CUDA_ARCH__) && BLS_FP2_LAZY
    uint32_t T2[24], T0[24];
    {
        uint32_t s0[12], s1[12];
#pragma unroll
        for (int i = 0; i < 12; i++) { s0[i] = a.c0.l[i]; s1[i] = b.c0.l[i]; }
        add12(s0, a.c1.l);                                   // < 2p < 2^382: no reduction
        add12(s1, b.c1.l);
        mul_wide12(T2, s0, s1);
    }
    mul_wide12(T0, a.c0.l, b.c0.l);
    sub24(T2, T0);
    {
        uint32_t T1[24];
        mul_wide12(T1, a.c1.l, b.c1.l);
        sub24(T2, T1);                                       // a0 b1 + a1 b0 in [0, 2 p^2)
        const uint32_t bw = sub24(T0, T1);                   // a0 b0 - a1 b1 in (-p^2, p^2): + p R when negative
        uint32_t u[12];
#pragma unroll
        for (int i = 0; i < 12; i++) u[i] = T0[12 + i];
        add_p12(u);
#pragma unroll
        for (int i = 0; i < 12; i++) T0[12 + i] = bw ? u[i] : T0[12 + i];
    }
    fp re, im;
    redc24(im.l, T2);
    redc24(re.l, T0);
    r.c0 = re;
    r.c1 = im;
    return;
#else
    fp2_mul_r3(r, a, b);
#endif
}

// (a0+a1)(a0-a1) + 2 a0 a1 u : 2 Fp multiplications
BLS_NOINLINE void fp2_sqr(fp2 &r, const fp2 &a) {
    fp s, d, t, u;
    FP_ADD(s, a.c0, a.c1);
    FP_SUB(d, a.c0, a.c1);
    FP_MUL(u, s, d);
    FP_MUL(t, a.c0, a.c1);
    FP_ADD(t, t, t);
    r.c0 = u;
    r.c1 = t;
}

// r = a * k, k in Fp
BLS_FN void fp2_mul_fp(fp2 &r, const fp2 &a, const fp &k) {
    fp_mul_ni(r.c0, a.c0, k);
    fp_mul_ni(r.c1, a.c1, k);
}

// small multiples
BLS_FN void fp2_mul3(fp2 &r, const fp2 &a) { fp2 t; fp2_dbl(t, a); fp2_add(r, t, a); }
BLS_FN void fp_mul3(fp &r, const fp &a) { fp t; fp_dbl(t, a); fp_add(r, t, a); }

}  // namespace bls
