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

// Karatsuba: 3 Fp multiplications.  Operands are read from memory right where they are used and the
// result is written last (r may alias a or b), which keeps the live register set near one multiplication.
BLS_NOINLINE void fp2_mul(fp2 &r, const fp2 &a, const fp2 &b) {
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
