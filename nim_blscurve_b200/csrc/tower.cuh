// tower.cuh — Fp6 = Fp2[v]/(v^3 - (1+u)), Fp12 = Fp6[w]/(w^2 - v)  (vendor/blst/src/fp12_tower.c:9-13).
// In-memory order of an Fp12 is a[j][i][k] (j: w power, i: v power, k: re/im), identical to the
// reference's vec384fp12 (fields.h:99-101), so 576-byte partials can be exchanged verbatim.
// Restates the results of mul_fp12 :309, mul_by_xy00z0_fp12 :372, sqr_fp12 :496, inverse_fp12 :569,
// cyclotomic_sqr_fp12 :630 (Granger-Scott), frobenius_map_fp12 :706, blst_bendian_from_fp12 :773.
#pragma once
#include "ec.cuh"

namespace bls {

struct fp6 { fp2 c0, c1, c2; };
struct fp12 { fp6 c0, c1; };

BLS_FN void fp6_add(fp6 &r, const fp6 &a, const fp6 &b) { fp2_add(r.c0, a.c0, b.c0); fp2_add(r.c1, a.c1, b.c1); fp2_add(r.c2, a.c2, b.c2); }
BLS_FN void fp6_sub(fp6 &r, const fp6 &a, const fp6 &b) { fp2_sub(r.c0, a.c0, b.c0); fp2_sub(r.c1, a.c1, b.c1); fp2_sub(r.c2, a.c2, b.c2); }
BLS_FN void fp6_neg(fp6 &r, const fp6 &a) { fp2_neg(r.c0, a.c0); fp2_neg(r.c1, a.c1); fp2_neg(r.c2, a.c2); }
BLS_FN void fp6_dbl(fp6 &r, const fp6 &a) { fp2_dbl(r.c0, a.c0); fp2_dbl(r.c1, a.c1); fp2_dbl(r.c2, a.c2); }
// r = a * v
BLS_FN void fp6_mul_v(fp6 &r, const fp6 &a) {
    fp2 t;
    fp2_mul_xi(t, a.c2);
    r.c2 = a.c1;
    r.c1 = a.c0;
    r.c0 = t;
}

BLS_NOINLINE void fp6_mul(fp6 &r, const fp6 &a, const fp6 &b) {
    fp2 t0, t1, t2, s, u, c0, c1, c2;
    fp2_mul(t0, a.c0, b.c0);
    fp2_mul(t1, a.c1, b.c1);
    fp2_mul(t2, a.c2, b.c2);
    fp2_add(s, a.c1, a.c2);
    fp2_add(u, b.c1, b.c2);
    fp2_mul(c0, s, u);
    fp2_sub(c0, c0, t1);
    fp2_sub(c0, c0, t2);
    fp2_mul_xi(c0, c0);
    fp2_add(c0, c0, t0);
    fp2_add(s, a.c0, a.c1);
    fp2_add(u, b.c0, b.c1);
    fp2_mul(c1, s, u);
    fp2_sub(c1, c1, t0);
    fp2_sub(c1, c1, t1);
    fp2_mul_xi(s, t2);
    fp2_add(c1, c1, s);
    fp2_add(s, a.c0, a.c2);
    fp2_add(u, b.c0, b.c2);
    fp2_mul(c2, s, u);
    fp2_sub(c2, c2, t0);
    fp2_sub(c2, c2, t2);
    fp2_add(c2, c2, t1);
    r.c0 = c0;
    r.c1 = c1;
    r.c2 = c2;
}

// r = a * (c0 + c1 v)   — 5 Fp2 multiplications
BLS_NOINLINE void fp6_mul_by_01(fp6 &r, const fp6 &a, const fp2 &c0, const fp2 &c1) {
    fp2 t0, t1, t2, s, u, r0, r1, r2;
    fp2_mul(t0, a.c0, c0);
    fp2_mul(t1, a.c1, c1);
    fp2_mul(t2, a.c2, c1);
    fp2_mul_xi(t2, t2);
    fp2_add(r0, t0, t2);
    fp2_add(s, a.c0, a.c1);
    fp2_add(u, c0, c1);
    fp2_mul(r1, s, u);
    fp2_sub(r1, r1, t0);
    fp2_sub(r1, r1, t1);
    fp2_mul(r2, a.c2, c0);
    fp2_add(r2, r2, t1);
    r.c0 = r0;
    r.c1 = r1;
    r.c2 = r2;
}

// r = a * (c1 v)   — 3 Fp2 multiplications
BLS_NOINLINE void fp6_mul_by_1(fp6 &r, const fp6 &a, const fp2 &c1) {
    fp2 r0, r1, r2;
    fp2_mul(r0, a.c2, c1);
    fp2_mul_xi(r0, r0);
    fp2_mul(r1, a.c0, c1);
    fp2_mul(r2, a.c1, c1);
    r.c0 = r0;
    r.c1 = r1;
    r.c2 = r2;
}

BLS_NOINLINE void fp6_inv(fp6 &r, const fp6 &a) {
    fp2 c0, c1, c2, t, s;
    fp2_sqr(c0, a.c0);
    fp2_mul(t, a.c1, a.c2);
    fp2_mul_xi(t, t);
    fp2_sub(c0, c0, t);
    fp2_sqr(c1, a.c2);
    fp2_mul_xi(c1, c1);
    fp2_mul(t, a.c0, a.c1);
    fp2_sub(c1, c1, t);
    fp2_sqr(c2, a.c1);
    fp2_mul(t, a.c0, a.c2);
    fp2_sub(c2, c2, t);
    fp2_mul(t, a.c2, c1);
    fp2_mul(s, a.c1, c2);
    fp2_add(t, t, s);
    fp2_mul_xi(t, t);
    fp2_mul(s, a.c0, c0);
    fp2_add(t, t, s);
    fp2_inv(t, t);
    fp2_mul(r.c0, c0, t);
    fp2_mul(r.c1, c1, t);
    fp2_mul(r.c2, c2, t);
}

BLS_FN void fp12_set_one(fp12 &r) {
    fp2 *c = &r.c0.c0;
    for (int i = 0; i < 6; i++) fp2_set_zero(c[i]);
    r.c0.c0.c0 = FP_ONE;
}

BLS_FN bool fp12_is_one(const fp12 &a) {
    const fp2 *c = &a.c0.c0;
    bool ok = fp_eq(c[0].c0, FP_ONE) & fp_is_zero(c[0].c1);
    for (int i = 1; i < 6; i++) ok &= fp2_is_zero(c[i]);
    return ok;
}

BLS_FN void fp12_conj(fp12 &r, const fp12 &a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }

// 3 Fp6 multiplications (18 Fp2)
BLS_NOINLINE void fp12_mul(fp12 &r, const fp12 &a, const fp12 &b) {
    fp6 t0, t1, s, u;
    fp6_mul(t0, a.c0, b.c0);
    fp6_mul(t1, a.c1, b.c1);
    fp6_add(s, a.c0, a.c1);
    fp6_add(u, b.c0, b.c1);
    fp6_mul(s, s, u);
    fp6_sub(s, s, t0);
    fp6_sub(r.c1, s, t1);
    fp6_mul_v(t1, t1);
    fp6_add(r.c0, t0, t1);
}

// complex squaring: 2 Fp6 multiplications
BLS_NOINLINE void fp12_sqr(fp12 &r, const fp12 &a) {
    fp6 t, s, u;
    fp6_mul(t, a.c0, a.c1);
    fp6_add(s, a.c0, a.c1);
    fp6_mul_v(u, a.c1);
    fp6_add(u, u, a.c0);
    fp6_mul(s, s, u);
    fp6_sub(s, s, t);
    fp6_mul_v(u, t);
    fp6_sub(r.c0, s, u);
    fp6_dbl(r.c1, t);
}

// f *= (l0 + l1 v + l2 v w)  — sparse "xy00z0" element (pairing.c:155), 13 Fp2 multiplications
BLS_NOINLINE void fp12_mul_by_line(fp12 &f, const fp2 &l0, const fp2 &l1, const fp2 &l2) {
    fp6 t0, t1, t2, s;
    fp2 l12;
    fp6_mul_by_01(t0, f.c0, l0, l1);
    fp6_mul_by_1(t1, f.c1, l2);
    fp6_add(s, f.c0, f.c1);
    fp2_add(l12, l1, l2);
    fp6_mul_by_01(t2, s, l0, l12);
    fp6_sub(t2, t2, t0);
    fp6_sub(f.c1, t2, t1);
    fp6_mul_v(t1, t1);
    fp6_add(f.c0, t0, t1);
}

BLS_NOINLINE void fp12_inv(fp12 &r, const fp12 &a) {
    fp6 t0, t1;
    fp6_mul(t0, a.c0, a.c0);
    fp6_mul(t1, a.c1, a.c1);
    fp6_mul_v(t1, t1);
    fp6_sub(t0, t0, t1);
    fp6_inv(t0, t0);
    fp6_mul(r.c0, a.c0, t0);
    fp6_mul(t1, a.c1, t0);
    fp6_neg(r.c1, t1);
}

// a^(p^n), n = 1, 2, 3: coefficient of v^i w^j is multiplied by xi^((2i+j)(p^n-1)/6), conjugated for odd n
BLS_NOINLINE void fp12_frob(fp12 &r, const fp12 &a, int n) {
    const fp2 *g = n == 1 ? FROB1 : (n == 2 ? FROB2 : FROB3);
    const fp2 *src = &a.c0.c0;
    fp2 *dst = &r.c0.c0;
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 3; i++) {
            fp2 c = src[3 * j + i];
            if (n & 1) fp2_conj(c, c);
            int k = 2 * i + j;
            if (k) fp2_mul(c, c, g[k - 1]);
            dst[3 * j + i] = c;
        }
}

// (a + b s)^2 in Fp4 = Fp2[s]/(s^2 - xi): t0 = a^2 + xi b^2, t1 = 2ab
BLS_FN void fp4_sqr(fp2 &t0, fp2 &t1, const fp2 &a, const fp2 &b) {
    fp2 a2, b2, s;
    fp2_sqr(a2, a);
    fp2_sqr(b2, b);
    fp2_add(s, a, b);
    fp2_sqr(s, s);
    fp2_sub(s, s, a2);
    fp2_sub(t1, s, b2);
    fp2_mul_xi(b2, b2);
    fp2_add(t0, a2, b2);
}

// Granger-Scott squaring for elements of the cyclotomic subgroup: 9 Fp2 squarings
BLS_NOINLINE void fp12_cyc_sqr(fp12 &r, const fp12 &a) {
    fp2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2;
    fp2 t0, t1, t2, t3, t;
    fp4_sqr(t0, t1, z0, z1);
    // z0 = 3 t0 - 2 z0 ; z1 = 3 t1 + 2 z1
    fp2_sub(t, t0, z0); fp2_dbl(t, t); fp2_add(z0, t, t0);
    fp2_add(t, t1, z1); fp2_dbl(t, t); fp2_add(z1, t, t1);
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    // z4 = 3 t0 - 2 z4 ; z5 = 3 t1 + 2 z5
    fp2_sub(t, t0, z4); fp2_dbl(t, t); fp2_add(z4, t, t0);
    fp2_add(t, t1, z5); fp2_dbl(t, t); fp2_add(z5, t, t1);
    // z2 = 3 xi t3 + 2 z2 ; z3 = 3 t2 - 2 z3
    fp2_mul_xi(t3, t3);
    fp2_add(t, t3, z2); fp2_dbl(t, t); fp2_add(z2, t, t3);
    fp2_sub(t, t2, z3); fp2_dbl(t, t); fp2_add(z3, t, t2);
    r.c0.c0 = z0; r.c0.c1 = z4; r.c0.c2 = z3;
    r.c1.c0 = z2; r.c1.c1 = z1; r.c1.c2 = z5;
}

// canonical GT bytes (fp12_tower.c:773-786): for i in 0..2, j in 0..1: a[j][i].re || a[j][i].im, 48-byte BE each
BLS_FN void fp12_to_bytes(uint8_t *out, const fp12 &a) {
    const fp2 *c = &a.c0.c0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) {
            const fp2 &e = c[3 * j + i];
            for (int k = 0; k < 2; k++) {
                fp t;
                fp_from_mont(t, k ? e.c1 : e.c0);
                for (int b = 0; b < 48; b++) *out++ = (uint8_t)(t.l[(47 - b) >> 2] >> (8 * ((47 - b) & 3)));
            }
        }
}

}  // namespace bls
