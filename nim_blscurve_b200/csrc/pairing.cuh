// pairing.cuh — optimal-ate Miller loop and final exponentiation for BLS12-381.
// Restates the results of vendor/blst/src/pairing.c: line_dbl :78, line_add :14, line_by_Px2 :128,
// miller_loop_n :220-261 (f_{|z|,Q}(P), conjugated because z < 0) and final_exp :371-404.
//
// Line functions are derived here from the untwisting map (x', y') -> (x'/w^2, y'/w^3): the line
// through T with slope lambda (twist coordinates), evaluated at P and scaled by w^3, is
//     l = (lambda x_T - y_T) + (-lambda x_P) v + y_P (v w)          ["xy00z0": a[0][0], a[0][1], a[1][1]]
// and is further scaled by Fp2 factors (2YZ^3 for a tangent, Z3 for a chord), which the final
// exponentiation kills; Miller-loop values therefore differ from BLST's by subfield factors while
// everything after final_exp — the only thing ever compared or serialised — is identical.
// The final exponentiation raises to 3 (p^12-1)/r exactly like the reference:
//     hard part 3 (p^4-p^2+1)/r = (z-1)^2 (z+p) (z^2+p^2-1) + 3.
#pragma once
#include "tower.cuh"

namespace bls {

// tangent at T, T <- 2T.  Outputs l0 and the P-independent parts of l1, l2 (l1 = l1p * (-x_P), l2 = l2p * y_P).
BLS_NOINLINE void line_dbl(g2_jac &T, fp2 &l0, fp2 &l1p, fp2 &l2p) {
    fp2 A, B, C, ZZ, E, D, Fq, t;
    fp2_sqr(A, T.x);
    fp2_sqr(B, T.y);
    fp2_sqr(C, B);
    fp2_sqr(ZZ, T.z);
    fp2_mul3(E, A);                       // 3X^2
    fp2_mul(l0, E, T.x);
    fp2_sub(l0, l0, B);
    fp2_sub(l0, l0, B);                   // 3X^3 - 2Y^2
    fp2_mul(l1p, E, ZZ);                  // 3X^2 Z^2        (times -x_P)
    fp2_add(t, T.x, B);
    fp2_sqr(t, t);
    fp2_sub(t, t, A);
    fp2_sub(t, t, C);
    fp2_dbl(D, t);
    fp2_sqr(Fq, E);
    fp2_mul(t, T.y, T.z);
    fp2_dbl(T.z, t);                      // Z3 = 2YZ
    fp2_mul(l2p, T.z, ZZ);                // 2YZ^3           (times y_P)
    fp2_sub(Fq, Fq, D);
    fp2_sub(Fq, Fq, D);
    T.x = Fq;
    fp2_sub(t, D, Fq);
    fp2_mul(t, E, t);
    fp2_dbl(C, C);
    fp2_dbl(C, C);
    fp2_dbl(C, C);
    fp2_sub(T.y, t, C);
}

// chord through T and Q, T <- T + Q (Q affine).  l1 = l1p * (-x_P), l2 = l2p * y_P.
BLS_NOINLINE void line_add(g2_jac &T, const g2_aff &Q, fp2 &l0, fp2 &l1p, fp2 &l2p) {
    fp2 Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    fp2_sqr(Z1Z1, T.z);
    fp2_mul(U2, Q.x, Z1Z1);
    fp2_mul(t, T.z, Z1Z1);
    fp2_mul(S2, Q.y, t);
    fp2_sub(H, U2, T.x);
    fp2_sub(rr, S2, T.y);
    fp2_dbl(rr, rr);
    fp2_sqr(HH, H);
    fp2_dbl(I, HH);
    fp2_dbl(I, I);
    fp2_mul(J, H, I);
    fp2_mul(V, T.x, I);
    fp2_add(t, T.z, H);
    fp2_sqr(t, t);
    fp2_sub(t, t, Z1Z1);
    fp2_sub(T.z, t, HH);                  // Z3 = 2 Z H
    fp2_sqr(t, rr);
    fp2_sub(t, t, J);
    fp2_sub(t, t, V);
    fp2_sub(t, t, V);
    T.x = t;
    fp2_sub(t, V, t);
    fp2_mul(t, rr, t);
    fp2_mul(J, T.y, J);
    fp2_dbl(J, J);
    fp2_sub(T.y, t, J);
    fp2_mul(l0, rr, Q.x);
    fp2_mul(t, Q.y, T.z);
    fp2_sub(l0, l0, t);                   // r x_Q - y_Q Z3
    l1p = rr;
    l2p = T.z;
}

BLS_FN void line_apply(fp12 &f, const fp2 &l0, const fp2 &l1p, const fp2 &l2p, const fp &neg_px, const fp &py) {
    fp2 l1, l2;
    fp2_mul_fp(l1, l1p, neg_px);
    fp2_mul_fp(l2, l2p, py);
    fp12_mul_by_line(f, l0, l1, l2);
}

// f = prod_k f_{|z|,Q_k}(P_k), conjugated; the Fp12 squarings are shared by the n pairs.
// Pairs with P or Q at infinity contribute 1.  T is caller-provided scratch of n entries.
BLS_NOINLINE void miller_loop_n(fp12 &f, const g2_aff *Q, const g1_aff *P, int n, g2_jac *T, fp *neg_px) {
    fp12_set_one(f);
    int live = 0;
    for (int k = 0; k < n; k++) {
        pt_from_affine(T[k], Q[k]);
        fp_neg(neg_px[k], P[k].x);
        // an infinite P is flagged by making T infinite as well
        if (aff_is_inf(P[k])) pt_set_inf(T[k]);
        live += !pt_is_inf(T[k]);
    }
    if (live == 0) return;
    const uint64_t z = BLS_Z_ABS;
    fp2 l0, l1p, l2p;
    for (int i = 62; i >= 0; i--) {
        if (i != 62) fp12_sqr(f, f);
        for (int k = 0; k < n; k++) {
            if (pt_is_inf(T[k])) continue;
            line_dbl(T[k], l0, l1p, l2p);
            line_apply(f, l0, l1p, l2p, neg_px[k], P[k].y);
        }
        if ((z >> i) & 1) {
            for (int k = 0; k < n; k++) {
                if (pt_is_inf(T[k])) continue;
                line_add(T[k], Q[k], l0, l1p, l2p);
                line_apply(f, l0, l1p, l2p, neg_px[k], P[k].y);
            }
        }
    }
    fp12_conj(f, f);
}

// r = a^z for a in the cyclotomic subgroup (z negative: conjugate of a^|z|)
BLS_NOINLINE void cyc_exp_z(fp12 &r, const fp12 &a) {
    fp12 acc = a;
    const uint64_t z = BLS_Z_ABS;
    for (int i = 62; i >= 0; i--) {
        fp12_cyc_sqr(acc, acc);
        if ((z >> i) & 1) fp12_mul(acc, acc, a);
    }
    fp12_conj(r, acc);
}

// r = f^(3 (p^12-1)/r)
BLS_NOINLINE void final_exp(fp12 &r, const fp12 &f) {
    fp12 t, a, b, c, d;
    // easy part: f^((p^6-1)(p^2+1))
    fp12_conj(a, f);
    fp12_inv(b, f);
    fp12_mul(t, a, b);
    fp12_frob(a, t, 2);
    fp12_mul(t, a, t);
    // hard part: t^((z-1)^2 (z+p) (z^2+p^2-1) + 3)
    cyc_exp_z(a, t);
    fp12_conj(b, t);
    fp12_mul(a, a, b);              // t^(z-1)
    cyc_exp_z(b, a);
    fp12_conj(c, a);
    fp12_mul(a, b, c);              // t^((z-1)^2)
    cyc_exp_z(b, a);
    fp12_frob(c, a, 1);
    fp12_mul(b, b, c);              // a^(z+p)
    cyc_exp_z(c, b);
    cyc_exp_z(c, c);                // b^(z^2)
    fp12_frob(d, b, 2);
    fp12_mul(c, c, d);
    fp12_conj(d, b);
    fp12_mul(c, c, d);              // b^(z^2+p^2-1)
    fp12_cyc_sqr(d, t);
    fp12_mul(d, d, t);              // t^3
    fp12_mul(r, c, d);
}

}  // namespace bls
