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

// The same two steps on HOMOGENEOUS projective coordinates (x = X/Z, y = Y/Z), the form the split Miller loop uses:
// tangent 2M + 7S instead of 5M + 6S (Costello-Lange-Naehrig, a = 0, b' = 4 xi).  With B = Y^2, C = Z^2, E = 3 b' C,
// the tangent scaled by 2YZ^2 / Z is  (B - E) + 3X^2 (-x_P) v + 2YZ y_P (v w); the doubled point is taken times 4
// (a projective rescale) so that no halving is needed:
//     X3 = 2XY (B - 3E),  Y3 = (B + 3E)^2 - 12 E^2,  Z3 = 4 B * 2YZ.
BLS_NOINLINE void line_dbl_proj(g2_jac &T, fp2 &l0, fp2 &l1p, fp2 &l2p) {
    fp2 A2, B, C, E, F, H, J, t, u;
    fp2_sqr(B, T.y);
    fp2_sqr(C, T.z);
    fp2_sqr(J, T.x);
    fp2_add(t, T.x, T.y);
    fp2_sqr(A2, t);
    fp2_sub(A2, A2, J);
    fp2_sub(A2, A2, B);                   // 2XY
    fp2_add(t, T.y, T.z);
    fp2_sqr(H, t);
    fp2_sub(H, H, B);
    fp2_sub(H, H, C);                     // 2YZ
    fp2_mul_xi(E, C);
    fp2_dbl(E, E);
    fp2_dbl(E, E);
    fp2_mul3(E, E);                       // E = 12 xi Z^2 = 3 b' Z^2
    fp2_mul3(F, E);                       // 3E
    fp2_sub(l0, B, E);
    fp2_mul3(l1p, J);                     // 3X^2            (times -x_P)
    l2p = H;                              // 2YZ             (times y_P)
    fp2_sub(t, B, F);
    fp2_mul(T.x, A2, t);
    fp2_add(t, B, F);
    fp2_sqr(t, t);
    fp2_dbl(u, E);
    fp2_sqr(u, u);
    fp2_mul3(u, u);                       // 12 E^2
    fp2_sub(T.y, t, u);
    fp2_mul(t, B, H);
    fp2_dbl(t, t);
    fp2_dbl(T.z, t);
}

// chord through T and the affine Q, T <- T + Q: theta = Y - y_Q Z, lambda = X - x_Q Z,
// line = (theta x_Q - lambda y_Q) + theta (-x_P) v + lambda y_P (v w)          (11M + 2S)
BLS_NOINLINE void line_add_proj(g2_jac &T, const g2_aff &Q, fp2 &l0, fp2 &l1p, fp2 &l2p) {
    fp2 th, la, c, d, e, f, g, h, t;
    fp2_mul(t, Q.y, T.z);
    fp2_sub(th, T.y, t);
    fp2_mul(t, Q.x, T.z);
    fp2_sub(la, T.x, t);
    fp2_sqr(c, th);
    fp2_sqr(d, la);
    fp2_mul(e, la, d);
    fp2_mul(f, T.z, c);
    fp2_mul(g, T.x, d);
    fp2_add(h, e, f);
    fp2_sub(h, h, g);
    fp2_sub(h, h, g);
    fp2_mul(T.x, la, h);
    fp2_sub(t, g, h);
    fp2_mul(t, th, t);
    fp2_mul(g, e, T.y);
    fp2_sub(T.y, t, g);
    fp2_mul(T.z, T.z, e);
    fp2_mul(l0, th, Q.x);
    fp2_mul(t, la, Q.y);
    fp2_sub(l0, l0, t);
    l1p = th;
    l2p = la;
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

// ---------------------------------------------------------------------------------------------
// Split Miller loop: lines first, accumulation second.
//
// prod_k f_{|z|,Q_k}(P_k) = prod_i ( prod_k line_{k,i} )^(2^i): the per-pair state (T) only feeds the
// line coefficients, and the Fp12 accumulator only consumes them.  miller_lines() walks T for ONE pair
// and emits its 68 sparse line triples (63 tangents + 5 chords, in execution order); miller_accumulate()
// folds the lines of a GROUP of pairs over a SEGMENT of the loop (bits i_hi..i_lo of |z|) into an Fp12;
// miller_combine() stitches the per-segment products with the remaining squarings.  The squarings are
// shared by every pair of a group, the groups and segments are independent work items, and the two
// halves have half the live state each.
//
// Line storage is word-major over pairs ("SoA"): word w (0..71: l0, l1, l2 as 3 x fp2) of line s of pair p
// lives at lines[(s*72 + w)*stride + p], so that both producers and consumers touch 128 contiguous bytes per warp.
#define ML_NLINES 68
#define ML_LINE_WORDS 72

BLS_FN int ml_bit(int i) { return (int)((BLS_Z_ABS >> i) & 1); }
// index of the tangent line of loop iteration i (i = 62..0); the chord of the same iteration, if any, follows it
BLS_FN int ml_line_index(int i) { return (62 - i) + (i < 62) + (i < 60) + (i < 57) + (i < 48) + (i < 16); }
// segment j of nseg covers iterations seg_hi >= i >= seg_lo
BLS_HD int ml_seg_hi(int j, int nseg) { return 62 - (63 * j) / nseg; }
BLS_HD int ml_seg_lo(int j, int nseg) { return 62 - (63 * (j + 1)) / nseg + 1; }

BLS_FN void line_store(uint32_t *dst, size_t stride, int s, const fp2 &l0, const fp2 &l1, const fp2 &l2) {
    uint32_t *d = dst + (size_t)s * ML_LINE_WORDS * stride;
    for (int w = 0; w < 12; w++) {
        d[(size_t)w * stride] = l0.c0.l[w];
        d[(size_t)(12 + w) * stride] = l0.c1.l[w];
        d[(size_t)(24 + w) * stride] = l1.c0.l[w];
        d[(size_t)(36 + w) * stride] = l1.c1.l[w];
        d[(size_t)(48 + w) * stride] = l2.c0.l[w];
        d[(size_t)(60 + w) * stride] = l2.c1.l[w];
    }
}

BLS_FN void line_load(fp2 &l0, fp2 &l1, fp2 &l2, const uint32_t *src, size_t stride, int s) {
    const uint32_t *d = src + (size_t)s * ML_LINE_WORDS * stride;
    for (int w = 0; w < 12; w++) {
        l0.c0.l[w] = d[(size_t)w * stride];
        l0.c1.l[w] = d[(size_t)(12 + w) * stride];
        l1.c0.l[w] = d[(size_t)(24 + w) * stride];
        l1.c1.l[w] = d[(size_t)(36 + w) * stride];
        l2.c0.l[w] = d[(size_t)(48 + w) * stride];
        l2.c1.l[w] = d[(size_t)(60 + w) * stride];
    }
}

// All 68 lines of one pair, already scaled by (-x_P, y_P).  A pair with P or Q at infinity emits the
// neutral line (1, 0, 0) everywhere, i.e. contributes 1 to the product (pairing.c:233-241).
BLS_NOINLINE void miller_lines(const g2_aff &Q, const g1_aff &P, uint32_t *dst, size_t stride) {
    fp2 l0, l1, l2;
    if (aff_is_inf(Q) | aff_is_inf(P)) {
        f_set_one(l0);
        fp2_set_zero(l1);
        fp2_set_zero(l2);
        for (int s = 0; s < ML_NLINES; s++) line_store(dst, stride, s, l0, l1, l2);
        return;
    }
    g2_jac T;
    pt_from_affine(T, Q);
    fp npx, py = P.y;
    fp_neg(npx, P.x);
    int s = 0;
    for (int i = 62; i >= 0; i--) {
        line_dbl_proj(T, l0, l1, l2);
        fp2_mul_fp(l1, l1, npx);
        fp2_mul_fp(l2, l2, py);
        line_store(dst, stride, s++, l0, l1, l2);
        if (ml_bit(i)) {
            line_add_proj(T, Q, l0, l1, l2);
            fp2_mul_fp(l1, l1, npx);
            fp2_mul_fp(l2, l2, py);
            line_store(dst, stride, s++, l0, l1, l2);
        }
    }
}

// f = prod over iterations i_hi..i_lo and over the pairs {g, g+ngroups, g+2 ngroups, ...} (at most G, < np)
// of line^(2^(i-i_lo)); `lines` points at pair 0 of the tile.
BLS_NOINLINE void miller_accumulate(fp12 &f, const uint32_t *lines, size_t stride, size_t np, size_t g, size_t ngroups,
                                    int G, int i_hi, int i_lo) {
    fp2 l0, l1, l2;
    bool first = true;
    fp12_set_one(f);
    for (int i = i_hi; i >= i_lo; i--) {
        if (!first) fp12_sqr(f, f);
        const int s = ml_line_index(i), nl = 1 + ml_bit(i);
        for (int a = 0; a < nl; a++)
            for (int k = 0; k < G; k++) {
                size_t p = g + (size_t)k * ngroups;
                if (p >= np) break;
                line_load(l0, l1, l2, lines + p, stride, s + a);
                if (first) {
                    f.c0.c0 = l0;
                    f.c0.c1 = l1;
                    f.c1.c1 = l2;
                    first = false;
                } else {
                    fp12_mul_by_line(f, l0, l1, l2);
                }
            }
    }
}

// F = conj( prod_j seg[j]^(2^seg_lo(j)) ): Horner over the segments, 63 squarings in total
BLS_NOINLINE void miller_combine(fp12 &F, const fp12 *seg, int nseg) {
    fp12 acc = seg[0];
    for (int j = 1; j < nseg; j++) {
        int len = ml_seg_hi(j, nseg) - ml_seg_lo(j, nseg) + 1;
        for (int t = 0; t < len; t++) fp12_sqr(acc, acc);
        fp12 b = seg[j];
        fp12_mul(acc, acc, b);
    }
    fp12_conj(F, acc);
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
