// h2c.cuh — hash_to_G2 for BLS12381G2_XMD:SHA-256_SSWU_RO_ (RFC 9380 8.8.2).
// Restates the results of vendor/blst/src/map_to_g2.c: map_to_isogenous_E2 :173-290 (SSWU on
// E2': y^2 = x^3 + 240u x + 1012(1+u), Z = -(2+u)), the addition of the two mapped points on E2'
// (:363), isogeny_map_to_E2 :43-171 (3-isogeny, evaluated projectively), clear_cofactor :327-349
// (Budroni-Pintore: [x^2-x-1]P + [x-1]psi(P) + psi^2(2P)) and Hash_to_G2 :388-396.
// The output is only canonical after to-affine; everything here is inversion-free.
#pragma once
#include "ec.cuh"
#include "sha256.cuh"

namespace bls {

// SSWU: u -> Jacobian point on E2'
BLS_NOINLINE void sswu_g2(g2_jac &out, const fp2 &u) {
    fp2 tv1, tv2, x1n, x1d, d2, d3, gxn, t, rs, y, xn, one;
    f_set_one(one);
    fp2_sqr(tv1, u);
    fp2_mul(tv1, tv1, SSWU_Z);             // Z u^2
    fp2_sqr(tv2, tv1);
    fp2_add(tv2, tv2, tv1);                // Z^2 u^4 + Z u^2
    fp2_add(x1n, tv2, one);
    fp2_mul(x1n, x1n, SSWU_B);             // B (tv2 + 1)
    if (fp2_is_zero(tv2)) {
        x1d = SSWU_ZA;                     // exceptional case: x1 = B / (Z A)
    } else {
        fp2_mul(x1d, tv2, SSWU_NEG_A);     // -A tv2
    }
    fp2_sqr(d2, x1d);
    fp2_mul(d3, d2, x1d);
    fp2_sqr(gxn, x1n);
    fp2_mul(t, d2, SSWU_A);
    fp2_add(gxn, gxn, t);
    fp2_mul(gxn, gxn, x1n);
    fp2_mul(t, d3, SSWU_B);
    fp2_add(gxn, gxn, t);                  // x1n^3 + A x1n x1d^2 + B x1d^3   (g(x1) = gxn / d3)
    fp2_mul(t, gxn, d3);
    bool is_sq = fp2_rsqrt_or_z(rs, t);
    if (is_sq) {
        fp2_mul(y, gxn, rs);               // sqrt(gxn/d3) = gxn / sqrt(gxn d3)
        xn = x1n;
    } else {
        fp2_mul(y, gxn, rs);
        fp2_mul(y, y, SSWU_Z);             // sqrt(Z gxn/d3) = Z gxn / sqrt(Z gxn d3)
        fp2_mul(y, y, tv1);
        fp2_mul(y, y, u);                  // y2 = Z u^3 * sqrt(Z g(x1))
        fp2_mul(xn, tv1, x1n);             // x2 = Z u^2 x1
    }
    bool flip = fp2_sgn0(u) != fp2_sgn0(y);
    fp2_cneg(y, y, flip);
    // affine (xn/x1d, y) -> Jacobian with Z = x1d
    fp2_mul(out.x, xn, x1d);
    fp2_mul(out.y, y, d3);
    out.z = x1d;
}

// 3-isogeny E2' -> E2 on Jacobian coordinates, no inversion:
//   W = Z^2;  XN = sum k1i X^i W^(3-i), XD = sum k2i X^i W^(2-i), YN, YD likewise (monic XD, YD)
//   x' = XN / (XD W),  y' = (Y/Z^3) YN / YD ;  Z' = Z XD YD, X' = XN XD YD^2, Y' = Y YN XD^3 YD^2
BLS_NOINLINE void iso3_g2(g2_jac &out, const g2_jac &p) {
    fp2 W, W2, W3, X2, X3, XN, XD, YN, YD, t, X2W, XW2, XW;
    fp2_sqr(W, p.z);
    fp2_sqr(W2, W);
    fp2_mul(W3, W2, W);
    fp2_sqr(X2, p.x);
    fp2_mul(X3, X2, p.x);
    fp2_mul(X2W, X2, W);
    fp2_mul(XW, p.x, W);
    fp2_mul(XW2, p.x, W2);
    // XN
    fp2_mul(XN, X3, ISO3_XNUM[3]);
    fp2_mul(t, X2W, ISO3_XNUM[2]); fp2_add(XN, XN, t);
    fp2_mul(t, XW2, ISO3_XNUM[1]); fp2_add(XN, XN, t);
    fp2_mul(t, W3, ISO3_XNUM[0]);  fp2_add(XN, XN, t);
    // XD (x_den * W^2)
    fp2_mul(XD, XW, ISO3_XDEN[1]);
    fp2_add(XD, XD, X2);
    fp2_mul(t, W2, ISO3_XDEN[0]);  fp2_add(XD, XD, t);
    // YN
    fp2_mul(YN, X3, ISO3_YNUM[3]);
    fp2_mul(t, X2W, ISO3_YNUM[2]); fp2_add(YN, YN, t);
    fp2_mul(t, XW2, ISO3_YNUM[1]); fp2_add(YN, YN, t);
    fp2_mul(t, W3, ISO3_YNUM[0]);  fp2_add(YN, YN, t);
    // YD
    fp2_mul(YD, X2W, ISO3_YDEN[2]);
    fp2_add(YD, YD, X3);
    fp2_mul(t, XW2, ISO3_YDEN[1]); fp2_add(YD, YD, t);
    fp2_mul(t, W3, ISO3_YDEN[0]);  fp2_add(YD, YD, t);
    fp2 XDYD, YD2, XD2;
    fp2_mul(XDYD, XD, YD);
    fp2_sqr(YD2, YD);
    fp2_sqr(XD2, XD);
    fp2_mul(t, XN, XDYD);
    fp2_mul(out.x, t, YD);                 // XN XD YD^2
    fp2_mul(t, p.y, YN);
    fp2_mul(t, t, XD2);
    fp2_mul(t, t, XDYD);
    fp2_mul(out.y, t, YD);                 // Y YN XD^3 YD^2
    fp2_mul(out.z, p.z, XDYD);
}

BLS_FN void g2_psi(g2_jac &r, const g2_jac &p) {
    fp2 t;
    fp2_conj(t, p.x); fp2_mul(r.x, t, PSI_CX);
    fp2_conj(t, p.y); fp2_mul(r.y, t, PSI_CY);
    fp2_conj(r.z, p.z);
}

BLS_FN void g2_psi2(g2_jac &r, const g2_jac &p) {
    fp2_mul_fp(r.x, p.x, PSI2_CX);
    fp2_neg(r.y, p.y);
    r.z = p.z;
}

// r = [x]P, x = -0xd201000000010000 (63 doublings + 5 additions, then negate)
BLS_NOINLINE void g2_mul_by_x(g2_jac &r, const g2_jac &p) {
    g2_jac acc = p;
    const uint64_t z = BLS_Z_ABS;
    for (int i = 62; i >= 0; i--) {
        pt_dbl(acc, acc);
        if ((z >> i) & 1) pt_add(acc, acc, p);
    }
    pt_neg(r, acc);
}

BLS_NOINLINE void g2_clear_cofactor(g2_jac &out, const g2_jac &p) {
    g2_jac t1, t2, t3, n;
    g2_mul_by_x(t1, p);          // [x]P
    g2_psi(t2, p);               // psi(P)
    pt_dbl(t3, p);
    g2_psi2(t3, t3);             // psi^2(2P)
    pt_neg(n, t2);
    pt_add(t3, t3, n);           // psi^2(2P) - psi(P)
    pt_add(t2, t1, t2);          // [x]P + psi(P)
    g2_mul_by_x(t2, t2);         // [x^2]P + [x]psi(P)
    pt_add(t3, t3, t2);
    pt_neg(n, t1);
    pt_add(t3, t3, n);           // - [x]P
    pt_neg(n, p);
    pt_add(out, t3, n);          // - P
}

// u0,u1 -> Jacobian point in G2 (not yet affine)
BLS_FN void map_to_g2(g2_jac &out, const fp2 &u0, const fp2 &u1) {
    g2_jac q0, q1;
    sswu_g2(q0, u0);
    sswu_g2(q1, u1);
    pt_add(q0, q0, q1, &SSWU_A);
    iso3_g2(q0, q0);
    g2_clear_cofactor(out, q0);
}

BLS_FN void hash_to_g2_jac(g2_jac &out, const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len) {
    fp2 u0, u1;
    hash_to_field_fp2x2(u0, u1, msg, msg_len, dst, dst_len);
    map_to_g2(out, u0, u1);
}

#ifdef __CUDACC__
// H(msg) for the 32-byte message of a SignatureSet under DST_ETH2 (bls_sig_min_pubkey.nim:31)
BLS_FN void hash_to_g2_jac_eth2(g2_jac &out, const uint8_t *msg32) {
    fp2 u0, u1;
    hash_to_field_fp2x2_eth2(u0, u1, msg32);
    map_to_g2(out, u0, u1);
}
#endif

// Zcash compressed encoding of an affine G2 point (e2.c:231-253): x.im || x.re big-endian, flags in byte 0
BLS_FN void g2_compress(uint8_t *out, const g2_aff &p) {
    if (aff_is_inf(p)) {
        out[0] = 0xc0;
        for (int i = 1; i < 96; i++) out[i] = 0;
        return;
    }
    fp c0, c1;
    fp_from_mont(c0, p.x.c0);
    fp_from_mont(c1, p.x.c1);
    for (int i = 0; i < 48; i++) {
        out[i] = (uint8_t)(c1.l[(47 - i) >> 2] >> (8 * ((47 - i) & 3)));
        out[48 + i] = (uint8_t)(c0.l[(47 - i) >> 2] >> (8 * ((47 - i) & 3)));
    }
    out[0] |= 0x80;
    if (fp2_is_lexically_largest(p.y)) out[0] |= 0x20;
}

}  // namespace bls
