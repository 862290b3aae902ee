// pair_route.cuh — the per-set G2 stages on lane PAIRS (fp2h.cuh): H(m_i) and the 68 Miller-loop line triples.
//
// Same results, bit for bit, as the one-thread-per-set forms in h2c.cuh / pairing.cuh (every value that leaves a stage
// is canonical: fully reduced field elements of the same formulas):
//   hash_to_g2_pair   restates Hash_to_G2 (map_to_g2.c:388-396) like hash_to_g2_jac: XMD, two SSWU maps, sum on E2',
//                     3-isogeny, cofactor clearing.  The Fp2 stretches run on the pair; the two Fp square-root chains of
//                     the two maps — the part of the hash that does not split — run one map per lane, side by side.
//   miller_lines_pair restates miller_lines() (pairing.cuh; line_dbl/line_add/line_by_Px2 of pairing.c:14-135).
#pragma once
#include "fp2h.cuh"
#include "h2c.cuh"
#include "pairing.cuh"

namespace bls {
#ifdef __CUDACC__

__device__ __forceinline__ void hk(fp2h &r, const fp2 &c) { h_take(r, c); }      // this lane's half of a constant

// sgn0 of an Fp2 element held as halves (RFC 9380 4.1): parity(re) unless re == 0, then parity(im)
__device__ __forceinline__ uint32_t h_sgn0(const fp2h &a) {
    const uint32_t m = h_mask();
    const uint32_t par = fp_parity(a.v), zero = fp_is_zero(a.v) ? 1u : 0u;
    const uint32_t mine = par | (zero << 1);
    const uint32_t other = __shfl_xor_sync(m, mine, 1);
    const uint32_t re = h_odd() ? other : mine, im = h_odd() ? mine : other;
    return (re & 2) ? (im & 1) : (re & 1);
}

// state of one SSWU map between its two halves (before / after the square root)
struct sswu_mid { fp2h u, tv1, x1n, x1d, d3, gxn, t; };

// first half of sswu_g2 (h2c.cuh): everything up to the argument t = gxn * d3 of the reciprocal square root
BLS_NOINLINE void sswu_pre_pair(sswu_mid &s, const fp2h &u) {
    fp2h tv2, d2, t, one, c;
    f_set_one(one);
    s.u = u;
    h_sqr(s.tv1, u);
    hk(c, SSWU_Z);
    h_mul(s.tv1, s.tv1, c);                 // Z u^2
    h_sqr(tv2, s.tv1);
    f_add(tv2, tv2, s.tv1);                 // Z^2 u^4 + Z u^2
    f_add(s.x1n, tv2, one);
    hk(c, SSWU_B);
    h_mul(s.x1n, s.x1n, c);                 // B (tv2 + 1)
    if (f_is_zero(tv2)) {
        hk(s.x1d, SSWU_ZA);                 // exceptional case: x1 = B / (Z A)
    } else {
        hk(c, SSWU_NEG_A);
        h_mul(s.x1d, tv2, c);               // -A tv2
    }
    h_sqr(d2, s.x1d);
    h_mul(s.d3, d2, s.x1d);
    h_sqr(s.gxn, s.x1n);
    hk(c, SSWU_A);
    h_mul(t, d2, c);
    f_add(s.gxn, s.gxn, t);
    h_mul(s.gxn, s.gxn, s.x1n);
    hk(c, SSWU_B);
    h_mul(t, s.d3, c);
    f_add(s.gxn, s.gxn, t);                 // x1n^3 + A x1n x1d^2 + B x1d^3
    h_mul(s.t, s.gxn, s.d3);
}

// second half: rs = 1/sqrt(t) (is_sq) or 1/sqrt(Z t) -> Jacobian point on E2'
BLS_NOINLINE void sswu_post_pair(g2h_jac &out, const sswu_mid &s, const fp2h &rs, bool is_sq) {
    fp2h y, xn, c;
    h_mul(y, s.gxn, rs);
    if (is_sq) {
        xn = s.x1n;
    } else {
        hk(c, SSWU_Z);
        h_mul(y, y, c);
        h_mul(y, y, s.tv1);
        h_mul(y, y, s.u);                   // y2 = Z u^3 * sqrt(Z g(x1))
        h_mul(xn, s.tv1, s.x1n);            // x2 = Z u^2 x1
    }
    const bool flip = h_sgn0(s.u) != h_sgn0(y);
    h_cneg(y, y, flip);
    h_mul(out.x, xn, s.x1d);
    h_mul(out.y, y, s.d3);
    out.z = s.x1d;
}

// 3-isogeny E2' -> E2 on Jacobian coordinates (iso3_g2 of h2c.cuh on halves)
BLS_NOINLINE void iso3_pair(g2h_jac &out, const g2h_jac &p) {
    fp2h W, W2, W3, X2, X3, XN, XD, YN, YD, t, X2W, XW2, XW, c;
    h_sqr(W, p.z);
    h_sqr(W2, W);
    h_mul(W3, W2, W);
    h_sqr(X2, p.x);
    h_mul(X3, X2, p.x);
    h_mul(X2W, X2, W);
    h_mul(XW, p.x, W);
    h_mul(XW2, p.x, W2);
    hk(c, ISO3_XNUM[3]); h_mul(XN, X3, c);
    hk(c, ISO3_XNUM[2]); h_mul(t, X2W, c); f_add(XN, XN, t);
    hk(c, ISO3_XNUM[1]); h_mul(t, XW2, c); f_add(XN, XN, t);
    hk(c, ISO3_XNUM[0]); h_mul(t, W3, c);  f_add(XN, XN, t);
    hk(c, ISO3_XDEN[1]); h_mul(XD, XW, c);
    f_add(XD, XD, X2);
    hk(c, ISO3_XDEN[0]); h_mul(t, W2, c);  f_add(XD, XD, t);
    hk(c, ISO3_YNUM[3]); h_mul(YN, X3, c);
    hk(c, ISO3_YNUM[2]); h_mul(t, X2W, c); f_add(YN, YN, t);
    hk(c, ISO3_YNUM[1]); h_mul(t, XW2, c); f_add(YN, YN, t);
    hk(c, ISO3_YNUM[0]); h_mul(t, W3, c);  f_add(YN, YN, t);
    hk(c, ISO3_YDEN[2]); h_mul(YD, X2W, c);
    f_add(YD, YD, X3);
    hk(c, ISO3_YDEN[1]); h_mul(t, XW2, c); f_add(YD, YD, t);
    hk(c, ISO3_YDEN[0]); h_mul(t, W3, c);  f_add(YD, YD, t);
    fp2h XDYD, YD2, XD2;
    h_mul(XDYD, XD, YD);
    h_sqr(YD2, YD);
    h_sqr(XD2, XD);
    h_mul(t, XN, XDYD);
    h_mul(out.x, t, YD);                    // XN XD YD^2
    h_mul(t, p.y, YN);
    h_mul(t, t, XD2);
    h_mul(t, t, XDYD);
    h_mul(out.y, t, YD);                    // Y YN XD^3 YD^2
    h_mul(out.z, p.z, XDYD);
}

__device__ __forceinline__ void psi_pair(g2h_jac &r, const g2h_jac &p) {
    fp2h t, c;
    h_conj(t, p.x); hk(c, PSI_CX); h_mul(r.x, t, c);
    h_conj(t, p.y); hk(c, PSI_CY); h_mul(r.y, t, c);
    h_conj(r.z, p.z);
}
__device__ __forceinline__ void psi2_pair(g2h_jac &r, const g2h_jac &p) {
    h_mul_fp(r.x, p.x, PSI2_CX);
    f_neg(r.y, p.y);
    r.z = p.z;
}
// r = [x]P, x = -0xd201000000010000
BLS_NOINLINE void mul_by_x_pair(g2h_jac &r, const g2h_jac &p) {
    g2h_jac acc = p;
    const uint64_t z = BLS_Z_ABS;
    for (int i = 62; i >= 0; i--) {
        pt_dbl(acc, acc);
        if ((z >> i) & 1) pt_add(acc, acc, p);
    }
    pt_neg(r, acc);
}
// clear_cofactor (map_to_g2.c:327-349), the combination of g2_clear_cofactor in h2c.cuh
BLS_NOINLINE void clear_cofactor_pair(g2h_jac &out, const g2h_jac &p) {
    g2h_jac t1, t2, t3, n;
    mul_by_x_pair(t1, p);
    psi_pair(t2, p);
    pt_dbl(t3, p);
    psi2_pair(t3, t3);
    pt_neg(n, t2);
    pt_add(t3, t3, n);
    pt_add(t2, t1, t2);
    mul_by_x_pair(t2, t2);
    pt_add(t3, t3, t2);
    pt_neg(n, t1);
    pt_add(t3, t3, n);
    pt_neg(n, p);
    pt_add(out, t3, n);
}

// Everything of H(msg) up to the cofactor clearing — XMD, the two SSWU maps, their sum on E2', the 3-isogeny — as halves
// of a Jacobian point on E2.  Both lanes of the pair hash the message (the XMD is 18 SHA-256 blocks, ~1 % of the work)
// and each keeps its halves of u0, u1.
BLS_NOINLINE void hash_map_to_e2_pair(g2h_jac &out, const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len) {
    const bool odd = h_odd();
    fp2h u0, u1;
    {
        uint32_t xmd[64];
        if (dst) expand_message_xmd_256(xmd, msg, msg_len, dst, dst_len);
        else expand_message_xmd_256_eth2(xmd, msg);      // dst == nullptr: 32-byte message under DST_ETH2
        fp_from_be64(u0.v, xmd + (odd ? 16 : 0));
        fp_from_be64(u1.v, xmd + (odd ? 48 : 32));
    }
    sswu_mid s0, s1;
    sswu_pre_pair(s0, u0);
    sswu_pre_pair(s1, u1);
    // the two reciprocal square roots are Fp chains that do not split over lanes: the even lane takes map 0's whole
    // argument, the odd lane map 1's, and both run the one-thread routine of fpx.cuh side by side
    fp send, recv;
    fp_select(send, odd, s0.t.v, s1.t.v);           // even lane gives away its half of t1, odd lane its half of t0
    h_xchg(recv, send);
    fp2 arg, rs;
    fp_select(arg.c0, odd, recv, s0.t.v);           // even: t0 = (own, recv)      odd: t1 = (recv, own)
    fp_select(arg.c1, odd, s1.t.v, recv);
    const bool sq_mine = fp2_rsqrt_or_z(rs, arg);
    // back to halves: rs0 = (even.rs.c0, even.rs.c1) belongs to map 0, rs1 (odd lane's) to map 1
    fp2h r0, r1;
    fp_select(send, odd, rs.c0, rs.c1);             // even lane sends im of rs0, odd lane sends re of rs1
    h_xchg(recv, send);
    fp_select(r0.v, odd, recv, rs.c0);              // even: re of rs0 (own)       odd: im of rs0 (from even)
    fp_select(r1.v, odd, rs.c1, recv);              // even: re of rs1 (from odd)  odd: im of rs1 (own)
    const uint32_t m = h_mask();
    const int sq_other = __shfl_xor_sync(m, (int)sq_mine, 1);
    const bool sq0 = odd ? (sq_other != 0) : sq_mine, sq1 = odd ? sq_mine : (sq_other != 0);
    g2h_jac q0, q1;
    sswu_post_pair(q0, s0, r0, sq0);
    sswu_post_pair(q1, s1, r1, sq1);
    fp2h a;
    hk(a, SSWU_A);
    pt_add(q0, q0, q1, &a);
    iso3_pair(out, q0);
}
// H(msg) as halves of a Jacobian point
BLS_NOINLINE void hash_to_g2_pair(g2h_jac &out, const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len) {
    g2h_jac q;
    hash_map_to_e2_pair(q, msg, msg_len, dst, dst_len);
    clear_cofactor_pair(out, q);
}
// this lane's halves of the homogeneous form (X Z : Y : Z^3) of a Jacobian point; infinity -> (0 : 1 : 0)
__device__ __forceinline__ void jac_to_hom_pair(fp2h &X, fp2h &Y, fp2h &Z, const g2h_jac &p) {
    if (pt_is_inf(p)) {
        f_set_zero(X);
        f_set_one(Y);
        f_set_zero(Z);
    } else {
        fp2h z2;
        h_mul(X, p.x, p.z);
        Y = p.y;
        h_sqr(z2, p.z);
        h_mul(Z, z2, p.z);
    }
}

// ---- Miller-loop lines on a pair (line_dbl_proj / line_add_proj / miller_lines of pairing.cuh) -----------------------
BLS_NOINLINE void line_dbl_pair(g2h_jac &T, fp2h &l0, fp2h &l1p, fp2h &l2p) {
    fp2h A2, B, Cc, E, F, H, J, t, u;
    h_sqr(B, T.y);
    h_sqr(Cc, T.z);
    h_sqr(J, T.x);
    f_add(t, T.x, T.y);
    h_sqr(A2, t);
    f_sub(A2, A2, J);
    f_sub(A2, A2, B);                       // 2XY
    f_add(t, T.y, T.z);
    h_sqr(H, t);
    f_sub(H, H, B);
    f_sub(H, H, Cc);                        // 2YZ
    h_mul_xi(E, Cc);
    f_dbl(E, E);
    f_dbl(E, E);
    h_mul3(E, E);                           // E = 12 xi Z^2
    h_mul3(F, E);                           // 3E
    f_sub(l0, B, E);
    h_mul3(l1p, J);                         // 3X^2            (times -x_P)
    l2p = H;                                // 2YZ             (times y_P)
    f_sub(t, B, F);
    h_mul(T.x, A2, t);
    f_add(t, B, F);
    h_sqr(t, t);
    f_dbl(u, E);
    h_sqr(u, u);
    h_mul3(u, u);                           // 12 E^2
    f_sub(T.y, t, u);
    h_mul(t, B, H);
    f_dbl(t, t);
    f_dbl(T.z, t);
}
BLS_NOINLINE void line_add_pair(g2h_jac &T, const g2h_aff &Q, fp2h &l0, fp2h &l1p, fp2h &l2p) {
    fp2h th, la, c, d, e, f, g, h, t;
    h_mul(t, Q.y, T.z);
    f_sub(th, T.y, t);
    h_mul(t, Q.x, T.z);
    f_sub(la, T.x, t);
    h_sqr(c, th);
    h_sqr(d, la);
    h_mul(e, la, d);
    h_mul(f, T.z, c);
    h_mul(g, T.x, d);
    f_add(h, e, f);
    f_sub(h, h, g);
    f_sub(h, h, g);
    h_mul(T.x, la, h);
    f_sub(t, g, h);
    h_mul(t, th, t);
    h_mul(g, e, T.y);
    f_sub(T.y, t, g);
    h_mul(T.z, T.z, e);
    h_mul(l0, th, Q.x);
    h_mul(t, la, Q.y);
    f_sub(l0, l0, t);
    l1p = th;
    l2p = la;
}
// this lane's 36 words of line s: (l0, l1, l2) x 12 words at word offsets 24 k + 12 odd of the 72-word triple
__device__ __forceinline__ void line_store_pair(uint32_t *dst, size_t stride, int s, const fp2h &l0, const fp2h &l1, const fp2h &l2) {
    uint32_t *d = dst + ((size_t)s * ML_LINE_WORDS + (h_odd() ? 12 : 0)) * stride;
    for (int w = 0; w < 12; w++) {
        d[(size_t)w * stride] = l0.v.l[w];
        d[(size_t)(24 + w) * stride] = l1.v.l[w];
        d[(size_t)(48 + w) * stride] = l2.v.l[w];
    }
}
// Q, P: the pair's affine inputs (Q as halves; P in both lanes); dst points at this pair's column of the line array
BLS_NOINLINE void miller_lines_pair(const g2h_aff &Q, const g1_aff &P, bool inf, uint32_t *dst, size_t stride) {
    fp2h l0, l1, l2;
    if (inf) {                                            // neutral line (1, 0, 0): the pair contributes one
        f_set_one(l0);
        f_set_zero(l1);
        f_set_zero(l2);
        for (int s = 0; s < ML_NLINES; s++) line_store_pair(dst, stride, s, l0, l1, l2);
        return;
    }
    g2h_jac T;
    T.x = Q.x;
    T.y = Q.y;
    f_set_one(T.z);
    fp npx, py = P.y;
    fp_neg(npx, P.x);
    int s = 0;
    for (int i = 62; i >= 0; i--) {
        line_dbl_pair(T, l0, l1, l2);
        h_mul_fp(l1, l1, npx);
        h_mul_fp(l2, l2, py);
        line_store_pair(dst, stride, s++, l0, l1, l2);
        if (ml_bit(i)) {
            line_add_pair(T, Q, l0, l1, l2);
            h_mul_fp(l1, l1, npx);
            h_mul_fp(l2, l2, py);
            line_store_pair(dst, stride, s++, l0, l1, l2);
        }
    }
}

#endif  // __CUDACC__
}  // namespace bls
