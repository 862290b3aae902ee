// fp2h.cuh — Fp2 split over a PAIR of lanes: the even lane of a pair holds the real coordinate, the odd lane the
// imaginary one ("half" elements).  The mid-size route (a few thousand to a few ten thousand signature sets) gives every
// set two lanes instead of one thread (too few warps to hide the carry-chain latency) or one warp (lanes idle in the
// dataflow programs): the G2 work of a set — hash_to_G2 (map_to_g2.c:388-396) and the Miller-loop line evaluations
// (pairing.c:14-135) — is Fp2 arithmetic, and an Fp2 product splits over two lanes with NO extra work:
//
//     re = a0 b0 - a1 b1,  im = a0 b1 + a1 b0      one two-term Montgomery dot product per lane: 2 x 144 + 156 = 444
//                                                  multiply-adds, against 3 x 300 = 900 for Karatsuba on one thread
//     a^2: re = (a0 + a1)(a0 - a1), im = (2 a0) a1  one 300-IMAD product per lane, against two on one thread
//
// so the latency of a set's serial chain halves, the per-lane stack halves, and the total work stays the same.
// Partner values move by __shfl_xor_sync over a two-lane mask, so pairs of one warp may follow different branches (the
// point formulas test for infinity) as long as BOTH lanes of a pair take the same one — every predicate below is
// pair-uniform by construction.  The type plugs into the point templates of ec.cuh through the same f_* interface as fp
// and fp2 (pt_dbl, pt_add, ... over jac_t<fp2h>).
#pragma once
#include "tower.cuh"

namespace bls {
#ifdef __CUDACC__

struct fp2h { fp v; };

__device__ __forceinline__ bool h_odd() { return (threadIdx.x & 1) != 0; }
__device__ __forceinline__ uint32_t h_mask() { return 3u << (threadIdx.x & 30); }

// the partner lane's half
__device__ __forceinline__ void h_xchg(fp &r, const fp &a) {
    const uint32_t m = h_mask();
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_xor_sync(m, a.l[i], 1);
}
__device__ __forceinline__ bool h_both(bool mine) {
    const uint32_t m = h_mask();
    return mine & (__shfl_xor_sync(m, (int)mine, 1) != 0);
}
// this lane's half of a full element
__device__ __forceinline__ void h_take(fp2h &r, const fp2 &a) { r.v = h_odd() ? a.c1 : a.c0; }
// full element from the pair's halves (both lanes get it)
__device__ __forceinline__ void h_full(fp2 &r, const fp2h &a) {
    fp o;
    h_xchg(o, a.v);
    const bool odd = h_odd();
    fp_select(r.c0, odd, o, a.v);
    fp_select(r.c1, odd, a.v, o);
}

// r = x0 y0 + x1 y1 (Montgomery, fully reduced): the multiplier of fp.cuh with two products accumulated per row before
// the row's reduction (same bounds as fp_dot<K> in acc_team.cuh: inputs < p, result < 2p before the final subtraction)
__device__ __forceinline__ void fp_dot2(fp &r, const fp &x0, const fp &y0, const fp &x1, const fp &y1) {
#ifdef __CUDA_ARCH__
    uint32_t E[13], O[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { E[i] = 0; O[i] = 0; }
    E[12] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        if (i == 0) {
            mad6(O, x0.l[1], x0.l[3], x0.l[5], x0.l[7], x0.l[9], x0.l[11], y0.l[0]);
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
            mad6_stray(O, E[0], s, x0.l[1], x0.l[3], x0.l[5], x0.l[7], x0.l[9], x0.l[11], y0.l[i]);
        }
        mad6_top(E, x0.l[0], x0.l[2], x0.l[4], x0.l[6], x0.l[8], x0.l[10], y0.l[i]);
        mad6(O, x1.l[1], x1.l[3], x1.l[5], x1.l[7], x1.l[9], x1.l[11], y1.l[i]);
        mad6_top(E, x1.l[0], x1.l[2], x1.l[4], x1.l[6], x1.l[8], x1.l[10], y1.l[i]);
        const uint32_t m = E[0] * BLS_N0;
        mad6(O, P32(1), P32(3), P32(5), P32(7), P32(9), P32(11), m);
        mad6_top(E, P32(0), P32(2), P32(4), P32(6), P32(8), P32(10), m);
    }
    add12(O, E + 1);
    reduce_once12(r.l, O);
#endif
}

// ---- field operations on halves -------------------------------------------------------------------------------------
BLS_NOINLINE void h_mul(fp2h &r, const fp2h &a, const fp2h &b) {
    fp ao = a.v, bo = b.v, ap, bp, nap, x0, x1, t;
    h_xchg(ap, ao);
    h_xchg(bp, bo);
    fp_neg(nap, ap);
    const bool odd = h_odd();
    // even lane: a0 b0 + (-a1) b1       odd lane: a0 b1 + a1 b0   (own = a1, b1; partner = a0, b0)
    fp_select(x0, odd, ap, ao);
    fp_select(x1, odd, ao, nap);
    fp_dot2(t, x0, bo, x1, bp);
    r.v = t;
}
BLS_NOINLINE void h_sqr(fp2h &r, const fp2h &a) {
    fp ao = a.v, ap, x, y, d, t;
    h_xchg(ap, ao);
    const bool odd = h_odd();
    fp_select(t, odd, ao, ap);
    fp_add(x, ao, t);                   // even: a0 + a1    odd: 2 a1
    fp_sub(d, ao, ap);                  // even: a0 - a1
    fp_select(y, odd, ap, d);           // odd: a0
    fp_mul(t, x, y);
    r.v = t;
}
// r = a * k, k in Fp (both lanes hold the same k)
__device__ __forceinline__ void h_mul_fp(fp2h &r, const fp2h &a, const fp &k) { fp_mul_ni(r.v, a.v, k); }
// r = a * (1 + u) = (a0 - a1) + (a0 + a1) u
__device__ __forceinline__ void h_mul_xi(fp2h &r, const fp2h &a) {
    fp ap, s, d;
    h_xchg(ap, a.v);
    fp_add(s, a.v, ap);
    fp_sub(d, a.v, ap);                 // even lane: a0 - a1
    fp_select(r.v, h_odd(), s, d);
}
__device__ __forceinline__ void h_conj(fp2h &r, const fp2h &a) { fp_cneg(r.v, a.v, h_odd()); }
__device__ __forceinline__ void h_cneg(fp2h &r, const fp2h &a, bool c) { fp_cneg(r.v, a.v, c); }
__device__ __forceinline__ void h_mul3(fp2h &r, const fp2h &a) { fp t; fp_dbl(t, a.v); fp_add(r.v, t, a.v); }
// 1 / a (0 -> 0): conj(a) / (a0^2 + a1^2); both lanes run the same Fp inversion
BLS_NOINLINE void h_inv(fp2h &r, const fp2h &a) {
    fp n, np, t;
    fp_sqr_ni(n, a.v);
    h_xchg(np, n);
    fp_add(n, n, np);
    fp_inv(n, n);
    fp_mul_ni(t, a.v, n);
    fp_cneg(r.v, t, h_odd());
}

// uniform field interface for the point templates of ec.cuh
__device__ __forceinline__ void f_add(fp2h &r, const fp2h &a, const fp2h &b) { fp_add(r.v, a.v, b.v); }
__device__ __forceinline__ void f_sub(fp2h &r, const fp2h &a, const fp2h &b) { fp_sub(r.v, a.v, b.v); }
__device__ __forceinline__ void f_dbl(fp2h &r, const fp2h &a) { fp_add(r.v, a.v, a.v); }
__device__ __forceinline__ void f_neg(fp2h &r, const fp2h &a) { fp_neg(r.v, a.v); }
__device__ __forceinline__ void f_mul(fp2h &r, const fp2h &a, const fp2h &b) { h_mul(r, a, b); }
__device__ __forceinline__ void f_sqr(fp2h &r, const fp2h &a) { h_sqr(r, a); }
__device__ __forceinline__ bool f_is_zero(const fp2h &a) { return h_both(fp_is_zero(a.v)); }
__device__ __forceinline__ bool f_eq(const fp2h &a, const fp2h &b) { return h_both(fp_eq(a.v, b.v)); }
__device__ __forceinline__ void f_set_zero(fp2h &r) { fp_set_zero(r.v); }
__device__ __forceinline__ void f_set_one(fp2h &r) { if (h_odd()) fp_set_zero(r.v); else r.v = FP_ONE; }
__device__ __forceinline__ void f_inv(fp2h &r, const fp2h &a) { h_inv(r, a); }
__device__ __forceinline__ void f_inv_vt(fp2h &r, const fp2h &a) { h_inv(r, a); }

typedef jac_t<fp2h> g2h_jac;
typedef aff_t<fp2h> g2h_aff;

#endif  // __CUDACC__
}  // namespace bls
