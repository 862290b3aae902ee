// acc_team.cuh — Miller-loop accumulation by TEAMS of six lanes (one lane per Fp2 coefficient of the Fp12 value).
//
// Restates what miller_loop_n does with its accumulator (vendor/blst/src/pairing.c:253-260: sqr_fp12 + one
// mul_by_xy00z0_fp12 per line, fp12_tower.c:372-397, :496-528) on a different decomposition:
//
//   Fp12 = Fp2[w]/(w^6 - xi), xi = 1+u;  f = sum_k f_k w^k.  In the tower's memory order a[j][i] (w^j v^i) the
//   coefficient of w^k is a[k&1][k>>1].  A line is l0 + l1 w^2 + l2 w^3, so
//       (f * line)_k = f_k l0 + xi^[k<2] f_(k-2) l1 + xi^[k<3] f_(k-3) l2            (indices mod 6)
//   and (f^2)_k is a sum of at most four products f_i f_j (i+j = k mod 6, cross terms doubled, xi when i+j >= 6).
//
// Lane k of a team owns f_k.  Each output coefficient is ONE Fp2 "dot product" of K terms, evaluated as three Fp
// dot products (Karatsuba over the whole sum):  D0 = sum x0 y0, D1 = sum x1 y1, D2 = sum (x0+x1)(y0+y1),
// re = D0 - D1, im = D2 - D0 - D1.  fp_dot<K> is the Montgomery multiplier of fp.cuh with K products accumulated
// per row before the row's reduction (K*144 + 156 multiply-adds instead of K*300), so a line costs
// 6 * 3 * (3*144 + 156) = 10 584 IMAD against 13 * 900 = 11 700 for the sparse Karatsuba product — with no Fp6/Fp12
// temporaries at all.  The team's coefficients (and the derived values c0+c1, c0-c1, 2c0 that make multiplication by
// xi an address choice) and the current line live in shared memory, 1 872 bytes per team; nothing spills to local
// memory.  Five teams per warp (lanes 30, 31 idle), one __syncwarp between a team's reads and its writes.
#pragma once
#include "pairing.cuh"

namespace bls {

#define ACC_TPW 5                       // teams per warp
#define ACC_BS 128                      // threads per block
#define ACC_TPB (ACC_TPW * ACC_BS / 32) // teams per block

// per-coefficient variants kept in shared memory
enum { FV_C0 = 0, FV_C1 = 1, FV_S = 2, FV_D = 3, FV_DD = 4 };   // c0, c1, c0+c1, c0-c1, 2*c0
struct __align__(16) acc_team_sm {
    fp f[6][5];
    fp l[9];                             // l0.c0 l0.c1 l1.c0 l1.c1 l2.c0 l2.c1, then l0.c0+l0.c1, l1.., l2..
};

#ifdef __CUDACC__
// r = sum_t a[t]*b[t] / R mod p, fully reduced.  a[t] in registers, b[t] streamed from (shared) memory four limbs
// at a time.  Bounds: inputs < p; the running total stays below (K+1) p 2^32 < 2^416 for K <= 6 (13-limb E, 12-limb O
// as in fp_mul) and the result below p (1 + K p / 2^384) < 2p, so one conditional subtraction finishes.
template <int K>
__device__ __forceinline__ void fp_dot(fp &r, const uint32_t (&a)[K][12], const fp *const (&b)[K]) {
#ifdef __CUDA_ARCH__          // (device pass only: the PTX helpers of fp.cuh do not exist in the host pass)
    uint32_t E[13], O[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { E[i] = 0; O[i] = 0; }
    E[12] = 0;
    uint32_t bl[K][4];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        if ((i & 3) == 0) {
#pragma unroll
            for (int t = 0; t < K; t++) {
                uint4 v = ((const uint4 *)b[t])[i >> 2];
                bl[t][0] = v.x; bl[t][1] = v.y; bl[t][2] = v.z; bl[t][3] = v.w;
            }
        }
        if (i == 0) {
            mad6(O, a[0][1], a[0][3], a[0][5], a[0][7], a[0][9], a[0][11], bl[0][0]);
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
            mad6_stray(O, E[0], s, a[0][1], a[0][3], a[0][5], a[0][7], a[0][9], a[0][11], bl[0][i & 3]);
        }
        mad6_top(E, a[0][0], a[0][2], a[0][4], a[0][6], a[0][8], a[0][10], bl[0][i & 3]);
#pragma unroll
        for (int t = 1; t < K; t++) {
            mad6(O, a[t][1], a[t][3], a[t][5], a[t][7], a[t][9], a[t][11], bl[t][i & 3]);
            mad6_top(E, a[t][0], a[t][2], a[t][4], a[t][6], a[t][8], a[t][10], bl[t][i & 3]);
        }
        const uint32_t m = E[0] * BLS_N0;
        mad6(O, P32(1), P32(3), P32(5), P32(7), P32(9), P32(11), m);
        mad6_top(E, P32(0), P32(2), P32(4), P32(6), P32(8), P32(10), m);
    }
    add12(O, E + 1);
    reduce_once12(r.l, O);
#endif
}

__device__ __forceinline__ void load_fp_regs(uint32_t (&dst)[12], const fp *src) {
    const uint4 *s = (const uint4 *)src;
    uint4 v0 = s[0], v1 = s[1], v2 = s[2];
    dst[0] = v0.x; dst[1] = v0.y; dst[2] = v0.z; dst[3] = v0.w;
    dst[4] = v1.x; dst[5] = v1.y; dst[6] = v1.z; dst[7] = v1.w;
    dst[8] = v2.x; dst[9] = v2.y; dst[10] = v2.z; dst[11] = v2.w;
}

// three-term dot product, operands by address (a: register side, b: streamed side)
__device__ __noinline__ void team_dot3(fp &r, const fp *a0, const fp *a1, const fp *a2, const fp *b0, const fp *b1, const fp *b2) {
    uint32_t a[3][12];
    load_fp_regs(a[0], a0);
    load_fp_regs(a[1], a1);
    load_fp_regs(a[2], a2);
    const fp *const b[3] = {b0, b1, b2};
    fp t;
    fp_dot<3>(t, a, b);
    r = t;
}

// four-term dot product for the squaring; term t's register-side operand is doubled when bit t of dbl_mask is set
// and replaced by zero when bit t of zero_mask is set
__device__ __noinline__ void team_dot4(fp &r, const fp *const *ap, const fp *const *bp, uint32_t dbl_mask, uint32_t zero_mask) {
    uint32_t a[4][12];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        fp x;
        load_fp_regs(x.l, ap[t]);
        fp x2;
        fp_dbl(x2, x);
        const bool d = (dbl_mask >> t) & 1, z = (zero_mask >> t) & 1;
#pragma unroll
        for (int i = 0; i < 12; i++) a[t][i] = z ? 0u : (d ? x2.l[i] : x.l[i]);
    }
    const fp *const b[4] = {bp[0], bp[1], bp[2], bp[3]};
    fp t;
    fp_dot<4>(t, a, b);
    r = t;
}

// derived variants of a coefficient (c0, c1) -> shared memory
__device__ __forceinline__ void team_store_coeff(acc_team_sm &T, int k, const fp &c0, const fp &c1) {
    fp s, d, dd;
    fp_add(s, c0, c1);
    fp_sub(d, c0, c1);
    fp_dbl(dd, c0);
    T.f[k][FV_C0] = c0;
    T.f[k][FV_C1] = c1;
    T.f[k][FV_S] = s;
    T.f[k][FV_D] = d;
    T.f[k][FV_DD] = dd;
}

// operand xi^x * f_j as (re, im, re+im) addresses: xi*(c0 + c1 u) = (c0 - c1) + (c0 + c1) u, sum = 2 c0
__device__ __forceinline__ void team_operand(const acc_team_sm &T, int j, bool x, const fp *&re, const fp *&im, const fp *&sum) {
    re = x ? &T.f[j][FV_D] : &T.f[j][FV_C0];
    im = x ? &T.f[j][FV_S] : &T.f[j][FV_C1];
    sum = x ? &T.f[j][FV_DD] : &T.f[j][FV_S];
}

// f <- f * line (line staged in T.l); lane k computes coefficient k.  `mask` = lanes of the warp that take part.
__device__ __forceinline__ void team_mul_line(acc_team_sm &T, int k, uint32_t mask) {
    const fp *x0[3], *x1[3], *sx[3];
    team_operand(T, k, false, x0[0], x1[0], sx[0]);
    team_operand(T, (k + 4) % 6, k < 2, x0[1], x1[1], sx[1]);
    team_operand(T, (k + 3) % 6, k < 3, x0[2], x1[2], sx[2]);
    fp D0, D1, D2, c0, c1;
    team_dot3(D0, &T.l[0], &T.l[2], &T.l[4], x0[0], x0[1], x0[2]);
    team_dot3(D1, &T.l[1], &T.l[3], &T.l[5], x1[0], x1[1], x1[2]);
    team_dot3(D2, &T.l[6], &T.l[7], &T.l[8], sx[0], sx[1], sx[2]);
    fp_sub(c0, D0, D1);
    fp_sub(c1, D2, D0);
    fp_sub(c1, c1, D1);
    __syncwarp(mask);                   // every lane has read the old coefficients
    team_store_coeff(T, k, c0, c1);
    __syncwarp(mask);
}

// f <- f^2.  (f^2)_k = sum over {i,j}, i+j = k mod 6: squares once, cross terms twice, xi when i+j >= 6.
// Per lane at most four terms; the tables give, for coefficient k, the left index, right index, xi flag (on the
// left operand), the doubling mask and the unused-term mask.
__device__ __forceinline__ void team_square(acc_team_sm &T, int k, uint32_t mask) {
    //                        k:      0            1            2            3            4            5
    const int LI[6][4] = {{0, 3, 1, 2}, {0, 2, 3, 0}, {1, 4, 0, 3}, {0, 1, 4, 0}, {2, 5, 0, 1}, {0, 1, 2, 0}};
    const int RI[6][4] = {{0, 3, 5, 4}, {1, 5, 4, 0}, {1, 4, 2, 5}, {3, 2, 5, 0}, {2, 5, 4, 3}, {5, 4, 3, 0}};
    const int XI[6][4] = {{0, 1, 1, 1}, {0, 1, 1, 0}, {0, 1, 0, 1}, {0, 0, 1, 0}, {0, 1, 0, 0}, {0, 0, 0, 0}};
    const uint32_t DBL[6] = {0xc, 0x7, 0xc, 0x7, 0xc, 0x7}, ZERO[6] = {0, 8, 0, 8, 0, 8};
    const fp *a0[4], *a1[4], *as[4], *b0[4], *b1[4], *bs[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        team_operand(T, RI[k][t], false, a0[t], a1[t], as[t]);         // register side: plain f_j (doubled in team_dot4)
        team_operand(T, LI[k][t], XI[k][t] != 0, b0[t], b1[t], bs[t]); // streamed side: xi^x f_i
    }
    fp D0, D1, D2, c0, c1;
    team_dot4(D0, a0, b0, DBL[k], ZERO[k]);
    team_dot4(D1, a1, b1, DBL[k], ZERO[k]);
    team_dot4(D2, as, bs, DBL[k], ZERO[k]);
    fp_sub(c0, D0, D1);
    fp_sub(c1, D2, D0);
    fp_sub(c1, c1, D1);
    __syncwarp(mask);
    team_store_coeff(T, k, c0, c1);
    __syncwarp(mask);
}
#endif  // __CUDACC__

}  // namespace bls
