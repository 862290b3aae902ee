// fplin.cuh — small-coefficient linear combinations of field elements in ONE step:
//     r = sum_k m_k x_k  -  sum_k m'_k x'_k   (mod p),     up to 7 terms, magnitudes 1..15.
//
// The tail programs (fpprog.hpp) spend most of their rounds on additions: the tower formulas put three to six levels of
// a +- b between two levels of multiplications (a cyclotomic squaring: one level of 30 products, then 3 t -+ 2 z over
// sums of four or five of them).  A level costs a warp a round trip through shared memory and a barrier whatever it
// computes, so the compiler flattens every linear expression into one combination and the interpreter evaluates it here:
//   * accumulation without carries: one 64-bit accumulator per limb and sign, acc[i] += x[i] * m (twelve independent
//     IMAD.WIDE per term; 7 terms x 15 x 2^32 < 2^39 per limb),
//   * one carry pass forming D = P - N + 128 p > 0 (13 limbs),
//   * one quotient estimate from the top 64 bits, D - q p in [0, 3p), two conditional subtractions.
// Plain C++ (no PTX): the same source runs in tests/hostsim, which executes the programs on the CPU.
#pragma once
#include "fp.cuh"

namespace bls {

// limb i (0..12) of 128 p
BLS_FN constexpr uint32_t P128(int i) {
    return i == 0 ? (P32(0) << 7) : (i == 12 ? (P32(11) >> 25) : ((P32(i) << 7) | (P32(i - 1) >> 25)));
}

struct lin_acc { uint64_t P[12], N[12]; };

BLS_FN void lin_clear(lin_acc &a) {
#pragma unroll
    for (int i = 0; i < 12; i++) { a.P[i] = 0; a.N[i] = 0; }
}
// acc += x * m (one sign); m == 0 contributes nothing (inactive lanes of a round pass the zero slot and m = 0)
BLS_FN void lin_add_term(uint64_t (&A)[12], const fp &x, uint32_t m) {
#pragma unroll
    for (int i = 0; i < 12; i++) A[i] += (uint64_t)x.l[i] * m;
}

#ifdef __CUDA_ARCH__
// ---- device form: the same steps on hardware carry chains (a lone warp runs dependent instructions at one per ~4.5
// cycles, so the count of chained instructions is the cost) ----
#define BLS_R13(x) BLS_R12(x), "+r"(x[12])
// acc (13 limbs) -= b
BLS_FN void sub13(uint32_t *acc, const uint32_t *b) {
    asm("sub.cc.u32 %0,%0,%13;\n\tsubc.cc.u32 %1,%1,%14;\n\tsubc.cc.u32 %2,%2,%15;\n\tsubc.cc.u32 %3,%3,%16;\n\t"
        "subc.cc.u32 %4,%4,%17;\n\tsubc.cc.u32 %5,%5,%18;\n\tsubc.cc.u32 %6,%6,%19;\n\tsubc.cc.u32 %7,%7,%20;\n\t"
        "subc.cc.u32 %8,%8,%21;\n\tsubc.cc.u32 %9,%9,%22;\n\tsubc.cc.u32 %10,%10,%23;\n\tsubc.cc.u32 %11,%11,%24;\n\t"
        "subc.u32 %12,%12,%25;"
        : BLS_R13(acc)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]),
          "r"(b[10]), "r"(b[11]), "r"(b[12]));
}
// acc (13 limbs) += 128 p
BLS_FN void add_128p13(uint32_t *acc) {
    asm("add.cc.u32 %0,%0,%13;\n\taddc.cc.u32 %1,%1,%14;\n\taddc.cc.u32 %2,%2,%15;\n\taddc.cc.u32 %3,%3,%16;\n\t"
        "addc.cc.u32 %4,%4,%17;\n\taddc.cc.u32 %5,%5,%18;\n\taddc.cc.u32 %6,%6,%19;\n\taddc.cc.u32 %7,%7,%20;\n\t"
        "addc.cc.u32 %8,%8,%21;\n\taddc.cc.u32 %9,%9,%22;\n\taddc.cc.u32 %10,%10,%23;\n\taddc.cc.u32 %11,%11,%24;\n\t"
        "addc.u32 %12,%12,%25;"
        : BLS_R13(acc)
        : "n"(P128(0)), "n"(P128(1)), "n"(P128(2)), "n"(P128(3)), "n"(P128(4)), "n"(P128(5)), "n"(P128(6)), "n"(P128(7)),
          "n"(P128(8)), "n"(P128(9)), "n"(P128(10)), "n"(P128(11)), "n"(P128(12)));
}
// 13 normalised limbs of sum_i A[i] 2^(32 i), A[i] < 2^40: low halves, plus the high halves one limb up
BLS_FN void lin_norm13(uint32_t *o, const uint64_t (&A)[12]) {
    uint32_t hi[12];
    o[0] = (uint32_t)A[0];
#pragma unroll
    for (int i = 1; i < 12; i++) o[i] = (uint32_t)A[i];
    o[12] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) hi[i] = (uint32_t)(A[i] >> 32);
    add12(o + 1, hi);                                        // no carry out: the top limb is hi[11] + carry < 2^9
}
// t < 4p (12 limbs) -> t mod p
BLS_FN void reduce_4p12(uint32_t *r, const uint32_t *t) {
    uint32_t u[12], v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) u[i] = t[i];
    uint32_t bw;
    // u = t - 2p
    asm("sub.cc.u32 %0,%0,%13;\n\tsubc.cc.u32 %1,%1,%14;\n\tsubc.cc.u32 %2,%2,%15;\n\tsubc.cc.u32 %3,%3,%16;\n\t"
        "subc.cc.u32 %4,%4,%17;\n\tsubc.cc.u32 %5,%5,%18;\n\tsubc.cc.u32 %6,%6,%19;\n\tsubc.cc.u32 %7,%7,%20;\n\t"
        "subc.cc.u32 %8,%8,%21;\n\tsubc.cc.u32 %9,%9,%22;\n\tsubc.cc.u32 %10,%10,%23;\n\tsubc.cc.u32 %11,%11,%24;\n\t"
        "subc.u32 %12,0,0;"
        : BLS_R12(u), "=r"(bw)
        : "n"(P32(0) << 1), "n"((P32(1) << 1) | (P32(0) >> 31)), "n"((P32(2) << 1) | (P32(1) >> 31)),
          "n"((P32(3) << 1) | (P32(2) >> 31)), "n"((P32(4) << 1) | (P32(3) >> 31)), "n"((P32(5) << 1) | (P32(4) >> 31)),
          "n"((P32(6) << 1) | (P32(5) >> 31)), "n"((P32(7) << 1) | (P32(6) >> 31)), "n"((P32(8) << 1) | (P32(7) >> 31)),
          "n"((P32(9) << 1) | (P32(8) >> 31)), "n"((P32(10) << 1) | (P32(9) >> 31)), "n"((P32(11) << 1) | (P32(10) >> 31)));
#pragma unroll
    for (int i = 0; i < 12; i++) v[i] = bw ? t[i] : u[i];    // < 2p
    reduce_once12(r, v);
}
BLS_FN void lin_finish(fp &r, const lin_acc &a) {
    uint32_t D[13], T[13];
    lin_norm13(D, a.P);
    lin_norm13(T, a.N);
    add_128p13(D);
    sub13(D, T);                                             // D = P - N + 128 p in (0, 233 p)
    // quotient estimate from the top 21 bits: h = D >> 368, p >> 368 = 6657; h / 6658 as a multiply-high never overshoots
    // D / p and undershoots by at most 2
    const uint32_t h = (D[12] << 16) | (D[11] >> 16);
    const uint32_t q = __umulhi(h, 1321131424u) >> 11;       // floor(2^43 / 6658) = 1321131424
    uint64_t qp[12];
#pragma unroll
    for (int i = 0; i < 12; i++) qp[i] = (uint64_t)P32(i) * q;
    lin_norm13(T, qp);
    sub13(D, T);                                             // in [0, 4p): twelve limbs
    reduce_4p12(r.l, D);
}
#else
// r = (P - N) mod p, fully reduced.  Requires sum of magnitudes on each side <= 120 (N < 128 p keeps D positive,
// D < 248 p < 2^389 keeps the quotient below 256).
BLS_FN void lin_finish(fp &r, const lin_acc &a) {
    uint32_t D[13];
    int64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const int64_t t = (int64_t)a.P[i] - (int64_t)a.N[i] + (int64_t)P128(i) + carry;
        D[i] = (uint32_t)t;
        carry = t >> 32;                                     // arithmetic shift: floor division by 2^32
    }
    D[12] = (uint32_t)(carry + (int64_t)P128(12));
    // quotient estimate: h = D >> 352, p >> 352 = 0x1a0111ea; dividing by one more than that never overshoots and
    // undershoots by at most 2 for D < 256 p
    const uint64_t h = ((uint64_t)D[12] << 32) | D[11];
    const uint32_t q = (uint32_t)(h / 0x1a0111ebull);
    // D -= q p  (the result is < 3 p < 2^383: twelve limbs)
    uint64_t mc = 0;
    int64_t bw = 0;
    uint32_t R[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint64_t m = (uint64_t)P32(i) * q + mc;
        mc = m >> 32;
        const int64_t t = (int64_t)D[i] - (int64_t)(uint32_t)m + bw;
        R[i] = (uint32_t)t;
        bw = t >> 32;
    }
    // two conditional subtractions of p
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        uint32_t U[12];
        int64_t b = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const int64_t t = (int64_t)R[i] - (int64_t)P32(i) + b;
            U[i] = (uint32_t)t;
            b = t >> 32;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) R[i] = b ? R[i] : U[i];
    }
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = R[i];
}
#endif

}  // namespace bls
