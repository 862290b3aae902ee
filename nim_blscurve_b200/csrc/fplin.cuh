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

}  // namespace bls
