// fpprog.hpp — host-side compiler for the warp-cooperative tail of the batch verifier.
//
// The per-batch tail (stitching the Miller-loop segments: 63 Fp12 squarings; the product of the rank partials; the
// final exponentiation of vendor/blst/src/pairing.c:371-404) is a few thousand Fp operations with a long dependency
// chain and plenty of instruction-level parallelism (an Fp12 product is 54 independent Fp multiplications).  One
// thread wastes that; 32 lanes can share it.  This header TRACES the tower formulas into a dataflow graph of
// Fp operations (mul / add / sub), list-schedules the graph into rounds of at most 32 independent operations of
// one kind, assigns shared-memory slots by liveness, and serialises the result.  The device interpreter
// (k_fp_program in kernels.cuh) executes one round per step: lane l performs operation l of the round on
// 48-byte slots in shared memory.
//
// Pure C++ (no CUDA): included by blsgpu.cu to build the programs at context creation and by tests/hostsim to
// execute them on the CPU against the straight-line formulas of tower.cuh / pairing.cuh.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <queue>
#include <utility>
#include <vector>

namespace fpprog {

enum { OP_NOP = 0, OP_MUL = 1, OP_ADD = 2, OP_SUB = 3, OP_LEAF = 4 };
enum { LANES = 32, MAX_SLOTS = 1024, ROUND_WORDS = LANES, LIN_DEFER = 1 };
// buffer ids of the I/O tables
enum { BUF_IN0 = 0, BUF_IN1 = 1, BUF_CONST = 2, BUF_OUT0 = 3 };
// fp indices inside the constant pool (BUF_CONST): Frobenius coefficients as fp2 = 2 fp each
enum { CONST_FROB1 = 0, CONST_FROB2 = 10, CONST_FROB3 = 20, CONST_PSI_CX = 30, CONST_PSI_CY = 32, CONST_PSI2_CX = 34, CONST_ONE = 35,
       CONST_COUNT = 36 };

struct Node { uint8_t op; int32_t a, b; };
struct IoRef { int32_t node, buf, idx; bool stream = false; };

struct Builder {
    std::vector<Node> nodes;
    std::vector<IoRef> inputs, outputs;
    Builder() { nodes.push_back({OP_LEAF, -1, -1}); }            // node 0 = the constant zero (slot 0)
    int leaf(int buf, int idx) {
        for (const IoRef &r : inputs) if (r.buf == buf && r.idx == idx) return r.node;
        nodes.push_back({OP_LEAF, -1, -1});
        inputs.push_back({(int32_t)nodes.size() - 1, buf, idx});
        return (int)nodes.size() - 1;
    }
    int emit(int op, int a, int b) {
        if ((op == OP_ADD || op == OP_SUB) && b == 0) return a;
        if (op == OP_ADD && a == 0) return b;
        if (op == OP_MUL && (a == 0 || b == 0)) return 0;
        nodes.push_back({(uint8_t)op, a, b});
        return (int)nodes.size() - 1;
    }
    // stream: the value is written to OUT0 by a STORE operation in the round after it is computed and its slot is freed,
    // instead of staying resident until the end of the program (programs with hundreds of outputs)
    void output(int node, int buf, int idx, bool stream = false) {
        IoRef r{node, buf, idx};
        r.stream = stream && node != 0 && buf == BUF_OUT0 && nodes[node].op != OP_LEAF;
        outputs.push_back(r);
    }
};

static thread_local Builder *g_b = nullptr;

// ---- traced field elements -----------------------------------------------------------------------
struct V { int id; };
inline V operator+(V a, V b) { return {g_b->emit(OP_ADD, a.id, b.id)}; }
inline V operator-(V a, V b) { return {g_b->emit(OP_SUB, a.id, b.id)}; }
inline V operator*(V a, V b) { return {g_b->emit(OP_MUL, a.id, b.id)}; }
inline V vzero() { return {0}; }
inline V vneg(V a) { return vzero() - a; }

struct V2 { V c0, c1; };
inline V2 operator+(V2 a, V2 b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
inline V2 operator-(V2 a, V2 b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
inline V2 neg(V2 a) { return {vneg(a.c0), vneg(a.c1)}; }
inline V2 dbl(V2 a) { return a + a; }
inline V2 conj(V2 a) { return {a.c0, vneg(a.c1)}; }
inline V2 mul_xi(V2 a) { return {a.c0 - a.c1, a.c0 + a.c1}; }          // * (1+u)
// Shallow mode (per-set G2 / line programs): a warp has lanes to spare there, so Fp2 products are taken schoolbook
// (4 multiplications, one addition level after) and squares as a0^2, a1^2, a0 a1 — one multiplication more than
// Karatsuba / complex squaring but two addition levels fewer per product on the critical path.  Results are the same
// field elements (every value is canonical), only the schedule changes.
static thread_local bool g_fp2_shallow = false;
inline V2 operator*(V2 a, V2 b) {
    if (g_fp2_shallow) {
        V t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = a.c0 * b.c1, t3 = a.c1 * b.c0;
        return {t0 - t1, t2 + t3};
    }
    V t2 = (a.c0 + a.c1) * (b.c0 + b.c1), t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;      // Karatsuba, as fp2_mul
    return {t0 - t1, t2 - t0 - t1};
}
inline V2 sqr(V2 a) {
    if (g_fp2_shallow) {
        V s0 = a.c0 * a.c0, s1 = a.c1 * a.c1, t = a.c0 * a.c1;
        return {s0 - s1, t + t};
    }
    V t = a.c0 * a.c1;                                                  // as fp2_sqr
    return {(a.c0 + a.c1) * (a.c0 - a.c1), t + t};
}
inline V2 mul_fp(V2 a, V k) { return {a.c0 * k, a.c1 * k}; }
inline V2 zero2() { return {vzero(), vzero()}; }

struct V6 { V2 c0, c1, c2; };
inline V6 operator+(V6 a, V6 b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
inline V6 operator-(V6 a, V6 b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
inline V6 neg(V6 a) { return {neg(a.c0), neg(a.c1), neg(a.c2)}; }
inline V6 dbl(V6 a) { return a + a; }
inline V6 mul_v(V6 a) { return {mul_xi(a.c2), a.c0, a.c1}; }
inline V6 operator*(V6 a, V6 b) {                                       // as fp6_mul
    V2 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = a.c2 * b.c2;
    V2 c0 = mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - t1 - t2) + t0;
    V2 c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1 + mul_xi(t2);
    V2 c2 = (a.c0 + a.c2) * (b.c0 + b.c2) - t0 - t2 + t1;
    return {c0, c1, c2};
}

struct V12 { V6 c0, c1; };
inline V12 operator*(V12 a, V12 b) {                                    // as fp12_mul
    V6 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    V6 s = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1;
    return {t0 + mul_v(t1), s};
}
inline V12 sqr(V12 a) {                                                 // complex squaring, as fp12_sqr
    V6 t = a.c0 * a.c1;
    V6 s = (a.c0 + a.c1) * (mul_v(a.c1) + a.c0) - t - mul_v(t);
    return {s, dbl(t)};
}
inline V12 conj(V12 a) { return {a.c0, neg(a.c1)}; }

// p - 2, little-endian 64-bit words (p: vendor/blst/src/consts.c:10-14)
static const uint64_t P_MINUS_2[6] = {0xb9feffffffffaaa9ull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                                      0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};

// The one Fp inversion of a tail can be taken out of the dataflow program: a 444-step Fermat chain is the longest
// serial stretch of the final exponentiation, and a single thread running the branchy binary Euclid of fpx.cuh
// (fp_inv_vartime) is about four times quicker.  Mode 1 traces up to the inversion and emits its argument as OUT0[0]
// (everything downstream is dead code and pruned); mode 2 reads the inverse from IN1[0].
enum { INV_INLINE = 0, INV_EMIT_ARG = 1, INV_EXTERNAL = 2 };
static thread_local int g_inv_mode = INV_INLINE;

// 1/a = a^(p-2), fixed 4-bit windows (0 -> 0)
inline V inv(V a) {
    if (g_inv_mode == INV_EMIT_ARG) { g_b->output(a.id, BUF_OUT0, 0); return {g_b->leaf(BUF_IN1, 0)}; }
    if (g_inv_mode == INV_EXTERNAL) return {g_b->leaf(BUF_IN1, 0)};
    V tbl[16];
    tbl[1] = a;
    for (int i = 2; i < 16; i++) tbl[i] = tbl[i - 1] * a;
    V acc = {0};
    bool started = false;
    for (int nib = 95; nib >= 0; nib--) {
        int d = (int)((P_MINUS_2[nib / 16] >> (4 * (nib % 16))) & 15);
        if (started) for (int k = 0; k < 4; k++) acc = acc * acc;
        if (d) {
            if (started) acc = acc * tbl[d]; else { acc = tbl[d]; started = true; }
        }
    }
    return acc;
}
inline V2 inv(V2 a) {
    V ni = inv(a.c0 * a.c0 + a.c1 * a.c1);
    return {a.c0 * ni, vneg(a.c1 * ni)};
}
inline V6 inv(V6 a) {                                                   // as fp6_inv
    V2 c0 = sqr(a.c0) - mul_xi(a.c1 * a.c2);
    V2 c1 = mul_xi(sqr(a.c2)) - a.c0 * a.c1;
    V2 c2 = sqr(a.c1) - a.c0 * a.c2;
    V2 t = inv(mul_xi(a.c2 * c1 + a.c1 * c2) + a.c0 * c0);
    return {c0 * t, c1 * t, c2 * t};
}
inline V12 inv(V12 a) {                                                 // as fp12_inv
    V6 t = inv(a.c0 * a.c0 - mul_v(a.c1 * a.c1));
    return {a.c0 * t, neg(a.c1 * t)};
}

inline V2 const2(int base, int k) { return {{g_b->leaf(BUF_CONST, base + 2 * k)}, {g_b->leaf(BUF_CONST, base + 2 * k + 1)}}; }

// a^(p^n), n = 1..3 (as fp12_frob): coefficient of v^i w^j times FROBn[2i+j-1], conjugated first for odd n
inline V12 frob(V12 a, int n) {
    const int base = n == 1 ? CONST_FROB1 : (n == 2 ? CONST_FROB2 : CONST_FROB3);
    V2 *src[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
    V2 dst[6];
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 3; i++) {
            V2 c = *src[3 * j + i];
            if (n & 1) c = conj(c);
            int k = 2 * i + j;
            if (k) c = c * const2(base, k - 1);
            dst[3 * j + i] = c;
        }
    return {{dst[0], dst[1], dst[2]}, {dst[3], dst[4], dst[5]}};
}

// (a + b v)^2 in Fp4 = Fp2[v]/(v^2 - xi): t0 = a^2 + xi b^2, t1 = 2 a b.  All ten Fp products come straight from the
// coordinates (schoolbook, 2ab as a product instead of (a+b)^2 - a^2 - b^2): the three Fp4 squarings of a cyclotomic
// squaring are 30 independent multiplications = ONE round of the 32 lanes with no addition level before it and three
// after, where the Karatsuba forms needed two levels before and three after.  Same field elements, shorter schedule.
inline void fp4_sqr(V2 &t0, V2 &t1, V2 a, V2 b) {
    V s0 = a.c0 * a.c0, s1 = a.c1 * a.c1, st = a.c0 * a.c1;
    V r0 = b.c0 * b.c0, r1 = b.c1 * b.c1, rt = b.c0 * b.c1;
    V p0 = a.c0 * b.c0, p1 = a.c1 * b.c1, p2 = a.c0 * b.c1, p3 = a.c1 * b.c0;
    V2 a2 = {s0 - s1, st + st}, b2 = {r0 - r1, rt + rt}, ab = {p0 - p1, p2 + p3};
    t1 = ab + ab;
    t0 = a2 + mul_xi(b2);
}
inline V12 cyc_sqr(V12 a) {                                             // Granger-Scott, as fp12_cyc_sqr
    V2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2, t0, t1, t2, t3;
    // 3 t -+ 2 z as (t + t) + (t -+ 2z): 2z does not depend on the products, so two addition levels follow t, not three
    fp4_sqr(t0, t1, z0, z1);
    z0 = dbl(t0) + (t0 - dbl(z0));
    z1 = dbl(t1) + (t1 + dbl(z1));
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    z4 = dbl(t0) + (t0 - dbl(z4));
    z5 = dbl(t1) + (t1 + dbl(z5));
    t3 = mul_xi(t3);
    z2 = dbl(t3) + (t3 + dbl(z2));
    z3 = dbl(t2) + (t2 - dbl(z3));
    return {{z0, z4, z3}, {z2, z1, z5}};
}

static const uint64_t Z_ABS = 0xd201000000010000ull;
inline V12 cyc_exp_z(V12 a) {
    V12 acc = a;
    for (int i = 62; i >= 0; i--) {
        acc = cyc_sqr(acc);
        if ((Z_ABS >> i) & 1) acc = acc * a;
    }
    return conj(acc);
}

// f^(3 (p^12-1)/r), the exponent of pairing.c:371-404 (same chain as final_exp in pairing.cuh)
inline V12 final_exp(V12 f) {
    V12 t = conj(f) * inv(f);
    t = frob(t, 2) * t;
    V12 a = cyc_exp_z(t) * conj(t);
    a = cyc_exp_z(a) * conj(a);
    V12 b = cyc_exp_z(a) * frob(a, 1);
    V12 c = cyc_exp_z(cyc_exp_z(b)) * frob(b, 2) * conj(b);
    V12 d = cyc_sqr(t) * t;
    return c * d;
}

inline V12 load12(int buf, int base) {
    V2 c[6];
    for (int k = 0; k < 6; k++) c[k] = {{g_b->leaf(buf, base + 2 * k)}, {g_b->leaf(buf, base + 2 * k + 1)}};
    return {{c[0], c[1], c[2]}, {c[3], c[4], c[5]}};
}
inline void store12(V12 a, int buf, int base) {
    V2 c[6] = {a.c0.c0, a.c0.c1, a.c0.c2, a.c1.c0, a.c1.c1, a.c1.c2};
    for (int k = 0; k < 6; k++) { g_b->output(c[k].c0.id, buf, base + 2 * k); g_b->output(c[k].c1.id, buf, base + 2 * k + 1); }
}

// ---- scheduling + slot allocation ------------------------------------------------------------------
struct Program {
    std::vector<uint32_t> words;      // [nrounds, nslots, n_in, n_out] + in table + out table + rounds
    int nrounds = 0, nmul_rounds = 0, nslots = 0, nops = 0;
    int version = 1;                  // 1: 32 words per round (mul / add / sub); 2: 32 x 4 words per round (mul / lincomb)
    bool ok = false;
};

// Format 2 ("lincomb" programs; executed by k_fp_program2 / hostsim run_program2).  A round is 32 x 4 words, lane l owns
// words [4 l, 4 l + 4):
//   w0 = op:2 | dst:10 | npos:3 | nneg:3 | term6:14        op 1 = MUL, 2 = LIN, 0 = idle (w0 == 0) or STORE (w0 != 0)
//   MUL   w1 = a:10 | b:10                                   slot[dst] = slot[a] * slot[b]
//   LIN   terms t0..t6 (14 bits each: slot:10 | magnitude:4) in w1 = t0 | t1 << 14, w2 = t2 | t3 << 14, w3 = t4 | t5 << 14,
//         t6 in w0; the first npos positions (from t0 up) are the positive terms, the last nneg positions (from t6 down)
//         the negative ones:    slot[dst] = sum_pos mag * slot - sum_neg mag * slot   (mod p; fplin.cuh)
//   STORE w1 = source slot, w2 = OUT0 index                 (streamed outputs)
enum { LIN_MAXT = 7, LIN_MAXMAG = 15 };
// Programs are compiled to format 1 unless BLSGPU_PROG_FORMAT=2 (A/B runs) or a test overrides it.  Measured on B200
// (profiles/r2/r2_summary.md): format 2 cuts the final exponentiation from 2 408 to 973 rounds, but a combination round
// costs a lone warp 1.35 us (accumulate + carry passes + quotient step are ~360 dependent instructions at ~4.5 cycles
// each) against 0.22 us for a two-operand round, so the shorter schedule runs SLOWER (final exponentiation 1.12 ms
// against 0.96 ms).  Kept as a tested alternative, not the default.
static thread_local int g_format_override = 0;
inline int program_format() {
    if (g_format_override) return g_format_override;
    static const int f = getenv("BLSGPU_PROG_FORMAT") ? atoi(getenv("BLSGPU_PROG_FORMAT")) : 1;
    return f == 2 ? 2 : 1;
}

inline uint32_t enc(int op, int d, int a, int b) { return ((uint32_t)op << 30) | ((uint32_t)d << 20) | ((uint32_t)a << 10) | (uint32_t)b; }

inline Program compile(const Builder &B, int slack = 40) {
    const int N = (int)B.nodes.size();
    Program P;
    std::vector<char> live(N, 0);
    for (const IoRef &o : B.outputs) live[o.node] = 1;
    for (int n = N - 1; n >= 0; n--)
        if (live[n] && B.nodes[n].op != OP_LEAF) { live[B.nodes[n].a] = 1; live[B.nodes[n].b] = 1; }
    // priority: longest weighted path to an output
    std::vector<int> height(N, 0);
    for (int n = N - 1; n >= 0; n--) {
        if (!live[n] || B.nodes[n].op == OP_LEAF) continue;
        int h = height[n] + (B.nodes[n].op == OP_MUL ? 10 : 1);
        height[B.nodes[n].a] = std::max(height[B.nodes[n].a], h);
        height[B.nodes[n].b] = std::max(height[B.nodes[n].b], h);
    }
    std::vector<std::vector<int>> users(N);
    std::vector<int> pending(N, 0), round_of(N, -1);
    typedef std::pair<int, int> PQE;       // (height, -node)
    std::priority_queue<PQE> q_mul, q_lin;
    auto push_ready = [&](int n) { (B.nodes[n].op == OP_MUL ? q_mul : q_lin).push({height[n], -n}); };
    for (int n = 0; n < N; n++) {
        if (!live[n] || B.nodes[n].op == OP_LEAF) continue;
        int a = B.nodes[n].a, b = B.nodes[n].b, cnt = 0;
        if (B.nodes[a].op != OP_LEAF) { users[a].push_back(n); cnt++; }
        if (b != a && B.nodes[b].op != OP_LEAF) { users[b].push_back(n); cnt++; }
        pending[n] = cnt;
        if (cnt == 0) push_ready(n);
    }
    // Rounds: operations far off the critical path (height more than `slack` below the most urgent ready one) wait,
    // which keeps the number of live values — shared-memory slots — bounded; cheap add/sub rounds go first so
    // that the expensive multiplication rounds are as full as possible.
    std::vector<std::vector<int>> rounds;
    std::vector<char> round_is_mul;
    while (!q_mul.empty() || !q_lin.empty()) {
        int hmax = std::max(q_mul.empty() ? -1 : q_mul.top().first, q_lin.empty() ? -1 : q_lin.top().first);
        // a ready add/sub that is less urgent than every ready multiplication (it is not on their way) does not get a
        // round of its own: it joins the next addition round
        bool is_mul = q_lin.empty() || q_lin.top().first < hmax - slack ||
                      (!q_mul.empty() && q_lin.top().first + LIN_DEFER <= q_mul.top().first);
        std::priority_queue<PQE> &q = is_mul ? q_mul : q_lin;
        std::vector<int> ops;
        while (!q.empty() && (int)ops.size() < LANES && q.top().first >= hmax - slack) { ops.push_back(-q.top().second); q.pop(); }
        const int r = (int)rounds.size();
        for (int n : ops) round_of[n] = r;
        for (int n : ops)
            for (int u : users[n]) if (--pending[u] == 0) push_ready(u);
        rounds.push_back(ops);
        round_is_mul.push_back(is_mul);
    }
    // streamed outputs: a STORE lane in the first round after the producer that has a free lane (or a new last round)
    std::vector<std::vector<std::pair<int, int>>> stores(rounds.size());      // per round: (node, OUT0 index)
    for (const IoRef &o : B.outputs) {
        if (!o.stream) continue;
        size_t q = (size_t)round_of[o.node] + 1;
        while (q < rounds.size() && rounds[q].size() + stores[q].size() >= (size_t)LANES) q++;
        if (q >= rounds.size()) { rounds.push_back({}); round_is_mul.push_back(0); stores.push_back({}); q = rounds.size() - 1; }
        stores[q].push_back({o.node, o.idx});
    }
    // liveness: last round in which each value is read
    const int NR = (int)rounds.size();
    std::vector<int> last_use(N, -1);
    for (int r = 0; r < NR; r++) {
        for (int n : rounds[r]) { last_use[B.nodes[n].a] = r; last_use[B.nodes[n].b] = r; }
        for (const auto &st : stores[r]) last_use[st.first] = std::max(last_use[st.first], r);
    }
    for (const IoRef &o : B.outputs) if (!o.stream) last_use[o.node] = NR + 1;
    last_use[0] = NR + 1;
    std::vector<int> slot(N, -1);
    slot[0] = 0;
    std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;
    int next_slot = 1;
    auto alloc = [&]() { if (!free_slots.empty()) { int s = free_slots.top(); free_slots.pop(); return s; } return next_slot++; };
    std::vector<std::vector<int>> expire(NR + 2);
    for (const IoRef &in : B.inputs)
        if (live[in.node]) {
            slot[in.node] = alloc();
            if (last_use[in.node] >= 0 && last_use[in.node] <= NR) expire[last_use[in.node]].push_back(in.node);
        }
    for (int r = 0; r < NR; r++) {
        for (int n : rounds[r]) {
            slot[n] = alloc();
            if (last_use[n] <= NR) expire[last_use[n] < r ? r : last_use[n]].push_back(n);
        }
        for (int n : expire[r]) free_slots.push(slot[n]);
    }
    P.nslots = next_slot;
    if (next_slot > MAX_SLOTS) return P;
    // serialise
    std::vector<IoRef> ins;
    for (const IoRef &in : B.inputs) if (live[in.node]) ins.push_back(in);
    std::vector<IoRef> outs;
    for (const IoRef &o : B.outputs) if (!o.stream) outs.push_back(o);
    P.words = {(uint32_t)NR, (uint32_t)P.nslots, (uint32_t)ins.size(), (uint32_t)outs.size()};
    for (const IoRef &in : ins) { P.words.push_back((uint32_t)slot[in.node]); P.words.push_back(((uint32_t)in.buf << 24) | (uint32_t)in.idx); }
    for (const IoRef &o : outs) { P.words.push_back((uint32_t)slot[o.node]); P.words.push_back(((uint32_t)o.buf << 24) | (uint32_t)o.idx); }
    for (int r = 0; r < NR; r++) {
        const int nops = (int)rounds[r].size();
        for (int l = 0; l < LANES; l++) {
            if (l < nops) {
                int n = rounds[r][l];
                P.words.push_back(enc(B.nodes[n].op, slot[n], slot[B.nodes[n].a], slot[B.nodes[n].b]));
                P.nops++;
            } else if (l - nops < (int)stores[r].size()) {
                // STORE: opcode 0 with a non-zero word; OUT0 index split over the d and b fields (20 bits)
                const std::pair<int, int> &st = stores[r][l - nops];
                P.words.push_back(enc(OP_NOP, (st.second >> 10) & 1023, slot[st.first], st.second & 1023));
            } else {
                P.words.push_back(0u);
            }
        }
        P.nmul_rounds += round_is_mul[r] ? 1 : 0;
    }
    P.nrounds = NR;
    P.ok = true;
    return P;
}

// ---- format 2: every linear expression flattened into one combination ---------------------------------------------
// Linear nodes (add / sub) are not operations of their own any more: each is expressed as a combination
// sum coef * base over MATERIALISED nodes (leaves, products, and those linear nodes that had to be given a slot: operands
// of a product, program outputs, and cut points where a combination would exceed 7 terms or a magnitude of 15).  Between
// two levels of products there is then ONE round of combinations instead of three to six rounds of two-operand additions.
inline Program compile2(const Builder &B, int slack = 40) {
    const int N = (int)B.nodes.size();
    Program P;
    P.version = 2;
    typedef std::vector<std::pair<int, int>> Form;            // (materialised node, coefficient), sorted by node
    std::vector<Form> form(N);
    std::vector<char> mat(N, 0), is_lin(N, 0);
    auto combine = [&](const Form &fa, const Form &fb, int sb) {
        Form r;
        size_t i = 0, j = 0;
        while (i < fa.size() || j < fb.size()) {
            if (j >= fb.size() || (i < fa.size() && fa[i].first < fb[j].first)) r.push_back(fa[i++]);
            else if (i >= fa.size() || fb[j].first < fa[i].first) { r.push_back({fb[j].first, sb * fb[j].second}); j++; }
            else { int c = fa[i].second + sb * fb[j].second; if (c) r.push_back({fa[i].first, c}); i++; j++; }
        }
        return r;
    };
    auto fits = [&](const Form &f) {
        if ((int)f.size() > LIN_MAXT) return false;
        for (const auto &t : f) if (t.second > LIN_MAXMAG || t.second < -LIN_MAXMAG) return false;
        return true;
    };
    auto ref = [&](int x) -> Form {                            // how a consumer sees node x
        if (x == 0) return Form();
        if (!is_lin[x] || mat[x]) return Form{{x, 1}};
        return form[x];
    };
    for (int n = 1; n < N; n++) {
        const Node &nd = B.nodes[n];
        if (nd.op == OP_LEAF) { mat[n] = 1; continue; }
        if (nd.op == OP_MUL) {
            mat[n] = 1;
            if (is_lin[nd.a]) mat[nd.a] = 1;                   // operands of a product need a slot
            if (is_lin[nd.b]) mat[nd.b] = 1;
            continue;
        }
        is_lin[n] = 1;
        const int sb = nd.op == OP_SUB ? -1 : 1;
        Form f = combine(ref(nd.a), ref(nd.b), sb);
        if (!fits(f)) {                                        // cut: give the larger operand a slot of its own, then the other
            const bool ca = is_lin[nd.a] && !mat[nd.a], cb = is_lin[nd.b] && !mat[nd.b];
            const size_t sa = ca ? form[nd.a].size() : 0, sbz = cb ? form[nd.b].size() : 0;
            if (ca && (!cb || sa >= sbz)) mat[nd.a] = 1; else if (cb) mat[nd.b] = 1;
            f = combine(ref(nd.a), ref(nd.b), sb);
            if (!fits(f)) {
                if (is_lin[nd.a]) mat[nd.a] = 1;
                if (is_lin[nd.b]) mat[nd.b] = 1;
                f = combine(ref(nd.a), ref(nd.b), sb);
            }
        }
        form[n] = f;
    }
    for (const IoRef &o : B.outputs) if (is_lin[o.node]) mat[o.node] = 1;
    // operations = products and materialised combinations that an output needs
    std::vector<char> need(N, 0);
    for (const IoRef &o : B.outputs) need[o.node] = 1;
    for (int n = N - 1; n >= 1; n--) {
        if (!need[n]) continue;
        const Node &nd = B.nodes[n];
        if (nd.op == OP_MUL) { need[nd.a] = 1; need[nd.b] = 1; }
        else if (is_lin[n]) for (const auto &t : form[n]) need[t.first] = 1;
    }
    need[0] = 0;
    auto is_op = [&](int n) { return need[n] && (B.nodes[n].op == OP_MUL || (is_lin[n] && mat[n])); };
    auto deps = [&](int n, std::vector<int> &out) {
        out.clear();
        if (B.nodes[n].op == OP_MUL) { out.push_back(B.nodes[n].a); if (B.nodes[n].b != B.nodes[n].a) out.push_back(B.nodes[n].b); }
        else for (const auto &t : form[n]) out.push_back(t.first);
    };
    // priority: longest weighted path to an output (a product costs about twice a combination)
    std::vector<int> height(N, 0);
    std::vector<int> dl;
    for (int n = N - 1; n >= 1; n--) {
        if (!is_op(n)) continue;
        const int h = height[n] + (B.nodes[n].op == OP_MUL ? 10 : 5);
        deps(n, dl);
        for (int d : dl) height[d] = std::max(height[d], h);
    }
    std::vector<std::vector<int>> users(N);
    std::vector<int> pending(N, 0), round_of(N, -1);
    typedef std::pair<int, int> PQE;
    std::priority_queue<PQE> q_mul, q_lin;
    auto push_ready = [&](int n) { (B.nodes[n].op == OP_MUL ? q_mul : q_lin).push({height[n], -n}); };
    for (int n = 1; n < N; n++) {
        if (!is_op(n)) continue;
        deps(n, dl);
        int cnt = 0;
        for (int d : dl) if (d != 0 && B.nodes[d].op != OP_LEAF) { users[d].push_back(n); cnt++; }
        pending[n] = cnt;
        if (cnt == 0) push_ready(n);
    }
    std::vector<std::vector<int>> rounds;
    std::vector<char> round_is_mul;
    while (!q_mul.empty() || !q_lin.empty()) {
        const int hmax = std::max(q_mul.empty() ? -1 : q_mul.top().first, q_lin.empty() ? -1 : q_lin.top().first);
        // a ready combination that is less urgent than every ready product joins a later round of combinations
        const bool is_mul = q_lin.empty() || q_lin.top().first < hmax - slack ||
                            (!q_mul.empty() && q_lin.top().first + LIN_DEFER <= q_mul.top().first);
        std::priority_queue<PQE> &q = is_mul ? q_mul : q_lin;
        std::vector<int> ops;
        while (!q.empty() && (int)ops.size() < LANES && q.top().first >= hmax - slack) { ops.push_back(-q.top().second); q.pop(); }
        const int r = (int)rounds.size();
        for (int n : ops) round_of[n] = r;
        for (int n : ops)
            for (int u : users[n]) if (--pending[u] == 0) push_ready(u);
        rounds.push_back(ops);
        round_is_mul.push_back(is_mul);
    }
    // streamed outputs
    std::vector<std::vector<std::pair<int, int>>> stores(rounds.size());
    for (const IoRef &o : B.outputs) {
        if (!o.stream) continue;
        size_t q = (size_t)round_of[o.node] + 1;
        while (q < rounds.size() && rounds[q].size() + stores[q].size() >= (size_t)LANES) q++;
        if (q >= rounds.size()) { rounds.push_back({}); round_is_mul.push_back(0); stores.push_back({}); q = rounds.size() - 1; }
        stores[q].push_back({o.node, o.idx});
    }
    // liveness and slots
    const int NR = (int)rounds.size();
    std::vector<int> last_use(N, -1);
    for (int r = 0; r < NR; r++) {
        for (int n : rounds[r]) { deps(n, dl); for (int d : dl) last_use[d] = r; }
        for (const auto &st : stores[r]) last_use[st.first] = std::max(last_use[st.first], r);
    }
    for (const IoRef &o : B.outputs) if (!o.stream) last_use[o.node] = NR + 1;
    last_use[0] = NR + 1;
    std::vector<int> slot(N, -1);
    slot[0] = 0;
    std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;
    int next_slot = 1;
    auto alloc = [&]() { if (!free_slots.empty()) { int s = free_slots.top(); free_slots.pop(); return s; } return next_slot++; };
    std::vector<std::vector<int>> expire(NR + 2);
    std::vector<IoRef> ins;
    for (const IoRef &in : B.inputs)
        if (need[in.node]) {
            ins.push_back(in);
            slot[in.node] = alloc();
            if (last_use[in.node] >= 0 && last_use[in.node] <= NR) expire[last_use[in.node]].push_back(in.node);
        }
    for (int r = 0; r < NR; r++) {
        for (int n : rounds[r]) {
            slot[n] = alloc();
            if (last_use[n] <= NR) expire[last_use[n] < r ? r : last_use[n]].push_back(n);
        }
        for (int n : expire[r]) free_slots.push(slot[n]);
    }
    P.nslots = next_slot;
    if (next_slot > MAX_SLOTS) return P;
    std::vector<IoRef> outs;
    for (const IoRef &o : B.outputs) if (!o.stream) outs.push_back(o);
    for (const IoRef &o : outs) if (slot[o.node] < 0) return P;          // an output that is neither input nor operation
    P.words = {(uint32_t)NR, (uint32_t)P.nslots, (uint32_t)ins.size(), (uint32_t)outs.size()};
    for (const IoRef &in : ins) { P.words.push_back((uint32_t)slot[in.node]); P.words.push_back(((uint32_t)in.buf << 24) | (uint32_t)in.idx); }
    for (const IoRef &o : outs) { P.words.push_back((uint32_t)slot[o.node]); P.words.push_back(((uint32_t)o.buf << 24) | (uint32_t)o.idx); }
    while (P.words.size() % 4) P.words.push_back(0u);        // the rounds are read as 16-byte words
    for (int r = 0; r < NR; r++) {
        const int nops = (int)rounds[r].size();
        for (int l = 0; l < LANES; l++) {
            uint32_t w[4] = {0, 0, 0, 0};
            if (l < nops) {
                const int n = rounds[r][l];
                if (B.nodes[n].op == OP_MUL) {
                    w[0] = (1u << 30) | ((uint32_t)slot[n] << 20);
                    w[1] = (uint32_t)slot[B.nodes[n].a] | ((uint32_t)slot[B.nodes[n].b] << 10);
                } else {
                    uint32_t t[LIN_MAXT] = {0, 0, 0, 0, 0, 0, 0};
                    int npos = 0, nneg = 0;
                    for (const auto &tm : form[n]) {
                        const uint32_t code = (uint32_t)slot[tm.first] | ((uint32_t)(tm.second < 0 ? -tm.second : tm.second) << 10);
                        if (tm.second > 0) t[npos++] = code; else t[LIN_MAXT - 1 - nneg++] = code;
                    }
                    w[0] = (2u << 30) | ((uint32_t)slot[n] << 20) | ((uint32_t)npos << 17) | ((uint32_t)nneg << 14) | t[6];
                    w[1] = t[0] | (t[1] << 14);
                    w[2] = t[2] | (t[3] << 14);
                    w[3] = t[4] | (t[5] << 14);
                }
                P.nops++;
            } else if (l - nops < (int)stores[r].size()) {
                const std::pair<int, int> &st = stores[r][l - nops];
                w[0] = 1u;                                   // op 0, non-zero: STORE
                w[1] = (uint32_t)slot[st.first];
                w[2] = (uint32_t)st.second;
            }
            for (int k = 0; k < 4; k++) P.words.push_back(w[k]);
        }
        P.nmul_rounds += round_is_mul[r] ? 1 : 0;
    }
    P.nrounds = NR;
    P.ok = true;
    return P;
}
inline Program compile_any(const Builder &B, int slack = 40) { return program_format() == 2 ? compile2(B, slack) : compile(B, slack); }

// ---- the three tail programs -------------------------------------------------------------------------
// Horner over nseg Miller-loop segment products (IN0: nseg x 12 fp) -> rank partial (OUT0: 12 fp); mirrors
// miller_combine() in pairing.cuh.  seg_len[j] = number of loop iterations of segment j.
inline Program build_combine(int nseg, const int *seg_len) {
    Builder b;
    g_b = &b;
    V12 acc = load12(BUF_IN0, 0);
    for (int j = 1; j < nseg; j++) {
        for (int t = 0; t < seg_len[j]; t++) acc = sqr(acc);
        acc = acc * load12(BUF_IN0, 12 * j);
    }
    store12(conj(acc), BUF_OUT0, 0);
    g_b = nullptr;
    return compile_any(b);
}

// product of `count` partials (IN0: count x 12 fp) followed by the final exponentiation -> OUT0: 12 fp.
// inv_mode INV_INLINE: one self-contained program.  INV_EMIT_ARG + INV_EXTERNAL: a pair of programs around an
// external inversion (OUT0[0] of the first = the Fp norm to invert, IN1[0] of the second = its inverse).
inline Program build_final(int count, int inv_mode = INV_INLINE) {
    Builder b;
    g_b = &b;
    g_inv_mode = inv_mode;
    std::vector<V12> v;
    for (int i = 0; i < count; i++) v.push_back(load12(BUF_IN0, 12 * i));
    while (v.size() > 1) {                                  // balanced product tree
        std::vector<V12> w;
        for (size_t i = 0; i + 1 < v.size(); i += 2) w.push_back(v[i] * v[i + 1]);
        if (v.size() & 1) w.push_back(v.back());
        v.swap(w);
    }
    V12 r = final_exp(v[0]);
    if (inv_mode != INV_EMIT_ARG) store12(r, BUF_OUT0, 0);
    g_inv_mode = INV_INLINE;
    g_b = nullptr;
    return compile_any(b);
}

// product of `count` Fp12 values (IN0: count x 12 fp) -> OUT0: 12 fp; balanced tree.  One warp per segment row of a small
// batch (blsgpu_merge-style GT product, aggregate.c:410-458): 54 multiplications per node spread over the 32 lanes
// instead of one thread per node.
inline Program build_fp12_product(int count) {
    Builder b;
    g_b = &b;
    std::vector<V12> v;
    for (int i = 0; i < count; i++) v.push_back(load12(BUF_IN0, 12 * i));
    while (v.size() > 1) {
        std::vector<V12> w;
        for (size_t i = 0; i + 1 < v.size(); i += 2) w.push_back(v[i] * v[i + 1]);
        if (v.size() & 1) w.push_back(v.back());
        v.swap(w);
    }
    store12(v[0], BUF_OUT0, 0);
    g_b = nullptr;
    return compile_any(b);
}

// ---- G1 in homogeneous projective coordinates, complete formulas --------------------------------------------------
// Renes-Costello-Batina (2016) algorithms 7 and 9 for y^2 = x^3 + b with a = 0, b = 4 (b3 = 12).  E(Fp) has odd order
// (cofactor (z-1)^2/3 and r are odd), so the formulas have no exceptional inputs: infinity is (0 : y : 0), P + P,
// P - P and P + infinity all come out of the same straight line — which is what a branch-free dataflow program needs.
// The point carries N = 2Z and M = 6Z so that the doubling needs no constant multiplications and only shallow
// addition chains between its two multiplication levels (12 Z^2 = M N, 36 Z^2 = M^2, 8 Y^2 Y Z = 4 Y^2 (Y N)).
// Written over the field type: V for G1 (b3 = 12), V2 for G2 on the twist y^2 = x^3 + 4(1+u) (b3 = 12(1+u)).
template <class T> struct PT { T x, y, z, n, m; };
typedef PT<V> VP;
inline V t_sqr(V a) { return a * a; }
inline V2 t_sqr(V2 a) { return sqr(a); }
template <class T> inline T mul12(T a) {
    T a2 = a + a, a4 = a2 + a2, a8 = a4 + a4;
    return a8 + a4;
}
inline V t_xi(V a) { return a; }                 // the (1+u) factor of the G2 curve constant; 1 on G1
inline V2 t_xi(V2 a) { return mul_xi(a); }
inline V t_neg(V a) { return vneg(a); }
inline V2 t_neg(V2 a) { return neg(a); }
template <class T> inline PT<T> vp_make(T x, T y, T z) {
    T n = z + z, n2 = n + n;
    return {x, y, z, n, n2 + n};
}
template <class T> inline PT<T> rcb_neg(PT<T> p) { return {p.x, t_neg(p.y), p.z, p.n, p.m}; }
template <class T> inline PT<T> rcb_dbl(PT<T> p) {
    // b3 Z^2 = xi * M N, 3 b3 Z^2 = xi * M^2
    T A = t_sqr(p.y), B2 = p.y * p.n, D = p.x * p.y, E36 = t_xi(t_sqr(p.m)), E12 = t_xi(p.m * p.n);
    T t0 = A - E36, ys = A + E12;
    T A2 = A + A, A4 = A2 + A2, A8 = A4 + A4;
    T xh = D * t0, yp = t0 * ys, yq = A8 * E12, z3 = A4 * B2, n3 = A8 * B2;
    T n6 = n3 + n3;
    return {xh + xh, yp + yq, z3, n3, n6 + n3};
}
template <class T> inline PT<T> rcb_add(PT<T> p, PT<T> q) {
    T t0 = p.x * q.x, t1 = p.y * q.y, t2 = p.z * q.z;
    T t3 = (p.x + p.y) * (q.x + q.y) - (t0 + t1);
    T t4 = (p.y + p.z) * (q.y + q.z) - (t1 + t2);
    T y3 = (p.x + p.z) * (q.x + q.z) - (t0 + t2);
    T t0x3 = t0 + t0 + t0;
    T t2b = t_xi(mul12(t2));
    T z3 = t1 + t2b, t1m = t1 - t2b;
    T y3b = t_xi(mul12(y3));
    T x3 = t3 * t1m - t4 * y3b;
    T yy = y3b * t0x3 + t1m * z3;
    T zz = z3 * t4 + t0x3 * t3;
    return vp_make(x3, yy, zz);
}

// Horner over the window sums of a G1 MSM (multi_scalar.c:295-311 integrates and combines windows the same way):
// IN0 = nwin homogeneous points (3 fp each, window 0 first; infinity as (0, 1, 0)), acc = [2^c] acc + W_w from the top
// window down.  OUT0[0..2] = the homogeneous result (X : Y : Z); the caller normalises it (one inversion).
inline Program build_msm_horner_g1(int nwin, int c) {
    Builder b;
    g_b = &b;
    auto load = [&](int w) { return vp_make<V>({b.leaf(BUF_IN0, 3 * w)}, {b.leaf(BUF_IN0, 3 * w + 1)}, {b.leaf(BUF_IN0, 3 * w + 2)}); };
    VP acc = load(nwin - 1);
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = rcb_dbl(acc);
        acc = rcb_add(acc, load(w));
    }
    b.output(acc.x.id, BUF_OUT0, 0);
    b.output(acc.y.id, BUF_OUT0, 1);
    b.output(acc.z.id, BUF_OUT0, 2);
    g_b = nullptr;
    return compile_any(b);
}
// the same over Fp2 for a G2 MSM (the signature sum of a batch): IN0 = nwin x 6 fp, OUT0[0..5]
inline PT<V2> load_g2(int buf, int base);
inline void store_g2(PT<V2> p, int buf, int base);
inline Program build_msm_horner_g2(int nwin, int c) {
    Builder b;
    g_b = &b;
    g_fp2_shallow = true;
    PT<V2> acc = load_g2(BUF_IN0, 6 * (nwin - 1));
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = rcb_dbl(acc);
        acc = rcb_add(acc, load_g2(BUF_IN0, 6 * w));
    }
    store_g2(acc, BUF_OUT0, 0);
    g_b = nullptr;
    g_fp2_shallow = false;
    return compile_any(b);
}

// ---- per-set G2 programs of the small-batch route (one warp per signature set, k_fp_program_many) ----------------
inline PT<V2> load_g2(int buf, int base) {
    V2 x = {{g_b->leaf(buf, base)}, {g_b->leaf(buf, base + 1)}}, y = {{g_b->leaf(buf, base + 2)}, {g_b->leaf(buf, base + 3)}};
    V2 z = {{g_b->leaf(buf, base + 4)}, {g_b->leaf(buf, base + 5)}};
    return vp_make(x, y, z);
}
inline void store_g2(PT<V2> p, int buf, int base) {
    V2 c[3] = {p.x, p.y, p.z};
    for (int k = 0; k < 3; k++) { g_b->output(c[k].c0.id, buf, base + 2 * k); g_b->output(c[k].c1.id, buf, base + 2 * k + 1); }
}
inline V2 cfp2(int idx) { return {{g_b->leaf(BUF_CONST, idx)}, {g_b->leaf(BUF_CONST, idx + 1)}}; }
// psi(X : Y : Z) = (conj(X) cx : conj(Y) cy : conj(Z)) (e2.c:455-482 on homogeneous coordinates), psi^2 likewise
inline PT<V2> g2_psi(PT<V2> p) { return vp_make(conj(p.x) * cfp2(CONST_PSI_CX), conj(p.y) * cfp2(CONST_PSI_CY), conj(p.z)); }
inline PT<V2> g2_psi2(PT<V2> p) { return {mul_fp(p.x, {g_b->leaf(BUF_CONST, CONST_PSI2_CX)}), neg(p.y), p.z, p.n, p.m}; }
// [x]P, x = -0xd201000000010000
inline PT<V2> g2_mul_by_x(PT<V2> p) {
    PT<V2> acc = p;
    for (int i = 62; i >= 0; i--) {
        acc = rcb_dbl(acc);
        if ((Z_ABS >> i) & 1) acc = rcb_add(acc, p);
    }
    return rcb_neg(acc);
}
// clear_cofactor (map_to_g2.c:327-349), same combination as g2_clear_cofactor in h2c.cuh.
// IN0[0..5] = (X : Y : Z) homogeneous on E2, OUT0[0..5] = [h_eff] of it.
inline Program build_g2_clear_cofactor() {
    Builder b;
    g_b = &b;
    g_fp2_shallow = true;
    PT<V2> p = load_g2(BUF_IN0, 0);
    PT<V2> t1 = g2_mul_by_x(p);
    PT<V2> t2 = g2_psi(p);
    PT<V2> t3 = g2_psi2(rcb_dbl(p));
    t3 = rcb_add(t3, rcb_neg(t2));
    t2 = g2_mul_by_x(rcb_add(t1, t2));
    t3 = rcb_add(t3, t2);
    t3 = rcb_add(t3, rcb_neg(t1));
    store_g2(rcb_add(t3, rcb_neg(p)), BUF_OUT0, 0);
    g_b = nullptr;
    g_fp2_shallow = false;
    return compile_any(b);
}
// [k]Q for a 64-bit k given as 64 field elements 0 / 1 (IN1[0..63], least significant first; Montgomery form), Q
// homogeneous in IN0[0..5]: double-and-always-add with the addend (b X : b Y + (1 - b) : b Z), which is Q for b = 1 and
// the point at infinity (0 : 1 : 0) for b = 0 — complete formulas make the zero bits cost nothing but their lanes.
inline Program build_g2_mul64() {
    Builder b;
    g_b = &b;
    g_fp2_shallow = true;
    PT<V2> q = load_g2(BUF_IN0, 0);
    V one = {b.leaf(BUF_CONST, CONST_ONE)};
    auto addend = [&](int bit) {
        V bb = {b.leaf(BUF_IN1, bit)};
        V2 y = mul_fp(q.y, bb);
        y.c0 = y.c0 + (one - bb);
        return vp_make(mul_fp(q.x, bb), y, mul_fp(q.z, bb));
    };
    PT<V2> acc = addend(63);
    for (int i = 62; i >= 0; i--) acc = rcb_add(rcb_dbl(acc), addend(i));
    store_g2(acc, BUF_OUT0, 0);
    g_b = nullptr;
    g_fp2_shallow = false;
    return compile_any(b);
}

// All 68 line triples of one Miller-loop pair (pairing.cuh miller_lines: 63 tangents + 5 chords in execution order, each
// already scaled by (-x_P, y_P)), the same formulas traced over V2.  IN0[0..3] = Q affine (x.re, x.im, y.re, y.im),
// IN1[0..1] = P affine (x, y); OUT0[6 s .. 6 s + 5] = (l0.re, l0.im, l1.re, l1.im, l2.re, l2.im) of line s.
// The caller substitutes the neutral line for pairs with a point at infinity (no branches in a dataflow program).
struct TP { V2 x, y, z; };
inline V2 mul3(V2 a) { return dbl(a) + a; }
inline void line_dbl_t(TP &T, V2 &l0, V2 &l1, V2 &l2) {                 // line_dbl_proj
    V2 B = sqr(T.y), C = sqr(T.z), J = sqr(T.x);
    V2 A2 = sqr(T.x + T.y) - J - B;
    V2 H = sqr(T.y + T.z) - B - C;
    V2 E = mul3(dbl(dbl(mul_xi(C))));
    V2 F = mul3(E);
    l0 = B - E;
    l1 = mul3(J);
    l2 = H;
    V2 nx = A2 * (B - F);
    V2 ny = sqr(B + F) - mul3(sqr(dbl(E)));
    V2 nz = dbl(dbl(B * H));
    T = {nx, ny, nz};
}
inline void line_add_t(TP &T, V2 qx, V2 qy, V2 &l0, V2 &l1, V2 &l2) {   // line_add_proj
    V2 th = T.y - qy * T.z, la = T.x - qx * T.z;
    V2 c = sqr(th), d = sqr(la);
    V2 e = la * d, f = T.z * c, g = T.x * d;
    V2 h = e + f - g - g;
    V2 nx = la * h, ny = th * (g - h) - e * T.y, nz = T.z * e;
    l0 = th * qx - la * qy;
    l1 = th;
    l2 = la;
    T = {nx, ny, nz};
}
inline Program build_miller_lines() {
    Builder b;
    g_b = &b;
    g_fp2_shallow = true;
    V2 qx = {{b.leaf(BUF_IN0, 0)}, {b.leaf(BUF_IN0, 1)}}, qy = {{b.leaf(BUF_IN0, 2)}, {b.leaf(BUF_IN0, 3)}};
    V px = {b.leaf(BUF_IN1, 0)}, py = {b.leaf(BUF_IN1, 1)};
    V npx = vneg(px);
    V2 one = {{b.leaf(BUF_CONST, CONST_ONE)}, vzero()};
    TP T = {qx, qy, one};
    int s = 0;
    auto emit = [&](V2 l0, V2 l1, V2 l2) {
        l1 = mul_fp(l1, npx);
        l2 = mul_fp(l2, py);
        V2 c[3] = {l0, l1, l2};
        for (int k = 0; k < 3; k++) {
            // outputs must be computed nodes or leaves with their own slot; "+ 0" folds away, so route through the table
            b.output(c[k].c0.id, BUF_OUT0, 6 * s + 2 * k, true);
            b.output(c[k].c1.id, BUF_OUT0, 6 * s + 2 * k + 1, true);
        }
        s++;
    };
    for (int i = 62; i >= 0; i--) {
        V2 l0, l1, l2;
        line_dbl_t(T, l0, l1, l2);
        emit(l0, l1, l2);
        if ((Z_ABS >> i) & 1) {
            line_add_t(T, qx, qy, l0, l1, l2);
            emit(l0, l1, l2);
        }
    }
    g_b = nullptr;
    g_fp2_shallow = false;
    // no deferral of off-critical-path operations: the line scalings are ready early, few, and their streamed results
    // free their slots at once — deferring them to the end is what would keep hundreds of values alive
    return compile_any(b, 1 << 28);
}

}  // namespace fpprog
