// sha256.cuh — SHA-256, expand_message_xmd and hash_to_field (RFC 9380 5.2/5.3.1).
// Restates vendor/blst/src/sha256.h:40-137, hash_to_field.c:51-114 (expand_message_xmd, with the
// Z_pad block pre-absorbed like :22-35) and :120-154 (hash_to_field: 64-byte big-endian chunks
// reduced mod p and put in Montgomery form; here via two Montgomery products by R^2 and 2^256*R^2).
#pragma once
#include "consts.cuh"

namespace bls {

BLS_TABLE uint32_t SHA256_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

BLS_FN uint32_t rotr32(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

// one compression; w[16] big-endian words of the block (clobbered)
BLS_NOINLINE void sha256_block(uint32_t *h, uint32_t *w) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 16
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i & 15];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// The same compression, force-inlined and fully unrolled: every index is static, so the sixteen schedule words and the
// eight state words live in registers (the noinline form above keeps w[] in local memory: ~2.9 us per block for a lone
// thread).  For the strictly sequential RLC scalar chains (blst_min_pubkey_sig_core.nim:507), where one thread's
// latency per block is the whole cost.
BLS_FN void sha256_block_regs(uint32_t (&h)[8], uint32_t (&w)[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i & 15];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

struct sha256_ctx {
    uint32_t h[8];
    uint32_t w[16];   // current block as big-endian words
    uint32_t off;     // bytes in w
    uint64_t len;     // total bytes absorbed
};

BLS_FN void sha256_init(sha256_ctx &c) {
    c.h[0] = 0x6a09e667u; c.h[1] = 0xbb67ae85u; c.h[2] = 0x3c6ef372u; c.h[3] = 0xa54ff53au;
    c.h[4] = 0x510e527fu; c.h[5] = 0x9b05688cu; c.h[6] = 0x1f83d9abu; c.h[7] = 0x5be0cd19u;
    for (int i = 0; i < 16; i++) c.w[i] = 0;
    c.off = 0;
    c.len = 0;
}

BLS_FN void sha256_put(sha256_ctx &c, uint8_t byte) {
    uint32_t o = c.off;
    c.w[o >> 2] |= (uint32_t)byte << (24 - 8 * (o & 3));
    c.off = o + 1;
    c.len++;
    if (c.off == 64) {
        sha256_block(c.h, c.w);
        for (int i = 0; i < 16; i++) c.w[i] = 0;
        c.off = 0;
    }
}

BLS_FN void sha256_update(sha256_ctx &c, const uint8_t *p, size_t n) {
    for (size_t i = 0; i < n; i++) sha256_put(c, p[i]);
}

// absorb a 32-byte digest held as 8 big-endian words
BLS_FN void sha256_update_words(sha256_ctx &c, const uint32_t *d, int nwords) {
    for (int i = 0; i < nwords; i++) {
        sha256_put(c, (uint8_t)(d[i] >> 24));
        sha256_put(c, (uint8_t)(d[i] >> 16));
        sha256_put(c, (uint8_t)(d[i] >> 8));
        sha256_put(c, (uint8_t)d[i]);
    }
}

// digest as 8 big-endian words
BLS_FN void sha256_final(sha256_ctx &c, uint32_t *out) {
    uint64_t bits = c.len * 8;
    sha256_put(c, 0x80);
    while (c.off != 56) sha256_put(c, 0);
    c.w[14] = (uint32_t)(bits >> 32);
    c.w[15] = (uint32_t)bits;
    sha256_block(c.h, c.w);
    for (int i = 0; i < 8; i++) out[i] = c.h[i];
}

// expand_message_xmd(msg, DST, 256) -> 64 big-endian words (8 digests)
BLS_FN void expand_message_xmd_256(uint32_t *out, const uint8_t *msg, size_t msg_len, const uint8_t *dst,
                                   uint32_t dst_len) {
    sha256_ctx c;
    uint32_t b0[8];
    sha256_init(c);
    for (int i = 0; i < 8; i++) c.h[i] = SHA256_ZPAD_STATE[i];   // Z_pad already absorbed
    c.len = 64;
    sha256_update(c, msg, msg_len);
    sha256_put(c, 0x01);   // l_i_b_str = I2OSP(256, 2)
    sha256_put(c, 0x00);
    sha256_put(c, 0x00);   // I2OSP(0, 1)
    sha256_update(c, dst, dst_len);
    sha256_put(c, (uint8_t)dst_len);
    sha256_final(c, b0);
    for (int i = 1; i <= 8; i++) {
        uint32_t t[8];
        for (int k = 0; k < 8; k++) t[k] = (i == 1) ? b0[k] : (b0[k] ^ out[8 * (i - 2) + k]);
        sha256_init(c);
        sha256_update_words(c, t, 8);
        sha256_put(c, (uint8_t)i);
        sha256_update(c, dst, dst_len);
        sha256_put(c, (uint8_t)dst_len);
        sha256_final(c, out + 8 * (i - 1));
    }
}

// blscurve/bls_sig_min_pubkey.nim:31
BLS_TABLE uint8_t DST_ETH2[43] = {
    'B','L','S','_','S','I','G','_','B','L','S','1','2','3','8','1','G','2','_','X','M','D',':','S','H','A','-',
    '2','5','6','_','S','S','W','U','_','R','O','_','P','O','P','_'};

#ifdef __CUDACC__
// expand_message_xmd(msg, DST_ETH2, 256) for the 32-byte messages of a SignatureSet: the 18 blocks are assembled as
// words (no byte-wise absorption) and compressed in registers.  Same 64 output words as expand_message_xmd_256.
//   b0 = H(Z_pad | msg | 01 00 | 00 | DST | 2b): after the pre-absorbed Z_pad block, 79 bytes = blocks
//        [msg(32) 01 00 00 DST[0..28]] [DST[29..42] 2b 80 0.. len=1144]
//   bi = H(x(32) | i | DST | 2b), 77 bytes = [x(32) i DST[0..30]] [DST[31..42] 2b 80 0.. len=616]
__device__ __forceinline__ uint32_t dst_eth2_byte(int k) { return k < 43 ? (uint32_t)DST_ETH2[k] : (k == 43 ? 0x2bu : (k == 44 ? 0x80u : 0u)); }
// big-endian word of the virtual string DST | 2b | 80 | 00.. starting at byte offset k
__device__ __forceinline__ uint32_t dst_eth2_word(int k) {
    return (dst_eth2_byte(k) << 24) | (dst_eth2_byte(k + 1) << 16) | (dst_eth2_byte(k + 2) << 8) | dst_eth2_byte(k + 3);
}
BLS_NOINLINE void expand_message_xmd_256_eth2(uint32_t *out, const uint8_t *msg32) {
    uint32_t h[8], w[16], b0[8];
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = SHA256_ZPAD_STATE[i];
#pragma unroll
    for (int i = 0; i < 8; i++)
        w[i] = ((uint32_t)msg32[4 * i] << 24) | ((uint32_t)msg32[4 * i + 1] << 16) | ((uint32_t)msg32[4 * i + 2] << 8) | msg32[4 * i + 3];
    w[8] = 0x01000000u | dst_eth2_byte(0);                       // 01 00 00 DST[0]
#pragma unroll
    for (int i = 9; i < 16; i++) w[i] = dst_eth2_word(1 + 4 * (i - 9));       // DST[1..28]
    sha256_block_regs(h, w);
#pragma unroll
    for (int i = 0; i < 14; i++) w[i] = dst_eth2_word(29 + 4 * i);            // DST[29..42] 2b 80 00..
    w[14] = 0;
    w[15] = 1144;
    sha256_block_regs(h, w);
#pragma unroll
    for (int i = 0; i < 8; i++) b0[i] = h[i];
    for (int blk = 1; blk <= 8; blk++) {
        h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
        h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = blk == 1 ? b0[i] : (b0[i] ^ out[8 * (blk - 2) + i]);
        w[8] = ((uint32_t)blk << 24) | (dst_eth2_byte(0) << 16) | (dst_eth2_byte(1) << 8) | dst_eth2_byte(2);
#pragma unroll
        for (int i = 9; i < 16; i++) w[i] = dst_eth2_word(3 + 4 * (i - 9));   // DST[3..30]
        sha256_block_regs(h, w);
#pragma unroll
        for (int i = 0; i < 14; i++) w[i] = dst_eth2_word(31 + 4 * i);        // DST[31..42] 2b 80 00..
        w[14] = 0;
        w[15] = 616;
        sha256_block_regs(h, w);
#pragma unroll
        for (int i = 0; i < 8; i++) out[8 * (blk - 1) + i] = h[i];
    }
}
#endif

// 64 bytes big-endian (16 BE words, most significant first) -> Fp in Montgomery form
BLS_FN void fp_from_be64(fp &r, const uint32_t *be) {
    fp lo, hi, t0, t1;
    for (int i = 0; i < 8; i++) { lo.l[i] = be[15 - i]; hi.l[i] = be[7 - i]; }
    for (int i = 8; i < 12; i++) { lo.l[i] = 0; hi.l[i] = 0; }
    fp_mul_ni(t0, lo, FP_R2);
    fp_mul_ni(t1, hi, FP_R2_2_256);
    fp_add(r, t0, t1);
}

// hash_to_field(msg, count=2) over Fp2: u0 = (e0, e1), u1 = (e2, e3)
BLS_FN void hash_to_field_fp2x2(fp2 &u0, fp2 &u1, const uint8_t *msg, size_t msg_len, const uint8_t *dst,
                                uint32_t dst_len) {
    uint32_t xmd[64];
    expand_message_xmd_256(xmd, msg, msg_len, dst, dst_len);
    fp_from_be64(u0.c0, xmd);
    fp_from_be64(u0.c1, xmd + 16);
    fp_from_be64(u1.c0, xmd + 32);
    fp_from_be64(u1.c1, xmd + 48);
}

#ifdef __CUDACC__
// the same for the 32-byte message of a SignatureSet under DST_ETH2 (word-level XMD, blocks compressed in registers)
BLS_FN void hash_to_field_fp2x2_eth2(fp2 &u0, fp2 &u1, const uint8_t *msg32) {
    uint32_t xmd[64];
    expand_message_xmd_256_eth2(xmd, msg32);
    fp_from_be64(u0.c0, xmd);
    fp_from_be64(u0.c1, xmd + 16);
    fp_from_be64(u1.c0, xmd + 32);
    fp_from_be64(u1.c1, xmd + 48);
}
#endif

}  // namespace bls
