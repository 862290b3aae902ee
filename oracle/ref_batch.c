/*
 * oracle/ref_batch.c — TEST INFRASTRUCTURE ONLY (checker + CPU baseline), never shipped.
 *
 * C restatement of the thin Nim layer that nim-blscurve puts on top of BLST for the
 * batch-verification path, linked against the reference's own BLST build
 * (oracle/_ref/libblst_ref.so, compiled from /root/reference/vendor/blst where it lies).
 * Nim is not installed here, so the Nim driver itself cannot run; the arithmetic it calls can.
 *
 * What is restated (file:line under /root/reference):
 *   scalar chain + chunk tag   blscurve/blst/blst_min_pubkey_sig_core.nim:476-505, :545-556
 *   update() arguments         blscurve/blst/blst_min_pubkey_sig_core.nim:558-568
 *   chunk split                blscurve/parallel_chunks.nim:42-55
 *   serial driver              blscurve/bls_batch_verifier.nim:121-160
 *   parallel driver + merge    blscurve/bls_batch_verifier.nim:296-371
 *   aggregateAll               blscurve/blst/blst_min_pubkey_sig_core.nim:179-195
 *   combine (scalar order)     blscurve/blst/blst_min_pubkey_sig_core.nim:570-647
 *   test inputs                tests/t_batch_verifier.nim:34-47, :60
 *   MSM bench inputs/shape     benchmarks/bls12381_msm_g1.nim:22-63
 *   final verify (for GT)      vendor/blst/src/aggregate.c:460-501
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>
#include "blst_decl.h"

static const char DST[] = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"; /* bls_sig_min_pubkey.nim:31 */
#define DST_LEN 43

typedef struct { uint8_t b[320]; } sigset_t320;
#define SET_PK(s)  ((const blst_p1_affine *)((s)->b))
#define SET_MSG(s) ((s)->b + 96)
#define SET_SIG(s) ((const blst_p2_affine *)((s)->b + 128))

int ref_ncores(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* parallel_chunks.nim:42-55 */
static void chunk_range(size_t nchunks, size_t total, size_t cid, size_t *off, size_t *len) {
    size_t base = total / nchunks, rem = total % nchunks;
    if (cid < rem) { *off = (base + 1) * cid; *len = base + 1; }
    else           { *off = base * cid + rem; *len = base; }
}

static void seed_init(uint8_t seed[32], const uint8_t srb[32], int tagged, uint64_t chunk) {
    if (tagged) {                       /* SHA256(srb || LE64(chunkID)) */
        uint8_t buf[40]; memcpy(buf, srb, 32);
        for (int i = 0; i < 8; i++) buf[32 + i] = (uint8_t)(chunk >> (8 * i));
        blst_sha256(seed, buf, 40);
    } else {
        blst_sha256(seed, srb, 32);     /* serial: SHA256(srb) */
    }
}

static uint64_t seed_next(uint8_t seed[32]) {
    uint64_t r;
    do {
        uint8_t t[32]; blst_sha256(t, seed, 32); memcpy(seed, t, 32);
        r = 0; for (int i = 7; i >= 0; i--) r = (r << 8) | seed[i];
    } while (r == 0);
    return r;
}

void ref_rlc_scalars(const uint8_t srb[32], size_t n, uint32_t chunks, uint64_t *out) {
    size_t nb = chunks == 0 ? 1 : (n < chunks ? n : chunks);
    for (size_t c = 0; c < nb; c++) {
        size_t off, len; uint8_t seed[32];
        chunk_range(nb, n, c, &off, &len);
        seed_init(seed, srb, chunks != 0, c);
        for (size_t i = off; i < off + len; i++) out[i] = seed_next(seed);
    }
}

/* one chunk: ctx.init + update* + commit.  Returns 1 on success. */
static int run_chunk(blst_pairing *ctx, const sigset_t320 *sets, size_t off, size_t len,
                     const uint8_t srb[32], int tagged, uint64_t chunk, const uint64_t *scalars) {
    uint8_t seed[32];
    blst_pairing_init(ctx, 1, (const uint8_t *)DST, DST_LEN);
    seed_init(seed, srb, tagged, chunk);
    for (size_t i = off; i < off + len; i++) {
        uint64_t r = scalars ? scalars[i] : seed_next(seed);
        uint8_t sc[32] = {0};
        for (int k = 0; k < 8; k++) sc[k] = (uint8_t)(r >> (8 * k));
        int err = blst_pairing_chk_n_mul_n_aggr_pk_in_g1(ctx, SET_PK(&sets[i]), 0, SET_SIG(&sets[i]), 0,
                                                         sc, 64, SET_MSG(&sets[i]), 32, NULL, 0);
        if (err != 0) return 0;
    }
    blst_pairing_commit(ctx);
    return 1;
}

/* Σ [r_i]·sig_i recomputed through the public point API so the GT bytes can be produced
   (aggregate.c:479-500 keeps S private). */
static void sum_rsig(blst_p2 *S, int *any, const sigset_t320 *sets, size_t n, const uint64_t *r) {
    *any = 0;
    for (size_t i = 0; i < n; i++) {
        static const uint8_t zero[192];
        if (memcmp(SET_SIG(&sets[i]), zero, 192) == 0) continue;
        blst_p2 p; uint8_t sc[8];
        for (int k = 0; k < 8; k++) sc[k] = (uint8_t)(r[i] >> (8 * k));
        blst_p2_from_affine(&p, SET_SIG(&sets[i]));
        blst_p2_mult(&p, &p, sc, 64);
        if (*any) blst_p2_add_or_double(S, S, &p); else { *S = p; *any = 1; }
    }
}

int ref_batch_verify(const uint8_t *sets_, size_t n, const uint8_t srb[32], uint32_t chunks,
                     const uint64_t *scalars_in, uint8_t gt_out[576]) {
    const sigset_t320 *sets = (const sigset_t320 *)sets_;
    memset(gt_out, 0, 576);
    if (n == 0) return 0;                                   /* bls_batch_verifier.nim:137, :312 */
    size_t nb = chunks == 0 ? 1 : (n < chunks ? n : chunks);/* :316 */
    size_t psz = blst_pairing_sizeof();
    uint8_t *ctxs = malloc(psz * nb);
    uint64_t *r = malloc(8 * n);
    if (scalars_in) memcpy(r, scalars_in, 8 * n); else ref_rlc_scalars(srb, n, chunks, r);
    int ok = 1;
    for (size_t c = 0; c < nb && ok; c++) {
        size_t off, len; chunk_range(nb, n, c, &off, &len);
        ok = run_chunk((blst_pairing *)(ctxs + psz * c), sets, off, len, srb, chunks != 0, c, scalars_in ? r : NULL);
    }
    if (ok) {
        for (size_t c = 1; c < nb && ok; c++)               /* merge order is immaterial (commutative) */
            ok = blst_pairing_merge((blst_pairing *)ctxs, (blst_pairing *)(ctxs + psz * c)) == 0;
    }
    if (ok) {
        int verdict = blst_pairing_finalverify((blst_pairing *)ctxs, NULL) ? 1 : 0;
        /* GT = FE( conj(ML(S, G1)) * ctx.GT )  — aggregate.c:479-496 */
        blst_p2 S; int any; blst_fp12 gt, acc;
        sum_rsig(&S, &any, sets, n, r);
        if (any) {
            blst_p2_affine Sa; blst_p2_to_affine(&Sa, &S);
            blst_miller_loop(&gt, &Sa, blst_p1_affine_generator());
        } else gt = *blst_fp12_one();
        blst_fp12_conjugate(&gt);
        acc = *blst_pairing_as_fp12((blst_pairing *)ctxs);
        blst_fp12_mul(&gt, &gt, &acc);
        blst_final_exp(&gt, &gt);
        blst_bendian_from_fp12(gt_out, &gt);
        if (blst_fp12_is_one(&gt) != (verdict != 0)) {
            fprintf(stderr, "ref_batch_verify: GT recomputation disagrees with blst_pairing_finalverify\n");
            abort();
        }
        ok = verdict;
    }
    free(ctxs); free(r);
    return ok;
}

/* ---------------- pthreads replica of batchVerifyParallel (CPU baseline) ---------------- */
typedef struct {
    blst_pairing *ctx; const sigset_t320 *sets; size_t off, len; const uint8_t *srb; uint64_t chunk; int ok;
} chunk_job;

static void *chunk_thread(void *p) {
    chunk_job *j = p;
    j->ok = run_chunk(j->ctx, j->sets, j->off, j->len, j->srb, 1, j->chunk, NULL);
    return NULL;
}

int ref_batch_verify_mt(const uint8_t *sets_, size_t n, const uint8_t srb[32], int threads) {
    const sigset_t320 *sets = (const sigset_t320 *)sets_;
    if (n == 0) return 0;
    if (threads <= 1 || n < 3) {                            /* bls_batch_verifier.nim:440, :468 */
        uint8_t gt[576]; return ref_batch_verify(sets_, n, srb, 0, NULL, gt);
    }
    size_t nb = n < (size_t)threads ? n : (size_t)threads;
    size_t psz = blst_pairing_sizeof();
    uint8_t *ctxs = malloc(psz * nb);
    chunk_job *jobs = calloc(nb, sizeof(chunk_job));
    pthread_t *th = calloc(nb, sizeof(pthread_t));
    for (size_t c = 0; c < nb; c++) {
        jobs[c].ctx = (blst_pairing *)(ctxs + psz * c); jobs[c].sets = sets; jobs[c].srb = srb; jobs[c].chunk = c;
        chunk_range(nb, n, c, &jobs[c].off, &jobs[c].len);
        pthread_create(&th[c], NULL, chunk_thread, &jobs[c]);
    }
    int ok = 1;
    for (size_t c = 0; c < nb; c++) { pthread_join(th[c], NULL); ok &= jobs[c].ok; }
    for (size_t c = 1; c < nb && ok; c++)
        ok = blst_pairing_merge((blst_pairing *)ctxs, (blst_pairing *)(ctxs + psz * c)) == 0;
    if (ok) ok = blst_pairing_finalverify((blst_pairing *)ctxs, NULL) ? 1 : 0;
    free(ctxs); free(jobs); free(th);
    return ok;
}

double ref_time_batch_verify(const uint8_t *sets, size_t n, const uint8_t srb[32], int threads, int reps) {
    double best = 1e30;
    for (int i = 0; i < reps; i++) {
        double t0 = now_s();
        int ok = ref_batch_verify_mt(sets, n, srb, threads);
        double t = now_s() - t0;
        if (!ok) return -1.0;
        if (t < best) best = t;
    }
    return best;
}

/* ---------------- input generation (t_batch_verifier.nim:34-47) ---------------- */
static void keygen_seed(blst_scalar *sk, uint64_t seed) {
    uint8_t ikm[32] = {0};
    for (int i = 0; i < 8; i++) ikm[i] = (uint8_t)(seed >> (8 * i));
    blst_keygen(sk, ikm, 32, NULL, 0);
}

void ref_make_set_hashed(uint64_t seed, const uint8_t hashed[32], uint8_t out[320]) {
    blst_scalar sk; blst_p1 pk; blst_p2 h;
    keygen_seed(&sk, seed);
    blst_sk_to_pk_in_g1(&pk, &sk);
    blst_p1_to_affine((blst_p1_affine *)out, &pk);
    memcpy(out + 96, hashed, 32);
    blst_hash_to_g2(&h, hashed, 32, (const uint8_t *)DST, DST_LEN, NULL, 0);
    blst_sign_pk_in_g1(&h, &h, &sk);
    blst_p2_to_affine((blst_p2_affine *)(out + 128), &h);
}

void ref_make_set(uint64_t seed, const uint8_t *message, size_t len, uint8_t out[320]) {
    uint8_t hashed[32];
    blst_sha256(hashed, message, len);
    ref_make_set_hashed(seed, hashed, out);
}

typedef struct { uint64_t start; size_t lo, hi; const uint8_t *prefix; size_t plen; uint8_t *out; } gen_job;

static void *gen_thread(void *p) {
    gen_job *j = p;
    for (size_t i = j->lo; i < j->hi; i++) {
        char buf[96]; uint64_t seed = j->start + i;
        memcpy(buf, j->prefix, j->plen);
        int k = snprintf(buf + j->plen, sizeof(buf) - j->plen, "%llu", (unsigned long long)seed);
        ref_make_set(seed, (const uint8_t *)buf, j->plen + (size_t)k, j->out + 320 * i);
    }
    return NULL;
}

static void run_split(void *(*fn)(void *), void *jobs, size_t jobsz, size_t n, int threads,
                      void (*setrange)(void *, size_t, size_t)) {
    if (threads <= 0) threads = ref_ncores();
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t th[256]; if (threads > 256) threads = 256;
    for (int t = 0; t < threads; t++) {
        size_t off, len; chunk_range((size_t)threads, n, (size_t)t, &off, &len);
        setrange((char *)jobs + jobsz * t, off, off + len);
        pthread_create(&th[t], NULL, fn, (char *)jobs + jobsz * t);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

static void gen_setrange(void *j, size_t lo, size_t hi) { ((gen_job *)j)->lo = lo; ((gen_job *)j)->hi = hi; }

void ref_make_sets(uint64_t start, size_t n, const uint8_t *prefix, size_t plen, uint8_t *out, int threads) {
    gen_job jobs[256];
    if (plen > 64) plen = 64;
    for (int t = 0; t < 256; t++) { jobs[t].start = start; jobs[t].prefix = prefix; jobs[t].plen = plen; jobs[t].out = out; }
    run_split(gen_thread, jobs, sizeof(gen_job), n, threads, gen_setrange);
}

/* The benchmark workload, on the host: the same bytes blsgpu_make_sets generates on the device (include/blsgpu.h:
   sk_i = 1-or'ed (SHA256(LE64(seed) || 0^24 || LE64(first+i)) read as a big-endian integer, mod 2^250),
   pk = [sk]G1, msg = SHA256("blsgpu" || LE64(first+i)), sig = [sk]H(msg)), so that the reference arm of bench.py can
   verify the first n sets of the SAME workload without a GPU.  blst_scalar is 32 little-endian bytes (blst.h:61). */
typedef struct { uint64_t seed; size_t first, lo, hi; uint8_t *out; } dgen_job;
static void *dgen_thread(void *p) {
    dgen_job *j = p;
    for (size_t i = j->lo; i < j->hi; i++) {
        uint64_t idx = j->first + i;
        uint8_t pre[40] = {0}, dig[32], m14[14] = {'b', 'l', 's', 'g', 'p', 'u'}, msg[32];
        for (int k = 0; k < 8; k++) { pre[k] = (uint8_t)(j->seed >> (8 * k)); pre[32 + k] = (uint8_t)(idx >> (8 * k)); m14[6 + k] = (uint8_t)(idx >> (8 * k)); }
        blst_sha256(dig, pre, 40);
        blst_scalar sk;
        for (int k = 0; k < 32; k++) sk.b[k] = dig[31 - k];
        sk.b[31] &= 0x03;
        sk.b[0] |= 1;
        blst_sha256(msg, m14, 14);
        uint8_t *out = j->out + 320 * i;
        blst_p1 pk; blst_p2 h;
        blst_sk_to_pk_in_g1(&pk, &sk);
        blst_p1_to_affine((blst_p1_affine *)out, &pk);
        memcpy(out + 96, msg, 32);
        blst_hash_to_g2(&h, msg, 32, (const uint8_t *)DST, DST_LEN, NULL, 0);
        blst_sign_pk_in_g1(&h, &h, &sk);
        blst_p2_to_affine((blst_p2_affine *)(out + 128), &h);
    }
    return NULL;
}
static void dgen_setrange(void *j, size_t lo, size_t hi) { ((dgen_job *)j)->lo = lo; ((dgen_job *)j)->hi = hi; }
void ref_make_sets_device_recipe(uint64_t seed, size_t first, size_t n, uint8_t *out, int threads) {
    dgen_job jobs[256];
    for (int t = 0; t < 256; t++) { jobs[t].seed = seed; jobs[t].first = first; jobs[t].out = out; }
    run_split(dgen_thread, jobs, sizeof(dgen_job), n, threads, dgen_setrange);
}

/* committee of nkeys signers on one message: aggregate pubkey + aggregate signature as ONE set
   (bls_batch_verifier.nim:38-40; fastAggregateVerify caller pattern bls_sig_min_pubkey.nim:234-258) */
int ref_aggregate_g1(const uint8_t *pts, size_t n, uint8_t out[96]);
int ref_aggregate_g2(const uint8_t *pts, size_t n, uint8_t out[192]);

typedef struct { uint64_t start; size_t lo, hi; const uint8_t *hashed; uint8_t *sets; } fa_job;
static void *fa_thread(void *p) {
    fa_job *j = p;
    for (size_t i = j->lo; i < j->hi; i++) ref_make_set_hashed(j->start + i, j->hashed, j->sets + 320 * i);
    return NULL;
}
static void fa_setrange(void *j, size_t lo, size_t hi) { ((fa_job *)j)->lo = lo; ((fa_job *)j)->hi = hi; }

int ref_fast_aggregate_set(uint64_t start, size_t nkeys, const uint8_t hashed[32], uint8_t *pks_out,
                           uint8_t set_out[320], int threads) {
    uint8_t *sets = malloc(320 * nkeys), *sigs = malloc(192 * nkeys);
    fa_job jobs[256];
    for (int t = 0; t < 256; t++) { jobs[t].start = start; jobs[t].hashed = hashed; jobs[t].sets = sets; }
    run_split(fa_thread, jobs, sizeof(fa_job), nkeys, threads, fa_setrange);
    for (size_t i = 0; i < nkeys; i++) {
        memcpy(pks_out + 96 * i, sets + 320 * i, 96);
        memcpy(sigs + 192 * i, sets + 320 * i + 128, 192);
    }
    int ok = ref_aggregate_g1(pks_out, nkeys, set_out);
    memcpy(set_out + 96, hashed, 32);
    ok &= ref_aggregate_g2(sigs, nkeys, set_out + 128);
    free(sets); free(sigs);
    return ok;
}

/* ---------------- hash_to_G2 ---------------- */
void ref_hash_to_g2(const uint8_t *msgs, size_t n, size_t msg_len, const uint8_t *dst, size_t dst_len,
                    uint8_t *comp_out, uint8_t *aff_out) {
    for (size_t i = 0; i < n; i++) {
        blst_p2 h; blst_p2_affine a;
        blst_hash_to_g2(&h, msgs + i * msg_len, msg_len, dst, dst_len, NULL, 0);
        blst_p2_to_affine(&a, &h);
        memcpy(aff_out + 192 * i, &a, 192);
        blst_p2_affine_compress(comp_out + 96 * i, &a);
    }
}

/* ---------------- aggregateAll (blst_min_pubkey_sig_core.nim:179-195) ---------------- */
int ref_aggregate_g1(const uint8_t *pts, size_t n, uint8_t out[96]) {
    if (n == 0) return 0;
    blst_p1 acc; blst_p1_from_affine(&acc, (const blst_p1_affine *)pts);
    for (size_t i = 1; i < n; i++) blst_p1_add_or_double_affine(&acc, &acc, (const blst_p1_affine *)(pts + 96 * i));
    blst_p1_to_affine((blst_p1_affine *)out, &acc);
    return 1;
}

int ref_aggregate_g2(const uint8_t *pts, size_t n, uint8_t out[192]) {
    if (n == 0) return 0;
    blst_p2 acc; blst_p2_from_affine(&acc, (const blst_p2_affine *)pts);
    for (size_t i = 1; i < n; i++) blst_p2_add_or_double_affine(&acc, &acc, (const blst_p2_affine *)(pts + 192 * i));
    blst_p2_to_affine((blst_p2_affine *)out, &acc);
    return 1;
}

/* ---------------- subtractAll (blst_min_pubkey_sig_core.nim:197-209) ---------------- */
void ref_subtract_g1(uint8_t dst[96], const uint8_t *pts, size_t n) {
    if (n == 0) return;                                                                 /* :199-200 */
    blst_p1 acc; blst_p1_from_affine(&acc, (const blst_p1_affine *)pts);                /* :202 */
    for (size_t i = 1; i < n; i++) blst_p1_add_or_double_affine(&acc, &acc, (const blst_p1_affine *)(pts + 96 * i));
    blst_p1_cneg(&acc, 1);                                                              /* :204-207 */
    blst_p1_add_or_double_affine(&acc, &acc, (const blst_p1_affine *)dst);              /* :208 */
    blst_p1_to_affine((blst_p1_affine *)dst, &acc);                                     /* :209 */
}

void ref_subtract_g2(uint8_t dst[192], const uint8_t *pts, size_t n) {
    if (n == 0) return;
    blst_p2 acc; blst_p2_from_affine(&acc, (const blst_p2_affine *)pts);
    for (size_t i = 1; i < n; i++) blst_p2_add_or_double_affine(&acc, &acc, (const blst_p2_affine *)(pts + 192 * i));
    blst_p2_cneg(&acc, 1);
    blst_p2_add_or_double_affine(&acc, &acc, (const blst_p2_affine *)dst);
    blst_p2_to_affine((blst_p2_affine *)dst, &acc);
}

void ref_g2_neg(const uint8_t in[192], uint8_t out[192]) {
    blst_p2 p; blst_p2_from_affine(&p, (const blst_p2_affine *)in);
    blst_p2_cneg(&p, 1);
    blst_p2_to_affine((blst_p2_affine *)out, &p);
}

/* ---------------- G1 MSM (benchmarks/bls12381_msm_g1.nim:48-63) ---------------- */
int ref_msm_g1(const uint8_t *pts, const uint8_t *scalars, size_t n, size_t nbits, uint8_t out[96]) {
    if (n == 0) { memset(out, 0, 96); return 0; }
    const blst_p1_affine *pp[2] = { (const blst_p1_affine *)pts, NULL };
    const uint8_t *ss[2] = { scalars, NULL };
    void *scratch = malloc(blst_p1s_mult_pippenger_scratch_sizeof(n));
    blst_p1 r;
    blst_p1s_mult_pippenger(&r, pp, n, ss, nbits, scratch);
    blst_p1_to_affine((blst_p1_affine *)out, &r);
    free(scratch);
    return 1;
}

/* G2 MSM: blst_p2s_mult_pippenger (multi_scalar.c:442-446), the shape combine() uses with nbits = 64
 * (blst_min_pubkey_sig_core.nim:637-644) */
int ref_msm_g2(const uint8_t *pts, const uint8_t *scalars, size_t n, size_t nbits, uint8_t out[192]) {
    if (n == 0) { memset(out, 0, 192); return 0; }
    const blst_p2_affine *pp[2] = { (const blst_p2_affine *)pts, NULL };
    const uint8_t *ss[2] = { scalars, NULL };
    void *scratch = malloc(blst_p2s_mult_pippenger_scratch_sizeof(n));
    blst_p2 r;
    blst_p2s_mult_pippenger(&r, pp, n, ss, nbits, scratch);
    blst_p2_to_affine((blst_p2_affine *)out, &r);
    free(scratch);
    return 1;
}

double ref_time_msm_g1(const uint8_t *pts, const uint8_t *scalars, size_t n, size_t nbits, int reps) {
    const blst_p1_affine *pp[2] = { (const blst_p1_affine *)pts, NULL };
    const uint8_t *ss[2] = { scalars, NULL };
    void *scratch = malloc(blst_p1s_mult_pippenger_scratch_sizeof(n));
    blst_p1 r; double best = 1e30;
    for (int i = 0; i < reps; i++) {
        double t0 = now_s();
        blst_p1s_mult_pippenger(&r, pp, n, ss, nbits, scratch);
        double t = now_s() - t0;
        if (t < best) best = t;
    }
    free(scratch);
    return best;
}

static uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

typedef struct { uint64_t seed; size_t lo, hi; uint8_t *pts, *sc; } msm_job;
static void *msm_thread(void *p) {
    msm_job *j = p;
    for (size_t i = j->lo; i < j->hi; i++) {
        uint64_t s = j->seed ^ (0xFACADEULL + i * 0x100000001b3ULL);
        uint8_t k[16] = {0}; uint64_t a = splitmix(&s), b = splitmix(&s);
        memcpy(k, &a, 8); memcpy(k + 8, &b, 4);              /* 96-bit point multiplier */
        blst_p1 t; blst_p1_from_affine(&t, blst_p1_affine_generator());
        blst_p1_mult(&t, &t, k, 96);
        blst_p1_to_affine((blst_p1_affine *)(j->pts + 96 * i), &t);
        for (int w = 0; w < 4; w++) { uint64_t v = splitmix(&s); memcpy(j->sc + 32 * i + 8 * w, &v, 8); }
        j->sc[32 * i + 31] &= 0x7f;                           /* 255-bit coefficients */
    }
    return NULL;
}
static void msm_setrange(void *j, size_t lo, size_t hi) { ((msm_job *)j)->lo = lo; ((msm_job *)j)->hi = hi; }

void ref_msm_inputs(uint64_t seed, size_t n, uint8_t *pts, uint8_t *scalars, int threads) {
    msm_job jobs[256];
    for (int t = 0; t < 256; t++) { jobs[t].seed = seed; jobs[t].pts = pts; jobs[t].sc = scalars; }
    run_split(msm_thread, jobs, sizeof(msm_job), n, threads, msm_setrange);
}

/* ---------------- combine (blst_min_pubkey_sig_core.nim:570-647) ---------------- */
void ref_combine(const uint8_t srb[32], const uint8_t *pks, const uint8_t *sigs, size_t n,
                 uint8_t pk_out[96], uint8_t sig_out[192]) {
    if (n == 1) { memcpy(pk_out, pks, 96); memcpy(sig_out, sigs, 192); return; }
    uint8_t seed[32]; memcpy(seed, srb, 32);
    uint64_t *sc = malloc(8 * n); int avail = 0;
    for (size_t i = 0; i < n; i++) {
        for (;;) {
            if (avail == 0) { uint8_t t[32]; blst_sha256(t, seed, 32); memcpy(seed, t, 32); avail = 4; }
            avail--;
            uint64_t v = 0; for (int k = 7; k >= 0; k--) v = (v << 8) | seed[8 * avail + k];
            if (v != 0) { sc[i] = v; break; }
        }
    }
    /* scalars are consumed as one contiguous LE byte string, 8 bytes per entry, nbits=64 */
    const blst_p1_affine *pp[2] = { (const blst_p1_affine *)pks, NULL };
    const blst_p2_affine *qq[2] = { (const blst_p2_affine *)sigs, NULL };
    const uint8_t *ss[2] = { (const uint8_t *)sc, NULL };
    size_t s1 = blst_p1s_mult_pippenger_scratch_sizeof(n), s2 = blst_p2s_mult_pippenger_scratch_sizeof(n);
    void *scratch = malloc(s1 > s2 ? s1 : s2);
    blst_p1 r1; blst_p2 r2;
    blst_p1s_mult_pippenger(&r1, pp, n, ss, 64, scratch);
    blst_p1_to_affine((blst_p1_affine *)pk_out, &r1);
    blst_p2s_mult_pippenger(&r2, qq, n, ss, 64, scratch);
    blst_p2_to_affine((blst_p2_affine *)sig_out, &r2);
    free(scratch); free(sc);
}

/* ---------------- multi-rank decomposition ----------------
 * One rank's share [first, first+n) of a batch of total_n sets: scalars come from the GLOBAL derivation, the
 * partial is conj(ML(S_k, G1)) * prod ML([r_i]pk_i, H(m_i)) in BLST's in-memory fp12 layout.  By bilinearity the
 * product of all partials, finally exponentiated, equals the single-context result of ref_batch_verify. */
int ref_partial(const uint8_t *sets_, size_t n, size_t first, size_t total_n, const uint8_t srb[32], uint32_t chunks,
                uint8_t partial_out[576], int *flag) {
    const sigset_t320 *sets = (const sigset_t320 *)sets_;
    blst_fp12 acc = *blst_fp12_one();
    *flag = 0;
    if (n) {
        uint64_t *r = malloc(8 * total_n);
        ref_rlc_scalars(srb, total_n, chunks, r);
        blst_pairing *ctx = malloc(blst_pairing_sizeof());
        if (!run_chunk(ctx, sets, 0, n, srb, 0, 0, r + first)) { *flag = 1; }
        else {
            blst_p2 S; int any; blst_fp12 gs;
            sum_rsig(&S, &any, sets, n, r + first);
            if (any) { blst_p2_affine Sa; blst_p2_to_affine(&Sa, &S); blst_miller_loop(&gs, &Sa, blst_p1_affine_generator()); }
            else gs = *blst_fp12_one();
            blst_fp12_conjugate(&gs);
            blst_fp12_mul(&acc, &gs, blst_pairing_as_fp12(ctx));
        }
        free(ctx); free(r);
    }
    memcpy(partial_out, &acc, 576);
    return 1;
}

int ref_finalize(const uint8_t *partials, size_t count, uint8_t gt_out[576]) {
    memset(gt_out, 0, 576);
    if (count == 0) return 0;
    blst_fp12 acc, t;
    memcpy(&acc, partials, 576);
    for (size_t i = 1; i < count; i++) { memcpy(&t, partials + 576 * i, 576); blst_fp12_mul(&acc, &acc, &t); }
    blst_final_exp(&acc, &acc);
    blst_bendian_from_fp12(gt_out, &acc);
    return blst_fp12_is_one(&acc) ? 1 : 0;
}

/* ---------------- aggregateVerify / fastAggregateVerify / verify ----------------
 * bls_sig_min_pubkey.nim:127-273 over ContextCoreAggregateVerify (blst_min_pubkey_sig_core.nim:310-398): update() per
 * (publicKey, message) with no signature, finish() adds the signature through the same call with PK = nil, then
 * commit + finalverify.  n pairs, arbitrary message lengths (msg i = msgs[offs[i] .. offs[i+1])), caller's DST.
 * GT bytes are recomputed like in ref_batch_verify (informative on failing inputs). */
int ref_aggregate_verify(const uint8_t *pks, size_t n, const uint8_t *msgs, const uint32_t *offs,
                         const uint8_t *dst, size_t dst_len, const uint8_t sig[192], uint8_t gt_out[576]) {
    static const uint8_t zero[192];
    memset(gt_out, 0, 576);
    if (n == 0) return 0;                                   /* bls_sig_min_pubkey.nim:140, :167, :189, :214 */
    blst_pairing *ctx = malloc(blst_pairing_sizeof());
    blst_pairing_init(ctx, 1, dst, dst_len);
    int ok = 1;
    for (size_t i = 0; i < n && ok; i++)
        ok = blst_pairing_chk_n_aggr_pk_in_g1(ctx, (const blst_p1_affine *)(pks + 96 * i), 0, NULL, 0,
                                              msgs + offs[i], offs[i + 1] - offs[i], NULL, 0) == 0;
    if (ok) ok = blst_pairing_chk_n_aggr_pk_in_g1(ctx, NULL, 0, (const blst_p2_affine *)sig, 0, NULL, 0, NULL, 0) == 0;
    if (ok) {
        blst_pairing_commit(ctx);
        int verdict = blst_pairing_finalverify(ctx, NULL) ? 1 : 0;
        blst_fp12 gt, acc;
        if (memcmp(sig, zero, 192) != 0) blst_miller_loop(&gt, (const blst_p2_affine *)sig, blst_p1_affine_generator());
        else gt = *blst_fp12_one();
        blst_fp12_conjugate(&gt);
        acc = *blst_pairing_as_fp12(ctx);
        blst_fp12_mul(&gt, &gt, &acc);
        blst_final_exp(&gt, &gt);
        blst_bendian_from_fp12(gt_out, &gt);
        if (blst_fp12_is_one(&gt) != (verdict != 0)) {
            fprintf(stderr, "ref_aggregate_verify: GT recomputation disagrees with blst_pairing_finalverify\n");
            abort();
        }
        ok = verdict;
    }
    free(ctx);
    return ok;
}

/* fastAggregateVerify(publicKeys, message, signature), bls_sig_min_pubkey.nim:238-258: aggregateAll + coreVerify */
int ref_fast_aggregate_verify(const uint8_t *pks, size_t n, const uint8_t *msg, size_t msg_len,
                              const uint8_t *dst, size_t dst_len, const uint8_t sig[192], uint8_t gt_out[576]) {
    uint8_t agg[96];
    uint32_t offs[2] = {0, (uint32_t)msg_len};
    memset(gt_out, 0, 576);
    if (!ref_aggregate_g1(pks, n, agg)) return 0;
    return ref_aggregate_verify(agg, 1, msg, offs, dst, dst_len, sig, gt_out);
}

/* ---------------- deserialisation + checks ----------------
 * PublicKey.fromBytes / Signature.fromBytes, bls_sig_io.nim:42-122: uncompress (48/96 B) or deserialize (96/192 B),
 * public keys reject infinity, then the subgroup check.  Returns the BLST_ERROR (0 = success; 1 bad encoding,
 * 2 not on curve, 3 not in group, 6 public key is infinity) and writes the affine point on success. */
int ref_pubkey_from_bytes(const uint8_t *in, size_t len, int group_check, uint8_t out[96]) {
    blst_p1_affine a;
    int err = len == 48 ? blst_p1_uncompress(&a, in) : blst_p1_deserialize(&a, in);
    memset(out, 0, 96);
    if (err) return err;
    if (blst_p1_affine_is_inf(&a)) return 6;                /* BLST_PK_IS_INFINITY */
    if (group_check && !blst_p1_affine_in_g1(&a)) return 3; /* BLST_POINT_NOT_IN_GROUP */
    memcpy(out, &a, 96);
    return 0;
}

int ref_signature_from_bytes(const uint8_t *in, size_t len, int group_check, uint8_t out[192]) {
    blst_p2_affine a;
    int err = len == 96 ? blst_p2_uncompress(&a, in) : blst_p2_deserialize(&a, in);
    memset(out, 0, 192);
    if (err) return err;
    if (group_check && !blst_p2_affine_in_g2(&a)) return 3;
    memcpy(out, &a, 192);
    return 0;
}

void ref_g1_compress(const uint8_t in[96], uint8_t out48[48], uint8_t out96[96]) {
    blst_p1_affine_compress(out48, (const blst_p1_affine *)in);
    blst_p1_affine_serialize(out96, (const blst_p1_affine *)in);
}
void ref_g2_compress(const uint8_t in[192], uint8_t out96[96], uint8_t out192[192]) {
    blst_p2_affine_compress(out96, (const blst_p2_affine *)in);
    blst_p2_affine_serialize(out192, (const blst_p2_affine *)in);
}

/* sign an arbitrary-length message under the caller's DST with the key of `seed` (coreSign, blst_min_pubkey_sig_core.nim:250-262) */
void ref_sign(uint64_t seed, const uint8_t *msg, size_t len, const uint8_t *dst, size_t dst_len, uint8_t pk_out[96],
              uint8_t sig_out[192]) {
    blst_scalar sk; blst_p1 pk; blst_p2 h;
    keygen_seed(&sk, seed);
    blst_sk_to_pk_in_g1(&pk, &sk);
    blst_p1_to_affine((blst_p1_affine *)pk_out, &pk);
    blst_hash_to_g2(&h, msg, len, dst, dst_len, NULL, 0);
    blst_sign_pk_in_g1(&h, &h, &sk);
    blst_p2_to_affine((blst_p2_affine *)sig_out, &h);
}
