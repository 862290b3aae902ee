/*
 * c_abi_multi.c — plain-C consumer of include/blsgpu.h: ONE blsgpu_batch_verify call over every visible GPU.
 *
 * What a Nim/C caller of the drop-in does: create a multi-device context, hand it a host array of
 * SignatureSets, get the reference's boolean back.  The multi-device result must equal the one-device result bit for
 * bit — verdict and the 576 GT bytes — for a valid batch and for a batch with one corrupted message, because the RLC
 * scalars come from the global (n, chunks) derivation (blst_min_pubkey_sig_core.nim:476-505) whatever the sharding.
 *
 * usage: c_abi_multi [n_sets=32768] [shares=0 (0 = one per visible GPU, at least 2: a device is then listed twice)]
 * exit code 0 = every check passed.  Compiled by __graft_entry__.build(); run by tests/test_gpu_abi_multi.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/blsgpu.h"

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) {                                     \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);                  \
            fprintf(stderr, "\n");                         \
            return 1;                                      \
        }                                                  \
    } while (0)

int main(int argc, char **argv) {
    size_t n = argc > 1 ? (size_t)atoll(argv[1]) : 32768;
    int shares = argc > 2 ? atoi(argv[2]) : 0;
    int ndev = blsgpu_device_count();
    CHECK(ndev > 0, "no CUDA device: libblsgpu has no CPU path");
    if (shares <= 0) shares = ndev > 1 ? ndev : 2;
    if (shares > 64) shares = 64;
    int devices[64];
    for (int k = 0; k < shares; k++) devices[k] = k % ndev;

    uint8_t srb[32];
    for (int i = 0; i < 32; i++) srb[i] = (uint8_t)(0xA5 ^ (7 * i));
    const uint32_t chunks = 16;                              /* tp.numThreads of the reference caller */

    blsgpu_ctx *one = blsgpu_create(0, n);
    CHECK(one, "blsgpu_create: %s", blsgpu_last_error(NULL));
    blsgpu_ctx *multi = blsgpu_create_multi(devices, shares, n);
    CHECK(multi, "blsgpu_create_multi: %s", blsgpu_last_error(NULL));
    CHECK(blsgpu_device_span(multi) == shares && blsgpu_device_span(one) == 1, "device span");
    CHECK(blsgpu_capacity(multi) >= n, "capacity %zu < %zu", blsgpu_capacity(multi), n);

    uint8_t *sets = (uint8_t *)malloc(n * BLSGPU_SET_BYTES);  /* pageable host memory, like a Nim seq */
    CHECK(sets, "malloc");
    CHECK(blsgpu_make_sets(one, 2026, 0, n, sets, 0) == 0, "make_sets: %s", blsgpu_last_error(one));

    uint8_t gt1[576], gtm[576];
    int r1 = blsgpu_batch_verify(one, sets, n, srb, chunks, NULL, gt1);
    CHECK(r1 == 1, "one-device verdict on a valid batch: %d (%s)", r1, blsgpu_last_error(one));
    int rm = blsgpu_batch_verify(multi, sets, n, srb, chunks, NULL, gtm);    /* first call: lazy allocations, programs */
    CHECK(rm == 1, "multi-device verdict on a valid batch: %d (%s)", rm, blsgpu_last_error(multi));
    double t0 = now_ms();
    for (int rep = 0; rep < 3; rep++) rm = blsgpu_batch_verify(multi, sets, n, srb, chunks, NULL, gtm);
    double t_multi = (now_ms() - t0) / 3;
    t0 = now_ms();
    for (int rep = 0; rep < 3; rep++) r1 = blsgpu_batch_verify(one, sets, n, srb, chunks, NULL, gt1);
    double t_one = (now_ms() - t0) / 3;
    CHECK(rm == 1 && r1 == 1, "repeated calls: %d / %d", r1, rm);
    CHECK(memcmp(gt1, gtm, 576) == 0, "GT of the valid batch differs between 1 and %d shares", shares);

    /* one corrupted message in the last share: both must reject with the same GT bytes */
    sets[(n - 1) * BLSGPU_SET_BYTES + 100] ^= 0x01;
    r1 = blsgpu_batch_verify(one, sets, n, srb, chunks, NULL, gt1);
    rm = blsgpu_batch_verify(multi, sets, n, srb, chunks, NULL, gtm);
    CHECK(r1 == 0 && rm == 0, "corrupted batch: verdicts %d / %d", r1, rm);
    CHECK(memcmp(gt1, gtm, 576) == 0, "GT of the corrupted batch differs between 1 and %d shares", shares);
    int nonzero = 0;
    for (int i = 0; i < 576; i++) nonzero |= gtm[i];
    CHECK(nonzero, "GT of a failing batch must be reported");
    sets[(n - 1) * BLSGPU_SET_BYTES + 100] ^= 0x01;

    /* serial derivation (chunks = 0) and a batch smaller than the number of shares */
    size_t small = (size_t)shares - 1;
    r1 = blsgpu_batch_verify(one, sets, small, srb, 0, NULL, gt1);
    rm = blsgpu_batch_verify(multi, sets, small, srb, 0, NULL, gtm);
    CHECK(r1 == 1 && rm == 1 && memcmp(gt1, gtm, 576) == 0, "batch of %zu sets over %d shares: %d / %d", small, shares, r1, rm);
    /* empty batch -> false (bls_batch_verifier.nim:312-314); an infinite public key -> false (aggregate.c:296) */
    CHECK(blsgpu_batch_verify(multi, sets, 0, srb, chunks, NULL, NULL) == 0, "empty batch must be false");
    uint8_t keep[96];
    memcpy(keep, sets + (n / 2) * BLSGPU_SET_BYTES, 96);
    memset(sets + (n / 2) * BLSGPU_SET_BYTES, 0, 96);
    rm = blsgpu_batch_verify(multi, sets, n, srb, chunks, NULL, gtm);
    r1 = blsgpu_batch_verify(one, sets, n, srb, chunks, NULL, gt1);
    CHECK(rm == 0 && r1 == 0, "infinite public key: verdicts %d / %d", r1, rm);
    memcpy(sets + (n / 2) * BLSGPU_SET_BYTES, keep, 96);
    /* the flag of that batch must not leak into the next call on the same contexts */
    rm = blsgpu_batch_verify(multi, sets, n, srb, chunks, NULL, NULL);
    CHECK(rm == 1, "valid batch after a flagged one: %d", rm);
    /* over capacity: a distinct negative code, never `false` */
    CHECK(blsgpu_batch_verify(multi, sets, blsgpu_capacity(multi) + 1, srb, chunks, NULL, NULL) == BLSGPU_ERR_CAPACITY,
          "capacity error code");

    /* sharded G1 MSM: same affine point as one device */
    size_t nm = n < 4096 ? n : 4096;
    uint8_t *pts = (uint8_t *)malloc(nm * 96), *sc = (uint8_t *)malloc(nm * 32);
    CHECK(pts && sc, "malloc");
    for (size_t i = 0; i < nm; i++) {
        memcpy(pts + 96 * i, sets + i * BLSGPU_SET_BYTES, 96);
        for (int b = 0; b < 32; b++) sc[32 * i + b] = (uint8_t)(i * 131 + b * 17 + 3);
        sc[32 * i + 31] &= 0x7f;
    }
    uint8_t m1[96], mm[96];
    CHECK(blsgpu_msm_g1(one, pts, sc, nm, 255, m1) == 1, "msm one: %s", blsgpu_last_error(one));
    CHECK(blsgpu_msm_g1(multi, pts, sc, nm, 255, mm) == 1, "msm multi: %s", blsgpu_last_error(multi));
    CHECK(memcmp(m1, mm, 96) == 0, "sharded MSM differs from the one-device MSM");

    printf("c_abi_multi OK: %zu sets, %d shares over %d device(s); blsgpu_batch_verify from a pageable host buffer: %.2f ms on the "
           "multi-device context, %.2f ms on one device; valid + corrupted GT identical to the one-device result; sharded MSM "
           "identical\n", n, shares, ndev, t_multi, t_one);
    free(pts);
    free(sc);
    free(sets);
    blsgpu_destroy(multi);
    blsgpu_destroy(one);
    return 0;
}
