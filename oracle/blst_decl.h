/*
 * oracle/blst_decl.h — TEST INFRASTRUCTURE ONLY.
 * Minimal prototypes for the BLST entry points the oracle driver calls, written out here so that
 * oracle/ref_batch.c compiles without the reference tree on the include path (the GPU box only
 * has the prebuilt oracle/_ref/libblst_ref.so).  The authoritative declarations are
 * /root/reference/vendor/blst/bindings/blst.h (types :59-73, :169-170, :196-197; functions cited
 * per line) and blst_aux.h:85-88, :116.
 */
#ifndef ORACLE_BLST_DECL_H
#define ORACLE_BLST_DECL_H
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>

typedef struct { uint8_t b[32]; } blst_scalar;               /* blst.h:61 */
typedef struct { uint64_t l[6]; } blst_fp;                   /* blst.h:63 */
typedef struct { blst_fp fp[2]; } blst_fp2;                  /* blst.h:65 */
typedef struct { blst_fp2 fp2[3]; } blst_fp6;
typedef struct { blst_fp6 fp6[2]; } blst_fp12;
typedef struct { blst_fp x, y, z; } blst_p1;                 /* blst.h:169 */
typedef struct { blst_fp x, y; } blst_p1_affine;             /* blst.h:170 */
typedef struct { blst_fp2 x, y, z; } blst_p2;
typedef struct { blst_fp2 x, y; } blst_p2_affine;
typedef struct blst_pairing_st blst_pairing;                 /* opaque, blst_pairing_sizeof() bytes */

void blst_sha256(uint8_t out[32], const uint8_t *msg, size_t msg_len);                 /* blst_aux.h:116 */
void blst_keygen(blst_scalar *out_SK, const uint8_t *IKM, size_t IKM_len,
                 const uint8_t *info, size_t info_len);                                /* blst.h:330 */
void blst_sk_to_pk_in_g1(blst_p1 *out_pk, const blst_scalar *SK);                      /* blst.h:332 */
void blst_sign_pk_in_g1(blst_p2 *out_sig, const blst_p2 *hash, const blst_scalar *SK); /* blst.h:333 */
void blst_hash_to_g2(blst_p2 *out, const uint8_t *msg, size_t msg_len, const uint8_t *DST, size_t DST_len,
                     const uint8_t *aug, size_t aug_len);                              /* blst.h:296 */

void blst_p1_to_affine(blst_p1_affine *out, const blst_p1 *in);                        /* blst.h:182 */
void blst_p1_from_affine(blst_p1 *out, const blst_p1_affine *in);
void blst_p1_add_or_double_affine(blst_p1 *out, const blst_p1 *a, const blst_p1_affine *b);
void blst_p1_mult(blst_p1 *out, const blst_p1 *p, const uint8_t *scalar, size_t nbits);
const blst_p1_affine *blst_p1_affine_generator(void);                                  /* blst.h:194 */
void blst_p2_to_affine(blst_p2_affine *out, const blst_p2 *in);                        /* blst.h:209 */
void blst_p2_from_affine(blst_p2 *out, const blst_p2_affine *in);
void blst_p2_add_or_double(blst_p2 *out, const blst_p2 *a, const blst_p2 *b);
void blst_p2_add_or_double_affine(blst_p2 *out, const blst_p2 *a, const blst_p2_affine *b);
void blst_p2_mult(blst_p2 *out, const blst_p2 *p, const uint8_t *scalar, size_t nbits);
void blst_p2_cneg(blst_p2 *p, bool cbit);
void blst_p1_cneg(blst_p1 *p, bool cbit);
void blst_p2_affine_compress(uint8_t out[96], const blst_p2_affine *in);               /* blst.h:314 */
void blst_p1_affine_compress(uint8_t out[48], const blst_p1_affine *in);               /* blst.h:310 */
void blst_p1_affine_serialize(uint8_t out[96], const blst_p1_affine *in);
void blst_p2_affine_serialize(uint8_t out[192], const blst_p2_affine *in);
int blst_p1_uncompress(blst_p1_affine *out, const uint8_t in[48]);                     /* blst.h:311 */
int blst_p1_deserialize(blst_p1_affine *out, const uint8_t in[96]);
int blst_p2_uncompress(blst_p2_affine *out, const uint8_t in[96]);                     /* blst.h:315 */
int blst_p2_deserialize(blst_p2_affine *out, const uint8_t in[192]);
bool blst_p1_affine_is_inf(const blst_p1_affine *a);                                   /* blst.h:192 */
bool blst_p1_affine_in_g1(const blst_p1_affine *p);                                    /* blst.h:191 */
bool blst_p2_affine_in_g2(const blst_p2_affine *p);                                    /* blst.h:218 */
bool blst_p1_affine_on_curve(const blst_p1_affine *p);
bool blst_p2_affine_on_curve(const blst_p2_affine *p);

size_t blst_p1s_mult_pippenger_scratch_sizeof(size_t npoints);                         /* blst.h:242 */
void blst_p1s_mult_pippenger(blst_p1 *ret, const blst_p1_affine *const points[], size_t npoints,
                             const uint8_t *const scalars[], size_t nbits, void *scratch);
size_t blst_p2s_mult_pippenger_scratch_sizeof(size_t npoints);                         /* blst.h:266 */
void blst_p2s_mult_pippenger(blst_p2 *ret, const blst_p2_affine *const points[], size_t npoints,
                             const uint8_t *const scalars[], size_t nbits, void *scratch);

void blst_miller_loop(blst_fp12 *ret, const blst_p2_affine *Q, const blst_p1_affine *P); /* blst.h:343 */
void blst_final_exp(blst_fp12 *ret, const blst_fp12 *f);                               /* blst.h:348 */
void blst_fp12_mul(blst_fp12 *ret, const blst_fp12 *a, const blst_fp12 *b);            /* blst.h:153 */
void blst_fp12_conjugate(blst_fp12 *a);
bool blst_fp12_is_one(const blst_fp12 *a);
const blst_fp12 *blst_fp12_one(void);
void blst_bendian_from_fp12(uint8_t out[576], const blst_fp12 *a);                     /* blst_aux.h:88 */

size_t blst_pairing_sizeof(void);                                                      /* blst.h:363 */
void blst_pairing_init(blst_pairing *ctx, bool hash_or_encode, const uint8_t *DST, size_t DST_len);
void blst_pairing_commit(blst_pairing *ctx);
int blst_pairing_chk_n_mul_n_aggr_pk_in_g1(blst_pairing *ctx, const blst_p1_affine *PK, bool pk_grpchk,
                                           const blst_p2_affine *sig, bool sig_grpchk,
                                           const uint8_t *scalar, size_t nbits,
                                           const uint8_t *msg, size_t msg_len,
                                           const uint8_t *aug, size_t aug_len);        /* blst.h:425 */
int blst_pairing_chk_n_aggr_pk_in_g1(blst_pairing *ctx, const blst_p1_affine *PK, bool pk_grpchk,
                                     const blst_p2_affine *sig, bool sig_grpchk,
                                     const uint8_t *msg, size_t msg_len,
                                     const uint8_t *aug, size_t aug_len);              /* blst.h:417 */
int blst_pairing_merge(blst_pairing *ctx, const blst_pairing *ctx1);                   /* blst.h:436 */
bool blst_pairing_finalverify(const blst_pairing *ctx, const blst_fp12 *gtsig);        /* blst.h:437 */
blst_fp12 *blst_pairing_as_fp12(blst_pairing *ctx);                                    /* blst_aux.h:87 */
#endif
