// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the device arithmetic headers as plain C++ (no __CUDACC__) so that the tower / curve /
// hash / pairing logic can be checked against the oracle on a machine without a GPU.  The PTX
// bodies of fp.cuh are replaced by the portable branch there; the PTX itself is checked on the GPU
// by tests/test_gpu_*.py.  Never linked into libblsgpu.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../../nim_blscurve_b200/csrc/h2c.cuh"
#include "../../nim_blscurve_b200/csrc/pairing.cuh"
using namespace bls;

extern "C" {
void hs_fp_mul(const fp *a, const fp *b, fp *r) { fp_mul(*r, *a, *b); }
void hs_fp_add(const fp *a, const fp *b, fp *r) { fp_add(*r, *a, *b); }
void hs_fp_sub(const fp *a, const fp *b, fp *r) { fp_sub(*r, *a, *b); }
void hs_fp_inv(const fp *a, fp *r) { fp_inv(*r, *a); }
int hs_fp2_rsqrt(const fp2 *a, fp2 *r) { return fp2_rsqrt_or_z(*r, *a) ? 1 : 0; }
void hs_hash_to_g2(const uint8_t *msg, size_t len, const uint8_t *dst, uint32_t dst_len, g2_aff *aff, uint8_t *comp) {
    g2_jac j;
    hash_to_g2_jac(j, msg, len, dst, dst_len);
    pt_to_affine(*aff, j);
    g2_compress(comp, *aff);
}
void hs_g1_mul_u64(const g1_aff *p, uint64_t k, g1_aff *r) { g1_jac j; pt_mul_u64(j, *p, k); pt_to_affine(*r, j); }
void hs_g2_mul_u64(const g2_aff *p, uint64_t k, g2_aff *r) { g2_jac j; pt_mul_u64(j, *p, k); pt_to_affine(*r, j); }

void hs_fp12_mul(const fp12 *a, const fp12 *b, fp12 *r) { fp12_mul(*r, *a, *b); }
void hs_fp12_sqr(const fp12 *a, fp12 *r) { fp12_sqr(*r, *a); }
void hs_fp12_cyc_sqr(const fp12 *a, fp12 *r) { fp12_cyc_sqr(*r, *a); }
void hs_fp12_inv(const fp12 *a, fp12 *r) { fp12_inv(*r, *a); }
void hs_fp12_frob(const fp12 *a, int n, fp12 *r) { fp12_frob(*r, *a, n); }
void hs_fp12_bytes(const fp12 *a, uint8_t *out) { fp12_to_bytes(out, *a); }
void hs_fp12_mul_by_line(fp12 *f, const fp2 *l) { fp12_mul_by_line(*f, l[0], l[1], l[2]); }
void hs_miller_loop_n(const g2_aff *Q, const g1_aff *P, int n, fp12 *f) {
    g2_jac T[16]; fp npx[16];
    miller_loop_n(*f, Q, P, n, T, npx);
}
void hs_final_exp(const fp12 *f, fp12 *r) { final_exp(*r, *f); }

struct h2c_trace { fp2 u0, u1; g2_jac q0, q1, sum, iso, out; g2_aff aff; g2_jac alt; g2_aff alt_aff; };
int hs_h2c_trace(const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len, h2c_trace *t) {
    hash_to_field_fp2x2(t->u0, t->u1, msg, msg_len, dst, dst_len);
    sswu_g2(t->q0, t->u0);
    sswu_g2(t->q1, t->u1);
    pt_add(t->sum, t->q0, t->q1, &SSWU_A);
    iso3_g2(t->iso, t->sum);
    g2_clear_cofactor(t->out, t->iso);
    pt_to_affine(t->aff, t->out);
    hash_to_g2_jac(t->alt, msg, msg_len, dst, dst_len);
    pt_to_affine(t->alt_aff, t->alt);
    return (int)sizeof(h2c_trace);
}
}
