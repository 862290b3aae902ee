// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the device arithmetic headers as plain C++ (no __CUDACC__) so that the tower / curve /
// hash / pairing logic can be checked against the oracle on a machine without a GPU.  The PTX
// bodies of fp.cuh are replaced by the portable branch there; the PTX itself is checked on the GPU
// by tests/test_gpu_*.py.  Never linked into libblsgpu.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../../nim_blscurve_b200/csrc/h2c.cuh"
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
}
