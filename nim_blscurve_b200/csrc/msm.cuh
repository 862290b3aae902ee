// msm.cuh — G1 multi-scalar multiplication: signed-digit Pippenger with a sorted bucket scatter.
// Restates the result of blst_p1s_mult_pippenger (vendor/blst/src/multi_scalar.c:415-434; window loop
// :370-397, tile :332-368, bucket integration :295-311) — the affine sum  sum_i [k_i] P_i  is canonical,
// so the GPU decomposition is free to differ:
//   1. k_msm_count    signed window digits (Booth-style carry, digits in (-2^(c-1), 2^(c-1)]), histogram
//                     of (window, |digit|) with global atomics
//   2. k_msm_scan     exclusive prefix sum of the histogram (one block)
//   3. k_msm_scatter  counting-sort scatter of (point index, sign) into bucket order
//   4. k_msm_bucket   one thread per (window, bucket): gathers its points, mixed Jacobian additions
//   5. k_msm_segment  running-sum integration of 32-bucket segments, weighted by the segment base
//   6. k_g1_tree_rows per-window tree sum of the segment results
//   7. k_msm_horner   Horner over windows (c doublings each), to affine
// Scalars are NOT reduced mod r (as in the reference); nbits low bits of each little-endian scalar are used.
#pragma once
#include <string>
#include "kernels.cuh"

namespace bls {

struct msm_state {
    uint8_t *buf = nullptr;
    size_t bytes = 0;
};

static inline void msm_free(msm_state &m) {
    if (m.buf) cudaFree(m.buf);
    m.buf = nullptr;
    m.bytes = 0;
}

#define MSM_SEG 32

// signed digit of window w (width c) of the nbits-bit little-endian scalar at sc (sb bytes)
__device__ __forceinline__ int msm_digit(const uint8_t *sc, int sb, int nbits, int c, int w, int &carry) {
    int bit = w * c;
    uint32_t raw = 0;
    // gather c bits starting at `bit`, masking everything at or above nbits
    for (int k = 0; k < c; k += 8) {
        int b0 = bit + k;
        if (b0 >= nbits) break;
        int byte = b0 >> 3, sh = b0 & 7;
        uint32_t v = sc[byte];
        if (byte + 1 < sb) v |= (uint32_t)sc[byte + 1] << 8;
        v >>= sh;
        int take = c - k < 8 ? c - k : 8;
        if (b0 + take > nbits) take = nbits - b0;
        v &= (1u << take) - 1;
        raw |= v << k;
    }
    int d = (int)raw + carry;
    if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
    return d;
}

__global__ void __launch_bounds__(256) k_msm_count(const uint8_t *scalars, size_t n, int sb, int nbits, int c, int nwin,
                                                   uint32_t *counts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *sc = scalars + i * sb;
    int carry = 0;
    const uint32_t B = 1u << (c - 1);
    for (int w = 0; w < nwin; w++) {
        int d = msm_digit(sc, sb, nbits, c, w, carry);
        if (d != 0) atomicAdd(&counts[(size_t)w * B + (uint32_t)(d < 0 ? -d : d) - 1], 1u);
    }
}

// exclusive scan of m counters into offsets[0..m] (offsets[m] = total); single block of 1024 threads
__global__ void __launch_bounds__(1024) k_msm_scan(const uint32_t *counts, size_t m, uint32_t *offsets, uint32_t *cursor) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (size_t base = 0; base < m; base += 1024) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < m ? counts[i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sums[lane], t = s;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_sums[lane] = t - s;      // exclusive warp offsets
        }
        __syncthreads();
        uint32_t excl = carry_s + warp_sums[wid] + x - v;
        if (i < m) { offsets[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[m] = carry_s;
}

__global__ void __launch_bounds__(256) k_msm_scatter(const uint8_t *scalars, size_t n, int sb, int nbits, int c, int nwin,
                                                     uint32_t *cursor, uint32_t *entries) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *sc = scalars + i * sb;
    int carry = 0;
    const uint32_t B = 1u << (c - 1);
    for (int w = 0; w < nwin; w++) {
        int d = msm_digit(sc, sb, nbits, c, w, carry);
        if (d != 0) {
            uint32_t pos = atomicAdd(&cursor[(size_t)w * B + (uint32_t)(d < 0 ? -d : d) - 1], 1u);
            entries[pos] = ((uint32_t)i << 1) | (d < 0 ? 1u : 0u);
        }
    }
}

__global__ void BLS_LB k_msm_bucket(const g1_aff *points, const uint32_t *offsets, const uint32_t *entries,
                                                    size_t nbuckets, g1_jac *buckets) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nbuckets) return;
    uint32_t lo = offsets[g], hi = offsets[g + 1];
    g1_jac acc;
    pt_set_inf(acc);
    for (uint32_t e = lo; e < hi; e++) {
        uint32_t ent = entries[e];
        g1_aff p = points[ent >> 1];
        if (ent & 1) fp_neg(p.y, p.y);
        pt_add_affine(acc, acc, p);
    }
    buckets[g] = acc;
}

// segment j of window w covers digit magnitudes b in [j*SEG+1, j*SEG+SEG] (bucket index b-1):
//   T = sum_b (b - j*SEG) B_b  +  [j*SEG] sum_b B_b
__global__ void BLS_LB k_msm_segment(const g1_jac *buckets, int c, int nwin, g1_jac *segs) {
    const uint32_t B = 1u << (c - 1);
    const uint32_t L = B < MSM_SEG ? B : MSM_SEG;
    const uint32_t nseg = B / L;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nwin * nseg) return;
    uint32_t w = (uint32_t)(t / nseg), j = (uint32_t)(t % nseg);
    const g1_jac *bk = buckets + (size_t)w * B + (size_t)j * L;
    g1_jac running, acc;
    pt_set_inf(running);
    pt_set_inf(acc);
    for (int b = (int)L - 1; b >= 0; b--) {
        g1_jac x = bk[b];
        pt_add(running, running, x);
        pt_add(acc, acc, running);
    }
    uint32_t k = j * L;
    if (k) {
        g1_jac m;
        pt_mul_words(m, running, &k, 1);
        pt_add(acc, acc, m);
    }
    segs[t] = acc;
}

// row-wise pairwise tree step over `rows` rows of `stride` entries
__global__ void BLS_LB k_g1_tree_rows(g1_jac *S, int rows, size_t stride, size_t m, size_t half) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * half) return;
    size_t r = t / half, i = t % half;
    if (i + half >= m) return;
    g1_jac a = S[r * stride + i], b = S[r * stride + i + half];
    pt_add(a, a, b);
    S[r * stride + i] = a;
}

__global__ void k_msm_horner(const g1_jac *W, size_t stride, int nwin, int c, g1_aff *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    g1_jac acc = W[(size_t)(nwin - 1) * stride];
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) pt_dbl(acc, acc);
        g1_jac x = W[(size_t)w * stride];
        pt_add(acc, acc, x);
    }
    g1_aff a;
    pt_to_affine(a, acc);
    *out = a;
}

// synthetic MSM inputs (benchmark only): P_i = [k_i]G1 with a 96-bit k_i, 255-bit coefficients
// (shape of benchmarks/bls12381_msm_g1.nim:22-44)
__global__ void BLS_LB k_msm_make_inputs(uint64_t seed, size_t n, g1_aff *points, uint8_t *scalars) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed ^ (0xFACADEull + i * 0x100000001b3ull);
    uint64_t v[6];
    for (int k = 0; k < 6; k++) {
        s += 0x9e3779b97f4a7c15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        v[k] = z ^ (z >> 31);
    }
    uint32_t kw[3] = {(uint32_t)v[0], (uint32_t)(v[0] >> 32), (uint32_t)v[1]};
    g1_jac g, r;
    g.x = G1_GEN_X; g.y = G1_GEN_Y; g.z = FP_ONE;
    pt_mul_words(r, g, kw, 3);
    g1_aff a;
    pt_to_affine(a, r);
    points[i] = a;
    for (int w = 0; w < 4; w++)
        for (int b = 0; b < 8; b++) scalars[32 * i + 8 * w + b] = (uint8_t)(v[2 + w] >> (8 * b));
    scalars[32 * i + 31] &= 0x7f;
}

static inline int msm_window_bits(size_t n) {
    int lg = 0;
    while ((n >> (lg + 1)) != 0) lg++;
    int c = lg - 4;
    if (c < 2) c = 2;
    if (c > 16) c = 16;
    return c;
}

// d_points / d_scalars are device pointers; the affine result lands in h_out (pinned, 96 bytes)
static inline int msm_g1_run(msm_state &st, const g1_aff *d_points, const uint8_t *d_scalars, size_t n, int nbits,
                             cudaStream_t s, uint8_t *h_out, std::string &err) {
    if (n >= ((size_t)1 << 31)) { err = "msm: too many points"; return -2; }
    const int sb = (nbits + 7) / 8;
    const int c = msm_window_bits(n);
    const int nwin = (nbits + 1 + c - 1) / c;
    const size_t B = (size_t)1 << (c - 1);
    const size_t nb = (size_t)nwin * B;
    const size_t L = B < MSM_SEG ? B : MSM_SEG;
    const size_t nseg = B / L;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_counts = 0;
    size_t o_offsets = o_counts + al(nb * 4);
    size_t o_cursor = o_offsets + al((nb + 1) * 4);
    size_t o_entries = o_cursor + al(nb * 4);
    size_t o_buckets = o_entries + al(n * (size_t)nwin * 4);
    size_t o_segs = o_buckets + al(nb * sizeof(g1_jac));
    size_t o_out = o_segs + al((size_t)nwin * nseg * sizeof(g1_jac));
    size_t total = o_out + 256;
    cudaError_t e;
#define MCK(call) if ((e = (call)) != cudaSuccess) { err = std::string(#call ": ") + cudaGetErrorString(e); return -1; }
    if (st.bytes < total) {
        if (st.buf) cudaFree(st.buf);
        st.buf = nullptr;
        st.bytes = 0;
        MCK(cudaMalloc((void **)&st.buf, total));
        st.bytes = total;
    }
    uint32_t *counts = (uint32_t *)(st.buf + o_counts), *offsets = (uint32_t *)(st.buf + o_offsets);
    uint32_t *cursor = (uint32_t *)(st.buf + o_cursor), *entries = (uint32_t *)(st.buf + o_entries);
    g1_jac *buckets = (g1_jac *)(st.buf + o_buckets), *segs = (g1_jac *)(st.buf + o_segs);
    g1_aff *d_out = (g1_aff *)(st.buf + o_out);
    MCK(cudaMemsetAsync(counts, 0, nb * 4, s));
    k_msm_count<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_scalars, n, sb, nbits, c, nwin, counts);
    k_msm_scan<<<1, 1024, 0, s>>>(counts, nb, offsets, cursor);
    k_msm_scatter<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_scalars, n, sb, nbits, c, nwin, cursor, entries);
    k_msm_bucket<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(d_points, offsets, entries, nb, buckets);
    size_t nt = (size_t)nwin * nseg;
    k_msm_segment<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(buckets, c, nwin, segs);
    for (size_t m = nseg; m > 1;) {
        size_t half = (m + 1) / 2;
        k_g1_tree_rows<<<(unsigned)(((size_t)nwin * half + 127) / 128), 128, 0, s>>>(segs, nwin, nseg, m, half);
        m = half;
    }
    k_msm_horner<<<1, 32, 0, s>>>(segs, nseg, nwin, c, d_out);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(h_out, d_out, 96, cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
#undef MCK
    return 0;
}

}  // namespace bls
