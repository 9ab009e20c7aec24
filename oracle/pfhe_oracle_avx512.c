/*
 * pfhe_oracle_avx512.c -- AVX-512 IFMA restatement of the reference's fast CPU path for the headline transform
 * (forward negacyclic NTT, u64 words, q < 2^50, "BIT_SHIFT = 52").
 *
 * TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE: used only by bench.py's cpu_baseline and --impl reference legs (and checked
 * against the scalar oracle in tests/test_oracle.py).  On a CPU with AVX-512 IFMA the reference does not run its scalar
 * transform: U64NttTable picks the Intel-HEXL-style back-end (primus_ntt/src/ntt/prime64/table.rs:166-231), i.e.
 *   forward_transform_to_bit_reverse_avx512::<52>   primus_ntt/src/ntt/prime64/avx512/transform.rs:14-203
 *   stage kernels T8+ / T4 / T2 / T1                 primus_ntt/src/ntt/prime64/avx512/stages.rs:8-319
 *   Harvey butterfly on 8 lanes                      primus_ntt/src/ntt/prime64/avx512/butterfly.rs:11-60
 *   52-bit multiply helpers                          primus_ntt/src/ntt/prime64/avx512/utils/arithmetic.rs:69-73,145-163,182-184
 *   MultiplyFactor with shift 52                     primus_factor/src/mul_factor/mod.rs:4-88
 * so a scalar port under-states the reference on such hosts by the factor this file measures.
 * Same radix-2 Cooley-Tukey order, same twiddles (roots[m + i], bit-reversed table), same lazy ranges ([0,4q) between stages), canonical
 * output; the T4 / T2 / T1 stages regroup lanes with two-source permutes instead of the reference's pre-expanded root layouts
 * (same arithmetic, different data movement).  Built only when the compiler accepts -mavx512ifma; used only when the CPU reports it.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#if defined(__AVX512F__) && defined(__AVX512IFMA__) && defined(__AVX512DQ__)
#include <immintrin.h>
#define PFHE_HAVE_IFMA 1
#else
#define PFHE_HAVE_IFMA 0
#endif

typedef struct {
    size_t n;
    uint64_t q;
    uint64_t *w, *wp;     /* roots[k] (bit-reversed table) and floor(roots[k] 2^52 / q) */
    uint64_t *w4, *wp4;   /* stage t = 4: each block's twiddle repeated 4x            */
    uint64_t *w2, *wp2;   /* stage t = 2: each block's twiddle repeated 2x            */
} o_ifma_ntt;

int o_ifma_supported(void)
{
#if PFHE_HAVE_IFMA
    return __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512f");
#else
    return 0;
#endif
}

void o_ifma_destroy(o_ifma_ntt *t)
{
    if (!t) return;
    free(t->w); free(t->wp); free(t->w4); free(t->wp4); free(t->w2); free(t->wp2);
    free(t);
}

/* roots: the scalar table's roots[] (roots[bitrev(k)] = psi^k), n entries; q < 2^50, n >= 16 */
o_ifma_ntt *o_ifma_create(unsigned log_n, uint64_t q, const uint64_t *roots)
{
    const size_t n = (size_t)1 << log_n;
    if (n < 16 || (q >> 50) != 0) return NULL;
    o_ifma_ntt *t = calloc(1, sizeof(*t));
    if (!t) return NULL;
    t->n = n; t->q = q;
    t->w = aligned_alloc(64, n * 8); t->wp = aligned_alloc(64, n * 8);
    t->w4 = aligned_alloc(64, n / 2 * 8); t->wp4 = aligned_alloc(64, n / 2 * 8);
    t->w2 = aligned_alloc(64, n / 2 * 8); t->wp2 = aligned_alloc(64, n / 2 * 8);
    if (!t->w || !t->wp || !t->w4 || !t->wp4 || !t->w2 || !t->wp2) { o_ifma_destroy(t); return NULL; }
    for (size_t k = 0; k < n; k++) {
        t->w[k] = roots[k];
        t->wp[k] = (uint64_t)((((unsigned __int128)roots[k]) << 52) / q);   /* MultiplyFactor::new(w, 52, q) */
    }
    for (size_t b = 0; b < n / 8; b++)       /* stage t = 4: m = n/8 blocks, twiddle roots[n/8 + b] */
        for (int r = 0; r < 4; r++) { t->w4[4 * b + r] = t->w[n / 8 + b]; t->wp4[4 * b + r] = t->wp[n / 8 + b]; }
    for (size_t b = 0; b < n / 4; b++)       /* stage t = 2: m = n/4 blocks */
        for (int r = 0; r < 2; r++) { t->w2[2 * b + r] = t->w[n / 4 + b]; t->wp2[2 * b + r] = t->wp[n / 4 + b]; }
    return t;
}

#if PFHE_HAVE_IFMA
static inline __m512i small_mod(__m512i x, __m512i m) { return _mm512_min_epu64(x, _mm512_sub_epi64(x, m)); }

/* Harvey forward butterfly, BIT_SHIFT = 52 (butterfly.rs:11-60): X, Y in [0,4q) -> [0,4q) */
static inline void fwd_bfly52(__m512i *x, __m512i *y, __m512i w, __m512i wp, __m512i neg_q, __m512i two_q, __m512i mask52)
{
    const __m512i zero = _mm512_setzero_si512();
    *x = small_mod(*x, two_q);
    const __m512i qv = _mm512_madd52hi_epu64(zero, wp, *y);
    const __m512i wy = _mm512_madd52lo_epu64(zero, w, *y);
    const __m512i t = _mm512_and_si512(_mm512_madd52lo_epu64(wy, qv, neg_q), mask52);
    *y = _mm512_add_epi64(*x, _mm512_sub_epi64(two_q, t));
    *x = _mm512_add_epi64(*x, t);
}

void o_ifma_forward(const o_ifma_ntt *tb, uint64_t *v)
{
    const size_t n = tb->n;
    const __m512i q = _mm512_set1_epi64((long long)tb->q), two_q = _mm512_set1_epi64((long long)(2 * tb->q));
    const __m512i neg_q = _mm512_set1_epi64((long long)(0 - tb->q)), mask52 = _mm512_set1_epi64((long long)((1ull << 52) - 1));
    size_t ri = 1;
    size_t gap = n >> 1, m = 1;
    for (; gap >= 8; gap >>= 1, m <<= 1) {                     /* stages with t >= 8 (stages.rs: fwd_t8) */
        for (size_t blk = 0; blk < m; blk++, ri++) {
            const __m512i w = _mm512_set1_epi64((long long)tb->w[ri]), wp = _mm512_set1_epi64((long long)tb->wp[ri]);
            uint64_t *x = v + blk * 2 * gap, *y = x + gap;
            for (size_t j = 0; j < gap; j += 8) {
                __m512i vx = _mm512_loadu_si512(x + j), vy = _mm512_loadu_si512(y + j);
                fwd_bfly52(&vx, &vy, w, wp, neg_q, two_q, mask52);
                _mm512_storeu_si512(x + j, vx);
                _mm512_storeu_si512(y + j, vy);
            }
        }
    }
    {   /* t = 4 (fwd_t4): two blocks of 8 words per iteration */
        const __m512i ix = _mm512_set_epi64(11, 10, 9, 8, 3, 2, 1, 0), iy = _mm512_set_epi64(15, 14, 13, 12, 7, 6, 5, 4);
        for (size_t i = 0; i < n; i += 16) {
            const __m512i a = _mm512_loadu_si512(v + i), b = _mm512_loadu_si512(v + i + 8);
            __m512i vx = _mm512_permutex2var_epi64(a, ix, b), vy = _mm512_permutex2var_epi64(a, iy, b);
            const __m512i w = _mm512_loadu_si512(tb->w4 + i / 2), wp = _mm512_loadu_si512(tb->wp4 + i / 2);
            fwd_bfly52(&vx, &vy, w, wp, neg_q, two_q, mask52);
            _mm512_storeu_si512(v + i, _mm512_permutex2var_epi64(vx, ix, vy));
            _mm512_storeu_si512(v + i + 8, _mm512_permutex2var_epi64(vx, iy, vy));
        }
    }
    {   /* t = 2 (fwd_t2): four blocks of 4 words per iteration */
        const __m512i ix = _mm512_set_epi64(13, 12, 9, 8, 5, 4, 1, 0), iy = _mm512_set_epi64(15, 14, 11, 10, 7, 6, 3, 2);
        const __m512i o0 = _mm512_set_epi64(11, 10, 3, 2, 9, 8, 1, 0), o1 = _mm512_set_epi64(15, 14, 7, 6, 13, 12, 5, 4);
        for (size_t i = 0; i < n; i += 16) {
            const __m512i a = _mm512_loadu_si512(v + i), b = _mm512_loadu_si512(v + i + 8);
            __m512i vx = _mm512_permutex2var_epi64(a, ix, b), vy = _mm512_permutex2var_epi64(a, iy, b);
            const __m512i w = _mm512_loadu_si512(tb->w2 + i / 2), wp = _mm512_loadu_si512(tb->wp2 + i / 2);
            fwd_bfly52(&vx, &vy, w, wp, neg_q, two_q, mask52);
            _mm512_storeu_si512(v + i, _mm512_permutex2var_epi64(vx, o0, vy));
            _mm512_storeu_si512(v + i + 8, _mm512_permutex2var_epi64(vx, o1, vy));
        }
    }
    {   /* t = 1 (fwd_t1): eight pairs per iteration, twiddles roots[n/2 + i/2 ...]; canonical output (transform.rs:104-124) */
        const __m512i ix = _mm512_set_epi64(14, 12, 10, 8, 6, 4, 2, 0), iy = _mm512_set_epi64(15, 13, 11, 9, 7, 5, 3, 1);
        const __m512i o0 = _mm512_set_epi64(11, 3, 10, 2, 9, 1, 8, 0), o1 = _mm512_set_epi64(15, 7, 14, 6, 13, 5, 12, 4);
        for (size_t i = 0; i < n; i += 16) {
            const __m512i a = _mm512_loadu_si512(v + i), b = _mm512_loadu_si512(v + i + 8);
            __m512i vx = _mm512_permutex2var_epi64(a, ix, b), vy = _mm512_permutex2var_epi64(a, iy, b);
            const __m512i w = _mm512_loadu_si512(tb->w + n / 2 + i / 2), wp = _mm512_loadu_si512(tb->wp + n / 2 + i / 2);
            fwd_bfly52(&vx, &vy, w, wp, neg_q, two_q, mask52);
            vx = small_mod(small_mod(vx, two_q), q);
            vy = small_mod(small_mod(vy, two_q), q);
            _mm512_storeu_si512(v + i, _mm512_permutex2var_epi64(vx, o0, vy));
            _mm512_storeu_si512(v + i + 8, _mm512_permutex2var_epi64(vx, o1, vy));
        }
    }
    (void)ri; (void)m;
}
#else
void o_ifma_forward(const o_ifma_ntt *tb, uint64_t *v) { (void)tb; (void)v; abort(); }
#endif

void o_ifma_forward_batch(const o_ifma_ntt *t, uint64_t *v, size_t batch, int threads)
{
    long long b;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (b = 0; b < (long long)batch; b++) o_ifma_forward(t, v + (size_t)b * t->n);
}
