// lattice32.cu -- blind rotation for the bootstrapping shape of BASELINE config 5 (u32 words, N = 1024, k = 1),
// rebuilt around the instruction-issue / IMAD-pipe limits measured in round 1 (profiles/r01_ncu_blind_rotate_u32.txt:
// 4159 instructions per thread per CMux, of which ~1970 were modular arithmetic).
//
// Same mathematics as blind_rotate_kernel (lattice.cu; external product =
// CrtGlwe::mul_dcrt_ggsw_to, primus_lattice/src/glwe/crt.rs:200-227; rotation = mul_monomial_assign,
// primus_poly/src/poly/mul.rs:74-99; digits = ApproxSignedBasis, primus_decompose/src/primitive/basis.rs:254-283 +
// primitive/common.rs:246-259), every canonical word bit-identical.  What changed is the schedule:
//
//  * one ciphertext = 128 threads x 8 coefficients; a transform is 3 register passes of 3 radix-2 stages
//    (index bits 9..7 | 6..4 | 3..1) plus one stage (bit 0) done with a lane^1 shuffle, so that the only CTA barrier of
//    a transform is the first exchange; the second exchange stays inside 16-thread groups (__syncwarp);
//  * the two digit polynomials of a level (input components r = 0, 1) and the two output components are transformed
//    in lock step: twiddles, addresses and barriers are shared by the pair (5 CTA barriers per CMux instead of ~35);
//  * exchange buffers are padded, not XOR-swizzled: word j*144 + t (first exchange) and 18*(idx >> 4) + (idx & 15)
//    (second exchange) are bank-conflict free for both the storing and the loading pattern and every element address is
//    `thread base + compile-time offset`;
//  * gadget digits without the carry chain: the balanced digits of the reference's sequential rule
//    (t = window + carry; carry = t >= B/2; digit = t - carry*B) are the unique balanced representation, so they equal
//    window_l(v + R) - B/2 with R = 2^(drop-1) + sum_l (B/2) 2^(drop + l*beta); the digit enters the transform as
//    window + (q - B/2) (congruent, < q + B), 3 instructions per digit;
//  * key multiply-accumulate in double-word lazy sums (reduce_dot_product, primus_modulus/src/common/compact/slice.rs:371-401)
//    followed by a Montgomery reduction (3 instructions) instead of the two-word Barrett reduction (12); the factor 2^-32 is
//    cancelled exactly by folding 2^32 into the n^-1 constants of the last inverse stage (table.rs:397-400);
//  * the pass over index bits 9..7 reads its seven twiddles from the kernel-parameter bank (uniform), the others from
//    the table in the reference's index order (prime32/table.rs:222-257) with 128-bit loads.
//
// Preconditions (checked on the host, otherwise the generic kernel in lattice.cu runs): u32 words, N = 1024,
// (2 log2 N + 2) q < 2^32 (forward values never wrap), 2 (2 levels + 1) q < 2^32 (Montgomery outputs feed the first inverse stage).
#include <cstdlib>

#include "host_math.hpp"
#include "internal.hpp"

namespace pfhe {
namespace br32 {

constexpr int LOGN = 10, N = 1 << LOGN, TPP = 128;
constexpr int E1W = 1152;  // words of one first-exchange buffer  (max index 7*144 + 127 = 1135)
constexpr int E2W = 1152;  // words of one second-exchange buffer (max index 18*63 + 15 = 1149)
// shared memory words: E1 [parity 2][poly 2][E1W] | E2 [poly 2][E2W] | ACC [2][N]
constexpr int SMEM_WORDS = 4 * E1W + 2 * E2W + 2 * N;

struct Params {
    uint32_t q, two_q;
    uint32_t qinv;           // q^-1 mod 2^32 (Montgomery reduction of the lazy sums)
    uint32_t one_q;          // floor(2^32 / q): Shoup quotient of 1 (folds any 32-bit value to [0, 2q))
    uint32_t invn_r, invn_r_q;    // n^-1 * 2^32 mod q and its Shoup quotient
    uint32_t invnw_r, invnw_r_q;  // n^-1 * inv_roots[N-1] * 2^32 mod q
    uint32_t redc_bias;      // q: r = hi - mulhi(m, q) + q in (0, (terms + 1) q]
    uint32_t first_inv_bias; // (terms + 1) q >= every Montgomery output
    uint32_t threshold, add_r, r;  // digit preparation: W = v + (v >= threshold ? add + R : R)
    uint32_t mask, digit_off;      // digit = ((W >> shift) & mask) + digit_off
    uint32_t drop_bits, log_basis, levels;
    uint2 fwd_head[8];       // fwd[0..7]: twiddles of stages 0..2 (uniform over the polynomial)
    uint2 inv_tail[8];       // inv[N-8 .. N-1]: twiddles of the last three inverse stages (entry 7 unused here)
    const uint2 *fwd, *inv;  // tables in the reference's index order, (w, floor(w 2^32 / q)) pairs
    const uint32_t *ordinal; // psi^k, k < 2N (monomial transforms, prime32/table.rs; ternary rotation only)
    uint32_t r32, r32_q;     // 2^32 mod q and its Shoup quotient (Montgomery form of the monomial factors)
};

__device__ __forceinline__ uint32_t shoup_lazy32(uint32_t y, uint2 w, uint32_t q) { return y * w.x - __umulhi(y, w.y) * q; }

// forward butterfly without the per-stage fold (values grow by at most 2q per stage, never wrap: host-checked)
__device__ __forceinline__ void bf_fwd(uint32_t &x, uint32_t &y, uint2 w, uint32_t q, uint32_t two_q) {
    const uint32_t t = shoup_lazy32(y, w, q);
    y = x + two_q - t;
    x = x + t;
}
// Harvey inverse butterfly, [0, 2q) -> [0, 2q)
__device__ __forceinline__ void bf_inv(uint32_t &x, uint32_t &y, uint2 w, uint32_t q, uint32_t two_q) {
    const uint32_t tx = x + y, ty = x + two_q - y;
    x = min(tx, tx - two_q);
    y = shoup_lazy32(ty, w, q);
}

// three forward stages on the 8 registers of NP polynomials; j bit 2 is the most significant index bit of the pass
template <int NP>
__device__ __forceinline__ void fwd_pass8(uint32_t (&x)[NP][8], uint2 w0, uint2 w10, uint2 w11, uint2 w20, uint2 w21, uint2 w22, uint2 w23,
                                          uint32_t q, uint32_t two_q) {
#pragma unroll
    for (int p = 0; p < NP; p++) {
#pragma unroll
        for (int j = 0; j < 4; j++) bf_fwd(x[p][j], x[p][j + 4], w0, q, two_q);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            bf_fwd(x[p][j], x[p][j + 2], w10, q, two_q);
            bf_fwd(x[p][4 + j], x[p][6 + j], w11, q, two_q);
        }
        bf_fwd(x[p][0], x[p][1], w20, q, two_q);
        bf_fwd(x[p][2], x[p][3], w21, q, two_q);
        bf_fwd(x[p][4], x[p][5], w22, q, two_q);
        bf_fwd(x[p][6], x[p][7], w23, q, two_q);
    }
}
// three inverse stages (j bit 0 first)
template <int NP>
__device__ __forceinline__ void inv_pass8(uint32_t (&x)[NP][8], uint2 w00, uint2 w01, uint2 w02, uint2 w03, uint2 w10, uint2 w11, uint2 w2,
                                          uint32_t q, uint32_t two_q) {
#pragma unroll
    for (int p = 0; p < NP; p++) {
        bf_inv(x[p][0], x[p][1], w00, q, two_q);
        bf_inv(x[p][2], x[p][3], w01, q, two_q);
        bf_inv(x[p][4], x[p][5], w02, q, two_q);
        bf_inv(x[p][6], x[p][7], w03, q, two_q);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            bf_inv(x[p][j], x[p][j + 2], w10, q, two_q);
            bf_inv(x[p][4 + j], x[p][6 + j], w11, q, two_q);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) bf_inv(x[p][j], x[p][j + 4], w2, q, two_q);
    }
}

__device__ __forceinline__ uint2 lo2(uint4 v) { return make_uint2(v.x, v.y); }
__device__ __forceinline__ uint2 hi2(uint4 v) { return make_uint2(v.z, v.w); }
__device__ __forceinline__ uint4 ldg4(const uint2 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// Forward transform of NP polynomials held as x[p][j] = coefficient j*128 + t; on exit x[p][m] = value m of the
// thread's contiguous output words [8t, 8t + 8) (bit-reversed order = the reference's forward output order).
template <int NP>
__device__ __forceinline__ void forward_pair(uint32_t (&x)[NP][8], const Params &P, uint32_t *e1, uint32_t *e2, int t, uint32_t q, uint32_t two_q) {
    const int h = t >> 4, l = t & 15, g = t >> 1, beta = t & 1;
    // pass A: index bits 9..7, twiddles fwd[1], fwd[2..3], fwd[4..7] (kernel parameters)
    fwd_pass8<NP>(x, P.fwd_head[1], P.fwd_head[2], P.fwd_head[3], P.fwd_head[4], P.fwd_head[5], P.fwd_head[6], P.fwd_head[7], q, two_q);
    {
        uint32_t *s = e1 + t;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * E1W + j * 144] = x[p][j];
    }
    // twiddles of pass B (bits 6..4): fwd[8 + h], fwd[16 + 2h ..], fwd[32 + 4h ..] -- issued before the barrier
    const uint2 b0 = __ldg(P.fwd + 8 + h);
    const uint4 b1 = ldg4(P.fwd + 16 + 2 * h), b2 = ldg4(P.fwd + 32 + 4 * h), b3 = ldg4(P.fwd + 34 + 4 * h);
    __syncthreads();
    {
        const uint32_t *s = e1 + h * 144 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * E1W + j * 16];
    }
    fwd_pass8<NP>(x, b0, lo2(b1), hi2(b1), lo2(b2), hi2(b2), lo2(b3), hi2(b3), q, two_q);
    {
        uint32_t *s = e2 + h * 144 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * E2W + j * 18] = x[p][j];
    }
    // twiddles of pass C (bits 3..1): fwd[64 + g], fwd[128 + 2g ..], fwd[256 + 4g ..]
    const uint2 c0 = __ldg(P.fwd + 64 + g);
    const uint4 c1 = ldg4(P.fwd + 128 + 2 * g), c2 = ldg4(P.fwd + 256 + 4 * g), c3 = ldg4(P.fwd + 258 + 4 * g);
    __syncwarp();
    {
        const uint32_t *s = e2 + 18 * g + beta;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * E2W + j * 2];
    }
    fwd_pass8<NP>(x, c0, lo2(c1), hi2(c1), lo2(c2), hi2(c2), lo2(c3), hi2(c3), q, two_q);
    // last stage (bit 0) across lane pairs: thread beta keeps the pairs with bit 3 == beta
    const uint4 d0 = ldg4(P.fwd + 512 + 4 * t), d1 = ldg4(P.fwd + 514 + 4 * t);
    const uint2 dw[4] = {lo2(d0), hi2(d0), lo2(d1), hi2(d1)};
#pragma unroll
    for (int p = 0; p < NP; p++) {
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t send = beta ? x[p][k] : x[p][k + 4];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
            uint32_t X = beta ? recv : x[p][k];
            uint32_t Y = beta ? x[p][k + 4] : recv;
            bf_fwd(X, Y, dw[k], q, two_q);
            o[2 * k] = X;
            o[2 * k + 1] = Y;
        }
#pragma unroll
        for (int m = 0; m < 8; m++) x[p][m] = o[m];
    }
}

// Inverse transform of NP polynomials: x[p][m] = word 8t + m, any value <= P.first_inv_bias, on entry;
// x[p][j] = canonical coefficient j*128 + t on exit (includes n^-1 and the compensating 2^32).
template <int NP>
__device__ __forceinline__ void inverse_pair(uint32_t (&x)[NP][8], const Params &P, uint32_t *e1, uint32_t *e2, int t, uint32_t q, uint32_t two_q) {
    const int h = t >> 4, l = t & 15, g = t >> 1, beta = t & 1;
    // first stage (gap 1): inputs up to first_inv_bias; the sum is folded with the Shoup quotient of 1.
    // twiddles inv[1 + 4t + k] (odd pair index: one 8-byte, one 16-byte, one 8-byte load)
    const uint4 dmid = ldg4(P.inv + 2 + 4 * t);
    const uint2 dw[4] = {__ldg(P.inv + 1 + 4 * t), lo2(dmid), hi2(dmid), __ldg(P.inv + 4 + 4 * t)};
    const uint2 one = make_uint2(1u, P.one_q);
    const uint32_t bias = P.first_inv_bias;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t a = x[p][2 * k], b = x[p][2 * k + 1];
            const uint32_t X = shoup_lazy32(a + b, one, q);          // [0, 2q)
            const uint32_t Y = shoup_lazy32(a + bias - b, dw[k], q);  // [0, 2q)
            const uint32_t send = beta ? X : Y;
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
            o[k] = beta ? recv : X;
            o[k + 4] = beta ? Y : recv;
        }
#pragma unroll
        for (int m = 0; m < 8; m++) x[p][m] = o[m];
    }
    // pass C' (bits 1..3): inv[base1 + 4g ..], inv[base2 + 2g ..], inv[base3 + g], base_lg = 1 + N - (N >> lg)
    {
        const uint2 *i1 = P.inv + (1 + N - (N >> 1)) + 4 * g, *i2 = P.inv + (1 + N - (N >> 2)) + 2 * g, *i3 = P.inv + (1 + N - (N >> 3)) + g;
        const uint4 cmid = ldg4(i1 + 1);
        inv_pass8<NP>(x, __ldg(i1), lo2(cmid), hi2(cmid), __ldg(i1 + 3), __ldg(i2), __ldg(i2 + 1), __ldg(i3), q, two_q);
    }
    __syncwarp();  // every lane of the 16-thread group is done with the previous transform's second exchange
    {
        uint32_t *s = e2 + 18 * g + beta;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * E2W + j * 2] = x[p][j];
    }
    const uint2 *j4 = P.inv + (1 + N - (N >> 4)) + 4 * h, *j5 = P.inv + (1 + N - (N >> 5)) + 2 * h, *j6 = P.inv + (1 + N - (N >> 6)) + h;
    const uint4 bmid = ldg4(j4 + 1);
    const uint2 b00 = __ldg(j4), b01 = lo2(bmid), b02 = hi2(bmid), b03 = __ldg(j4 + 3), b10 = __ldg(j5), b11 = __ldg(j5 + 1), b2 = __ldg(j6);
    __syncwarp();
    {
        const uint32_t *s = e2 + h * 144 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * E2W + j * 18];
    }
    inv_pass8<NP>(x, b00, b01, b02, b03, b10, b11, b2, q, two_q);
    {
        uint32_t *s = e1 + h * 144 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * E1W + j * 16] = x[p][j];
    }
    __syncthreads();
    {
        const uint32_t *s = e1 + t;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * E1W + j * 144];
    }
    // pass A' (bits 7..9): inv[N-7 .. N-4], inv[N-3 .. N-2], final stage fused with n^-1 (table.rs:397-400), times 2^32
#pragma unroll
    for (int p = 0; p < NP; p++) {
        bf_inv(x[p][0], x[p][1], P.inv_tail[1], q, two_q);
        bf_inv(x[p][2], x[p][3], P.inv_tail[2], q, two_q);
        bf_inv(x[p][4], x[p][5], P.inv_tail[3], q, two_q);
        bf_inv(x[p][6], x[p][7], P.inv_tail[4], q, two_q);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            bf_inv(x[p][j], x[p][j + 2], P.inv_tail[5], q, two_q);
            bf_inv(x[p][4 + j], x[p][6 + j], P.inv_tail[6], q, two_q);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t tx = x[p][j] + x[p][j + 4], ty = x[p][j] + two_q - x[p][j + 4];
            const uint32_t a = shoup_lazy32(tx, make_uint2(P.invn_r, P.invn_r_q), q);
            const uint32_t b = shoup_lazy32(ty, make_uint2(P.invnw_r, P.invnw_r_q), q);
            x[p][j] = min(a, a - q);
            x[p][j + 4] = min(b, b - q);
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(TPP, MINB)
blind_rotate_n1024_kernel(const __grid_constant__ Params P, const uint32_t *__restrict__ bsk, uint32_t n_lwe, const uint32_t *__restrict__ lwe,
                          const uint32_t *__restrict__ test_vector, uint32_t *__restrict__ acc_out) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *e1 = smem, *e2 = smem + 4 * E1W, *accs = smem + 4 * E1W + 2 * E2W;  // accs: [2][N] canonical words, natural order
    const int t = threadIdx.x;
    const size_t ct = blockIdx.x;
    const uint32_t q = P.q, two_q = P.two_q;
    const uint32_t *my_lwe = lwe + ct * (size_t)(n_lwe + 1);
    constexpr uint32_t kMask2N = 2 * N - 1;
    // ACC <- (0, tv * X^(2N - b))     (mul_monomial_assign, primus_poly/src/poly/mul.rs:74-99)
    {
        const uint32_t b = __ldg(my_lwe + n_lwe) & kMask2N;
        const uint32_t rot = (2 * N - b) & kMask2N;
        for (int i = t; i < N; i += TPP) {
            accs[i] = 0;
            const uint32_t srcw = ((uint32_t)i - rot) & kMask2N;
            const uint32_t v = __ldg(test_vector + (srcw & (N - 1)));
            accs[N + i] = (srcw >= (uint32_t)N) ? (v == 0 ? 0u : q - v) : v;
        }
    }
    __syncthreads();
    const uint32_t levels = P.levels;
    const size_t rgsw_len = (size_t)2 * levels * 2 * N;
    int par = 0;
    uint32_t a_next = n_lwe ? __ldg(my_lwe) & kMask2N : 0u;
#pragma unroll 1
    for (uint32_t i = 0; i < n_lwe; i++) {
        const uint32_t a = a_next;
        if (i + 1 < n_lwe) a_next = __ldg(my_lwe + i + 1) & kMask2N;
        const uint32_t *key = bsk + (size_t)i * rgsw_len + (size_t)t * 8;
        // D = ACC * X^a - ACC (both components), prepared for digit extraction: W = adj(D) + R
        uint32_t W[2][8];
        {
            const uint32_t base = ((uint32_t)t - a) & kMask2N;
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t src = (base + 128u * j) & kMask2N;
                    const uint32_t v = accs[r * N + (src & (N - 1))];
                    const uint32_t p = accs[r * N + j * 128 + t];
                    const uint32_t s = (src & N) ? q - v : v;  // [0, q]
                    uint32_t d = s + q - p;                      // (0, 2q]
                    d = min(d, d - q);
                    d = min(d, d - q);                           // canonical [0, q)
                    W[r][j] = d + (d >= P.threshold ? P.add_r : P.r);
                }
        }
        uint64_t acc[2][8];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int m = 0; m < 8; m++) acc[c][m] = 0;
#pragma unroll 1
        for (uint32_t lv = 0; lv < levels; lv++) {
            const uint32_t shift = P.drop_bits + lv * P.log_basis;
            uint32_t x[2][8];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) x[r][j] = ((W[r][j] >> shift) & P.mask) + P.digit_off;
            forward_pair<2>(x, P, e1 + par * 2 * E1W, e2, t, q, two_q);
            par ^= 1;
            // acc[c] += fwd(digit_l(D_r)) * key[r][l][c]
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const uint32_t *kp = key + ((size_t)(r * levels + lv) * 2 + c) * N;
                    const uint4 k0 = __ldg(reinterpret_cast<const uint4 *>(kp)), k1 = __ldg(reinterpret_cast<const uint4 *>(kp) + 1);
                    acc[c][0] += (uint64_t)x[r][0] * k0.x;
                    acc[c][1] += (uint64_t)x[r][1] * k0.y;
                    acc[c][2] += (uint64_t)x[r][2] * k0.z;
                    acc[c][3] += (uint64_t)x[r][3] * k0.w;
                    acc[c][4] += (uint64_t)x[r][4] * k1.x;
                    acc[c][5] += (uint64_t)x[r][5] * k1.y;
                    acc[c][6] += (uint64_t)x[r][6] * k1.z;
                    acc[c][7] += (uint64_t)x[r][7] * k1.w;
                }
        }
        // Montgomery reduction of the lazy sums: r = (acc - m q) / 2^32 + q,  m = lo * q^-1 mod 2^32
        uint32_t y[2][8];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const uint32_t lo = (uint32_t)acc[c][m], hi = (uint32_t)(acc[c][m] >> 32);
                y[c][m] = hi + P.redc_bias - __umulhi(lo * P.qinv, q);
            }
        inverse_pair<2>(y, P, e1 + par * 2 * E1W, e2, t, q, two_q);
        par ^= 1;
        // ACC_c += result (each coefficient owned by exactly one thread)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t s = accs[c * N + j * 128 + t] + y[c][j];
                accs[c * N + j * 128 + t] = min(s, s - q);
            }
        __syncthreads();
    }
    uint32_t *o = acc_out + ct * 2 * N;
    for (int i = t; i < 2 * N; i += TPP) o[i] = accs[i];
}

// Ternary-secret ("monomial combination") blind rotation, SURVEY.md 8(f)2.  Per LWE coefficient a_i the accumulator is multiplied by
// RGSW(X^(a_i s_i)), s_i in {-1, 0, 1}, through ONE external product with the combined key
//     K_i = (NTT(X^a_i) - 1) .* BSK+_i + (NTT(X^-a_i) - 1) .* BSK-_i
// (NTT(X^d) = NttTable::transform_coeff_one_monomial(d), primus_ntt/src/ntt/prime64/table.rs:611-651; NTT(1) = all ones):
//     ACC <- ACC + INTT( sum_{r,l} fwd(digit_l(ACC_r)) .* K_i[r][l][c] ).
// K_i is never formed: the monomial factors are constant over (r, l), so the lazy sums against BSK+ and BSK- are kept apart and
// combined once per output coefficient:  sum fwd(d) .* K = m+ .* (sum fwd(d) .* BSK+) + m- .* (sum fwd(d) .* BSK-)  (exact mod q).
template <int MINB>
__global__ void __launch_bounds__(TPP, MINB)
blind_rotate_ternary_n1024_kernel(const __grid_constant__ Params P, const uint32_t *__restrict__ bsk_plus, const uint32_t *__restrict__ bsk_minus,
                                  uint32_t n_lwe, const uint32_t *__restrict__ lwe, const uint32_t *__restrict__ test_vector,
                                  uint32_t *__restrict__ acc_out) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *e1 = smem, *e2 = smem + 4 * E1W, *accs = smem + 4 * E1W + 2 * E2W;
    const int t = threadIdx.x;
    const size_t ct = blockIdx.x;
    const uint32_t q = P.q, two_q = P.two_q;
    const uint32_t *my_lwe = lwe + ct * (size_t)(n_lwe + 1);
    constexpr uint32_t kMask2N = 2 * N - 1;
    {
        const uint32_t b = __ldg(my_lwe + n_lwe) & kMask2N;
        const uint32_t rot = (2 * N - b) & kMask2N;
        for (int i = t; i < N; i += TPP) {
            accs[i] = 0;
            const uint32_t srcw = ((uint32_t)i - rot) & kMask2N;
            const uint32_t v = __ldg(test_vector + (srcw & (N - 1)));
            accs[N + i] = (srcw >= (uint32_t)N) ? (v == 0 ? 0u : q - v) : v;
        }
    }
    __syncthreads();
    const uint32_t levels = P.levels;
    const size_t rgsw_len = (size_t)2 * levels * 2 * N;
    // exponent base of this thread's NTT-domain words i = 8t + m: (2 brv(i) + 1)
    uint32_t odd[8];
#pragma unroll
    for (int m = 0; m < 8; m++) odd[m] = 2u * (__brev((uint32_t)(8 * t + m)) >> (32 - LOGN)) + 1u;
    const uint2 r32 = make_uint2(P.r32, P.r32_q);
    int par = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < n_lwe; i++) {
        const uint32_t a = __ldg(my_lwe + i) & kMask2N;
        const uint32_t *kp0 = bsk_plus + (size_t)i * rgsw_len + (size_t)t * 8, *km0 = bsk_minus + (size_t)i * rgsw_len + (size_t)t * 8;
        uint32_t W[2][8];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t d = accs[r * N + j * 128 + t];  // canonical
                W[r][j] = d + (d >= P.threshold ? P.add_r : P.r);
            }
        uint64_t accp[2][8], accm[2][8];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int m = 0; m < 8; m++) accp[c][m] = accm[c][m] = 0;
#pragma unroll 1
        for (uint32_t lv = 0; lv < levels; lv++) {
            const uint32_t shift = P.drop_bits + lv * P.log_basis;
            uint32_t x[2][8];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) x[r][j] = ((W[r][j] >> shift) & P.mask) + P.digit_off;
            forward_pair<2>(x, P, e1 + par * 2 * E1W, e2, t, q, two_q);
            par ^= 1;
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const size_t off = ((size_t)(r * levels + lv) * 2 + c) * N;
                    const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(kp0 + off)), p1 = __ldg(reinterpret_cast<const uint4 *>(kp0 + off) + 1);
                    const uint4 m0 = __ldg(reinterpret_cast<const uint4 *>(km0 + off)), m1 = __ldg(reinterpret_cast<const uint4 *>(km0 + off) + 1);
                    const uint32_t kp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w}, km[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        accp[c][m] += (uint64_t)x[r][m] * kp[m];
                        accm[c][m] += (uint64_t)x[r][m] * km[m];
                    }
                }
        }
        // monomial factors of this step in Montgomery form: (NTT(X^{+-a}) - 1) * 2^32 mod q
        uint32_t y[2][8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const uint32_t e = (odd[m] * a) & kMask2N;
            const uint32_t wp = __ldg(P.ordinal + e) - 1u, wm = __ldg(P.ordinal + ((2 * N - e) & kMask2N)) - 1u;   // canonical, >= 0
            const uint32_t mp = shoup_lazy32(wp, r32, q), mm = shoup_lazy32(wm, r32, q);                           // [0, 2q)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const uint32_t yp = (uint32_t)(accp[c][m] >> 32) + q - __umulhi((uint32_t)accp[c][m] * P.qinv, q);   // (0, (terms+1) q]
                const uint32_t ym = (uint32_t)(accm[c][m] >> 32) + q - __umulhi((uint32_t)accm[c][m] * P.qinv, q);
                const uint64_t s = (uint64_t)yp * mp + (uint64_t)ym * mm;                                           // < 4 (terms+1) q^2 < 2^64
                y[c][m] = (uint32_t)(s >> 32) + q - __umulhi((uint32_t)s * P.qinv, q);                               // <= first_inv_bias
            }
        }
        inverse_pair<2>(y, P, e1 + par * 2 * E1W, e2, t, q, two_q);
        par ^= 1;
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t s = accs[c * N + j * 128 + t] + y[c][j];
                accs[c * N + j * 128 + t] = min(s, s - q);
            }
        __syncthreads();
    }
    uint32_t *o = acc_out + ct * 2 * N;
    for (int i = t; i < 2 * N; i += TPP) o[i] = accs[i];
}

static uint32_t inv_mod_2_32(uint32_t q) {  // q odd
    uint32_t x = q;  // 3 correct bits
    for (int i = 0; i < 5; i++) x *= 2u - q * x;
    return x;
}

}  // namespace br32

static bool build_params(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g, uint64_t terms, br32::Params &P) {
    using namespace br32;
    if (tb.log_n != LOGN) return false;
    const uint64_t q = tb.q;
    if ((2ull * LOGN + 2) * q >= (1ull << 32) || 2 * (terms + 1) * q >= (1ull << 32)) return false;
    P.q = tb.q;
    P.two_q = tb.two_q;
    P.qinv = br32::inv_mod_2_32(tb.q);
    P.one_q = (uint32_t)((1ull << 32) / q);
    const uint32_t r32 = (uint32_t)((1ull << 32) % q);
    P.r32 = r32;
    P.r32_q = host::shoup_quot<uint32_t>(r32, tb.q);
    P.invn_r = host::mulmod<uint32_t>(tb.inv_n, r32, tb.q);
    P.invn_r_q = host::shoup_quot<uint32_t>(P.invn_r, tb.q);
    P.invnw_r = host::mulmod<uint32_t>(head.inv_tail[7].x, r32, tb.q);
    P.invnw_r_q = host::shoup_quot<uint32_t>(P.invnw_r, tb.q);
    P.redc_bias = tb.q;
    P.first_inv_bias = (uint32_t)((terms + 1) * q);
    const uint32_t half = g.log_basis == 1 ? 0u : (1u << (g.log_basis - 1));
    uint32_t R = g.drop_bits ? (1u << (g.drop_bits - 1)) : 0u;
    for (uint32_t l = 0; l < g.levels; l++) R += half << (g.drop_bits + l * g.log_basis);
    P.r = R;
    P.threshold = g.has_threshold ? g.threshold : 0xffffffffu;
    P.add_r = g.add + R;
    P.mask = g.basis_m1;
    P.digit_off = half ? tb.q - half : 0u;
    P.drop_bits = g.drop_bits;
    P.log_basis = g.log_basis;
    P.levels = g.levels;
    for (int i = 0; i < 8; i++) {
        P.fwd_head[i] = head.fwd_head[i];
        P.inv_tail[i] = head.inv_tail[i];
    }
    P.fwd = tb.fwd;
    P.inv = tb.inv;
    P.ordinal = tb.ordinal;
    return true;
}

// Ternary-secret blind rotation (pfhe_blind_rotate_ternary32_batch); cudaErrorNotSupported outside the u32 / N = 1024 shape.
cudaError_t launch_blind_rotate_ternary32(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g,
                                          const uint32_t *bsk_plus, const uint32_t *bsk_minus, uint32_t n_lwe, const uint32_t *lwe,
                                          const uint32_t *tv, uint32_t *acc_out, size_t batch, cudaStream_t stream) {
    using namespace br32;
    Params P{};
    // every Montgomery output is at most (terms + 1) q with terms = 2 levels; the combined value is below 2q
    if (!build_params(tb, head, g, 2ull * g.levels, P)) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    constexpr size_t smem = sizeof(uint32_t) * SMEM_WORDS;
    auto k = blind_rotate_ternary_n1024_kernel<3>;
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k<<<(unsigned)batch, TPP, smem, stream>>>(P, bsk_plus, bsk_minus, n_lwe, lwe, tv, acc_out);
    count_launch();
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the shape / modulus does not qualify (the caller then uses the generic kernel).
cudaError_t launch_blind_rotate_fast32(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g,
                                       const uint32_t *bsk, uint32_t n_lwe, const uint32_t *lwe, const uint32_t *tv, uint32_t *acc_out,
                                       size_t batch, cudaStream_t stream) {
    using namespace br32;
    static const bool off = getenv("PFHE_BR_FAST") && getenv("PFHE_BR_FAST")[0] == '0';  // A/B tuning hook
    Params P{};
    if (off || !build_params(tb, head, g, 2ull * g.levels, P)) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    constexpr size_t smem = sizeof(uint32_t) * SMEM_WORDS;
    const char *e = getenv("PFHE_BR_MINB");  // tuning hook: resident CTAs per SM the register allocator is asked for
    const int mb = e ? atoi(e) : 4;
    auto launch = [&](auto k) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        k<<<(unsigned)batch, TPP, smem, stream>>>(P, bsk, n_lwe, lwe, tv, acc_out);
        count_launch();
        return cudaGetLastError();
    };
    if (mb == 3) return launch(blind_rotate_n1024_kernel<3>);
    if (mb == 5) return launch(blind_rotate_n1024_kernel<5>);
    if (mb == 6) return launch(blind_rotate_n1024_kernel<6>);
    return launch(blind_rotate_n1024_kernel<4>);
}

}  // namespace pfhe
