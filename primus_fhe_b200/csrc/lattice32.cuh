// lattice32.cuh -- building blocks of the re-scheduled u32 lattice kernels (lattice32.cu: blind rotation, N = 1024;
// lattice32_ep.cu: external products, N = 1024 / 2048): kernel-parameter block, butterflies, 8-word register passes and the
// paired transforms of the N = 1024 schedule.  See the header comment of lattice32.cu for the design.
#pragma once
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

}  // namespace br32

// host: fills the parameter block; false when (table, gadget, number of accumulated terms) does not satisfy the preconditions
// ((2 log2 N + 2) q < 2^32, 2 (terms + 1) q < 2^32) or the degree is not log_n
bool br32_build_params(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g, uint64_t terms, uint32_t log_n,
                       br32::Params &P);

}  // namespace pfhe
