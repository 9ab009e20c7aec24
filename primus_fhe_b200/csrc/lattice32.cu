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
#include "lattice32.cuh"

namespace pfhe {
namespace br32 {

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

bool br32_build_params(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g, uint64_t terms, uint32_t log_n,
                       br32::Params &P) {
    if (tb.log_n != log_n) return false;
    const uint64_t q = tb.q;
    if ((2ull * log_n + 2) * q >= (1ull << 32) || 2 * (terms + 1) * q >= (1ull << 32)) return false;
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
    if (!br32_build_params(tb, head, g, 2ull * g.levels, LOGN, P)) return cudaErrorNotSupported;
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
    if (off || !br32_build_params(tb, head, g, 2ull * g.levels, LOGN, P)) return cudaErrorNotSupported;
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
