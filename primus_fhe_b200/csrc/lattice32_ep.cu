// lattice32_ep.cu -- fused RLWE external products (k = 1) on u32 words with the schedule of lattice32.cu (paired transforms, padded
// exchange buffers, carry-free digits, Montgomery-reduced lazy sums):
//   N = 1024: 128 threads per ciphertext, the transforms of lattice32.cuh;
//   N = 2048 (BASELINE config 4, instance A): 256 threads per ciphertext, 11 stages as 3 + 3 + 3 + 2 register passes.
// out_c = [inv]( sum_{r<=1} sum_l fwd(digit_l(in_r)) .* key[r][l][c] ) = CrtGlwe::mul_dcrt_ggsw_to [+ into_coeff_form]
// (primus_lattice/src/glwe/crt.rs:200-227, macros/mod.rs:892-937), digits = ApproxSignedBasis (primus_decompose/src/primitive/
// basis.rs:254-283, common.rs:246-259).  Canonical outputs are bit-identical to external_product_kernel (lattice.cu).
//
// N = 2048 schedule (thread t of 256, 8 words per thread):
//   pass A  index bits 10..8   thread = bits 7..0                 words idx = j*256 + t          twiddles fwd[1..7] (parameters)
//   pass B  bits 7..5          thread = (b10..8)<<5 | b4..0       idx = h*256 + j*32 + l          fwd[8+h], fwd[16+2h..], fwd[32+4h..]
//   pass C  bits 4..2          thread = (b10..5)<<2 | b1..0       idx = g*32 + j*4 + lam          fwd[64+g], fwd[128+2g..], fwd[256+4g..]
//   pass D  bits 1..0          thread = b10..3                    idx = 8t + m                    fwd[512+2t..], fwd[1024+4t..]
// Exchange 1 (A->B) is the only CTA barrier of a transform and needs no padding (both patterns are unit stride over the lanes);
// exchange 2 (B->C, inside one warp) and exchange 3 (C->D, inside 4-thread groups) share a buffer padded as idx + 4*(idx >> 5)
// (36 words per 32): conflict free for the stride-4 words of pass C and for the 128-bit loads of pass D.
#include "lattice32.cuh"

namespace pfhe {
namespace ep32 {
using namespace br32;

// ---- N = 2048 transforms ---------------------------------------------------------------------------------------------------
constexpr int LOGN2 = 11, N2 = 1 << LOGN2, TPP2 = 256;
constexpr int F1W = N2;        // exchange-1 buffer words per polynomial
constexpr int F2W = 36 * 64;   // exchange-2/3 buffer words per polynomial (idx + 4 * (idx >> 5))

__device__ __forceinline__ void radix4_fwd(uint32_t (&x)[8], uint2 wa0, uint2 wa1, uint2 wb0, uint2 wb1, uint2 wb2, uint2 wb3, uint32_t q, uint32_t two_q) {
    // index bit 1 (m bit 1), then bit 0, on the 8 contiguous words of a thread (m bit 2 is a spectator)
    bf_fwd(x[0], x[2], wa0, q, two_q);
    bf_fwd(x[1], x[3], wa0, q, two_q);
    bf_fwd(x[4], x[6], wa1, q, two_q);
    bf_fwd(x[5], x[7], wa1, q, two_q);
    bf_fwd(x[0], x[1], wb0, q, two_q);
    bf_fwd(x[2], x[3], wb1, q, two_q);
    bf_fwd(x[4], x[5], wb2, q, two_q);
    bf_fwd(x[6], x[7], wb3, q, two_q);
}

// x[p][j] = coefficient j*256 + t  ->  x[p][m] = output word 8t + m (bit-reversed order)
template <int NP>
__device__ __forceinline__ void forward_pair_2048(uint32_t (&x)[NP][8], const Params &P, uint32_t *f1, uint32_t *f2, int t, uint32_t q, uint32_t two_q) {
    const int h = t >> 5, l = t & 31, g = t >> 2, lam = t & 3;
    fwd_pass8<NP>(x, P.fwd_head[1], P.fwd_head[2], P.fwd_head[3], P.fwd_head[4], P.fwd_head[5], P.fwd_head[6], P.fwd_head[7], q, two_q);
    {
        uint32_t *s = f1 + t;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * F1W + j * 256] = x[p][j];
    }
    const uint2 b0 = __ldg(P.fwd + 8 + h);
    const uint4 b1 = ldg4(P.fwd + 16 + 2 * h), b2 = ldg4(P.fwd + 32 + 4 * h), b3 = ldg4(P.fwd + 34 + 4 * h);
    __syncthreads();
    {
        const uint32_t *s = f1 + h * 256 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * F1W + j * 32];
    }
    fwd_pass8<NP>(x, b0, lo2(b1), hi2(b1), lo2(b2), hi2(b2), lo2(b3), hi2(b3), q, two_q);
    {
        uint32_t *s = f2 + h * 288 + l;  // idx + 4*(idx>>5) with idx = h*256 + j*32 + l
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * F2W + j * 36] = x[p][j];
    }
    const uint2 c0 = __ldg(P.fwd + 64 + g);
    const uint4 c1 = ldg4(P.fwd + 128 + 2 * g), c2 = ldg4(P.fwd + 256 + 4 * g), c3 = ldg4(P.fwd + 258 + 4 * g);
    __syncwarp();
    {
        const uint32_t *s = f2 + 36 * g + lam;  // idx = g*32 + j*4 + lam
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * F2W + j * 4];
    }
    fwd_pass8<NP>(x, c0, lo2(c1), hi2(c1), lo2(c2), hi2(c2), lo2(c3), hi2(c3), q, two_q);
    {
        uint32_t *s = f2 + 36 * g + lam;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * F2W + j * 4] = x[p][j];
    }
    const uint4 d0 = ldg4(P.fwd + 512 + 2 * t), d1 = ldg4(P.fwd + 1024 + 4 * t), d2 = ldg4(P.fwd + 1026 + 4 * t);
    __syncwarp();
    {
        const uint32_t *s = f2 + 36 * g + 8 * lam;  // words 8t .. 8t+7: idx = 32g + 8 lam + m
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(s + p * F2W), v1 = *reinterpret_cast<const uint4 *>(s + p * F2W + 4);
            x[p][0] = v0.x; x[p][1] = v0.y; x[p][2] = v0.z; x[p][3] = v0.w;
            x[p][4] = v1.x; x[p][5] = v1.y; x[p][6] = v1.z; x[p][7] = v1.w;
        }
    }
#pragma unroll
    for (int p = 0; p < NP; p++) radix4_fwd(x[p], lo2(d0), hi2(d0), lo2(d1), hi2(d1), lo2(d2), hi2(d2), q, two_q);
}

// x[p][m] = word 8t + m (any value <= P.first_inv_bias)  ->  x[p][j] = canonical coefficient j*256 + t (n^-1 and 2^32 included)
template <int NP>
__device__ __forceinline__ void inverse_pair_2048(uint32_t (&x)[NP][8], const Params &P, uint32_t *f1, uint32_t *f2, int t, uint32_t q, uint32_t two_q) {
    const int h = t >> 5, l = t & 31, g = t >> 2, lam = t & 3;
    constexpr int B1 = 1 + N2 - (N2 >> 1), B2 = 1 + N2 - (N2 >> 2), B3 = 1 + N2 - (N2 >> 3), B4 = 1 + N2 - (N2 >> 4), B5 = 1 + N2 - (N2 >> 5),
                  B6 = 1 + N2 - (N2 >> 6), B7 = 1 + N2 - (N2 >> 7);
    {
        // gap 1 (inputs up to first_inv_bias, sums folded with the Shoup quotient of 1), then gap 2
        const uint4 dmid = ldg4(P.inv + 2 + 4 * t);
        const uint2 w0[4] = {__ldg(P.inv + 1 + 4 * t), lo2(dmid), hi2(dmid), __ldg(P.inv + 4 + 4 * t)};
        const uint2 w1a = __ldg(P.inv + B1 + 2 * t), w1b = __ldg(P.inv + B1 + 2 * t + 1);
        const uint2 one = make_uint2(1u, P.one_q);
        const uint32_t bias = P.first_inv_bias;
#pragma unroll
        for (int p = 0; p < NP; p++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t a = x[p][2 * k], b = x[p][2 * k + 1];
                x[p][2 * k] = shoup_lazy32(a + b, one, q);
                x[p][2 * k + 1] = shoup_lazy32(a + bias - b, w0[k], q);
            }
            bf_inv(x[p][0], x[p][2], w1a, q, two_q);
            bf_inv(x[p][1], x[p][3], w1a, q, two_q);
            bf_inv(x[p][4], x[p][6], w1b, q, two_q);
            bf_inv(x[p][5], x[p][7], w1b, q, two_q);
        }
    }
    __syncwarp();  // the 4-thread group is done with the previous transform's reads of f2
    {
        uint32_t *s = f2 + 36 * g + 8 * lam;
#pragma unroll
        for (int p = 0; p < NP; p++) {
            *reinterpret_cast<uint4 *>(s + p * F2W) = make_uint4(x[p][0], x[p][1], x[p][2], x[p][3]);
            *reinterpret_cast<uint4 *>(s + p * F2W + 4) = make_uint4(x[p][4], x[p][5], x[p][6], x[p][7]);
        }
    }
    const uint2 *i2 = P.inv + B2 + 4 * g, *i3 = P.inv + B3 + 2 * g, *i4 = P.inv + B4 + g;
    const uint4 cmid = ldg4(i2 + 1);
    const uint2 c00 = __ldg(i2), c03 = __ldg(i2 + 3), c10 = __ldg(i3), c11 = __ldg(i3 + 1), c2 = __ldg(i4);
    __syncwarp();
    {
        const uint32_t *s = f2 + 36 * g + lam;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * F2W + j * 4];
    }
    inv_pass8<NP>(x, c00, lo2(cmid), hi2(cmid), c03, c10, c11, c2, q, two_q);
    {
        uint32_t *s = f2 + 36 * g + lam;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * F2W + j * 4] = x[p][j];
    }
    const uint2 *j5 = P.inv + B5 + 4 * h, *j6 = P.inv + B6 + 2 * h, *j7 = P.inv + B7 + h;
    const uint4 bmid = ldg4(j5 + 1);
    const uint2 b00 = __ldg(j5), b03 = __ldg(j5 + 3), b10 = __ldg(j6), b11 = __ldg(j6 + 1), b2 = __ldg(j7);
    __syncwarp();
    {
        const uint32_t *s = f2 + h * 288 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * F2W + j * 36];
    }
    inv_pass8<NP>(x, b00, lo2(bmid), hi2(bmid), b03, b10, b11, b2, q, two_q);
    {
        uint32_t *s = f1 + h * 256 + l;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[p * F1W + j * 32] = x[p][j];
    }
    __syncthreads();
    {
        const uint32_t *s = f1 + t;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[p][j] = s[p * F1W + j * 256];
    }
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

// ---- the fused external product ----------------------------------------------------------------------------------------------
// BIG = false: N = 1024 (128 threads); BIG = true: N = 2048 (256 threads).  One ciphertext per CTA.
template <bool BIG, int MINB>
__global__ void __launch_bounds__(BIG ? TPP2 : TPP, MINB)
external_product_u32_kernel(const __grid_constant__ Params P, const uint32_t *__restrict__ key, const uint32_t *__restrict__ in,
                            uint32_t *__restrict__ out, int to_coeff) {
    constexpr int NN = BIG ? N2 : N, THREADS = BIG ? TPP2 : TPP;
    constexpr int W1 = BIG ? F1W : E1W, W2 = BIG ? F2W : E2W;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *x1 = smem, *x2 = smem + 4 * W1;  // x1: [parity 2][poly 2][W1], x2: [poly 2][W2]
    const int t = threadIdx.x;
    const size_t ct = blockIdx.x;
    const uint32_t q = P.q, two_q = P.two_q;
    const uint32_t *cin = in + ct * 2 * NN;
    uint32_t *cout = out + ct * 2 * NN;
    const uint32_t levels = P.levels;
    uint32_t W[2][8];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t d = __ldg(cin + r * NN + j * THREADS + t);  // canonical
            W[r][j] = d + (d >= P.threshold ? P.add_r : P.r);
        }
    uint64_t acc[2][8];
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int m = 0; m < 8; m++) acc[c][m] = 0;
    int par = 0;
    const uint32_t *kt = key + (size_t)t * 8;
#pragma unroll 1
    for (uint32_t lv = 0; lv < levels; lv++) {
        const uint32_t shift = P.drop_bits + lv * P.log_basis;
        uint32_t x[2][8];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int j = 0; j < 8; j++) x[r][j] = ((W[r][j] >> shift) & P.mask) + P.digit_off;
        if constexpr (BIG) forward_pair_2048<2>(x, P, x1 + par * 2 * W1, x2, t, q, two_q);
        else forward_pair<2>(x, P, x1 + par * 2 * W1, x2, t, q, two_q);
        par ^= 1;
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const uint32_t *kp = kt + ((size_t)(r * levels + lv) * 2 + c) * NN;
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
    uint32_t y[2][8];
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int m = 0; m < 8; m++) y[c][m] = (uint32_t)(acc[c][m] >> 32) + P.redc_bias - __umulhi((uint32_t)acc[c][m] * P.qinv, q);
    if (to_coeff) {
        if constexpr (BIG) inverse_pair_2048<2>(y, P, x1 + par * 2 * W1, x2, t, q, two_q);
        else inverse_pair<2>(y, P, x1 + par * 2 * W1, x2, t, q, two_q);
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) cout[c * NN + j * THREADS + t] = y[c][j];
    } else {
        // NTT-domain output: undo the Montgomery factor (times 2^32 mod q), canonical, 8 contiguous words per thread
        const uint2 r32 = make_uint2(P.r32, P.r32_q);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            uint32_t w[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const uint32_t v = shoup_lazy32(y[c][m], r32, q);
                w[m] = min(v, v - q);
            }
            uint4 *o = reinterpret_cast<uint4 *>(cout + c * NN + t * 8);
            o[0] = make_uint4(w[0], w[1], w[2], w[3]);
            o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

}  // namespace ep32

// k = 1 external product on u32 words, N = 1024 / 2048; cudaErrorNotSupported when the shape / modulus does not qualify.
cudaError_t launch_external_product_fast32(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g, uint32_t k,
                                           const uint32_t *key, const uint32_t *in, uint32_t *out, size_t batch, bool to_coeff,
                                           cudaStream_t stream) {
    using namespace ep32;
    static const bool off = getenv("PFHE_EP_FAST") && getenv("PFHE_EP_FAST")[0] == '0';  // A/B tuning hook
    if (off || k != 1 || (tb.log_n != 10 && tb.log_n != 11)) return cudaErrorNotSupported;
    br32::Params P{};
    if (!br32_build_params(tb, head, g, 2ull * g.levels, tb.log_n, P)) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    auto launch = [&](auto kern, int threads, size_t smem) -> cudaError_t {
        cudaError_t err = smem > 48 * 1024 ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
        if (err != cudaSuccess) return err;
        kern<<<(unsigned)batch, threads, smem, stream>>>(P, key, in, out, to_coeff ? 1 : 0);
        count_launch();
        return cudaGetLastError();
    };
    if (tb.log_n == 11) return launch(external_product_u32_kernel<true, 2>, TPP2, sizeof(uint32_t) * (4 * F1W + 2 * F2W));
    return launch(external_product_u32_kernel<false, 4>, br32::TPP, sizeof(uint32_t) * (4 * br32::E1W + 2 * br32::E2W));
}

}  // namespace pfhe
