// rns.cu -- RNS limb handling and the multi-limb (L > 1) gadget / external product, plus the remaining
// element-wise helpers of the path.
//
// Replaces:
//   RNSBase::{compose_multiple_values_to, decompose_big_uint_values_to}      primus_rns/src/base.rs:457-481, :609-673
//   RNSBase::wrapping_decompose_small_values_scaled_add_to (fused)            primus_rns/src/base.rs:326-386, :739-756
//   BigUintApproxSignedBasis init + unsigned levels + centred lift            primus_decompose/src/big_integer/basis.rs:326-367,
//                                                                             big_integer/common.rs:83-141,275-325; base.rs:279-315
//   DcrtGlwe::add_dcrt_glev_mul_crt_poly_assign / CrtGlwe::mul_dcrt_ggsw_to   primus_lattice/src/glwe/dcrt.rs:178-255, glwe/crt.rs:200-227
//   Polynomial::mul_monomial_assign / CrtGlwe::mul_monic_monomial_assign      primus_poly/src/poly/mul.rs:74-99, primus_lattice/src/glwe/crt.rs:76-114
//   reduce_dot_product                                                         primus_modulus/src/common/compact/slice.rs:371-401
// The L > 1 external product is composed from kernels (digits -> DCRT NTT -> fused MAC -> INTT) with a
// stream-ordered scratch buffer; the L = 1 case has the fully fused kernel in lattice.cu.
#include <cstring>
#include <type_traits>
#include <vector>

#include "host_math.hpp"
#include "internal.hpp"
#include "rns.hpp"

namespace pfhe {

static unsigned grid_for(size_t items, int threads) {
    size_t blocks = (items + threads - 1) / threads;
    const size_t cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

template <typename T> struct WideOf;
template <> struct WideOf<uint32_t> { using type = uint64_t; };
template <> struct WideOf<uint64_t> { using type = unsigned __int128; };

// ---- multiword helpers on little-endian word arrays (register / local arrays of kRnsMaxWords) ----------
template <typename T> __device__ __forceinline__ bool big_ge(const T *a, const T *b, int len) {
    for (int i = len - 1; i >= 0; i--)
        if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
template <typename T> __device__ __forceinline__ void big_sub(T *a, const T *b, int len) {
    T borrow = 0;
    for (int i = 0; i < len; i++) {
        const T bi = b[i], ai = a[i];
        const T d = ai - bi - borrow;
        borrow = (ai < bi) || (ai == bi && borrow) ? 1 : 0;
        a[i] = d;
    }
}
template <typename T> __device__ __forceinline__ void big_add(T *a, const T *b, int len) {
    T carry = 0;
    for (int i = 0; i < len; i++) {
        const T s = a[i] + b[i], s2 = s + carry;
        carry = (s < a[i]) || (s2 < s) ? 1 : 0;
        a[i] = s2;
    }
}
// acc += m * v, returns the carry word
template <typename T> __device__ __forceinline__ T big_mul_add(const T *m, T v, T *acc, int len) {
    using W = typename WideOf<T>::type;
    T carry = 0;
    for (int i = 0; i < len; i++) {
        const W s = (W)m[i] * v + acc[i] + carry;
        acc[i] = (T)s;
        carry = (T)(s >> (sizeof(T) * 8));
    }
    return carry;
}

// ---- register-resident multi-word arithmetic: the limb count L is a template parameter, the composed value is held in L words
// (value_len <= L; the unused top words of product / punct / threshold / add are zero, so uniform L-word arithmetic is exact) -----
template <typename T, int L> struct BigRegs {
    using W = typename WideOf<T>::type;
    static constexpr int BITS = sizeof(T) * 8;
    // d = a - b over L words; returns the final borrow
    __device__ __forceinline__ static T sub(const T (&a)[L], const T *b, T (&d)[L]) {
        T borrow = 0;
#pragma unroll
        for (int k = 0; k < L; k++) {
            const T t = a[k] - b[k];
            const T nb = (T)((a[k] < b[k]) | (t < borrow));
            d[k] = t - borrow;
            borrow = nb;
        }
        return borrow;
    }
    __device__ __forceinline__ static void add(T (&a)[L], const T *b) {
        T carry = 0;
#pragma unroll
        for (int k = 0; k < L; k++) {
            const T s = a[k] + b[k], s2 = s + carry;
            carry = (T)((s < a[k]) | (s2 < s));
            a[k] = s2;
        }
    }
    __device__ __forceinline__ static void shr_word(T (&a)[L]) {
#pragma unroll
        for (int k = 0; k + 1 < L; k++) a[k] = a[k + 1];
        a[L - 1] = 0;
    }
    __device__ __forceinline__ static void shr_bits(T (&a)[L], uint32_t b) {  // 0 < b < BITS
#pragma unroll
        for (int k = 0; k + 1 < L; k++) a[k] = (T)((a[k] >> b) | (a[k + 1] << (BITS - b)));
        a[L - 1] >>= b;
    }
    // compose_to (base.rs:609-636): value = sum_i (Q/q_i) * (x_i * (Q/q_i)^-1 mod q_i) mod Q, one conditional subtraction per term
    __device__ __forceinline__ static void compose(const RnsDev<T> &r, const T (&res)[L], T (&value)[L]) {
#pragma unroll
        for (int k = 0; k < L; k++) value[k] = 0;
#pragma unroll
        for (int i = 0; i < L; i++) {
            const T prod = shoup<T>(res[i], r.inv_punct[i], r.inv_punct_q[i], r.q[i]);
            T carry = 0;
#pragma unroll
            for (int k = 0; k < L; k++) {
                const W s = (W)r.punct[i][k] * prod + value[k] + carry;
                value[k] = (T)s;
                carry = (T)(s >> BITS);
            }
            if (i == 0) continue;  // the first term is below Q
            T d[L];
            const T borrow = sub(value, r.product, d);
            const bool ge = (carry != 0) | (borrow == 0);
#pragma unroll
            for (int k = 0; k < L; k++) value[k] = ge ? d[k] : value[k];
        }
    }
};

template <typename T, int L>
__global__ void __launch_bounds__(256) rns_compose_kernel(const __grid_constant__ RnsDev<T> r, const T *__restrict__ residues,
                                                          T *__restrict__ big, size_t count) {
    const int vl = r.value_len;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T res[L], value[L];
#pragma unroll
        for (int l = 0; l < L; l++) res[l] = ldg_stream(residues + (size_t)l * count + i);
        BigRegs<T, L>::compose(r, res, value);
        T *o = big + i * (size_t)vl;
#pragma unroll
        for (int k = 0; k < L; k++)
            if (k < vl) o[k] = value[k];
    }
}
// decompose_big_uint_values_to (base.rs:457-481): value mod q_l = sum_k word_k * (2^(BITS k) mod q_l) mod q_l, every term a Shoup
// product with a precomputed quotient (exact for arbitrary words)
template <typename T, int L>
__global__ void __launch_bounds__(256) rns_decompose_kernel(const __grid_constant__ RnsDev<T> r, const T *__restrict__ big,
                                                            T *__restrict__ residues, size_t count) {
    const int vl = r.value_len;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T w[L];
        const T *v = big + i * (size_t)vl;
#pragma unroll
        for (int k = 0; k < L; k++) w[k] = k < vl ? v[k] : (T)0;
#pragma unroll
        for (int l = 0; l < L; l++) {
            const T q = r.q[l];
            T acc = 0;
#pragma unroll
            for (int k = 0; k < L; k++)
                if (k < vl) acc = mod_add<T>(acc, shoup<T>(w[k], r.pw[l][k], r.pw_q[l][k], q), q);
            stg_stream(residues + (size_t)l * count + i, acc);
        }
    }
}

// residues[limbs][count] -> digits[levels][limbs][count]: compose, init_value_carry, unsigned digits, centred lift.
// One thread owns VEC consecutive coefficients so that every load / store is a 16-byte vector (the kernel is bound by
// its levels*limbs stores per coefficient).  The composed value lives in L registers per coefficient and is shifted down by
// log_basis per level, so the digit window is always the low bits of word 0 (no dynamically indexed word array).
template <typename T, int L>
__global__ void __launch_bounds__(128) rns_gadget_kernel(const __grid_constant__ RnsDev<T> r, const T *__restrict__ residues,
                                                         T *__restrict__ digits, size_t count, size_t polys, size_t in_stride,
                                                         size_t out_stride) {
    // `polys` independent CRT polynomials: residues + p*in_stride, digits + p*out_stride
    using BR = BigRegs<T, L>;
    constexpr int BITS = sizeof(T) * 8, VB = L <= 4 ? 16 : 8, VEC = VB / sizeof(T);  // many limbs: fewer coefficients per thread (registers)
    using Raw = typename std::conditional<VB == 16, uint4, uint64_t>::type;
    struct alignas(VB) V {
        T v[VEC];
    };
    const size_t cv = count / VEC;
    const T bm1 = r.basis_m1, half = (T)((r.basis_m1 + 2) / 2);  // ceil(B/2)
    const uint32_t pre = r.drop_bits ? r.drop_bits - 1 : 0;       // bits below the rounding bit
    for (size_t p = blockIdx.y; p < polys; p += gridDim.y) {
        const T *res_in = residues + p * in_stride;
        T *dig = digits + p * out_stride;
        for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < cv; g += (size_t)gridDim.x * blockDim.x) {
            const size_t i = g * VEC;
            T value[VEC][L];
            uint32_t carry[VEC];
            {
                V in[L];
#pragma unroll
                for (int l = 0; l < L; l++) *reinterpret_cast<Raw *>(&in[l]) = ldg_stream(reinterpret_cast<const Raw *>(res_in + (size_t)l * count + i));
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    T res[L];
#pragma unroll
                    for (int l = 0; l < L; l++) res[l] = in[l].v[k];
                    BR::compose(r, res, value[k]);
                    if (r.has_threshold) {
                        T d[L];
                        if (BR::sub(value[k], r.threshold, d) == 0) BR::add(value[k], r.add);  // value >= threshold
                    }
                    // init_value_carry: the rounding bit (bit drop-1), then drop the low bits
                    carry[k] = 0;
                    if (r.drop_bits) {
                        for (uint32_t s = 0; s < pre / BITS; s++) BR::shr_word(value[k]);
                        if (pre % BITS) BR::shr_bits(value[k], pre % BITS);
                        carry[k] = (uint32_t)(value[k][0] & 1);
                        BR::shr_bits(value[k], 1);
                    }
                }
            }
            for (uint32_t lv = 0; lv < r.levels; lv++) {
                T d[VEC];
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    const T t = (value[k][0] & bm1) + carry[k];
                    carry[k] = (t & r.carry_mask) != 0;
                    d[k] = t & bm1;
                    BR::shr_bits(value[k], r.log_basis);
                }
#pragma unroll
                for (int l = 0; l < L; l++) {
                    V o;
#pragma unroll
                    for (int k = 0; k < VEC; k++) o.v[k] = (r.basis_m1 == 1 || d[k] < half) ? d[k] : r.q[l] - (bm1 + 1) + d[k];
                    stg_stream(reinterpret_cast<Raw *>(dig + ((size_t)lv * L + l) * count + i), *reinterpret_cast<const Raw *>(&o));
                }
            }
        }
    }
}

// out[ct][c][limb][i] = sum_{r,l} digits[ct][r][l][limb][i] * key[r][l][c][limb][i] mod q_limb  (NTT domain)
// One thread owns VEC consecutive coefficients of one (ciphertext, limb) and ALL output components: every digit word is
// read from HBM exactly once (16-byte loads), the key comes from L2, sums are lazy double words (<= 16 terms per Barrett
// reduction, reduce_dot_product, primus_modulus/src/common/compact/slice.rs:371-401).
template <typename T, int COMPS>
__global__ void __launch_bounds__(256) rns_key_mac_kernel(const __grid_constant__ LimbConsts<T> lc, int limbs, uint32_t levels,
                                                          const T *__restrict__ digits, const T *__restrict__ key, T *__restrict__ out,
                                                          size_t n, size_t batch) {
    using W = typename WideOf<T>::type;
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V {
        T v[VEC];
    };
    const size_t nv = n / VEC, per_ct = (size_t)limbs * nv, total = batch * per_ct;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t ct = gid / per_ct, rem = gid % per_ct;
        const int limb = (int)(rem / nv);
        const size_t i = (rem % nv) * VEC;
        const Barrett<T> br = lc.br[limb];
        W acc[COMPS][VEC];
#pragma unroll
        for (int c = 0; c < COMPS; c++)
#pragma unroll
            for (int k = 0; k < VEC; k++) acc[c][k] = 0;
        uint32_t terms = 0;
        for (int r = 0; r < COMPS; r++) {
            for (uint32_t l = 0; l < levels; l++) {
                const V d = *reinterpret_cast<const V *>(digits + ((((ct * COMPS + r) * levels + l) * limbs + limb) * n + i));
                if (terms == 16) {
#pragma unroll
                    for (int c = 0; c < COMPS; c++)
#pragma unroll
                        for (int k = 0; k < VEC; k++) acc[c][k] = barrett_reduce_wide(br, (T)acc[c][k], (T)(acc[c][k] >> (sizeof(T) * 8)));
                    terms = 1;
                }
                terms++;
#pragma unroll
                for (int c = 0; c < COMPS; c++) {
                    const V kv = *reinterpret_cast<const V *>(key + (((((size_t)r * levels + l) * COMPS + c) * limbs + limb) * n + i));
#pragma unroll
                    for (int k = 0; k < VEC; k++) acc[c][k] += (W)d.v[k] * kv.v[k];
                }
            }
        }
#pragma unroll
        for (int c = 0; c < COMPS; c++) {
            V o;
#pragma unroll
            for (int k = 0; k < VEC; k++) o.v[k] = barrett_reduce_wide(br, (T)acc[c][k], (T)(acc[c][k] >> (sizeof(T) * 8)));
            *reinterpret_cast<V *>(out + (((ct * COMPS + c) * limbs + limb) * n + i)) = o;
        }
    }
}

// fused centred lift * scale + accumulate (base.rs:326-386)
template <typename T> struct LiftScaleConsts {
    T q[kMaxLimbs], temp[kMaxLimbs], f[kMaxLimbs], fq[kMaxLimbs];
    T half;
    int limbs, unsigned_mode;
};
template <typename T>
__global__ void __launch_bounds__(256) rns_lift_scaled_acc_kernel(const __grid_constant__ LiftScaleConsts<T> lc, const T *__restrict__ small,
                                                                  T *__restrict__ acc, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const T v = small[i];
        for (int l = 0; l < lc.limbs; l++) {
            const T centred = (lc.unsigned_mode || v < lc.half) ? v : lc.temp[l] + v;
            const size_t o = (size_t)l * count + i;
            acc[o] = mod_add<T>(acc[o], shoup_exact<T>(centred, lc.f[l], lc.fq[l], lc.q[l]), lc.q[l]);  // q may be as large as 2^63 - 1 here
        }
    }
}
// 16 bytes per access (count a multiple of the vector length, 16-byte aligned pointers)
template <typename T>
__global__ void __launch_bounds__(256) rns_lift_scaled_acc_vec_kernel(const __grid_constant__ LiftScaleConsts<T> lc, const T *__restrict__ small,
                                                                      T *__restrict__ acc, size_t count) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V {
        T v[VEC];
    };
    const size_t cv = count / VEC;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cv; i += (size_t)gridDim.x * blockDim.x) {
        V in;
        *reinterpret_cast<uint4 *>(&in) = ldg_stream(reinterpret_cast<const uint4 *>(small) + i);
        for (int l = 0; l < lc.limbs; l++) {
            uint4 *p = reinterpret_cast<uint4 *>(acc + (size_t)l * count) + i;
            V a;
            *reinterpret_cast<uint4 *>(&a) = ldg_stream(p);
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const T centred = (lc.unsigned_mode || in.v[k] < lc.half) ? in.v[k] : lc.temp[l] + in.v[k];
                a.v[k] = mod_add<T>(a.v[k], shoup_exact<T>(centred, lc.f[l], lc.fq[l], lc.q[l]), lc.q[l]);
            }
            stg_stream(p, *reinterpret_cast<const uint4 *>(&a));
        }
    }
}

// p * X^r in Z_q[X]/(X^N+1), r in [0, 2N): rotate right by r mod N, negate the wrapped part, flip all when r >= N.
// polys are [batch][limbs][N]; degrees[batch]
template <typename T>
__global__ void __launch_bounds__(256) mul_monomial_kernel(const __grid_constant__ LimbConsts<T> lc, int limbs, const uint32_t *__restrict__ degrees,
                                                           const T *__restrict__ in, T *__restrict__ out, uint32_t log_n, size_t batch) {
    const uint32_t n = 1u << log_n, mask2 = 2 * n - 1;
    const size_t total = batch * (size_t)limbs * n;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = gid >> log_n, b = poly / limbs;
        const int limb = (int)(poly % limbs);
        const uint32_t i = (uint32_t)(gid & (n - 1));
        const uint32_t r = degrees[b] & mask2;
        const uint32_t srcw = (i - r) & mask2;  // out[i] = sign * in[(i - r) mod 2N]
        const T v = in[(poly << log_n) + (srcw & (n - 1))];
        out[gid] = srcw >= n ? mod_neg<T>(v, lc.br[limb].q) : v;
    }
}

// out[row] = sum_i a[row][i] * b[row][i] mod q ; one CTA per row
template <typename T>
__global__ void __launch_bounds__(256) dot_product_kernel(const Barrett<T> br, const T *__restrict__ a, const T *__restrict__ b,
                                                          T *__restrict__ out, size_t n) {
    using W = typename WideOf<T>::type;
    __shared__ T partial[256];
    const size_t row = blockIdx.x;
    const T *pa = a + row * n, *pb = b + row * n;
    T acc = 0;
    W wide = 0;
    uint32_t terms = 0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        wide += (W)pa[i] * pb[i];
        if (++terms == 16) {
            acc = mod_add<T>(acc, barrett_reduce_wide(br, (T)wide, (T)(wide >> (sizeof(T) * 8))), br.q);
            wide = 0;
            terms = 0;
        }
    }
    acc = mod_add<T>(acc, barrett_reduce_wide(br, (T)wide, (T)(wide >> (sizeof(T) * 8))), br.q);
    partial[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) partial[threadIdx.x] = mod_add<T>(partial[threadIdx.x], partial[threadIdx.x + s], br.q);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[row] = partial[0];
}


// ---- BaseConverter (primus_rns/src/converter.rs) ------------------------------------------------------------------
// One thread per coefficient: adjusted residues y_i = x_i * (Q/q_i)^-1 mod q_i (converter.rs:141-184), then one lazy
// double-word dot product per output modulus (converter.rs:203-212; <= 8 terms, no intermediate reduction needed).
// EXACT (converter.rs:257-365): the floating correction v = trunc(sum_i (double)y_i / (double)q_i + 0.5) is evaluated
// with the reference's operation order (IEEE division, sequential sum), so the result is bit-identical to the CPU path.
template <typename T> __device__ __forceinline__ double word_to_double(T v);
template <> __device__ __forceinline__ double word_to_double<uint32_t>(uint32_t v) { return __uint2double_rn(v); }
template <> __device__ __forceinline__ double word_to_double<uint64_t>(uint64_t v) { return __ull2double_rn(v); }
template <typename T> __device__ __forceinline__ T double_to_word(double v);
template <> __device__ __forceinline__ uint32_t double_to_word<uint32_t>(double v) { return __double2uint_rz(v); }
template <> __device__ __forceinline__ uint64_t double_to_word<uint64_t>(double v) { return __double2ull_rz(v); }

// NIN (input limbs) is a template parameter: the adjusted residues stay in registers.  The coefficient index is split with a shift when the
// polynomial length is a power of two (log_n >= 0), by one 64-bit division otherwise.
// F64 (u64 words, every modulus below 2^50 - 2^10): all modular products run on the FP64 pipe -- x*inv mod q_i and y_i * (Q/q_i mod p_k) mod p_k
// are exact-integer-in-a-double products (6 instructions, |result| <= 0.75 p: the quotient estimate of a product below 2^100 is off by at
// most 0.25 + the rounding to an integer), the <= 8 terms per output sum exactly (6.5 p < 2^53, walked with exact rationals by
// tools/f64_bounds.py::canonical_product_budget) and one fold gives the canonical residue.  The
// integer formulation needs ~28 64-bit multiplies per coefficient of a 3 -> 2 conversion (0.31 of the HBM copy peak, integer-issue bound);
// the FP64 one 86 FP64 instructions (HBM bound).  An input word >= 2^50 (not a canonical residue) takes the integer path for that term.
template <typename T, int NIN, bool EXACT, bool F64>
__global__ void __launch_bounds__(256) baseconv_kernel(const __grid_constant__ BaseConvDev<T> c, const T *__restrict__ in, T *__restrict__ out,
                                                       size_t n, size_t polys, int log_n) {
    using W = typename WideOf<T>::type;
    const size_t total = polys * n;
    const int n_out = EXACT ? 1 : c.n_out;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t p = log_n >= 0 ? gid >> log_n : gid / n, j = log_n >= 0 ? gid & (n - 1) : gid % n;
        const T *x = in + p * (size_t)NIN * n + j;
        T xin[NIN];
#pragma unroll
        for (int i = 0; i < NIN; i++) xin[i] = ldg_stream(x + (size_t)i * n);
        double agg = 0.0;
        if constexpr (F64 && sizeof(T) == 8) {
            using F = F64LazyField;
            T any = 0;
#pragma unroll
            for (int i = 0; i < NIN; i++) any |= xin[i];
            if (any >> 50) {  // a word that is not a canonical residue (rare): reduce it first, (x mod q) * inv == x * inv (mod q)
#pragma unroll
                for (int i = 0; i < NIN; i++) xin[i] = barrett_reduce_wide(c.in_br[i], xin[i], (T)0);
            }
            double y[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                const F::Ctx cx{c.q_f[i], c.in_qinv_f[i], 0.0, 0.0, 0, 0};
                const double v = F::mulmod(F::from_u64(xin[i]), c.inv_f[i], cx, 0);   // == x * inv mod q_i, |v| <= 0.75 q_i
                y[i] = v < 0.0 ? __dadd_rn(v, c.q_f[i]) : v;                         // canonical, as the reference's y_i
                if (EXACT) agg = __dadd_rn(agg, __ddiv_rn(y[i], c.q_f[i]));
            }
            for (int k = 0; k < n_out; k++) {
                const F::Ctx cx{c.out_p_f[k], c.out_pinv_f[k], 0.0, F::kTwo52 + c.out_p_f[k], c.out_br[k].q, 0};
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < NIN; i++) acc = __dadd_rn(acc, F::mulmod(y[i], c.matrix_f[k][i], cx, 0));
                if (EXACT) {
                    const double v = (double)double_to_word<T>(__dadd_rn(agg, 0.5));     // <= NIN
                    acc = __dsub_rn(acc, F::mulmod(v, c.q_mod_p_f[0], cx, 0));
                    stg_stream(out + p * n + j, (T)F::canon(acc, cx));
                } else {
                    stg_stream(out + (p * (size_t)c.n_out + k) * n + j, (T)F::canon(acc, cx));
                }
            }
        } else {
            T y[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                y[i] = c.inv[i] == 1 ? barrett_reduce_wide(c.in_br[i], xin[i], (T)0) : shoup<T>(xin[i], c.inv[i], c.inv_q[i], c.in_br[i].q);
                if (EXACT) agg = __dadd_rn(agg, __ddiv_rn(word_to_double<T>(y[i]), c.q_f[i]));
            }
            for (int k = 0; k < n_out; k++) {
                T r;
                if constexpr (NIN <= 4) {
                    // few terms: lazy Shoup products (1 high + 2 low multiplies each, any word-sized y) kept in [0, 2p) beat the double-word
                    // dot product + two-word Barrett reduction (7 multiplies on its own); p < 2^(BITS-2), so 4p does not overflow
                    const T pk = c.out_br[k].q, two_p = pk + pk;
                    T acc = 0;
#pragma unroll
                    for (int i = 0; i < NIN; i++) acc = csub<T>(acc + shoup_lazy<T>(y[i], c.matrix[k][i], c.matrix_q[k][i], pk), two_p);
                    r = csub<T>(acc, pk);
                } else {
                    W acc = 0;
#pragma unroll
                    for (int i = 0; i < NIN; i++) acc += (W)y[i] * c.matrix[k][i];
                    r = barrett_reduce_wide(c.out_br[k], (T)acc, (T)(acc >> (sizeof(T) * 8)));
                }
                if (EXACT) {
                    const T v = double_to_word<T>(__dadd_rn(agg, 0.5));
                    r = mod_sub<T>(r, barrett_mul<T>(c.out_br[0], v, c.q_mod_p[0]), c.out_br[0].q);
                    stg_stream(out + p * n + j, r);
                } else {
                    stg_stream(out + (p * (size_t)c.n_out + k) * n + j, r);
                }
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
template <typename T> static void hbig_mul_word(std::vector<T> &a, T v) {
    using W = typename host::Wide<T>::type;
    T carry = 0;
    for (auto &w : a) {
        const W s = (W)w * v + carry;
        w = (T)s;
        carry = (T)(s >> host::Wide<T>::BITS);
    }
    if (carry) a.push_back(carry);
}
template <typename T> static T hbig_mod_word(const std::vector<T> &a, T q) {
    using W = typename host::Wide<T>::type;
    W r = 0;
    for (size_t i = a.size(); i-- > 0;) r = ((r << host::Wide<T>::BITS) | a[i]) % q;
    return (T)r;
}
template <typename T> static int hbig_bits(const std::vector<T> &a) {
    for (size_t i = a.size(); i-- > 0;)
        if (a[i]) return (int)(i * host::Wide<T>::BITS) + host::bit_length<T>(a[i]);
    return 0;
}
template <typename T> static void hbig_shl(std::vector<T> &a, unsigned s) {
    constexpr int B = host::Wide<T>::BITS;
    const size_t len = a.size();
    while (s >= (unsigned)B) {
        for (size_t i = len; i-- > 1;) a[i] = a[i - 1];
        a[0] = 0;
        s -= B;
    }
    if (!s) return;
    for (size_t i = len; i-- > 0;) a[i] = (T)((a[i] << s) | (i ? (a[i - 1] >> (B - s)) : 0));
}
template <typename T> static bool hbig_lt(const std::vector<T> &a, const std::vector<T> &b) {
    for (size_t i = a.size(); i-- > 0;)
        if (a[i] != b[i]) return a[i] < b[i];
    return false;
}
template <typename T> static T hgcd(T a, T b) {
    while (b) {
        const T t = a % b;
        a = b;
        b = t;
    }
    return a;
}
template <typename T> static T hmodinv(T a, T m) {
    using W = typename host::Wide<T>::type;
    using S = typename std::conditional<sizeof(T) == 8, __int128, int64_t>::type;
    S t = 0, nt = 1;
    W r = m, nr = a % m;
    while (nr) {
        const W qd = r / nr;
        const S tt = t - (S)qd * nt;
        t = nt;
        nt = tt;
        const W tr = r - qd * nr;
        r = nr;
        nr = tr;
    }
    if (r != 1) return 0;
    if (t < 0) t += (S)m;
    return (T)t;
}

// RNSBase::new + BigUintApproxSignedBasis::new (base.rs:79-122, big_integer/basis.rs:40-211)
template <typename T> int make_rns(const T *moduli, size_t limbs, uint32_t log_basis, uint32_t levels_in, RnsDev<T> &r) {
    constexpr int B = host::Wide<T>::BITS;
    memset(&r, 0, sizeof(r));
    if (limbs == 0) return 6;                                   // RNSError::EmptyBase
    if (limbs > (size_t)kRnsMaxLimbs) return 9;
    for (size_t i = 0; i < limbs; i++) {
        if (moduli[i] < 2) return 9;
        for (size_t j = i + 1; j < limbs; j++)
            if (hgcd<T>(moduli[i], moduli[j]) != 1) return 7;   // RNSError::CoPrimeError
    }
    std::vector<T> prod{1};
    for (size_t i = 0; i < limbs; i++) hbig_mul_word<T>(prod, moduli[i]);
    const int bits = hbig_bits<T>(prod);
    const int value_len = (bits + B - 1) / B;
    if (value_len > kRnsMaxWords) return 9;
    prod.resize(value_len, 0);
    r.limbs = (int)limbs;
    r.value_len = value_len;
    for (int k = 0; k < value_len; k++) r.product[k] = prod[k];
    for (size_t i = 0; i < limbs; i++) {
        r.q[i] = moduli[i];
        std::vector<T> p{1};
        for (size_t j = 0; j < limbs; j++)
            if (j != i) hbig_mul_word<T>(p, moduli[j]);
        p.resize(value_len, 0);
        for (int k = 0; k < value_len; k++) r.punct[i][k] = p[k];
        const T inv = hmodinv<T>(hbig_mod_word<T>(p, moduli[i]), moduli[i]);
        if (inv == 0 && moduli[i] != 1) return 7;
        r.inv_punct[i] = inv;
        r.inv_punct_q[i] = host::shoup_quot<T>(inv, moduli[i]);
        // 2^(BITS k) mod q_i and its Shoup quotient (rns_decompose_kernel)
        std::vector<T> pw{1};
        for (int k = 0; k < value_len; k++) {
            r.pw[i][k] = moduli[i] == 1 ? 0 : hbig_mod_word<T>(pw, moduli[i]);
            r.pw_q[i][k] = host::shoup_quot<T>(r.pw[i][k], moduli[i]);
            pw.insert(pw.begin(), (T)0);
        }
    }
    if (log_basis == 0) return 0;  // RNS base only (compose / decompose)
    if ((int)log_basis >= B) return 9;
    uint32_t levels = (uint32_t)bits / log_basis, drop = (uint32_t)bits - levels * log_basis;
    if (levels_in) {
        if (levels < levels_in) return 9;
        levels = levels_in;
        drop = (uint32_t)bits - levels * log_basis;
    }
    if (levels == 0) return 9;
    r.log_basis = log_basis;
    r.levels = levels;
    r.drop_bits = drop;
    r.basis_m1 = (T)(((T)1 << log_basis) - 1);
    r.carry_mask = log_basis == 1 ? (T)2 : (T)(((T)1 << log_basis) | ((T)1 << (log_basis - 1)));
    r.has_init_mask = drop > 0;
    if (drop > 0) {
        r.init_index = (int)((drop - 1) / B);
        r.init_mask = (T)1 << ((drop - 1) % B);
    }
    std::vector<T> value(value_len, 0);
    bool have = false;
    if (log_basis == 1) {
        if (drop != 0) {
            for (uint32_t i = 0; i < levels; i++) {
                hbig_shl<T>(value, 1);
                value[0] |= 1;
            }
            hbig_shl<T>(value, 1);
            value[0] |= 1;
            hbig_shl<T>(value, drop - 1);
            have = hbig_lt<T>(value, prod);
        }
    } else {
        for (uint32_t i = 0; i < levels; i++) {
            hbig_shl<T>(value, log_basis);
            value[0] |= (T)(r.basis_m1 >> 1);
        }
        if (drop > 0) {
            hbig_shl<T>(value, 1);
            value[0] |= 1;
            hbig_shl<T>(value, drop - 1);
        } else {
            T one = 1;
            for (int i = 0; i < value_len && one; i++) {
                value[i] = (T)(value[i] + one);
                one = value[i] == 0;
            }
        }
        have = hbig_lt<T>(value, prod);
    }
    r.has_threshold = have;
    for (int k = 0; k < value_len; k++) r.threshold[k] = value[k];
    // add = (2^bits - 1) - (Q - 1)
    std::vector<T> add(value_len, (T)~(T)0);
    const int unused = value_len * B - bits;
    if (unused) add[value_len - 1] = (T)(add[value_len - 1] >> unused);
    std::vector<T> qm1(prod);
    {
        T one = 1;
        for (int i = 0; i < value_len && one; i++) {
            const T o = qm1[i];
            qm1[i] = (T)(o - one);
            one = o == 0;
        }
    }
    T borrow = 0;
    for (int i = 0; i < value_len; i++) {
        const T ai = add[i], bi = qm1[i];
        add[i] = (T)(ai - bi - borrow);
        borrow = (ai < bi) || (ai == bi && borrow) ? 1 : 0;
    }
    for (int k = 0; k < value_len; k++) r.add[k] = add[k];
    return 0;
}
template int make_rns<uint32_t>(const uint32_t *, size_t, uint32_t, uint32_t, RnsDev<uint32_t> &);
template int make_rns<uint64_t>(const uint64_t *, size_t, uint32_t, uint32_t, RnsDev<uint64_t> &);


template <typename T> int make_baseconv(const T *in_moduli, size_t n_in, const T *out_moduli, size_t n_out, BaseConvDev<T> &c) {
    constexpr int B = host::Wide<T>::BITS;
    memset(&c, 0, sizeof(c));
    RnsDev<T> in, outb;
    int rc = make_rns<T>(in_moduli, n_in, 0, 0, in);
    if (rc) return rc;
    if ((rc = make_rns<T>(out_moduli, n_out, 0, 0, outb)) != 0) return rc;  // the output base is an RNSBase too
    c.n_in = (int)n_in;
    c.n_out = (int)n_out;
    for (size_t i = 0; i < n_in; i++) {
        if ((in_moduli[i] >> (B - 2)) != 0) return 5;  // BarrettModulus range
        c.inv[i] = in.inv_punct[i];
        c.inv_q[i] = in.inv_punct_q[i];
        c.in_br[i].q = in_moduli[i];
        host::barrett_ratio<T>(in_moduli[i], c.in_br[i].r0, c.in_br[i].r1);
        c.q_f[i] = (double)in_moduli[i];
        c.in_qinv_f[i] = 1.0 / (double)in_moduli[i];
        c.inv_f[i] = (double)c.inv[i];
    }
    constexpr uint64_t kF64Max = ((uint64_t)1 << 50) - 1024;  // the exactness budget of the FP64 products (ntt_core.cuh, tools/f64_bounds.py)
    bool f64 = B == 64 && !(getenv("PFHE_DISABLE_F64") && getenv("PFHE_DISABLE_F64")[0] == '1');
    for (size_t i = 0; i < n_in; i++) f64 = f64 && (uint64_t)in_moduli[i] <= kF64Max;
    for (size_t k = 0; k < n_out; k++) f64 = f64 && (uint64_t)out_moduli[k] <= kF64Max;
    c.f64_ok = f64 ? 1 : 0;
    std::vector<T> prod(in.product, in.product + in.value_len);
    for (size_t k = 0; k < n_out; k++) {
        if ((out_moduli[k] >> (B - 2)) != 0) return 5;
        c.out_br[k].q = out_moduli[k];
        host::barrett_ratio<T>(out_moduli[k], c.out_br[k].r0, c.out_br[k].r1);
        for (size_t i = 0; i < n_in; i++) {
            std::vector<T> p(in.punct[i], in.punct[i] + in.value_len);
            c.matrix[k][i] = hbig_mod_word<T>(p, out_moduli[k]);
            c.matrix_q[k][i] = host::shoup_quot<T>(c.matrix[k][i], out_moduli[k]);
        }
        c.q_mod_p[k] = hbig_mod_word<T>(prod, out_moduli[k]);
        c.out_p_f[k] = (double)out_moduli[k];
        c.out_pinv_f[k] = 1.0 / (double)out_moduli[k];
        c.q_mod_p_f[k] = (double)c.q_mod_p[k];
        for (size_t i = 0; i < n_in; i++) c.matrix_f[k][i] = (double)c.matrix[k][i];
    }
    return 0;
}
template int make_baseconv<uint32_t>(const uint32_t *, size_t, const uint32_t *, size_t, BaseConvDev<uint32_t> &);
template int make_baseconv<uint64_t>(const uint64_t *, size_t, const uint64_t *, size_t, BaseConvDev<uint64_t> &);

#define PFHE_LIMB_SWITCH(L_, CALL) \
    switch (L_) {                  \
        case 1: CALL(1); break;    \
        case 2: CALL(2); break;    \
        case 3: CALL(3); break;    \
        case 4: CALL(4); break;    \
        case 5: CALL(5); break;    \
        case 6: CALL(6); break;    \
        case 7: CALL(7); break;    \
        case 8: CALL(8); break;    \
        default: return cudaErrorNotSupported; \
    }

template <typename T, int NIN>
static void run_baseconv(const BaseConvDev<T> &c, const T *in, T *out, size_t n, size_t polys, int log_n, bool exact, unsigned grid, cudaStream_t s) {
    if constexpr (sizeof(T) == 8) {
        if (c.f64_ok) {
            if (exact) baseconv_kernel<T, NIN, true, true><<<grid, 256, 0, s>>>(c, in, out, n, polys, log_n);
            else baseconv_kernel<T, NIN, false, true><<<grid, 256, 0, s>>>(c, in, out, n, polys, log_n);
            return;
        }
    }
    if (exact) baseconv_kernel<T, NIN, true, false><<<grid, 256, 0, s>>>(c, in, out, n, polys, log_n);
    else baseconv_kernel<T, NIN, false, false><<<grid, 256, 0, s>>>(c, in, out, n, polys, log_n);
}
template <typename T>
cudaError_t launch_baseconv(const BaseConvDev<T> &c, const T *in, T *out, size_t n, size_t polys, bool exact, cudaStream_t s) {
    if (!n || !polys) return cudaSuccess;
    int log_n = -1;
    if ((n & (n - 1)) == 0) {
        log_n = 0;
        while (((size_t)1 << log_n) < n) log_n++;
    }
    const unsigned grid = grid_for(n * polys, 256);
#define PFHE_BC_CALL(NIN) run_baseconv<T, NIN>(c, in, out, n, polys, log_n, exact, grid, s)
    PFHE_LIMB_SWITCH(c.n_in, PFHE_BC_CALL)
#undef PFHE_BC_CALL
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_baseconv<uint32_t>(const BaseConvDev<uint32_t> &, const uint32_t *, uint32_t *, size_t, size_t, bool, cudaStream_t);
template cudaError_t launch_baseconv<uint64_t>(const BaseConvDev<uint64_t> &, const uint64_t *, uint64_t *, size_t, size_t, bool, cudaStream_t);

template <typename T> cudaError_t launch_rns_compose(const RnsDev<T> &r, const T *residues, T *big, size_t count, cudaStream_t s) {
    if (!count) return cudaSuccess;
#define PFHE_RC_CALL(L) rns_compose_kernel<T, L><<<grid_for(count, 256), 256, 0, s>>>(r, residues, big, count)
    PFHE_LIMB_SWITCH(r.limbs, PFHE_RC_CALL)
#undef PFHE_RC_CALL
    count_launch();
    return cudaGetLastError();
}
template <typename T> cudaError_t launch_rns_decompose(const RnsDev<T> &r, const T *big, T *residues, size_t count, cudaStream_t s) {
    if (!count) return cudaSuccess;
#define PFHE_RD_CALL(L) rns_decompose_kernel<T, L><<<grid_for(count, 256), 256, 0, s>>>(r, big, residues, count)
    PFHE_LIMB_SWITCH(r.limbs, PFHE_RD_CALL)
#undef PFHE_RD_CALL
    count_launch();
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_rns_gadget(const RnsDev<T> &r, const T *residues, T *digits, size_t count, size_t polys, size_t in_stride, size_t out_stride,
                              cudaStream_t s) {
    if (!count || !polys) return cudaSuccess;
    if (count % (16 / sizeof(T)) || (in_stride % (16 / sizeof(T))) || (out_stride % (16 / sizeof(T))) || (reinterpret_cast<uintptr_t>(residues) & 15) ||
        (reinterpret_cast<uintptr_t>(digits) & 15))
        return cudaErrorNotSupported;  // polynomial lengths on this path are powers of two >= 4
    const size_t cv = count / ((r.limbs <= 4 ? 16 : 8) / sizeof(T));
    unsigned gx = (unsigned)((cv + 127) / 128), gy = (unsigned)(polys < 65535 ? polys : 65535);
    const unsigned cap = 148 * 16;  // grid-stride beyond that
    if ((size_t)gx * gy > cap) {
        if (gx > cap) gx = cap;
        gy = cap / gx ? (gy < cap / gx ? gy : cap / gx) : 1;
    }
    const dim3 grid(gx, gy);
#define PFHE_RG_CALL(L) rns_gadget_kernel<T, L><<<grid, 128, 0, s>>>(r, residues, digits, count, polys, in_stride, out_stride)
    PFHE_LIMB_SWITCH(r.limbs, PFHE_RG_CALL)
#undef PFHE_RG_CALL
    count_launch();
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_rns_key_mac(const LimbConsts<T> &lc, int limbs, int comps, uint32_t levels, const T *digits, const T *key, T *out, size_t n,
                               size_t batch, cudaStream_t s) {
    if (!batch) return cudaSuccess;
    const size_t threads_total = batch * limbs * (n / (16 / sizeof(T)));
    if (n % (16 / sizeof(T))) return cudaErrorNotSupported;
    if (comps == 2)
        rns_key_mac_kernel<T, 2><<<grid_for(threads_total, 256), 256, 0, s>>>(lc, limbs, levels, digits, key, out, n, batch);
    else if (comps == 3)
        rns_key_mac_kernel<T, 3><<<grid_for(threads_total, 256), 256, 0, s>>>(lc, limbs, levels, digits, key, out, n, batch);
    else
        return cudaErrorNotSupported;
    count_launch();
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_rns_lift_scaled_acc(const T *moduli, int limbs, T small_modulus, const T *scalars, const T *small, T *acc, size_t count,
                                       cudaStream_t s) {
    if (!count) return cudaSuccess;
    LiftScaleConsts<T> lc;
    lc.limbs = limbs;
    lc.unsigned_mode = small_modulus == 2;
    lc.half = (T)((small_modulus + 1) / 2);
    for (int l = 0; l < limbs; l++) {
        lc.q[l] = moduli[l];
        lc.temp[l] = moduli[l] - small_modulus;
        lc.f[l] = scalars[l];
        lc.fq[l] = host::shoup_quot<T>(scalars[l], moduli[l]);
    }
    constexpr size_t VEC = 16 / sizeof(T);
    if (count % VEC == 0 && ((reinterpret_cast<uintptr_t>(small) | reinterpret_cast<uintptr_t>(acc)) & 15) == 0)
        rns_lift_scaled_acc_vec_kernel<T><<<grid_for(count / VEC, 256), 256, 0, s>>>(lc, small, acc, count);
    else
        rns_lift_scaled_acc_kernel<T><<<grid_for(count, 256), 256, 0, s>>>(lc, small, acc, count);
    count_launch();
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_mul_monomial(const LimbConsts<T> &lc, int limbs, const uint32_t *degrees, const T *in, T *out, uint32_t log_n, size_t batch,
                                cudaStream_t s) {
    if (!batch) return cudaSuccess;
    mul_monomial_kernel<T><<<grid_for((batch * limbs) << log_n, 256), 256, 0, s>>>(lc, limbs, degrees, in, out, log_n, batch);
    count_launch();
    return cudaGetLastError();
}
template <typename T> cudaError_t launch_dot_product(const Barrett<T> &br, const T *a, const T *b, T *out, size_t rows, size_t n, cudaStream_t s) {
    if (!rows) return cudaSuccess;
    dot_product_kernel<T><<<(unsigned)rows, 256, 0, s>>>(br, a, b, out, n);
    count_launch();
    return cudaGetLastError();
}

#define PFHE_INST(T)                                                                                                                         \
    template cudaError_t launch_rns_compose<T>(const RnsDev<T> &, const T *, T *, size_t, cudaStream_t);                                     \
    template cudaError_t launch_rns_decompose<T>(const RnsDev<T> &, const T *, T *, size_t, cudaStream_t);                                   \
    template cudaError_t launch_rns_gadget<T>(const RnsDev<T> &, const T *, T *, size_t, size_t, size_t, size_t, cudaStream_t);              \
    template cudaError_t launch_rns_key_mac<T>(const LimbConsts<T> &, int, int, uint32_t, const T *, const T *, T *, size_t, size_t,        \
                                               cudaStream_t);                                                                                \
    template cudaError_t launch_rns_lift_scaled_acc<T>(const T *, int, T, const T *, const T *, T *, size_t, cudaStream_t);                  \
    template cudaError_t launch_mul_monomial<T>(const LimbConsts<T> &, int, const uint32_t *, const T *, T *, uint32_t, size_t, cudaStream_t); \
    template cudaError_t launch_dot_product<T>(const Barrett<T> &, const T *, const T *, T *, size_t, size_t, cudaStream_t);
PFHE_INST(uint32_t)
PFHE_INST(uint64_t)

}  // namespace pfhe
