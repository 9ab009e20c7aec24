// lattice_dcrt.cu -- the multi-limb (L > 1) GGSW external product as ONE kernel per (ciphertext, limb): compose, multi-word gadget digits,
// forward transform, key multiply-accumulate and inverse transform without an HBM round trip of the digits.
// Restates CrtGlwe::mul_dcrt_ggsw_to -> add_dcrt_glev_mul_crt_poly_assign (primus_lattice/src/glwe/crt.rs:200-227, glwe/dcrt.rs:178-255).
#include "lattice_core.cuh"
#include "rns.hpp"

namespace pfhe {

// ---- multi-limb external product in ONE kernel (composed values of at most two words by default, up to four words opt-in) -------------
// The digits of CrtGlwe::mul_dcrt_ggsw_to couple all limbs of a coefficient (compose -> multi-word gadget,
// primus_lattice/src/glwe/dcrt.rs:219-236), which is why round 1 wrote them to HBM first (21 % of the product's time, profiles/
// r02_large_n_experiments.md).  For Q below two words the coupling is cheap enough to repeat per limb: the CTA of (ciphertext, limb)
// composes its 8 coefficients per thread itself (base.rs:609-636), adds the carry-free digit offset
//     R = 2^(drop-1) + sum_l (B/2) 2^(drop + l*beta)
// once per input component -- the balanced digits of init_value_carry_slice_inplace + unsigned_decompose_slice_to + the centred lift
// (big_integer/basis.rs:326-367, big_integer/common.rs:275-325, base.rs:279-315) are window_l(value + R) - B/2 -- and then runs the
// same transform / multiply-accumulate / inverse as dcrt_external_product_kernel with the digits produced in registers.
constexpr int kFusedMaxWords = 4;  // composed values of up to four words (five 50-bit limbs, four 31-bit limbs)
template <typename T> struct BigGadget {
    T q[kRnsMaxLimbs];
    T product[kFusedMaxWords], punct[kRnsMaxLimbs][kFusedMaxWords];
    T inv_punct[kRnsMaxLimbs], inv_punct_q[kRnsMaxLimbs];
    T threshold[kFusedMaxWords], add[kFusedMaxWords], offset[kFusedMaxWords];  // offset = R
    T mask, half;
    uint32_t drop_bits, log_basis, levels;
    int limbs, has_threshold;
};
template <typename T> struct TwoWords;
template <> struct TwoWords<uint32_t> { using U = uint64_t; };
template <> struct TwoWords<uint64_t> { using U = unsigned __int128; };

// VLEN <= 2: the composed value is one double word (u64 / unsigned __int128 arithmetic).  VLEN = 3, 4: VLEN registers per coefficient, shifted
// down by the basis per level so that the digit window is always the low bits of word 0; no register cap (the state is 8 x VLEN words per thread).
template <typename T, int VLEN> struct WideRegs {
    using W = typename TwoWords<T>::U;
    static constexpr int BITS = sizeof(T) * 8;
    __device__ __forceinline__ static T sub(const T (&a)[VLEN], const T *b, T (&d)[VLEN]) {  // d = a - b, returns the borrow
        T borrow = 0;
#pragma unroll
        for (int k = 0; k < VLEN; k++) {
            const T t = a[k] - b[k];
            const T nb = (T)((a[k] < b[k]) | (t < borrow));
            d[k] = t - borrow;
            borrow = nb;
        }
        return borrow;
    }
    __device__ __forceinline__ static void add(T (&a)[VLEN], const T *b) {
        T carry = 0;
#pragma unroll
        for (int k = 0; k < VLEN; k++) {
            const T s = a[k] + b[k], s2 = s + carry;
            carry = (T)((s < a[k]) | (s2 < s));
            a[k] = s2;
        }
    }
    __device__ __forceinline__ static void shr_word(T (&a)[VLEN]) {
#pragma unroll
        for (int k = 0; k + 1 < VLEN; k++) a[k] = a[k + 1];
        a[VLEN - 1] = 0;
    }
    __device__ __forceinline__ static void shr_bits(T (&a)[VLEN], uint32_t b) {  // 0 < b < BITS
#pragma unroll
        for (int k = 0; k + 1 < VLEN; k++) a[k] = (T)((a[k] >> b) | (a[k + 1] << (BITS - b)));
        a[VLEN - 1] >>= b;
    }
};

template <typename F, int LOGN, int LOGE, int COMPS, int VLEN>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)), ep_min_blocks<LOGN, LOGE, 1>())
dcrt_external_product_fused_kernel(const DevNtt<typename F::WordT> *__restrict__ tables, const __grid_constant__ BigGadget<typename F::WordT> bg,
                                   const typename F::WordT *__restrict__ key, const typename F::WordT *__restrict__ in,
                                   typename F::WordT *__restrict__ out, int to_coeff) {
    using EP = ExtProd<F, LOGN, LOGE, COMPS>;
    using Core = typename EP::Core;
    using T = typename F::WordT;
    using U = typename TwoWords<T>::U;
    using Elem = typename F::Elem;
    using LA = LatAcc<F>;
    constexpr int N = EP::N, E = EP::E, CW = EP::CW, NV = EP::NV, FB0 = EP::FB0, BITS = sizeof(T) * 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Elem *sm = reinterpret_cast<Elem *>(smem_raw);
    const int t = threadIdx.x, limbs = bg.limbs;
    const size_t ct = blockIdx.x / (unsigned)limbs;
    const int limb = (int)(blockIdx.x % (unsigned)limbs);
    const DevNtt<T> tb = tables[limb];
    const typename F::Ctx cx = F::ctx(tb);
    const T q = tb.q, mask = bg.mask, half = bg.half;
    const uint32_t levels = bg.levels;
    LSyncBlock sync;
    typename EP::Acc acc[COMPS][E];
#pragma unroll
    for (int c = 0; c < COMPS; c++)
#pragma unroll
        for (int j = 0; j < E; j++) LA::zero(acc[c][j]);
    uint32_t terms = 0;
    // transform of one level's digits + key multiply-accumulate (x holds the canonical digits on entry)
    auto level_mac = [&](Elem (&x)[E], int r, uint32_t l) {
        Core::template fwd_from<0, true>(x, sm, tb, cx, t, sync);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = LA::prepare(x[j], cx);
        if (LA::kRenorm) {
            if (terms == LA::kRenormEvery) {
#pragma unroll
                for (int c = 0; c < COMPS; c++)
#pragma unroll
                    for (int j = 0; j < E; j++) LA::renorm(acc[c][j], cx);
                terms = 1;
            }
            terms++;
        }
#pragma unroll
        for (int c = 0; c < COMPS; c++) {
            const T *kp = key + (((((size_t)r * levels + l) * COMPS + c) * limbs + limb) * (size_t)N) + (size_t)t * E;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const typename Core::WVec kv = ldg_vec(reinterpret_cast<const typename Core::WVec *>(kp) + v);
#pragma unroll
                for (int k = 0; k < CW; k++) LA::mac(acc[c][v * CW + k], x[v * CW + k], kv.v[k], cx);
            }
        }
    };
    if constexpr (VLEN <= 2) {
        const U bigq = VLEN == 2 ? (((U)bg.product[1] << BITS) | bg.product[0]) : (U)bg.product[0];
        const U thr = VLEN == 2 ? (((U)bg.threshold[1] << BITS) | bg.threshold[0]) : (U)bg.threshold[0];
        const U addv = VLEN == 2 ? (((U)bg.add[1] << BITS) | bg.add[0]) : (U)bg.add[0];
        const U offs = VLEN == 2 ? (((U)bg.offset[1] << BITS) | bg.offset[0]) : (U)bg.offset[0];
#pragma unroll 1
        for (int r = 0; r < COMPS; r++) {
            // composed coefficient + threshold adjustment + digit offset, 8 per thread
            U w[E];
            const T *cin = in + ((ct * COMPS + r) * limbs) * (size_t)N;
#pragma unroll
            for (int j = 0; j < E; j++) {
                const int idx = Core::elem_index(FB0, t, j);
                U v = 0;
                for (int i = 0; i < limbs; i++) {
                    const T prod = shoup<T>(__ldg(cin + (size_t)i * N + idx), bg.inv_punct[i], bg.inv_punct_q[i], bg.q[i]);
                    U term = (U)bg.punct[i][0] * prod;                        // (Q / q_i) * prod < Q
                    if (VLEN == 2) term += (U)(T)(bg.punct[i][1] * prod) << BITS;
                    const U s = v + term;
                    v = (s < v || s >= bigq) ? s - bigq : s;                   // one subtraction: both operands are below Q
                }
                if (bg.has_threshold && v >= thr) v += addv;
                w[j] = v + offs;  // a carry out of the top word is beyond every digit window
            }
#pragma unroll 1
            for (uint32_t l = 0; l < levels; l++) {
                const uint32_t pos = bg.drop_bits + l * bg.log_basis;
                Elem x[E];
#pragma unroll
                for (int j = 0; j < E; j++) {
                    const T win = (T)(w[j] >> pos) & mask;
                    const T d = win >= half ? win - half : win + (q - half);   // balanced digit window - B/2, canonical mod q_limb
                    x[j] = F::load(d, cx);
                }
                level_mac(x, r, l);
            }
        }
    } else {
        using WR = WideRegs<T, VLEN>;
        using W2 = typename TwoWords<T>::U;
#pragma unroll 1
        for (int r = 0; r < COMPS; r++) {
            T w[E][VLEN];
            const T *cin = in + ((ct * COMPS + r) * limbs) * (size_t)N;
#pragma unroll
            for (int j = 0; j < E; j++) {
                const int idx = Core::elem_index(FB0, t, j);
#pragma unroll
                for (int k = 0; k < VLEN; k++) w[j][k] = 0;
                for (int i = 0; i < limbs; i++) {   // compose_to (base.rs:609-636), one conditional subtraction per term
                    const T prod = shoup<T>(__ldg(cin + (size_t)i * N + idx), bg.inv_punct[i], bg.inv_punct_q[i], bg.q[i]);
                    T carry = 0;
#pragma unroll
                    for (int k = 0; k < VLEN; k++) {
                        const W2 s = (W2)bg.punct[i][k] * prod + w[j][k] + carry;
                        w[j][k] = (T)s;
                        carry = (T)(s >> BITS);
                    }
                    T d[VLEN];
                    const T borrow = WR::sub(w[j], bg.product, d);
                    const bool ge = (carry != 0) | (borrow == 0);
#pragma unroll
                    for (int k = 0; k < VLEN; k++) w[j][k] = ge ? d[k] : w[j][k];
                }
                if (bg.has_threshold) {
                    T d[VLEN];
                    if (WR::sub(w[j], bg.threshold, d) == 0) WR::add(w[j], bg.add);
                }
                WR::add(w[j], bg.offset);  // + R; a carry out of the top word is beyond every digit window
                for (uint32_t sft = 0; sft < bg.drop_bits / BITS; sft++) WR::shr_word(w[j]);
                if (bg.drop_bits % BITS) WR::shr_bits(w[j], bg.drop_bits % BITS);
            }
#pragma unroll 1
            for (uint32_t l = 0; l < levels; l++) {
                Elem x[E];
#pragma unroll
                for (int j = 0; j < E; j++) {
                    const T win = w[j][0] & mask;
                    const T d = win >= half ? win - half : win + (q - half);
                    x[j] = F::load(d, cx);
                    WR::shr_bits(w[j], bg.log_basis);
                }
                level_mac(x, r, l);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < COMPS; c++) {
        T *o = out + (((ct * COMPS + c) * limbs + limb) * (size_t)N);
        Elem x[E];
        if (to_coeff) {
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = F::from_mac(LA::final(acc[c][j], cx), cx);
            Core::template inv_from<Core::P::NPASS - 1>(x, sm, tb, cx, t, sync);
            Core::inv_regs_to_global(x, o, cx, t);
            sync();
        } else {
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = F::mac_bits(LA::final(acc[c][j], cx), cx);
            Core::template sm_store<Core::P::NPASS - 1>(x, sm, t);
            sync();
            Core::copy_s2g(sm, o, t);
            sync();
        }
    }
}

template <typename F, int LOGN, int COMPS, int VLEN>
static cudaError_t run_dcrt_ep_fused_f(const DevNtt<typename F::WordT> *tables, const BigGadget<typename F::WordT> &bg, const typename F::WordT *key,
                                       const typename F::WordT *in, typename F::WordT *out, size_t batch, bool to_coeff, cudaStream_t stream) {
    using T = typename F::WordT;
    constexpr int threads = 1 << (LOGN - 3);
    constexpr size_t smem = sizeof(T) * ((size_t)1 << LOGN);
    auto k = dcrt_external_product_fused_kernel<F, LOGN, 3, COMPS, VLEN>;
    cudaError_t e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)(batch * bg.limbs), threads, smem, stream>>>(tables, bg, key, in, out, to_coeff ? 1 : 0);
    count_launch();
    return cudaGetLastError();
}
template <typename T, int LOGN, int COMPS, int VLEN>
static cudaError_t run_dcrt_ep_fused(int policy, const DevNtt<T> *tables, const BigGadget<T> &bg, const T *key, const T *in, T *out, size_t batch,
                                     bool to_coeff, cudaStream_t stream) {
    if constexpr (sizeof(T) == 8) {
        if (policy == 1) return run_dcrt_ep_fused_f<F64LazyField, LOGN, COMPS, VLEN>(tables, bg, key, in, out, batch, to_coeff, stream);
    } else {
        if (policy == 2) return run_dcrt_ep_fused_f<IntWide32Field, LOGN, COMPS, VLEN>(tables, bg, key, in, out, batch, to_coeff, stream);
    }
    return run_dcrt_ep_fused_f<IntField<T>, LOGN, COMPS, VLEN>(tables, bg, key, in, out, batch, to_coeff, stream);
}

// cudaErrorNotSupported: composed value longer than two words (four with the opt-in below), k > 2 or a degree without a lattice tile (the
// caller then uses the gadget kernel + per-limb kernel pair)
template <typename T>
cudaError_t launch_dcrt_external_product_fused(int policy, const DevNtt<T> *tables, const RnsDev<T> &r, uint32_t log_n, uint32_t k, const T *key,
                                               const T *in, T *out, size_t batch, bool to_coeff, cudaStream_t s) {
    constexpr int BITS = sizeof(T) * 8;
    if (r.value_len > kFusedMaxWords || r.log_basis == 0 || k < 1 || k > 2 || log_n < 10 || log_n > 12) return cudaErrorNotSupported;
    // Three / four-word values (k = 1, N <= 2048): every (ciphertext, limb) CTA repeats the multi-word compose and digit shifts, which costs what
    // the digit round trip saves -- 433 K against 436 K products/s at L = 3 and 222 K against 279 K at L = 4 with the gadget kernel + per-limb
    // kernel pair (N = 2048, base 2^7; profiles/r02_dcrt_external_product_fused_ab.log).  Opt-in: PFHE_DCRT_EP_FUSED_WIDE=1.
    static const bool wide = getenv("PFHE_DCRT_EP_FUSED_WIDE") && getenv("PFHE_DCRT_EP_FUSED_WIDE")[0] == '1';
    if (r.value_len > 2 && (k != 1 || log_n > 11 || !wide)) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    BigGadget<T> bg{};
    bg.limbs = r.limbs;
    for (int i = 0; i < r.limbs; i++) {
        bg.q[i] = r.q[i];
        for (int w = 0; w < kFusedMaxWords; w++) bg.punct[i][w] = w < r.value_len ? r.punct[i][w] : 0;
        bg.inv_punct[i] = r.inv_punct[i];
        bg.inv_punct_q[i] = r.inv_punct_q[i];
    }
    for (int w = 0; w < kFusedMaxWords; w++) {
        bg.product[w] = w < r.value_len ? r.product[w] : 0;
        bg.threshold[w] = w < r.value_len ? r.threshold[w] : 0;
        bg.add[w] = w < r.value_len ? r.add[w] : 0;
    }
    bg.has_threshold = r.has_threshold;
    bg.mask = r.basis_m1;
    bg.half = r.log_basis == 1 ? 0 : (T)((T)1 << (r.log_basis - 1));
    bg.drop_bits = r.drop_bits;
    bg.log_basis = r.log_basis;
    bg.levels = r.levels;
    // R = 2^(drop-1) + sum_l half << (drop + l*beta), little-endian words (the level windows do not overlap: no carries between the terms)
    auto add_at = [&](T v, uint32_t pos) {
        const uint32_t wi = pos / BITS, sh = pos % BITS;
        T carry = 0;
        for (uint32_t w = wi, part = 0; w < (uint32_t)kFusedMaxWords; w++, part++) {
            const T piece = part == 0 ? (T)(v << sh) : (part == 1 && sh ? (T)(v >> (BITS - sh)) : (T)0);
            const T s0 = bg.offset[w] + piece, s1 = s0 + carry;
            carry = (T)((s0 < piece) | (s1 < s0));
            bg.offset[w] = s1;
        }
    };
    if (r.drop_bits) add_at((T)1, r.drop_bits - 1);
    for (uint32_t l = 0; l < r.levels; l++) add_at(bg.half, r.drop_bits + l * r.log_basis);
#define PFHE_DEPF_CASE(LOGN, WIDE)                                                                                                 \
    case LOGN:                                                                                                                     \
        if (k == 1) {                                                                                                              \
            if constexpr (WIDE) {                                                                                                  \
                if (r.value_len == 3) return run_dcrt_ep_fused<T, LOGN, 2, 3>(policy, tables, bg, key, in, out, batch, to_coeff, s); \
                if (r.value_len == 4) return run_dcrt_ep_fused<T, LOGN, 2, 4>(policy, tables, bg, key, in, out, batch, to_coeff, s); \
            }                                                                                                                      \
            return r.value_len == 2 ? run_dcrt_ep_fused<T, LOGN, 2, 2>(policy, tables, bg, key, in, out, batch, to_coeff, s)       \
                                    : run_dcrt_ep_fused<T, LOGN, 2, 1>(policy, tables, bg, key, in, out, batch, to_coeff, s);      \
        }                                                                                                                          \
        return r.value_len == 2 ? run_dcrt_ep_fused<T, LOGN, 3, 2>(policy, tables, bg, key, in, out, batch, to_coeff, s)           \
                                : run_dcrt_ep_fused<T, LOGN, 3, 1>(policy, tables, bg, key, in, out, batch, to_coeff, s);
    switch (log_n) {
        PFHE_DEPF_CASE(10, true)
        PFHE_DEPF_CASE(11, true)
        PFHE_DEPF_CASE(12, false)
    }
#undef PFHE_DEPF_CASE
    return cudaErrorNotSupported;
}
template cudaError_t launch_dcrt_external_product_fused<uint32_t>(int, const DevNtt<uint32_t> *, const RnsDev<uint32_t> &, uint32_t, uint32_t,
                                                                  const uint32_t *, const uint32_t *, uint32_t *, size_t, bool, cudaStream_t);
template cudaError_t launch_dcrt_external_product_fused<uint64_t>(int, const DevNtt<uint64_t> *, const RnsDev<uint64_t> &, uint32_t, uint32_t,
                                                                  const uint64_t *, const uint64_t *, uint64_t *, size_t, bool, cudaStream_t);

}  // namespace pfhe
