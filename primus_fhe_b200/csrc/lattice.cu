// lattice.cu -- fused GGSW external product and blind rotation (single modulus, L = 1) for sm_100a.
//
// External product (per ciphertext, digits never leave the SM):
//   out_c = [inv]( sum_{r<=k} sum_{l<levels} fwd(digit_l(in_r)) .* key[r][l][c] )
// restating CrtGlwe::mul_dcrt_ggsw_to            primus_lattice/src/glwe/crt.rs:200-227
//   -> add_dcrt_glev_mul_crt_poly_assign          primus_lattice/src/glwe/dcrt.rs:178-255
//   -> add_dcrt_glwe_mul_dcrt_polynomial_assign   primus_lattice/src/glwe/dcrt.rs:108-126
// with, for L = 1, the signed digits of ApproxSignedBasis (primus_decompose/src/primitive/basis.rs:254-283,
// primitive/common.rs:246-259) == unsigned digit + centred lift (big_integer/common.rs:275-285 +
// primus_rns/src/base.rs:721-731).  The key MAC accumulates double-word products lazily and reduces once
// (reduce_dot_product, primus_modulus/src/common/compact/slice.rs:371-401; safe for <= 16 terms since
// q < 2^(BITS-2)).
//
// Blind rotation (composed; not in the reference): the accumulator (2 polynomials)
// stays in shared memory for all n_lwe CMux steps; only the LWE sample, the test vector and the final
// accumulator touch HBM; the bootstrapping key streams through L2.
#include <cstdlib>

#include "internal.hpp"
#include "rns.hpp"

#include "lattice_core.cuh"

namespace pfhe {

template <typename F, int LOGN, int LOGE, int COMPS, int PPB, int KPREF = 0>  // key path: 0 direct loads, 1 cp.async staging, 2 L1 prefetch
__global__ void __launch_bounds__((1 << (LOGN - LOGE)) * PPB, ep_min_blocks<LOGN, LOGE, PPB>())
external_product_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb, const __grid_constant__ GadgetParams<typename F::WordT> g,
                        const typename F::WordT *__restrict__ key, const typename F::WordT *__restrict__ in,
                        typename F::WordT *__restrict__ out, size_t batch, int to_coeff) {
    using EP = ExtProd<F, LOGN, LOGE, COMPS>;
    using Core = typename EP::Core;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    using LA = LatAcc<F>;
    constexpr int N = EP::N, E = EP::E, TPP = EP::TPP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPP, t = threadIdx.x % TPP;
    Elem *sm = reinterpret_cast<Elem *>(smem_raw) + (size_t)grp * N;
    size_t ct = (size_t)blockIdx.x * PPB + grp;
    const bool active = ct < batch;
    if (!active) {
        if (TPP == 32) return;
        ct = batch - 1;
    }
    typename LSyncFor<TPP>::type sync;
    const typename F::Ctx cx = F::ctx(tb);
    const T *cin = in + ct * COMPS * N;
    T *cout = out + ct * COMPS * N;
    typename EP::Acc acc[COMPS][E];
#pragma unroll
    for (int c = 0; c < COMPS; c++)
#pragma unroll
        for (int j = 0; j < E; j++) LA::zero(acc[c][j]);
    if constexpr (KPREF == 2) {
        EP::accumulate([&](int r, int idx) { return __ldg(cin + (size_t)r * N + idx); }, key, g, tb, cx, acc, sm, t, sync, nullptr, 0, 0, true);
    } else if constexpr (KPREF == 1) {  // key staging area behind the exchange buffers of the CTA's groups
        uint4 *kstage = reinterpret_cast<uint4 *>(smem_raw + sizeof(T) * PPB * N);
        EP::accumulate([&](int r, int idx) { return __ldg(cin + (size_t)r * N + idx); }, key, g, tb, cx, acc, sm, t, sync, kstage, TPP * PPB,
                       (int)threadIdx.x);
    } else {
        EP::accumulate([&](int r, int idx) { return __ldg(cin + (size_t)r * N + idx); }, key, g, tb, cx, acc, sm, t, sync);
    }
#pragma unroll
    for (int c = 0; c < COMPS; c++) {
        Elem x[E];
        if (to_coeff) {
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = F::from_mac(LA::final(acc[c][j], cx), cx);
            Core::template inv_from<Core::P::NPASS - 1>(x, sm, tb, cx, t, sync);
            if (active) Core::inv_regs_to_global(x, cout + (size_t)c * N, cx, t);
            sync();
        } else {
            // NTT-domain output: canonical words, bit-reversed order, coalesced through the buffer
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = F::mac_bits(LA::final(acc[c][j], cx), cx);
            if (active) Core::template sm_store<Core::P::NPASS - 1>(x, sm, t);
            sync();
            if (active) Core::copy_s2g(sm, cout + (size_t)c * N, t);
            sync();
        }
    }
}

// Blind rotation: one ciphertext per thread group, accumulator resident in shared memory (canonical words).
template <typename F, int LOGN, int LOGE, int PPB, int MINB>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)) * PPB, MINB)
blind_rotate_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb, const __grid_constant__ GadgetParams<typename F::WordT> g,
                    const typename F::WordT *__restrict__ bsk, uint32_t n_lwe, const uint32_t *__restrict__ lwe,
                    const typename F::WordT *__restrict__ test_vector, typename F::WordT *__restrict__ acc_out, size_t batch) {
    using EP = ExtProd<F, LOGN, LOGE, 2>;
    using Core = typename EP::Core;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    using LA = LatAcc<F>;
    constexpr int N = EP::N, E = EP::E, TPP = EP::TPP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPP, t = threadIdx.x % TPP;
    Elem *sm = reinterpret_cast<Elem *>(smem_raw) + (size_t)grp * 3 * N;  // exchange buffer
    T *accs = reinterpret_cast<T *>(sm + N);                              // ACC: [2][N] canonical words, natural order
    size_t ct = (size_t)blockIdx.x * PPB + grp;
    const bool active = ct < batch;
    if (!active) {
        if (TPP == 32) return;
        ct = batch - 1;
    }
    typename LSyncFor<TPP>::type sync;
    const typename F::Ctx cx = F::ctx(tb);
    const T q = tb.q;
    const uint32_t *my_lwe = lwe + ct * (size_t)(n_lwe + 1);
    const uint32_t two_n_mask = 2 * N - 1;
    // ACC <- (0, tv * X^(2N - b))     (mul_monomial_assign, primus_poly/src/poly/mul.rs:74-99)
    {
        const uint32_t b = __ldg(my_lwe + n_lwe) & two_n_mask;
        const uint32_t rot = (2 * N - b) & two_n_mask;
        for (int i = t; i < N; i += TPP) {
            accs[i] = 0;
            const uint32_t srcw = ((uint32_t)i - rot) & two_n_mask;  // out[i] = sign * tv[(i - rot) mod 2N]
            const T v = __ldg(test_vector + (srcw & (N - 1)));
            accs[N + i] = (srcw >= (uint32_t)N) ? mod_neg<T>(v, q) : v;
        }
    }
    sync();
    const size_t rgsw_len = (size_t)2 * g.levels * 2 * N;
#pragma unroll 1
    for (uint32_t i = 0; i < n_lwe; i++) {
        const uint32_t a = __ldg(my_lwe + i) & two_n_mask;
        typename EP::Acc acc[2][E];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int j = 0; j < E; j++) LA::zero(acc[c][j]);
        // D = ACC * X^a - ACC, evaluated on the fly from the resident accumulator
        auto getD = [&](int r, int idx) -> T {
            const T *p = accs + r * N;
            const uint32_t srcw = ((uint32_t)idx - a) & two_n_mask;
            const T v = p[srcw & (N - 1)];
            const T rotated = (srcw >= (uint32_t)N) ? mod_neg<T>(v, q) : v;
            return mod_sub<T>(rotated, p[idx], q);
        };
        EP::accumulate(getD, bsk + (size_t)i * rgsw_len, g, tb, cx, acc, sm, t, sync);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            Elem x[E];
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = F::from_mac(LA::final(acc[c][j], cx), cx);
            Core::template inv_from<Core::P::NPASS - 1>(x, sm, tb, cx, t, sync);
            // ACC_c += result (each coefficient owned by exactly one thread)
#pragma unroll
            for (int j = 0; j < E; j++) {
                const int idx = Core::elem_index(EP::FB0, t, j);
                accs[c * N + idx] = mod_add<T>(accs[c * N + idx], F::inv_word(x[j], cx), q);
            }
            sync();
        }
    }
    if (active) {
        T *o = acc_out + ct * 2 * N;
        for (int i = t; i < 2 * N; i += TPP) o[i] = accs[i];
    }
}

// ---- dispatch -------------------------------------------------------------------------------------
template <typename F, int LOGN, int LOGE, int COMPS, int PPB>
static cudaError_t run_ep_f(const DevNtt<typename F::WordT> &tb, const GadgetParams<typename F::WordT> &g, const typename F::WordT *key,
                            const typename F::WordT *in, typename F::WordT *out, size_t batch, bool to_coeff, cudaStream_t stream) {
    using T = typename F::WordT;
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    constexpr size_t smem = sizeof(T) * PPB * ((size_t)1 << LOGN);
    cudaError_t e;
    if constexpr (sizeof(T) == 8 && COMPS == 2) {  // u64 words, k = 1: optional key staging through shared memory (cp.async), see ExtProd::accumulate
        // measured slower than the direct key loads (3.64 M against 4.87 M products/s at N = 2048, l = 7: the 64 KiB of staging halves the resident CTAs),
        // so it is opt-in; profiles/r02_large_n_experiments.md
        static const int kpref = getenv("PFHE_EP_KEY_PREFETCH") ? atoi(getenv("PFHE_EP_KEY_PREFETCH")) : 0;
        if (kpref == 2) {
            auto kk = external_product_kernel<F, LOGN, LOGE, COMPS, PPB, 2>;
            if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            kk<<<(unsigned)((batch + PPB - 1) / PPB), threads, smem, stream>>>(tb, g, key, in, out, batch, to_coeff ? 1 : 0);
            count_launch();
            return cudaGetLastError();
        }
        if (kpref == 1) {
            constexpr size_t smem_k = smem + (size_t)2 * COMPS * ((1 << LOGE) * sizeof(T) / 16) * threads * 16;
            auto kk = external_product_kernel<F, LOGN, LOGE, COMPS, PPB, 1>;
            if ((e = cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k)) != cudaSuccess) return e;
            kk<<<(unsigned)((batch + PPB - 1) / PPB), threads, smem_k, stream>>>(tb, g, key, in, out, batch, to_coeff ? 1 : 0);
            count_launch();
            return cudaGetLastError();
        }
    }
    auto k = external_product_kernel<F, LOGN, LOGE, COMPS, PPB>;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)((batch + PPB - 1) / PPB), threads, smem, stream>>>(tb, g, key, in, out, batch, to_coeff ? 1 : 0);
    count_launch();
    return cudaGetLastError();
}
template <typename F, int LOGN, int LOGE, int PPB, int MINB>
static cudaError_t run_br_f(const DevNtt<typename F::WordT> &tb, const GadgetParams<typename F::WordT> &g, const typename F::WordT *bsk,
                            uint32_t n_lwe, const uint32_t *lwe, const typename F::WordT *tv, typename F::WordT *acc_out, size_t batch,
                            cudaStream_t stream) {
    using T = typename F::WordT;
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    constexpr size_t smem = sizeof(T) * PPB * 3 * ((size_t)1 << LOGN);
    auto k = blind_rotate_kernel<F, LOGN, LOGE, PPB, MINB>;
    cudaError_t e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)((batch + PPB - 1) / PPB), threads, smem, stream>>>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch);
    count_launch();
    return cudaGetLastError();
}


// IntWide32Field preconditions: forward values stay below 2^32 and 16 lazy products below 2^64
static bool wide32_ok(uint64_t q, int logn) {
    static const bool off = getenv("PFHE_DISABLE_WIDE32") != nullptr;  // A/B tuning hook
    const uint64_t growth = 2 * (uint64_t)logn + 1;
    if (off || growth * q >= ((uint64_t)1 << 32)) return false;
    const unsigned __int128 worst = (unsigned __int128)16 * growth * q * q;
    return (worst >> 64) == 0;
}

template <typename T, int LOGN, int LOGE, int COMPS, int PPB>
static cudaError_t run_ep(const DevNtt<T> &tb, const GadgetParams<T> &g, const T *key, const T *in, T *out, size_t batch, bool to_coeff,
                          cudaStream_t stream) {
    if constexpr (sizeof(T) == 8) {
        static const bool lazy = !(getenv("PFHE_F64_LAZY") && getenv("PFHE_F64_LAZY")[0] == '0');  // A/B tuning hook
        if (tb.use_f64 && lazy) return run_ep_f<F64LazyField, LOGN, LOGE, COMPS, PPB>(tb, g, key, in, out, batch, to_coeff, stream);
        if (tb.use_f64) return run_ep_f<F64Field, LOGN, LOGE, COMPS, PPB>(tb, g, key, in, out, batch, to_coeff, stream);
    } else {
        if (wide32_ok(tb.q, LOGN)) return run_ep_f<IntWide32Field, LOGN, LOGE, COMPS, PPB>(tb, g, key, in, out, batch, to_coeff, stream);
    }
    return run_ep_f<IntField<T>, LOGN, LOGE, COMPS, PPB>(tb, g, key, in, out, batch, to_coeff, stream);
}
template <typename T, int LOGN, int LOGE, int PPB, int MINB = 1>
static cudaError_t run_br(const DevNtt<T> &tb, const GadgetParams<T> &g, const T *bsk, uint32_t n_lwe, const uint32_t *lwe, const T *tv,
                          T *acc_out, size_t batch, cudaStream_t stream) {
    if constexpr (sizeof(T) == 8) {
        static const bool lazy = !(getenv("PFHE_F64_LAZY") && getenv("PFHE_F64_LAZY")[0] == '0');
        if (tb.use_f64 && lazy) return run_br_f<F64LazyField, LOGN, LOGE, PPB, MINB>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, stream);
        if (tb.use_f64) return run_br_f<F64Field, LOGN, LOGE, PPB, MINB>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, stream);
    } else {
        if (wide32_ok(tb.q, LOGN)) return run_br_f<IntWide32Field, LOGN, LOGE, PPB, MINB>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, stream);
    }
    return run_br_f<IntField<T>, LOGN, LOGE, PPB, MINB>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, stream);
}

// ---- multi-limb (L > 1) external product, fused per (ciphertext, limb) ----------------------------------------------
// The digits of CrtGlwe::mul_dcrt_ggsw_to depend on ALL limbs of a coefficient (compose -> multi-word gadget), so they are
// produced once by rns_gadget_kernel (rns.cu) as lifted residues [ct][r][level][limb][N]; from there every limb is an
// independent single-modulus problem: this kernel reads each digit polynomial ONCE, transforms it in registers,
// multiply-accumulates it against key[r][level][c][limb] and writes only the k+1 output polynomials of its limb
// (add_dcrt_glev_mul_crt_poly_assign, primus_lattice/src/glwe/dcrt.rs:178-255, restricted to one limb).
template <typename F, int LOGN, int LOGE, int COMPS>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)), ep_min_blocks<LOGN, LOGE, 1>())
dcrt_external_product_kernel(const DevNtt<typename F::WordT> *__restrict__ tables, int limbs, uint32_t levels,
                             const typename F::WordT *__restrict__ key, const typename F::WordT *__restrict__ digits,
                             typename F::WordT *__restrict__ out, int to_coeff) {
    using EP = ExtProd<F, LOGN, LOGE, COMPS>;
    using Core = typename EP::Core;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    using LA = LatAcc<F>;
    constexpr int N = EP::N, E = EP::E, CW = EP::CW, NV = EP::NV, FB0 = EP::FB0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Elem *sm = reinterpret_cast<Elem *>(smem_raw);
    const int t = threadIdx.x;
    const size_t ct = blockIdx.x / (unsigned)limbs;
    const int limb = (int)(blockIdx.x % (unsigned)limbs);
    const DevNtt<T> tb = tables[limb];
    const typename F::Ctx cx = F::ctx(tb);
    LSyncBlock sync;
    typename EP::Acc acc[COMPS][E];
#pragma unroll
    for (int c = 0; c < COMPS; c++)
#pragma unroll
        for (int j = 0; j < E; j++) LA::zero(acc[c][j]);
    uint32_t terms = 0;
#pragma unroll 1
    for (int r = 0; r < COMPS; r++) {
#pragma unroll 1
        for (uint32_t l = 0; l < levels; l++) {
            const T *dig = digits + ((((ct * COMPS + r) * levels + l) * limbs + limb) * (size_t)N);
            Elem x[E];
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = F::load(ldg_stream(dig + Core::elem_index(FB0, t, j)), cx);
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
                    for (int w = 0; w < CW; w++) LA::mac(acc[c][v * CW + w], x[v * CW + w], kv.v[w], cx);
                }
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

template <typename F, int LOGN, int LOGE, int COMPS>
static cudaError_t run_dcrt_ep_f(const DevNtt<typename F::WordT> *tables, int limbs, uint32_t levels, const typename F::WordT *key,
                                 const typename F::WordT *digits, typename F::WordT *out, size_t batch, bool to_coeff, cudaStream_t stream) {
    using T = typename F::WordT;
    constexpr int threads = 1 << (LOGN - LOGE);
    constexpr size_t smem = sizeof(T) * ((size_t)1 << LOGN);
    auto k = dcrt_external_product_kernel<F, LOGN, LOGE, COMPS>;
    cudaError_t e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)(batch * limbs), threads, smem, stream>>>(tables, limbs, levels, key, digits, out, to_coeff ? 1 : 0);
    count_launch();
    return cudaGetLastError();
}
// policy: 0 = integer pipe, 1 = lazy FP64 (every limb < 2^50), 2 = wide u32 forward (every limb passes wide32_ok)
template <typename T, int LOGN, int COMPS>
static cudaError_t run_dcrt_ep(int policy, const DevNtt<T> *tables, int limbs, uint32_t levels, const T *key, const T *digits, T *out,
                               size_t batch, bool to_coeff, cudaStream_t stream) {
    if constexpr (sizeof(T) == 8) {
        if (policy == 1) return run_dcrt_ep_f<F64LazyField, LOGN, 3, COMPS>(tables, limbs, levels, key, digits, out, batch, to_coeff, stream);
    } else {
        if (policy == 2) return run_dcrt_ep_f<IntWide32Field, LOGN, 3, COMPS>(tables, limbs, levels, key, digits, out, batch, to_coeff, stream);
    }
    return run_dcrt_ep_f<IntField<T>, LOGN, 3, COMPS>(tables, limbs, levels, key, digits, out, batch, to_coeff, stream);
}
bool dcrt_wide32_ok(uint64_t q, int logn) { return wide32_ok(q, logn); }

template <typename T>
cudaError_t launch_dcrt_external_product(int policy, const DevNtt<T> *tables, int limbs, uint32_t log_n, uint32_t k, uint32_t levels,
                                         const T *key, const T *digits, T *out, size_t batch, bool to_coeff, cudaStream_t s) {
    if (batch == 0) return cudaSuccess;
#define PFHE_DEP_CASE(LOGN)                                                                                                           \
    case LOGN:                                                                                                                        \
        if (k == 1) return run_dcrt_ep<T, LOGN, 2>(policy, tables, limbs, levels, key, digits, out, batch, to_coeff, s);              \
        if (k == 2) return run_dcrt_ep<T, LOGN, 3>(policy, tables, limbs, levels, key, digits, out, batch, to_coeff, s);              \
        break;
    switch (log_n) {
        PFHE_DEP_CASE(10)
        PFHE_DEP_CASE(11)
        PFHE_DEP_CASE(12)
    }
#undef PFHE_DEP_CASE
    return cudaErrorNotSupported;
}
template cudaError_t launch_dcrt_external_product<uint32_t>(int, const DevNtt<uint32_t> *, int, uint32_t, uint32_t, uint32_t, const uint32_t *,
                                                            const uint32_t *, uint32_t *, size_t, bool, cudaStream_t);
template cudaError_t launch_dcrt_external_product<uint64_t>(int, const DevNtt<uint64_t> *, int, uint32_t, uint32_t, uint32_t, const uint64_t *,
                                                            const uint64_t *, uint64_t *, size_t, bool, cudaStream_t);

// The lattice kernels use their own (smaller) register tile: DevNtt::fwd_pass/inv_pass must have been laid
// out for lattice_loge(bits, log_n) -- capi.cu passes the matching DevNtt view.
int lattice_loge(int bits, int log_n) {
    if (log_n < 10 || log_n > 12) return 0;
    return 3;
}

#define PFHE_EP_CASE(LOGN, LOGE, PPB)                                                                          \
    case LOGN:                                                                                                  \
        if (k == 1) return run_ep<T, LOGN, LOGE, 2, PPB>(tb, g, key, in, out, batch, to_coeff, s);              \
        if (k == 2) return run_ep<T, LOGN, LOGE, 3, PPB>(tb, g, key, in, out, batch, to_coeff, s);              \
        break;

template <>
cudaError_t launch_external_product<uint64_t>(const DevNtt<uint64_t> &tb, const GadgetParams<uint64_t> &g, uint32_t k, const uint64_t *key,
                                              const uint64_t *in, uint64_t *out, size_t batch, bool to_coeff, cudaStream_t s) {
    using T = uint64_t;
    if (batch == 0) return cudaSuccess;
    switch (tb.log_n) {
        PFHE_EP_CASE(10, 3, 2)
        PFHE_EP_CASE(11, 3, 1)
        PFHE_EP_CASE(12, 3, 1)
    }
    return cudaErrorNotSupported;
}
template <>
cudaError_t launch_external_product<uint32_t>(const DevNtt<uint32_t> &tb, const GadgetParams<uint32_t> &g, uint32_t k, const uint32_t *key,
                                              const uint32_t *in, uint32_t *out, size_t batch, bool to_coeff, cudaStream_t s) {
    using T = uint32_t;
    if (batch == 0) return cudaSuccess;
    switch (tb.log_n) {
        PFHE_EP_CASE(10, 3, 1)
        PFHE_EP_CASE(11, 3, 1)
        PFHE_EP_CASE(12, 3, 1)
    }
    return cudaErrorNotSupported;
}

template <>
cudaError_t launch_blind_rotate<uint64_t>(const DevNtt<uint64_t> &tb, const GadgetParams<uint64_t> &g, const uint64_t *bsk, uint32_t n_lwe,
                                          const uint32_t *lwe, const uint64_t *tv, uint64_t *acc_out, size_t batch, cudaStream_t s) {
    using T = uint64_t;
    if (batch == 0) return cudaSuccess;
    switch (tb.log_n) {
        case 10: return run_br<T, 10, 3, 2>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
        case 11: return run_br<T, 11, 3, 1>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
        case 12: return run_br<T, 12, 3, 1>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
    }
    return cudaErrorNotSupported;
}
template <>
cudaError_t launch_blind_rotate<uint32_t>(const DevNtt<uint32_t> &tb, const GadgetParams<uint32_t> &g, const uint32_t *bsk, uint32_t n_lwe,
                                          const uint32_t *lwe, const uint32_t *tv, uint32_t *acc_out, size_t batch, cudaStream_t s) {
    using T = uint32_t;
    if (batch == 0) return cudaSuccess;
    switch (tb.log_n) {
        case 10: {
            const char *e = getenv("PFHE_BR_MINB");  // tuning hook
            const int mb = e ? atoi(e) : 4;
            if (mb == 5) return run_br<T, 10, 3, 1, 5>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
            if (mb == 8) return run_br<T, 10, 3, 1, 8>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
            if (mb == 6) return run_br<T, 10, 3, 1, 6>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
            return run_br<T, 10, 3, 1, 4>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
        }
        case 11: return run_br<T, 11, 3, 1>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
        case 12: return run_br<T, 12, 3, 1>(tb, g, bsk, n_lwe, lwe, tv, acc_out, batch, s);
    }
    return cudaErrorNotSupported;
}

}  // namespace pfhe
