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

namespace pfhe {

struct LSyncBlock {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct LSyncWarp {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
template <int TPP> struct LSyncFor {
    using type = LSyncBlock;
};
template <> struct LSyncFor<32> {
    using type = LSyncWarp;
};

template <typename T> struct Wide2 {
    T lo, hi;
};
__device__ __forceinline__ void mac_wide(uint64_t &acc, uint32_t a, uint32_t b) { acc += (uint64_t)a * b; }
__device__ __forceinline__ void mac_wide(Wide2<uint64_t> &acc, uint64_t a, uint64_t b) {
    const uint64_t lo = a * b, hi = __umul64hi(a, b);
    acc.lo += lo;
    acc.hi += hi + (acc.lo < lo);
}

// Key multiply-accumulate policy per field.
//  integer pipe: lazy double-word sums, <= 16 products before one Barrett reduction (reduce_dot_product,
//                primus_modulus/src/common/compact/slice.rs:371-401; safe because q < 2^(BITS-2));
//  FP64 pipe   : acc <- fold(acc + mulmod(x, key)) with everything an exact integer double in (-q, q).
template <typename F> struct LatAcc;
template <> struct LatAcc<IntField<uint32_t>> {
    using F = IntField<uint32_t>;
    using Acc = uint64_t;
    static constexpr bool kRenorm = true;
    static constexpr uint32_t kRenormEvery = 16;
    __device__ __forceinline__ static void zero(Acc &a) { a = 0; }
    __device__ __forceinline__ static uint32_t prepare(uint32_t x, const F::Ctx &c) { return F::fwd_word(x, c); }
    __device__ __forceinline__ static void mac(Acc &a, uint32_t x, uint32_t key, const F::Ctx &) { mac_wide(a, x, key); }
    __device__ __forceinline__ static void renorm(Acc &a, const F::Ctx &c) { a = barrett_reduce_wide(c.br, (uint32_t)a, (uint32_t)(a >> 32)); }
    __device__ __forceinline__ static uint32_t final(const Acc &a, const F::Ctx &c) { return barrett_reduce_wide(c.br, (uint32_t)a, (uint32_t)(a >> 32)); }
};
// wide forward outputs (< 2^32) go straight into the double-word sums: 16 * (2 log2 N + 1) * q^2 < 2^64 is checked on the host
template <> struct LatAcc<IntWide32Field> : LatAcc<IntField<uint32_t>> {
    using F = IntWide32Field;
    __device__ __forceinline__ static uint32_t prepare(uint32_t x, const F::Ctx &) { return x; }
};
template <> struct LatAcc<IntField<uint64_t>> {
    using F = IntField<uint64_t>;
    using Acc = Wide2<uint64_t>;
    static constexpr bool kRenorm = true;
    static constexpr uint32_t kRenormEvery = 16;
    __device__ __forceinline__ static void zero(Acc &a) { a.lo = 0; a.hi = 0; }
    __device__ __forceinline__ static uint64_t prepare(uint64_t x, const F::Ctx &c) { return F::fwd_word(x, c); }
    __device__ __forceinline__ static void mac(Acc &a, uint64_t x, uint64_t key, const F::Ctx &) { mac_wide(a, x, key); }
    __device__ __forceinline__ static void renorm(Acc &a, const F::Ctx &c) { a.lo = barrett_reduce_wide(c.br, a.lo, a.hi); a.hi = 0; }
    __device__ __forceinline__ static uint64_t final(const Acc &a, const F::Ctx &c) { return barrett_reduce_wide(c.br, a.lo, a.hi); }
};
template <> struct LatAcc<F64Field> {
    using F = F64Field;
    using Acc = double;
    static constexpr bool kRenorm = false;
    static constexpr uint32_t kRenormEvery = 16;
    __device__ __forceinline__ static void zero(Acc &a) { a = 0.0; }
    __device__ __forceinline__ static double prepare(double x, const F::Ctx &) { return x; }  // |x| < 2q is a valid multiplier input
    __device__ __forceinline__ static void mac(Acc &a, double x, uint64_t key, const F::Ctx &c) {
        a = F::fold(__dadd_rn(a, F::mulmod(x, F::from_u64(key), c)), c);
    }
    __device__ __forceinline__ static void renorm(Acc &, const F::Ctx &) {}
    __device__ __forceinline__ static double final(const Acc &a, const F::Ctx &) { return a; }  // (-q, q): inverse-transform input
};

// FP64 pipe, lazy folds (F64LazyField): the transformed digit is folded once to |x| <= q/2 + 1, every product is then
// below 0.625 q in magnitude (level-0 quotient), and the accumulator is folded after 8 terms (8 * 0.625 q + q/2 < 8 q <= 2^53,
// all sums exact).  No per-term fold, no integer-ALU work.
template <> struct LatAcc<F64LazyField> {
    using F = F64LazyField;
    using Acc = double;
    static constexpr bool kRenorm = true;
    static constexpr uint32_t kRenormEvery = 8;
    __device__ __forceinline__ static void zero(Acc &a) { a = 0.0; }
    __device__ __forceinline__ static double prepare(double x, const F::Ctx &c) {
        F::refold(x, c);
        return x;
    }
    __device__ __forceinline__ static void mac(Acc &a, double x, uint64_t key, const F::Ctx &c) {
        a = __dadd_rn(a, F::mulmod(x, F::from_u64(key), c, 0));
    }
    __device__ __forceinline__ static void renorm(Acc &a, const F::Ctx &c) { F::refold(a, c); }
    __device__ __forceinline__ static double final(const Acc &a, const F::Ctx &c) {  // centred: first inverse pass input
        double v = a;
        F::refold(v, c);
        return v;
    }
};

// init_value_carry (primus_decompose/src/primitive/basis.rs:254-283): adjusted value + initial carry
template <typename T> __device__ __forceinline__ T gadget_init(const GadgetParams<T> &g, T v, uint32_t &carry) {
    if (g.has_threshold && v >= g.threshold) v += g.add;
    carry = g.has_init_mask ? (uint32_t)((v & g.init_mask) != 0) : 0u;
    return v;
}
// OnceSignedDecomposer::decompose_to for level l (primitive/common.rs:246-259); updates the carry
template <typename T> __device__ __forceinline__ T gadget_level(const GadgetParams<T> &g, T adj, uint32_t shift, uint32_t &carry) {
    const T t = ((adj >> shift) & g.basis_m1) + carry;
    carry = (t & g.carry_mask) != 0;
    return carry ? (t > g.basis_m1 ? T(0) : t + g.q_minus_basis) : t;
}

// Vec loads through the read-only path
template <typename V> __device__ __forceinline__ V ldg_vec(const V *p) {
    static_assert(sizeof(V) == 16, "16-byte vectors only");
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    V v;
    *reinterpret_cast<uint4 *>(&v) = r;
    return v;
}

template <typename F, int LOGN, int LOGE, int COMPS> struct ExtProd {
    using Core = NttCore<F, LOGN, LOGE>;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    using LA = LatAcc<F>;
    using Acc = typename LA::Acc;
    static constexpr int N = Core::N, E = Core::E, TPP = Core::TPP, FB0 = Core::P::fb(0);
    static constexpr int CW = Core::CW, NV = Core::NV;

    // acc[c][j] (+)= sum_{r,l} fwd(digit_l(get(r, idx))) * key[r][l][c][t*E + j]
    // kstage != nullptr: the key words of term (r, l) -- COMPS x E words per thread -- are copied asynchronously (cp.async, no registers) into a
    // thread-private, double-buffered shared-memory slot BEFORE the digit of that term is transformed, so the L2 latency of the key hides
    // behind the transform instead of stalling the multiply-accumulate (r01 ncu: 20 % of the stall samples were long-scoreboard waits on the
    // key).  Layout [buffer][c][vector][thread] x 16 bytes: conflict free for the 128-bit reads.  `kthreads` = threads sharing kstage.
    template <typename GetIn, typename SyncF>
    __device__ __forceinline__ static void accumulate(GetIn get, const T *__restrict__ key, const GadgetParams<T> &g, const DevNtt<T> &tb,
                                                      const typename F::Ctx &cx, Acc (&acc)[COMPS][E], Elem *sm, int t, SyncF sync,
                                                      uint4 *kstage = nullptr, int kthreads = 0, int kt = 0, bool l1_prefetch = false) {
        uint32_t terms = 0;
        auto stage_key = [&](int r, uint32_t l, int buf) {
            const T *kp = key + ((size_t)(r * g.levels + l) * COMPS) * N + (size_t)t * E;
#pragma unroll
            for (int c = 0; c < COMPS; c++)
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(kstage + ((size_t)((buf * COMPS + c) * NV + v) * kthreads + kt));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(kp + (size_t)c * N + v * CW) : "memory");
                }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int kbuf = 0;
        if (kstage) stage_key(0, 0, 0);
#pragma unroll 1
        for (int r = 0; r < COMPS; r++) {
            // adjusted coefficient + digit offset stay in registers across the levels: the balanced digits of the reference's carry chain
            // (init_value_carry + OnceSignedDecomposer, primitive/basis.rs:254-283, common.rs:246-259) are unique, hence equal to
            // window_l(adjusted + offset) - half -- no carry state, three instructions per digit
            T adj[E];
#pragma unroll
            for (int j = 0; j < E; j++) {
                const T v = get(r, Core::elem_index(FB0, t, j));
                adj[j] = v + ((g.has_threshold && v >= g.threshold) ? (T)(g.add + g.offset) : g.offset);
            }
#pragma unroll 1
            for (uint32_t l = 0; l < g.levels; l++) {
                Elem x[E];
                const uint32_t shift = g.drop_bits + l * g.log_basis;
#pragma unroll
                for (int j = 0; j < E; j++) {
                    const T win = (adj[j] >> shift) & g.basis_m1;
                    x[j] = F::load(win >= g.half ? (T)(win - g.half) : (T)(win + (g.q - g.half)), cx);   // canonical digit mod q
                }
                const T *kp = key + ((size_t)(r * g.levels + l) * COMPS) * N + (size_t)t * E;
                if (l1_prefetch) {  // pull this term's key lines from L2 into L1 while the digit is transformed (no registers held)
#pragma unroll
                    for (int c = 0; c < COMPS; c++) asm volatile("prefetch.global.L1 [%0];" ::"l"(kp + (size_t)c * N));
                }
                Core::template fwd_from<0, true>(x, sm, tb, cx, t, sync);  // releases the exchange buffer for the next digit
#pragma unroll
                for (int j = 0; j < E; j++) x[j] = LA::prepare(x[j], cx);
                if (kstage) {  // the next term's key starts its trip now; this term's key has had the whole transform to arrive
                    const bool last = (r == COMPS - 1) && (l + 1 == g.levels);
                    if (!last) {
                        stage_key(l + 1 == g.levels ? r + 1 : r, l + 1 == g.levels ? 0u : l + 1, kbuf ^ 1);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                }
                if (LA::kRenorm) {
                    if (terms == LA::kRenormEvery) {  // keep the lazy sums inside their exact range
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
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        typename Core::WVec kv;
                        if (kstage) *reinterpret_cast<uint4 *>(&kv) = kstage[(size_t)((kbuf * COMPS + c) * NV + v) * kthreads + kt];
                        else kv = ldg_vec(reinterpret_cast<const typename Core::WVec *>(kp + (size_t)c * N) + v);
#pragma unroll
                        for (int w = 0; w < CW; w++) LA::mac(acc[c][v * CW + w], x[v * CW + w], kv.v[w], cx);
                    }
                }
                kbuf ^= 1;
            }
        }
    }
};

// at least 16 warps per SM: caps the allocator at 128 registers for the 256-thread configurations (the freer
// instruction scheduling after the barrier reduction otherwise grows to ~180 registers and halves the occupancy)
template <int LOGN, int LOGE, int PPB> constexpr int ep_min_blocks() {
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    return threads >= 512 ? 1 : 512 / threads;
}

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

// ---- multi-limb external product in ONE kernel (composed values of at most two words) ------------------------------------------------
// The digits of CrtGlwe::mul_dcrt_ggsw_to couple all limbs of a coefficient (compose -> multi-word gadget,
// primus_lattice/src/glwe/dcrt.rs:219-236), which is why round 1 wrote them to HBM first (21 % of the product's time, profiles/
// r02_large_n_experiments.md).  For Q below two words the coupling is cheap enough to repeat per limb: the CTA of (ciphertext, limb)
// composes its 8 coefficients per thread itself (base.rs:609-636), adds the carry-free digit offset
//     R = 2^(drop-1) + sum_l (B/2) 2^(drop + l*beta)
// once per input component -- the balanced digits of init_value_carry_slice_inplace + unsigned_decompose_slice_to + the centred lift
// (big_integer/basis.rs:326-367, big_integer/common.rs:275-325, base.rs:279-315) are window_l(value + R) - B/2 -- and then runs the
// same transform / multiply-accumulate / inverse as dcrt_external_product_kernel with the digits produced in registers.
template <typename T> struct BigGadget {
    T q[kRnsMaxLimbs];
    T product[2], punct[kRnsMaxLimbs][2];
    T inv_punct[kRnsMaxLimbs], inv_punct_q[kRnsMaxLimbs];
    T threshold[2], add[2], offset[2];  // offset = R
    T mask, half;
    uint32_t drop_bits, log_basis, levels;
    int limbs, has_threshold;
};
template <typename T> struct TwoWords;
template <> struct TwoWords<uint32_t> { using U = uint64_t; };
template <> struct TwoWords<uint64_t> { using U = unsigned __int128; };

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
    const U bigq = VLEN == 2 ? (((U)bg.product[1] << BITS) | bg.product[0]) : (U)bg.product[0];
    const U thr = VLEN == 2 ? (((U)bg.threshold[1] << BITS) | bg.threshold[0]) : (U)bg.threshold[0];
    const U addv = VLEN == 2 ? (((U)bg.add[1] << BITS) | bg.add[0]) : (U)bg.add[0];
    const U offs = VLEN == 2 ? (((U)bg.offset[1] << BITS) | bg.offset[0]) : (U)bg.offset[0];
    LSyncBlock sync;
    typename EP::Acc acc[COMPS][E];
#pragma unroll
    for (int c = 0; c < COMPS; c++)
#pragma unroll
        for (int j = 0; j < E; j++) LA::zero(acc[c][j]);
    uint32_t terms = 0;
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

// cudaErrorNotSupported: composed value longer than two words, k > 2 or a degree without a lattice tile (the caller then uses the
// gadget kernel + per-limb kernel pair)
template <typename T>
cudaError_t launch_dcrt_external_product_fused(int policy, const DevNtt<T> *tables, const RnsDev<T> &r, uint32_t log_n, uint32_t k, const T *key,
                                               const T *in, T *out, size_t batch, bool to_coeff, cudaStream_t s) {
    constexpr int BITS = sizeof(T) * 8;
    if (r.value_len > 2 || r.log_basis == 0 || k < 1 || k > 2 || log_n < 10 || log_n > 12) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    BigGadget<T> bg{};
    bg.limbs = r.limbs;
    for (int i = 0; i < r.limbs; i++) {
        bg.q[i] = r.q[i];
        bg.punct[i][0] = r.punct[i][0];
        bg.punct[i][1] = r.value_len > 1 ? r.punct[i][1] : 0;
        bg.inv_punct[i] = r.inv_punct[i];
        bg.inv_punct_q[i] = r.inv_punct_q[i];
    }
    for (int w = 0; w < 2; w++) {
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
    // R = 2^(drop-1) + sum_l half << (drop + l*beta), as two words
    unsigned __int128 R = r.drop_bits ? (unsigned __int128)1 << (r.drop_bits - 1) : 0;
    for (uint32_t l = 0; l < r.levels; l++) {
        const uint32_t pos = r.drop_bits + l * r.log_basis;
        if (pos < 128) R += (unsigned __int128)bg.half << pos;
    }
    bg.offset[0] = (T)R;
    bg.offset[1] = (T)(R >> BITS);
    const bool two = r.value_len == 2;
#define PFHE_DEPF_CASE(LOGN)                                                                                                       \
    case LOGN:                                                                                                                     \
        if (k == 1) return two ? run_dcrt_ep_fused<T, LOGN, 2, 2>(policy, tables, bg, key, in, out, batch, to_coeff, s)            \
                               : run_dcrt_ep_fused<T, LOGN, 2, 1>(policy, tables, bg, key, in, out, batch, to_coeff, s);           \
        return two ? run_dcrt_ep_fused<T, LOGN, 3, 2>(policy, tables, bg, key, in, out, batch, to_coeff, s)                        \
                   : run_dcrt_ep_fused<T, LOGN, 3, 1>(policy, tables, bg, key, in, out, batch, to_coeff, s);
    switch (log_n) {
        PFHE_DEPF_CASE(10)
        PFHE_DEPF_CASE(11)
        PFHE_DEPF_CASE(12)
    }
#undef PFHE_DEPF_CASE
    return cudaErrorNotSupported;
}
template cudaError_t launch_dcrt_external_product_fused<uint32_t>(int, const DevNtt<uint32_t> *, const RnsDev<uint32_t> &, uint32_t, uint32_t,
                                                                  const uint32_t *, const uint32_t *, uint32_t *, size_t, bool, cudaStream_t);
template cudaError_t launch_dcrt_external_product_fused<uint64_t>(int, const DevNtt<uint64_t> *, const RnsDev<uint64_t> &, uint32_t, uint32_t,
                                                                  const uint64_t *, const uint64_t *, uint64_t *, size_t, bool, cudaStream_t);

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
