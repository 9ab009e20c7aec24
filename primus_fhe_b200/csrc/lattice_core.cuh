// lattice_core.cuh -- building blocks shared by the lattice kernels (lattice.cu, lattice_dcrt.cu): barrier functors, the key
// multiply-accumulate policy per field, and ExtProd::accumulate (digits -> forward transform -> lazy key MAC).
#pragma once
#include <cstdlib>

#include "internal.hpp"

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

}  // namespace pfhe
