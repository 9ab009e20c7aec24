// ntt_core.cuh -- CTA-cooperative negacyclic NTT / INTT on one polynomial held in shared memory.
//
// Algorithm = the reference's in-place radix-2 Harvey transforms
// (primus_ntt/src/ntt/prime64/scalar/transform.rs:13-141 forward CT, :151-320 inverse GS with the
// fused n^-1 last stage), regrouped B200-first:
//   * a polynomial of N = 2^LOGN words is owned by TPP = N/E threads, E = 2^LOGE words per thread;
//   * the log2 N radix-2 stages are executed as ceil(LOGN/LOGE) register passes of up to LOGE stages
//     (radix-2^LOGE butterflies entirely in registers), with one shared-memory exchange between passes;
//   * the exchange buffer keeps the canonical in-place index order, XOR-swizzled on 16-byte chunks so
//     that both the strided (one word per lane) and the contiguous (128-bit per lane) access patterns
//     are bank-conflict free (tools/bank_sim.py);
//   * twiddles are (w, w' = floor(w 2^BITS / q)) pairs, re-laid-out per pass as [slot][high] so that the
//     lanes of a warp read consecutive pairs (coalesced 128-bit loads, L1/L2 resident tables);
//   * global loads/stores are coalesced; the contiguous side goes through the swizzled buffer as
//     128-bit vectors.
// Outputs are canonical ([0,q)), which satisfies both the canonical and the lazy trait contracts.
#pragma once
#include "modarith.cuh"

namespace pfhe {

// Device-side description of one NTT table (one modulus). Arrays live in global memory.
template <typename T> struct DevNtt {
    using Pair = typename Word<T>::Pair;
    T q, two_q;
    T inv_n, inv_n_q;      // n^-1 and its Shoup quotient
    Barrett<T> br;         // for pointwise products
    uint32_t log_n;
    uint32_t loge;         // LOGE the per-pass tables were laid out for (0: none)
    const Pair *fwd;       // fwd[k] = (roots[k], roots_q[k]),  roots[brv(k)] = psi^k        (table.rs:347-351)
    const Pair *inv;       // inv[k] = (inv_roots[k], ..),      inv_roots[brv(k)+1] = psi^-(k+1) (table.rs:354-358);
                           // inv[N-1] replaced by inv_n_w = inv_n*inv_roots[N-1] (table.rs:397-400)
    const Pair *fwd_pass;  // per-pass [slot][high] layouts (see Plan)
    const Pair *inv_pass;
    const T *ordinal;      // psi^k, k < 2N (monomial transforms, table.rs:330-338)
    // FP64-pipe path (u64 words, q < 2^50): twiddles as exact doubles, same [slot][high] pass layout
    const double *fwd_pass_f;
    const double *inv_pass_f;
    double q_f, qinv_f, inv_n_f;
    uint32_t use_f64;
};

template <typename T> __device__ __forceinline__ typename Word<T>::Pair ld_pair(const typename Word<T>::Pair *p) { return __ldg(p); }


// Streaming global accesses: polynomial data is touched exactly once per kernel, so it must not evict the twiddle
// tables from L1 (they are re-read by every polynomial): loads bypass L1 allocation, stores are marked streaming.
// The loads are ordinary coherent loads (no `.nc`): the transforms run in place (`forward_batch(dev, dev)`), i.e. the
// kernel writes the memory it reads, which the read-only (non-coherent) path does not allow.
#ifndef PFHE_STREAM_HINTS
#define PFHE_STREAM_HINTS 1
#endif
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t *p) {
#if PFHE_STREAM_HINTS
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint64_t ldg_stream(const uint64_t *p) {
#if PFHE_STREAM_HINTS
    uint64_t v;
    asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
#if PFHE_STREAM_HINTS
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void stg_stream(uint32_t *p, uint32_t v) {
#if PFHE_STREAM_HINTS
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void stg_stream(uint64_t *p, uint64_t v) {
#if PFHE_STREAM_HINTS
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void stg_stream(uint4 *p, uint4 v) {
#if PFHE_STREAM_HINTS
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    *p = v;
#endif
}

// Compile-time pass plan shared by host (table layout) and device.
template <int LOGN, int LOGE> struct Plan {
    static_assert(LOGE >= 1 && LOGN >= LOGE, "bad plan");
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, TPP = N / E;
    static constexpr int NPASS = (LOGN + LOGE - 1) / LOGE;
    static constexpr int FIRST = LOGN - (NPASS - 1) * LOGE;  // stages in pass 0 (1..LOGE)
    __host__ __device__ static constexpr int nstages(int p) { return p == 0 ? FIRST : LOGE; }
    __host__ __device__ static constexpr int s0(int p) { return p == 0 ? 0 : FIRST + (p - 1) * LOGE; }
    // low bit of the index field owned by a thread in pass p
    __host__ __device__ static constexpr int fb(int p) { return p == 0 ? LOGN - LOGE : LOGN - (s0(p) + LOGE); }
    // number of distinct `high` values (= threads sharing a twiddle set are those with equal t >> fb)
    __host__ __device__ static constexpr int nh(int p) { return (TPP >> fb(p)) > 0 ? (TPP >> fb(p)) : 1; }
    __host__ __device__ static constexpr int pass_entries(int p) { return ((1 << nstages(p)) - 1) * nh(p); }
    __host__ __device__ static constexpr int pass_offset(int p) {
        int o = 0;
        for (int i = 0; i < p; i++) o += pass_entries(i);
        return o;
    }
    static constexpr int TOTAL = pass_offset(NPASS);
};

// Runtime mirror of Plan for the host-side table builder.
struct PlanRt {
    int logn, loge, npass, first;
    __host__ PlanRt(int logn_, int loge_) : logn(logn_), loge(loge_) {
        npass = (logn + loge - 1) / loge;
        first = logn - (npass - 1) * loge;
    }
    __host__ int nstages(int p) const { return p == 0 ? first : loge; }
    __host__ int s0(int p) const { return p == 0 ? 0 : first + (p - 1) * loge; }
    __host__ int fb(int p) const { return p == 0 ? logn - loge : logn - (s0(p) + loge); }
    __host__ int nh(int p) const {
        int tpp = 1 << (logn - loge);
        return (tpp >> fb(p)) > 0 ? (tpp >> fb(p)) : 1;
    }
    __host__ int pass_entries(int p) const { return ((1 << nstages(p)) - 1) * nh(p); }
    __host__ int pass_offset(int p) const {
        int o = 0;
        for (int i = 0; i < p; i++) o += pass_entries(i);
        return o;
    }
    __host__ int total() const { return pass_offset(npass); }
};


// ======================================================================================================
// Field policies: how one butterfly is computed.  Both give bit-identical canonical results.
// ======================================================================================================

// Integer pipe: Harvey lazy butterflies with Shoup twiddles (the reference's scalar algorithm,
// primus_ntt/src/ntt/prime64/scalar/arithmetic.rs:32-79).  Any q < 2^(BITS-2).
template <typename T> struct IntField {
    using WordT = T;
    using Elem = T;
    using Tw = typename Word<T>::Pair;
    struct Ctx {
        T q, two_q, inv_n, inv_n_q;
        Barrett<T> br;
    };
    __device__ __forceinline__ static Ctx ctx(const DevNtt<T> &tb) { return Ctx{tb.q, tb.two_q, tb.inv_n, tb.inv_n_q, tb.br}; }
    __device__ __forceinline__ static const Tw *fwd_tw(const DevNtt<T> &tb) { return tb.fwd_pass; }
    __device__ __forceinline__ static const Tw *inv_tw(const DevNtt<T> &tb) { return tb.inv_pass; }
    __device__ __forceinline__ static Tw ld(const Tw *p) { return __ldg(p); }
    // lazy-fold interface (see F64LazyField): the integer butterflies renormalise every stage, nothing to do
    static constexpr bool kLazy = false;
    static constexpr int kFwdMax = 64, kInvMax = 64;
    __device__ __forceinline__ static void refold(Elem &, const Ctx &) {}
    __device__ __forceinline__ static Elem load(T w, const Ctx &) { return w; }          // fwd: < 4q, inv: < 2q
    __device__ __forceinline__ static Elem load_bits(Elem raw, const Ctx &) { return raw; }
    __device__ __forceinline__ static void fwd(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) { fwd_bfly<T>(x, y, w.x, w.y, c.q, c.two_q); }
    __device__ __forceinline__ static void inv(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) { inv_bfly<T>(x, y, w.x, w.y, c.q, c.two_q); }
    // final inverse stage fused with n^-1 (transform.rs:283-318); outputs canonical
    __device__ __forceinline__ static void inv_last(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) {
        const T tx = x + y, ty = x + c.two_q - y;
        x = shoup<T>(tx, c.inv_n, c.inv_n_q, c.q);
        y = shoup<T>(ty, w.x, w.y, c.q);
    }
    __device__ __forceinline__ static T fwd_word(Elem v, const Ctx &c) { return csub(csub(v, c.two_q), c.q); }
    __device__ __forceinline__ static Elem fwd_bits(Elem v, const Ctx &c) { return fwd_word(v, c); }
    __device__ __forceinline__ static T inv_word(Elem v, const Ctx &) { return v; }
    __device__ __forceinline__ static Elem inv_bits(Elem v, const Ctx &) { return v; }
    // forward outputs a, b -> a*b mod q as an inverse-transform input (BarrettModulus::reduce_mul)
    __device__ __forceinline__ static Elem pointwise(Elem a, Elem b, const Ctx &c) { return barrett_mul<T>(c.br, fwd_word(a, c), fwd_word(b, c)); }
    // key-MAC results (canonical words here) as inverse-transform inputs / as canonical output bits
    __device__ __forceinline__ static Elem from_mac(T v, const Ctx &) { return v; }
    __device__ __forceinline__ static Elem mac_bits(T v, const Ctx &) { return v; }
};


// Integer pipe, u32 words, forward transforms WITHOUT the per-stage conditional subtraction: the Harvey forward
// butterfly (prime32/scalar/arithmetic.rs:32-41) first folds X from [0,4q) to [0,2q); the Shoup product accepts any
// 32-bit Y, so when (2*log2(N) + 1) * q < 2^32 the fold can simply be dropped -- values grow by 2q per stage and never
// wrap (5 instead of 7 instructions per butterfly).  Used by the lattice kernels for digit transforms, whose outputs
// feed double-word multiply-accumulates that need no canonical input; inverse transforms are those of IntField.
struct IntWide32Field : IntField<uint32_t> {
    __device__ __forceinline__ static void fwd(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) {
        const uint32_t t = shoup_lazy<uint32_t>(y, w.x, w.y, c.q);
        y = x + c.two_q - t;
        x = x + t;
    }
    // any 32-bit value -> [0, q): one Barrett step with floor(2^32 / q) (= high word of the two-word ratio)
    __device__ __forceinline__ static uint32_t fwd_word(Elem v, const Ctx &c) {
        const uint32_t k = __umulhi(v, c.br.r1);
        return csub<uint32_t>(csub<uint32_t>(v - k * c.q, c.two_q), c.q);
    }
    __device__ __forceinline__ static Elem fwd_bits(Elem v, const Ctx &c) { return fwd_word(v, c); }
    __device__ __forceinline__ static Elem pointwise(Elem a, Elem b, const Ctx &c) { return barrett_mul<uint32_t>(c.br, fwd_word(a, c), fwd_word(b, c)); }
};

// FP64 pipe (B200 keeps full-rate FP64: 64 DFMA/clk/SM, while a 64x64-bit integer product costs ~4 half-rate
// IMAD.WIDE).  Values are integers held exactly in doubles, |v| < 2q < 2^51; q < 2^50.
//   mulmod(y, w): P = y*w as (h, l) = (RN(P), P - h)  [exact, fma];  c = rint(h * RN(1/q))  [magic-constant rounding];
//   r = (h - c*q) + l  is exact and |r| < q  because |c - P/q| < 1/2 + |P/q| 2^-52 < 1  for |y| < 2q <= 2^51 - 2.
// Results are the same residues the integer path produces; the canonical output is bit-identical.
struct F64Field {
    using WordT = uint64_t;
    using Elem = double;
    using Tw = double;
    struct Ctx {
        double q, qinv, inv_n;
        double off1, off2;  // 2^52 + q, 2^52 + 2q (exact): biases that make a lazy value a positive 52-bit mantissa
        uint64_t qi;
        uint32_t q_hi, q_lo;
    };
    static constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    static constexpr double kTwo52 = 4503599627370496.0;
    __device__ __forceinline__ static Ctx ctx(const DevNtt<uint64_t> &tb) {
        return Ctx{tb.q_f, tb.qinv_f, tb.inv_n_f, kTwo52 + tb.q_f, kTwo52 + 2.0 * tb.q_f, tb.q,
                   (uint32_t)__double2hiint(tb.q_f), (uint32_t)__double2loint(tb.q_f)};
    }
    __device__ __forceinline__ static const Tw *fwd_tw(const DevNtt<uint64_t> &tb) { return tb.fwd_pass_f; }
    __device__ __forceinline__ static const Tw *inv_tw(const DevNtt<uint64_t> &tb) { return tb.inv_pass_f; }
    __device__ __forceinline__ static Tw ld(const Tw *p) { return __ldg(p); }

    __device__ __forceinline__ static double mulmod(double y, double w, const Ctx &c) {
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double t = __fma_rn(h, c.qinv, kMagic);
        const double k = __dsub_rn(t, kMagic);
        const double d = __fma_rn(-k, c.q, h);
        return __dadd_rn(d, l);
    }
    // (-2q, 2q) -> (-q, q): subtract copysign(q, x) when |x| >= th, where th <= q is q with the low mantissa word
    // cleared (th > q - 2^29).  Folding a little early is harmless (|x - q| < q still holds for x in [th, 2q)), and it
    // lets the test look at the high word only: 1 ISETP + 2 LOP3 + 2 SEL on the integer ALU, 1 DADD on the FP64 pipe.
    __device__ __forceinline__ static double fold(double x, const Ctx &c) {
        const uint32_t hi = (uint32_t)__double2hiint(x);
        const bool ge = (hi & 0x7fffffffu) >= c.q_hi;
        const uint32_t shi = ge ? (c.q_hi | (hi & 0x80000000u)) : 0u;
        const uint32_t slo = ge ? c.q_lo : 0u;
        return __dsub_rn(x, __hiloint2double((int)shi, (int)slo));
    }
    // exact u64 (< 2^52) -> double without the slow conversion pipe
    __device__ __forceinline__ static double from_u64(uint64_t v) {
        return __dsub_rn(__hiloint2double((int)(0x43300000u | (uint32_t)(v >> 32)), (int)(uint32_t)v), kTwo52);
    }
    __device__ __forceinline__ static uint64_t to_u64(double v) {  // v integer in [0, 2^52)
        const double t = __dadd_rn(v, kTwo52);
        return ((uint64_t)((uint32_t)__double2hiint(t) & 0x000fffffu) << 32) | (uint32_t)__double2loint(t);
    }
    // inputs are canonical words in [0, q) (lazy-range callers are canonicalised by a pointwise pre-pass, capi.cu)
    __device__ __forceinline__ static Elem load(uint64_t w, const Ctx &) { return from_u64(w); }
    __device__ __forceinline__ static Elem load_bits(Elem raw, const Ctx &c) { return load((uint64_t)__double_as_longlong(raw), c); }
    static constexpr bool kLazy = false;
    static constexpr int kFwdMax = 64, kInvMax = 64;
    __device__ __forceinline__ static void refold(Elem &, const Ctx &) {}
    // |x|,|y| < 2q on entry and exit
    __device__ __forceinline__ static void fwd(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) {
        const double xt = fold(x, c);
        const double t = mulmod(y, w, c);
        x = __dadd_rn(xt, t);
        y = __dsub_rn(xt, t);
    }
    // |x|,|y| < q on entry and exit
    __device__ __forceinline__ static void inv(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) {
        const double s = __dadd_rn(x, y), d = __dsub_rn(x, y);
        x = fold(s, c);
        y = mulmod(d, w, c);
    }
    __device__ __forceinline__ static void inv_last(Elem &x, Elem &y, const Tw &w, const Ctx &c, int = 0) {
        const double s = __dadd_rn(x, y), d = __dsub_rn(x, y);
        x = mulmod(s, c.inv_n, c);
        y = mulmod(d, w, c);
    }
    // canonical words: one biased FP64 add exposes v + kq as a 52-bit integer mantissa; the final conditional
    // subtractions run on the integer ALU (keeps the FP64 pipe, the bottleneck, for butterflies)
    __device__ __forceinline__ static uint64_t mant(double t) {
        return ((uint64_t)((uint32_t)__double2hiint(t) & 0x000fffffu) << 32) | (uint32_t)__double2loint(t);
    }
    __device__ __forceinline__ static uint64_t fwd_word(Elem v, const Ctx &c) {  // v in (-2q, 2q)
        const uint64_t w = mant(__dadd_rn(v, c.off2));                           // v + 2q in (0, 4q)
        return csub<uint64_t>(csub<uint64_t>(w, c.qi + c.qi), c.qi);
    }
    __device__ __forceinline__ static Elem fwd_bits(Elem v, const Ctx &c) { return __longlong_as_double((long long)fwd_word(v, c)); }
    __device__ __forceinline__ static uint64_t inv_word(Elem v, const Ctx &c) {  // v in (-q, q)
        return csub<uint64_t>(mant(__dadd_rn(v, c.off1)), c.qi);
    }
    __device__ __forceinline__ static Elem inv_bits(Elem v, const Ctx &c) { return __longlong_as_double((long long)inv_word(v, c)); }
    __device__ __forceinline__ static Elem pointwise(Elem a, Elem b, const Ctx &c) { return mulmod(fold(a, c), fold(b, c), c); }
    // key-MAC results (doubles in (-q, q)) as inverse-transform inputs / as canonical output bits
    __device__ __forceinline__ static Elem from_mac(double v, const Ctx &) { return v; }
    __device__ __forceinline__ static Elem mac_bits(double v, const Ctx &c) { return __longlong_as_double((long long)inv_word(v, c)); }
};


// FP64 pipe, lazy folds.  Same exact-integer-in-a-double arithmetic as F64Field, but the per-stage conditional fold of
// the sum path (1 DADD + 5 integer-ALU instructions per butterfly) is gone: q < 2^50 leaves three spare mantissa bits,
// so values are allowed to grow for a whole register pass and every element is folded ONCE per pass with
// x - q*rint(x/q) (3 FP64 instructions, no ALU).  A butterfly is then 6 (product) + 2 (add/sub) FP64 instructions.
// The product rounds its quotient to a multiple of 2^c with the magic constant 1.5*2^(52+c); the level c grows with the
// number of stages since the last fold so that the quotient always fits (tools/f64_bounds.py walks every schedule used
// here with exact rationals and asserts the k range and the integer exactness of every sum for q <= 2^50 - 2^10):
//   forward (CT):  level by stage-since-fold {0,0,0,1,2}, |value| < 7.76 q after 5 stages
//   inverse (GS):  level {0,1,2,3}, the sum path doubles per stage: |value| <= 8 (q/2 + 2) <= 2^53 after 4 stages
// Loads centre canonical words exactly to [-(q-1)/2, (q+1)/2); stores fold, add q and subtract q conditionally.
struct F64LazyField : F64Field {
    struct Ctx {
        double q, qinv, inv_n;
        double off1;   // 2^52 + q
        uint64_t qi, half;  // q, (q+1)/2
    };
    static constexpr bool kLazy = true;
    static constexpr int kFwdMax = 5, kInvMax = 4;
    __device__ __forceinline__ static Ctx ctx(const DevNtt<uint64_t> &tb) {
        return Ctx{tb.q_f, tb.qinv_f, tb.inv_n_f, kTwo52 + tb.q_f, tb.q, (tb.q + 1) >> 1};
    }
    __device__ __forceinline__ static double magic(int level) {
        return level <= 0 ? 6755399441055744.0 : level == 1 ? 13510798882111488.0 : level == 2 ? 27021597764222976.0 : 54043195528445952.0;
    }
    __device__ __forceinline__ static double mulmod(double y, double w, const Ctx &c, int level) {
        const double m = magic(level);
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double t = __fma_rn(h, c.qinv, m);
        const double k = __dsub_rn(t, m);
        const double d = __fma_rn(-k, c.q, h);
        return __dadd_rn(d, l);
    }
    // any |x| <= 2^53 -> |x| <= q/2 + 1
    __device__ __forceinline__ static void refold(double &x, const Ctx &c) {
        const double t = __fma_rn(x, c.qinv, kMagic);
        const double k = __dsub_rn(t, kMagic);
        x = __fma_rn(-k, c.q, x);
    }
    __device__ __forceinline__ static int fwd_level(int since) { return since < 3 ? 0 : since - 2; }
    __device__ __forceinline__ static Elem load(uint64_t w, const Ctx &c) {
        const double bias = w >= c.half ? c.off1 : kTwo52;
        return __dsub_rn(__hiloint2double((int)(0x43300000u | (uint32_t)(w >> 32)), (int)(uint32_t)w), bias);
    }
    __device__ __forceinline__ static Elem load_bits(Elem raw, const Ctx &c) { return load((uint64_t)__double_as_longlong(raw), c); }
    __device__ __forceinline__ static void fwd(Elem &x, Elem &y, const Tw &w, const Ctx &c, int since) {
        const double t = mulmod(y, w, c, fwd_level(since));
        y = __dsub_rn(x, t);
        x = __dadd_rn(x, t);
    }
    __device__ __forceinline__ static void inv(Elem &x, Elem &y, const Tw &w, const Ctx &c, int since) {
        const double s = __dadd_rn(x, y), d = __dsub_rn(x, y);
        x = s;
        y = mulmod(d, w, c, since);
    }
    __device__ __forceinline__ static void inv_last(Elem &x, Elem &y, const Tw &w, const Ctx &c, int since) {
        const double s = __dadd_rn(x, y), d = __dsub_rn(x, y);
        x = mulmod(s, c.inv_n, c, since);
        y = mulmod(d, w, c, since);
    }
    __device__ __forceinline__ static uint64_t canon(Elem v, const Ctx &c) {
        refold(v, c);
        return csub<uint64_t>(mant(__dadd_rn(v, c.off1)), c.qi);
    }
    __device__ __forceinline__ static uint64_t fwd_word(Elem v, const Ctx &c) { return canon(v, c); }
    __device__ __forceinline__ static Elem fwd_bits(Elem v, const Ctx &c) { return __longlong_as_double((long long)canon(v, c)); }
    __device__ __forceinline__ static uint64_t inv_word(Elem v, const Ctx &c) { return canon(v, c); }
    __device__ __forceinline__ static Elem inv_bits(Elem v, const Ctx &c) { return __longlong_as_double((long long)canon(v, c)); }
    // key-MAC results (centred doubles) as inverse-transform inputs / as canonical output bits
    __device__ __forceinline__ static Elem from_mac(double v, const Ctx &) { return v; }
    __device__ __forceinline__ static Elem mac_bits(double v, const Ctx &c) { return __longlong_as_double((long long)canon(v, c)); }
    // forward outputs (|a|,|b| < 7.76 q) -> centred product, ready for the first inverse pass
    __device__ __forceinline__ static Elem pointwise(Elem a, Elem b, const Ctx &c) {
        refold(b, c);
        double r = mulmod(a, b, c, 2);
        refold(r, c);
        return r;
    }
};

template <typename F, int LOGN, int LOGE, bool TMA_SWZ = false> struct NttCore {
    using P = Plan<LOGN, LOGE>;
    using T = typename F::WordT;   // word type in global memory
    using Elem = typename F::Elem; // register / shared-memory representation
    using Tw = typename F::Tw;
    using Ctx = typename F::Ctx;
    static_assert(sizeof(Elem) == sizeof(T), "exchange buffer elements are word sized");
    static constexpr int N = P::N, E = P::E, TPP = P::TPP;
    static constexpr int CW = 16 / sizeof(T);                // words per 16-byte chunk
    static constexpr int SW_DST = (sizeof(T) == 8) ? 1 : 2;  // log2(CW)
    // XOR source: bits 7..9 of the byte address is exactly TMA's SWIZZLE_128B, so a tensor copy can move whole polynomials
    // between the swizzled buffer and linear HBM.  For thread rows of 128 bytes or less this is also the conflict-free
    // choice.  For 256-byte rows (u64 E = 32, u32 E = 64) the conflict-free source is one bit higher; TMA_SWZ = true
    // selects the TMA-compatible one anyway (2-way conflicts on the contiguous-per-thread pattern) -- worth it only for
    // the forward transform, whose copy-out it replaces (measured).
    static constexpr int SW_ROW = (TMA_SWZ || (LOGE - SW_DST) <= 3) ? 3 : (LOGE - SW_DST);
    static constexpr bool kTmaSwizzle = SW_ROW == 3;
    static constexpr int kRowWords = 128 / (int)sizeof(T);                     // one TMA row = 128 bytes
    static constexpr int kTmaRows = N / kRowWords;                             // rows per polynomial
    static constexpr int kTmaBoxRows = kTmaRows < 256 ? kTmaRows : 256;        // box height limit of a tensor map
    static constexpr int kTmaBoxes = kTmaRows / kTmaBoxRows;
    static constexpr int SW_SRC = SW_DST + SW_ROW;
    static constexpr int NV = E / CW > 0 ? E / CW : 1;       // 16-byte vectors per thread row
    static_assert(E >= CW, "a thread row must hold at least one 16-byte chunk");

    struct alignas(16) Vec {
        Elem v[CW];
    };
    struct alignas(16) WVec {
        T v[CW];
    };

    __device__ __forceinline__ static int swz(int idx) { return idx ^ (((idx >> SW_SRC) & 7) << SW_DST); }

    __device__ __forceinline__ static int elem_index(int p_fb, int t, int j) {
        const int low = t & ((1 << p_fb) - 1), high = t >> p_fb;
        return (high << (p_fb + LOGE)) | (j << p_fb) | low;
    }

    // ---- register passes -------------------------------------------------------------------
    template <int PASS> __device__ __forceinline__ static void fwd_pass_regs(Elem (&x)[E], const DevNtt<T> &tb, const Ctx &c, int t) {
        constexpr int NS = P::nstages(PASS), FB = P::fb(PASS), NH = P::nh(PASS);
        static_assert(NS <= F::kFwdMax, "too many forward stages between folds");
        const Tw *tw = F::fwd_tw(tb) + P::pass_offset(PASS);
        const int high = (t >> FB);
        if (F::kLazy && PASS > 0) {  // pass 0 starts from centred loads
#pragma unroll
            for (int j = 0; j < E; j++) F::refold(x[j], c);
        }
#pragma unroll
        for (int ls = 0; ls < NS; ls++) {
            const int jb = LOGE - 1 - ls;
#pragma unroll
            for (int jp = 0; jp < (1 << ls); jp++) {
                const Tw w = F::ld(tw + (((1 << ls) - 1 + jp) * NH + high));
#pragma unroll
                for (int jl = 0; jl < (1 << jb); jl++) {
                    const int j0 = (jp << (jb + 1)) | jl, j1 = j0 | (1 << jb);
                    F::fwd(x[j0], x[j1], w, c, ls);
                }
            }
        }
    }

    // inverse pass; pass 0 contains the final stage fused with n^-1
    template <int PASS> __device__ __forceinline__ static void inv_pass_regs(Elem (&x)[E], const DevNtt<T> &tb, const Ctx &c, int t) {
        constexpr int NS = P::nstages(PASS), FB = P::fb(PASS), NH = P::nh(PASS);
        const Tw *tw = F::inv_tw(tb) + P::pass_offset(PASS);
        const int high = (t >> FB);
        if (F::kLazy && PASS < P::NPASS - 1) {  // the first inverse pass starts from centred loads / folded products
#pragma unroll
            for (int j = 0; j < E; j++) F::refold(x[j], c);
        }
#pragma unroll
        for (int ls = NS - 1; ls >= 0; ls--) {
            const int jb = LOGE - 1 - ls;
            const int done = NS - 1 - ls;  // stages of this pass already executed
            if (F::kLazy && done > 0 && done % F::kInvMax == 0) {
#pragma unroll
                for (int j = 0; j < E; j++) F::refold(x[j], c);
            }
            const int since = done % F::kInvMax;
#pragma unroll
            for (int jp = 0; jp < (1 << ls); jp++) {
                const Tw w = F::ld(tw + (((1 << ls) - 1 + jp) * NH + high));
#pragma unroll
                for (int jl = 0; jl < (1 << jb); jl++) {
                    const int j0 = (jp << (jb + 1)) | jl, j1 = j0 | (1 << jb);
                    if (PASS == 0 && ls == 0)
                        F::inv_last(x[j0], x[j1], w, c, since);
                    else
                        F::inv(x[j0], x[j1], w, c, since);
                }
            }
        }
    }

    // ---- shared-memory exchange (canonical index order, swizzled) ---------------------------
    template <int PASS> __device__ __forceinline__ static void sm_store(const Elem (&x)[E], Elem *sm, int t) {
        constexpr int FB = P::fb(PASS);
        if constexpr (FB == 0) {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                Vec v;
#pragma unroll
                for (int k = 0; k < CW; k++) v.v[k] = x[c * CW + k];
                *reinterpret_cast<Vec *>(sm + swz(t * E + c * CW)) = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < E; j++) sm[swz(elem_index(FB, t, j))] = x[j];
        }
    }
    template <int PASS> __device__ __forceinline__ static void sm_load(Elem (&x)[E], const Elem *sm, int t) {
        constexpr int FB = P::fb(PASS);
        if constexpr (FB == 0) {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                Vec v = *reinterpret_cast<const Vec *>(sm + swz(t * E + c * CW));
#pragma unroll
                for (int k = 0; k < CW; k++) x[c * CW + k] = v.v[k];
            }
        } else {
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = sm[swz(elem_index(FB, t, j))];
        }
    }

    // coalesced 128-bit copy between global memory (natural order, raw words) and the swizzled buffer
    __device__ __forceinline__ static void copy_g2s(const T *g, Elem *sm, int t) {
#pragma unroll
        for (int v = t; v < N / CW; v += TPP) {
            const uint4 d = ldg_stream(reinterpret_cast<const uint4 *>(g + v * CW));
            *reinterpret_cast<uint4 *>(sm + swz(v * CW)) = d;
        }
    }
    __device__ __forceinline__ static void copy_s2g(const Elem *sm, T *g, int t) {
#pragma unroll
        for (int v = t; v < N / CW; v += TPP) {
            const uint4 d = *reinterpret_cast<const uint4 *>(sm + swz(v * CW));
            stg_stream(reinterpret_cast<uint4 *>(g + v * CW), d);
        }
    }

    // ---- recursive pass drivers ---------------------------------------------------------------
    // Synchronisation rules of the exchange buffer:
    //  * the elements a thread loads for pass P are exactly the elements it stores after pass P, so no barrier is
    //    needed between an sm_load and the following sm_store of the same pass;
    //  * the exchange between passes P and P+1 (P >= 1) only moves data inside groups of 2^fb(P) consecutive threads
    //    (the thread-id prefix `high` of pass P is shared by source and destination): when that group is at most a
    //    warp, __syncwarp replaces the CTA barrier;
    //  * callers that reuse the buffer with ANOTHER pass's pattern (a second transform, the coalesced copy-out) must
    //    synchronise the whole thread group themselves.
    template <int PASS_LO, typename SyncF> __device__ __forceinline__ static void xchg_sync(SyncF sync) {
        if constexpr (PASS_LO >= 1 && (1 << P::fb(PASS_LO)) <= 32)
            __syncwarp();
        else
            sync();
    }
    // forward passes PASS..NPASS-1 with x holding pass PASS's elements on entry; on exit x holds the
    // last pass's elements (thread t owns words [t*E, (t+1)*E)), lazy representation.
    // RELEASE: synchronise the whole group right after the LAST exchange load, i.e. hand the buffer back before the
    // final register pass -- for callers that run transforms back to back (the next transform's first store then
    // needs no barrier of its own, and the final pass + whatever follows overlaps other warps' exchanges).
    template <int PASS, bool RELEASE = false, typename SyncF>
    __device__ __forceinline__ static void fwd_from(Elem (&x)[E], Elem *sm, const DevNtt<T> &tb, const Ctx &c, int t, SyncF sync) {
        fwd_pass_regs<PASS>(x, tb, c, t);
        if constexpr (PASS + 1 < P::NPASS) {
            sm_store<PASS>(x, sm, t);
            xchg_sync<PASS>(sync);
            sm_load<PASS + 1>(x, sm, t);
            if constexpr (RELEASE && PASS + 2 == P::NPASS) sync();
            fwd_from<PASS + 1, RELEASE>(x, sm, tb, c, t, sync);
        }
    }
    // inverse passes PASS..0 with x holding pass PASS's elements on entry; on exit x holds pass 0's
    // elements in final form (F::inv_word gives the canonical word).
    template <int PASS, typename SyncF>
    __device__ __forceinline__ static void inv_from(Elem (&x)[E], Elem *sm, const DevNtt<T> &tb, const Ctx &c, int t, SyncF sync) {
        inv_pass_regs<PASS>(x, tb, c, t);
        if constexpr (PASS > 0) {
            sm_store<PASS>(x, sm, t);
            xchg_sync<PASS - 1>(sync);
            sm_load<PASS - 1>(x, sm, t);
            inv_from<PASS - 1>(x, sm, tb, c, t, sync);
        }
    }

    // ---- whole transforms -----------------------------------------------------------------------
    // forward: global (natural order, strided coalesced loads) -> registers of the last pass (lazy form)
    template <typename SyncF>
    __device__ __forceinline__ static void forward_g2r(const T *g, Elem (&x)[E], Elem *sm, const DevNtt<T> &tb, const Ctx &c, int t, SyncF sync) {
        constexpr int FB0 = P::fb(0);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load(ldg_stream(g + elem_index(FB0, t, j)), c);
        fwd_from<0>(x, sm, tb, c, t, sync);
    }
    // forward outputs in registers -> canonical words in the swizzled buffer (then copy_s2g)
    __device__ __forceinline__ static void fwd_regs_to_sm(const Elem (&x)[E], Elem *sm, const Ctx &c, int t) {
        Elem y[E];
#pragma unroll
        for (int j = 0; j < E; j++) y[j] = F::fwd_bits(x[j], c);
        sm_store<P::NPASS - 1>(y, sm, t);
    }
    // global (bit-reversed order) -> registers of the last pass, as inverse-transform inputs
    template <typename SyncF> __device__ __forceinline__ static void load_g2r(const T *g, Elem (&x)[E], Elem *sm, const Ctx &c, int t, SyncF sync) {
        copy_g2s(g, sm, t);
        sync();
        sm_load<P::NPASS - 1>(x, sm, t);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load_bits(x[j], c);
    }
    // inverse outputs (pass-0 layout) -> global, natural order, canonical
    __device__ __forceinline__ static void inv_regs_to_global(const Elem (&x)[E], T *g, const Ctx &c, int t) {
        constexpr int FB0 = P::fb(0);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(g + elem_index(FB0, t, j), F::inv_word(x[j], c));
    }
};

}  // namespace pfhe
