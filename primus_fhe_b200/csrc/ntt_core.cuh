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
};

template <typename T> __device__ __forceinline__ typename Word<T>::Pair ld_pair(const typename Word<T>::Pair *p) { return __ldg(p); }

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

template <typename T, int LOGN, int LOGE> struct NttCore {
    using P = Plan<LOGN, LOGE>;
    using Pair = typename Word<T>::Pair;
    static constexpr int N = P::N, E = P::E, TPP = P::TPP;
    static constexpr int CW = 16 / sizeof(T);                // words per 16-byte chunk
    static constexpr int SW_DST = (sizeof(T) == 8) ? 1 : 2;  // log2(CW)
    static constexpr int SW_ROW = (LOGE - SW_DST) > 3 ? (LOGE - SW_DST) : 3;
    static constexpr int SW_SRC = SW_DST + SW_ROW;
    static constexpr int NV = E / CW > 0 ? E / CW : 1;       // 16-byte vectors per thread row
    static_assert(E >= CW, "a thread row must hold at least one 16-byte chunk");

    struct alignas(16) Vec {
        T v[CW];
    };

    __device__ __forceinline__ static int swz(int idx) { return idx ^ (((idx >> SW_SRC) & 7) << SW_DST); }

    __device__ __forceinline__ static int elem_index(int p_fb, int t, int j) {
        const int low = t & ((1 << p_fb) - 1), high = t >> p_fb;
        return (high << (p_fb + LOGE)) | (j << p_fb) | low;
    }

    // ---- register passes -------------------------------------------------------------------
    template <int PASS> __device__ __forceinline__ static void fwd_pass_regs(T (&x)[E], const DevNtt<T> &tb, int t) {
        constexpr int NS = P::nstages(PASS), FB = P::fb(PASS), NH = P::nh(PASS);
        const Pair *tw = tb.fwd_pass + P::pass_offset(PASS);
        const int high = (t >> FB);
        const T q = tb.q, two_q = tb.two_q;
#pragma unroll
        for (int ls = 0; ls < NS; ls++) {
            const int jb = LOGE - 1 - ls;
#pragma unroll
            for (int jp = 0; jp < (1 << ls); jp++) {
                const Pair w = ld_pair<T>(tw + (((1 << ls) - 1 + jp) * NH + high));
#pragma unroll
                for (int jl = 0; jl < (1 << jb); jl++) {
                    const int j0 = (jp << (jb + 1)) | jl, j1 = j0 | (1 << jb);
                    fwd_bfly<T>(x[j0], x[j1], w.x, w.y, q, two_q);
                }
            }
        }
    }

    // inverse pass; LASTSCALE: this pass contains the final stage fused with n^-1
    template <int PASS> __device__ __forceinline__ static void inv_pass_regs(T (&x)[E], const DevNtt<T> &tb, int t) {
        constexpr int NS = P::nstages(PASS), FB = P::fb(PASS), NH = P::nh(PASS);
        const Pair *tw = tb.inv_pass + P::pass_offset(PASS);
        const int high = (t >> FB);
        const T q = tb.q, two_q = tb.two_q;
#pragma unroll
        for (int ls = NS - 1; ls >= 0; ls--) {
            const int jb = LOGE - 1 - ls;
#pragma unroll
            for (int jp = 0; jp < (1 << ls); jp++) {
                const Pair w = ld_pair<T>(tw + (((1 << ls) - 1 + jp) * NH + high));
#pragma unroll
                for (int jl = 0; jl < (1 << jb); jl++) {
                    const int j0 = (jp << (jb + 1)) | jl, j1 = j0 | (1 << jb);
                    if (PASS == 0 && ls == 0) {
                        // final stage: x' = inv_n*(x+y), y' = inv_n_w*(x+2q-y)  (transform.rs:283-318)
                        T tx = x[j0] + x[j1];
                        T ty = x[j0] + two_q - x[j1];
                        x[j0] = shoup<T>(tx, tb.inv_n, tb.inv_n_q, q);
                        x[j1] = shoup<T>(ty, w.x, w.y, q);
                    } else {
                        inv_bfly<T>(x[j0], x[j1], w.x, w.y, q, two_q);
                    }
                }
            }
        }
    }

    // ---- shared-memory exchange (canonical index order, swizzled) ---------------------------
    template <int PASS> __device__ __forceinline__ static void sm_store(const T (&x)[E], T *sm, int t) {
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
    template <int PASS> __device__ __forceinline__ static void sm_load(T (&x)[E], const T *sm, int t) {
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

    // coalesced 128-bit copy between global memory (natural order) and the swizzled buffer
    __device__ __forceinline__ static void copy_g2s(const T *g, T *sm, int t) {
#pragma unroll
        for (int v = t; v < N / CW; v += TPP) {
            Vec d = *reinterpret_cast<const Vec *>(g + v * CW);
            *reinterpret_cast<Vec *>(sm + swz(v * CW)) = d;
        }
    }
    __device__ __forceinline__ static void copy_s2g(const T *sm, T *g, int t) {
#pragma unroll
        for (int v = t; v < N / CW; v += TPP) {
            Vec d = *reinterpret_cast<const Vec *>(sm + swz(v * CW));
            *reinterpret_cast<Vec *>(g + v * CW) = d;
        }
    }

    // ---- recursive pass drivers ---------------------------------------------------------------
    // forward passes PASS..NPASS-1 with x holding pass PASS's elements on entry; on exit x holds the
    // last pass's elements (thread t owns words [t*E, (t+1)*E)), values in [0,4q).
    template <int PASS, typename SyncF> __device__ __forceinline__ static void fwd_from(T (&x)[E], T *sm, const DevNtt<T> &tb, int t, SyncF sync) {
        fwd_pass_regs<PASS>(x, tb, t);
        if constexpr (PASS + 1 < P::NPASS) {
            sm_store<PASS>(x, sm, t);
            sync();
            sm_load<PASS + 1>(x, sm, t);
            sync();  // buffer may be rewritten by the next exchange
            fwd_from<PASS + 1>(x, sm, tb, t, sync);
        }
    }
    // inverse passes PASS..0 with x holding pass PASS's elements (values < 2q) on entry; on exit x holds
    // pass 0's elements, canonical.
    template <int PASS, typename SyncF> __device__ __forceinline__ static void inv_from(T (&x)[E], T *sm, const DevNtt<T> &tb, int t, SyncF sync) {
        inv_pass_regs<PASS>(x, tb, t);
        if constexpr (PASS > 0) {
            sm_store<PASS>(x, sm, t);
            sync();
            sm_load<PASS - 1>(x, sm, t);
            sync();
            inv_from<PASS - 1>(x, sm, tb, t, sync);
        }
    }

    // ---- whole transforms -----------------------------------------------------------------------
    // forward: global (natural order, strided coalesced loads) -> registers of the last pass, canonical
    template <typename SyncF> __device__ __forceinline__ static void forward_g2r(const T *g, T (&x)[E], T *sm, const DevNtt<T> &tb, int t, SyncF sync) {
        constexpr int FB0 = P::fb(0);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = g[elem_index(FB0, t, j)];
        fwd_from<0>(x, sm, tb, t, sync);
        const T q = tb.q, two_q = tb.two_q;
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = csub(csub(x[j], two_q), q);
    }
    // registers of the last pass -> global (bit-reversed order == in-place order), coalesced via the buffer
    template <typename SyncF> __device__ __forceinline__ static void store_r2g(const T (&x)[E], T *g, T *sm, int t, SyncF sync) {
        sm_store<P::NPASS - 1>(x, sm, t);
        sync();
        copy_s2g(sm, g, t);
    }
    // global (bit-reversed order) -> registers of the last pass
    template <typename SyncF> __device__ __forceinline__ static void load_g2r(const T *g, T (&x)[E], T *sm, int t, SyncF sync) {
        copy_g2s(g, sm, t);
        sync();
        sm_load<P::NPASS - 1>(x, sm, t);
        sync();
    }
    // inverse: registers of the last pass -> global (natural order), canonical
    template <typename SyncF> __device__ __forceinline__ static void inverse_r2g(T (&x)[E], T *g, T *sm, const DevNtt<T> &tb, int t, SyncF sync) {
        inv_from<P::NPASS - 1>(x, sm, tb, t, sync);
        constexpr int FB0 = P::fb(0);
#pragma unroll
        for (int j = 0; j < E; j++) g[elem_index(FB0, t, j)] = x[j];
    }
};

}  // namespace pfhe
