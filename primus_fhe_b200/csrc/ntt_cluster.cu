// ntt_cluster.cu -- N = 16384 transforms and fused products (u64 words, q < 2^50, FP64 lazy-fold butterflies) on a 2-CTA thread-block cluster.
//
// Why: with one CTA per polynomial the register file (512 threads x 32 doubles) and shared memory (one 128 KiB exchange buffer) of an SM
// hold exactly ONE polynomial, so the load, FP64 and store phases of an SM never overlap (profiles/r02_ncu_ntt_fwd_n16384.txt: FP64 pipe
// 57 %, 24 % of the stall samples wait for the loads, 17 % sit at EXIT).  Here a polynomial is split over the two CTAs of a cluster:
// each CTA runs 256 of the 512 threads, keeps 32 doubles per thread and a 64 KiB buffer -- so TWO CTAs of different clusters are resident
// per SM and one's memory phases hide behind the other's arithmetic.  The one exchange that crosses the pair uses st.async with
// mbarrier::complete_tx on the destination CTA's mbarrier and a remote mbarrier.arrive as the "buffer free" signal: no cluster barrier, no
// GPU-scope fence and no L1 invalidation per exchange (the barrier.cluster form is kept behind PFHE_NTT_CLUSTER_ASYNC=0).
// Measured (profiles/r02_cluster_ntt_ab.log): forward 9.98 M against 9.32 M NTT/s, 8-limb fused product 353 K against 331 K (266 K inside the
// bench sequence).  The code is generic over the degree and the register tile -- Cl<14, 5>: 2 x 256 threads x 32 words (default);
// Cl<14, 4>: 2 x 512 x 16 and Cl<13, 5>: N = 8192, 2 x 128 x 32 were measured slower than their one-CTA kernels (opt-in).
//
// Index algebra for the default tile (Plan<14, 5>: passes of 4 | 5 | 5 stages over index bits 13..10 | 9..5 | 4..0; T = cluster-wide thread id, 9 bits):
//   pass 0: thread T owns indices  j*512 + T            (j = index bits 13..9)
//   pass 1: thread T owns indices  (T>>5)*1024 + j*32 + (T&31)
//   pass 2: thread T owns indices  T*32 + j
// CTA c runs the threads with T>>8 == c.  From pass 1 on a thread's indices all have bit 13 == c, so CTA c's buffer holds the index range
// [c*8192, (c+1)*8192) and the exchange between passes 1 and 2 is CTA-local (and warp-local).  Only the exchange between passes 0 and 1
// crosses the pair: a thread stores its words with j < 16 into CTA 0's buffer and those with j >= 16 into CTA 1's.
// The inverse transform mirrors this with a pass-0 buffer laid out as [k][T & 255] per CTA.
#include <cstdlib>

#include "internal.hpp"
#include "tma.cuh"

namespace pfhe {

namespace {

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// asynchronous remote store that signals `bytes written` on an mbarrier of the destination CTA (no fence, no L1 invalidation on the reader side)
__device__ __forceinline__ void st_async_cluster(uint32_t addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr), "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}

using F = F64LazyField;
using T = uint64_t;
constexpr int CW = 2;

// LOGE = 5: 2 x 256 threads x 32 words (128 registers, 16 warps per SM); LOGE = 4: 2 x 512 threads x 16 words (64 registers, 32 warps per SM)
template <int LOGN, int LOGE> struct Cl {
    static constexpr int N = 1 << LOGN, HALF = N / 2;
    using Core = NttCore<F, LOGN, LOGE>;
    using P = typename Core::P;
    static constexpr int E = 1 << LOGE, TPC = (N / E) / 2, LOG_TPC = LOGN - LOGE - 1, FB0 = P::fb(0), NPASS = P::NPASS;
    static_assert(FB0 == LOGN - LOGE && P::fb(1) + LOGE <= LOGN - 1, "from pass 1 on a thread's indices must stay inside one half of the index space");

    // The exchange-buffer swizzle of NttCore never touches the top index bit, so a CTA-local buffer addressed with (index & (N/2 - 1)) keeps the
    // conflict-free patterns of the single-CTA kernels.
    __device__ __forceinline__ static int lswz(int idx) { return Core::swz(idx) & (HALF - 1); }

    template <int PASS> __device__ __forceinline__ static void local_sync() {  // exchange PASS -> PASS + 1 moves data inside groups of 2^fb(PASS) threads
        if constexpr ((1 << P::fb(PASS)) <= 32) __syncwarp();
        else __syncthreads();
    }
    template <int PASS> __device__ __forceinline__ static void load_pass(double (&x)[E], const double *buf, int Tg) {
        if constexpr (P::fb(PASS) == 0) {
#pragma unroll
            for (int v = 0; v < E / CW; v++) {
                const double2 w = *reinterpret_cast<const double2 *>(buf + lswz(Tg * E + v * CW));
                x[v * CW] = w.x;
                x[v * CW + 1] = w.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < E; j++) x[j] = buf[lswz(Core::elem_index(P::fb(PASS), Tg, j))];
        }
    }
    template <int PASS> __device__ __forceinline__ static void store_pass(const double (&x)[E], double *buf, int Tg) {
        if constexpr (P::fb(PASS) == 0) {
#pragma unroll
            for (int v = 0; v < E / CW; v++) *reinterpret_cast<double2 *>(buf + lswz(Tg * E + v * CW)) = make_double2(x[v * CW], x[v * CW + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < E; j++) buf[lswz(Core::elem_index(P::fb(PASS), Tg, j))] = x[j];
        }
    }
    template <int PASS> __device__ __forceinline__ static void fwd_local(double (&x)[E], double *buf, const DevNtt<T> &tb, const F::Ctx &c, int Tg) {
        load_pass<PASS>(x, buf, Tg);
        Core::template fwd_pass_regs<PASS>(x, tb, c, Tg);
        if constexpr (PASS + 1 < NPASS) {
            store_pass<PASS>(x, buf, Tg);  // the slots this thread just read
            local_sync<PASS>();
            fwd_local<PASS + 1>(x, buf, tb, c, Tg);
        }
    }
    // ---- cross-CTA exchange by st.async + mbarriers (default): no cluster barrier, no fence, no L1 invalidation per exchange ------------------
    // Each CTA owns two mbarriers (count 1): `full` counts the bytes the peer writes into this CTA's buffer (st.async complete_tx; the one
    // arrival is this CTA's expect_tx), `ready` is arrived on REMOTELY by the peer when the peer's buffer may be overwritten.
    struct Xchg {
        uint64_t *full, *ready;
        uint32_t full_phase, ready_phase, peer_buf, peer_full, peer_ready;
        bool first;  // the first exchange of a kernel needs no `ready` handshake (buffers are free after the start barrier)
    };
    __device__ __forceinline__ static Xchg xchg_make(double *buf, uint64_t *bars, uint32_t rank) {
        Xchg xs;
        xs.full = bars;
        xs.ready = bars + 1;
        xs.full_phase = xs.ready_phase = 0u;
        xs.peer_buf = map_to_cta(smem_u32(buf), rank ^ 1u);
        xs.peer_full = map_to_cta(smem_u32(bars), rank ^ 1u);
        xs.peer_ready = map_to_cta(smem_u32(bars + 1), rank ^ 1u);
        xs.first = true;
        return xs;
    }
    // Call when every thread of this CTA has finished reading the buffer's current contents.  Afterwards the local half may be overwritten
    // at once and the peer's half once the peer has said the same about its buffer.
    __device__ __forceinline__ static void xchg_begin(Xchg &xs) {
        if (!xs.first) __syncthreads();
        if (threadIdx.x == 0) {
            if (!xs.first)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(xs.peer_ready) : "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(xs.full)), "r"((uint32_t)(TPC * (E / 2) * 8)) : "memory");
        }
        if (!xs.first) {
            mbar_wait(xs.ready, xs.ready_phase);
            xs.ready_phase ^= 1u;
        }
        xs.first = false;
    }
    __device__ __forceinline__ static void xchg_end(Xchg &xs) {
        __syncthreads();                     // this CTA's own half is in place
        mbar_wait(xs.full, xs.full_phase);   // the peer's half has landed
        xs.full_phase ^= 1u;
    }
    __device__ __forceinline__ static void forward_g2r_async(const T *g, double (&x)[E], double *buf, Xchg &xs, const DevNtt<T> &tb, const F::Ctx &c,
                                                             int Tg, uint32_t rank) {
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load(ldg_stream(g + Core::elem_index(FB0, Tg, j)), c);
        Core::template fwd_pass_regs<0>(x, tb, c, Tg);
        xchg_begin(xs);
#pragma unroll
        for (int j = 0; j < E; j++) {
            const int idx = Core::elem_index(FB0, Tg, j);
            if ((uint32_t)(j >> (LOGE - 1)) == rank) buf[lswz(idx)] = x[j];
            else st_async_cluster(xs.peer_buf + 8u * (uint32_t)lswz(idx), x[j], xs.peer_full);
        }
        xchg_end(xs);
        fwd_local<1>(x, buf, tb, c, Tg);
    }
    __device__ __forceinline__ static void inverse_r2r_async(double (&x)[E], double *buf, Xchg &xs, const DevNtt<T> &tb, const F::Ctx &c, int Tg,
                                                             uint32_t rank) {
        inv_local<NPASS - 1>(x, buf, tb, c, Tg);  // ends with the pass-1 outputs in registers
        xs.first = false;                         // the buffers have been in use: always handshake
        xchg_begin(xs);
#pragma unroll
        for (int j = 0; j < E; j++) {
            const int i = Core::elem_index(P::fb(1), Tg, j);
            const uint32_t owner = (uint32_t)(i >> LOG_TPC) & 1u;
            const int slot = ((i >> FB0) << LOG_TPC) | (i & (TPC - 1));
            if (owner == rank) buf[slot] = x[j];
            else st_async_cluster(xs.peer_buf + 8u * (uint32_t)slot, x[j], xs.peer_full);
        }
        xchg_end(xs);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = buf[(j << LOG_TPC) | (Tg & (TPC - 1))];
        Core::template inv_pass_regs<0>(x, tb, c, Tg);
    }
    // forward: global (natural order) -> registers of the last pass (lazy form); Tg = cluster-wide thread id, buf = this CTA's 64 KiB buffer
    __device__ __forceinline__ static void forward_g2r(const T *g, double (&x)[E], double *buf, const DevNtt<T> &tb, const F::Ctx &c, int Tg,
                                                       uint32_t rank) {
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load(ldg_stream(g + Core::elem_index(FB0, Tg, j)), c);
        Core::template fwd_pass_regs<0>(x, tb, c, Tg);
        // exchange 0 -> 1 across the pair: a word goes to the CTA that owns its index bit 13
        const uint32_t base = smem_u32(buf);
        const uint32_t peer = map_to_cta(base, rank ^ 1u), mine = map_to_cta(base, rank);
#pragma unroll
        for (int j = 0; j < E; j++) {
            const int idx = Core::elem_index(FB0, Tg, j);   // bit 13 == top bit of j: known at compile time
            const uint32_t dst = ((uint32_t)(j >> (LOGE - 1)) == rank) ? mine : peer;
            st_cluster(dst + 8u * (uint32_t)lswz(idx), x[j]);
        }
        cluster_sync();
        fwd_local<1>(x, buf, tb, c, Tg);
    }

    template <int PASS> __device__ __forceinline__ static void inv_local(double (&x)[E], double *buf, const DevNtt<T> &tb, const F::Ctx &c, int Tg) {
        Core::template inv_pass_regs<PASS>(x, tb, c, Tg);
        if constexpr (PASS > 1) {
            store_pass<PASS>(x, buf, Tg);
            local_sync<PASS - 1>();
            load_pass<PASS - 1>(x, buf, Tg);
            inv_local<PASS - 1>(x, buf, tb, c, Tg);
        }
    }
    // inverse: registers of the last pass (inverse-transform inputs) -> registers of pass 0 in final form (F::inv_word gives the canonical word)
    __device__ __forceinline__ static void inverse_r2r(double (&x)[E], double *buf, const DevNtt<T> &tb, const F::Ctx &c, int Tg, uint32_t rank) {
        inv_local<NPASS - 1>(x, buf, tb, c, Tg);  // ends with the pass-1 outputs in registers
        // exchange 1 -> 0 across the pair.  Every thread of BOTH CTAs must have finished reading its pass-1 words before anyone overwrites a
        // buffer with the pass-0 layout [j][T & (TPC-1)]: index i belongs to pass-0 thread i & (2 TPC - 1), word j = i >> FB0.
        cluster_sync();
        const uint32_t base = smem_u32(buf);
        const uint32_t peer = map_to_cta(base, rank ^ 1u), mine = map_to_cta(base, rank);
#pragma unroll
        for (int j = 0; j < E; j++) {
            const int i = Core::elem_index(P::fb(1), Tg, j);
            const uint32_t owner = (uint32_t)(i >> LOG_TPC) & 1u;
            const int slot = ((i >> FB0) << LOG_TPC) | (i & (TPC - 1));
            st_cluster((owner == rank ? mine : peer) + 8u * (uint32_t)slot, x[j]);
        }
        cluster_sync();
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = buf[(j << LOG_TPC) | (Tg & (TPC - 1))];
        Core::template inv_pass_regs<0>(x, tb, c, Tg);
    }
    // this CTA's half of the bit-reversed (= contiguous per thread) canonical words: buffer (last-pass layout) <-> global, coalesced 16-byte accesses
    __device__ __forceinline__ static void copy_half_s2g(const double *buf, T *g_half, int t) {
#pragma unroll
        for (int v = t; v < HALF / CW; v += TPC) {
            const uint4 d = *reinterpret_cast<const uint4 *>(buf + lswz(v * CW));
            stg_stream(reinterpret_cast<uint4 *>(g_half + v * CW), d);
        }
    }
    __device__ __forceinline__ static void copy_half_g2s(const T *g_half, double *buf, int t) {
#pragma unroll
        for (int v = t; v < HALF / CW; v += TPC) {
            const uint4 d = ldg_stream(reinterpret_cast<const uint4 *>(g_half + v * CW));
            *reinterpret_cast<uint4 *>(buf + lswz(v * CW)) = d;
        }
    }
};

// MODE 0: forward, 1: inverse, 2: fused product c = a * b (fwd(a) parked in the output polynomial, as in polymul_kernel<STASH>)
// ASYNC: cross-CTA exchanges by st.async + mbarriers (default); otherwise st.shared::cluster + barrier.cluster (PFHE_NTT_CLUSTER_ASYNC=0)
template <int LOGN, int LOGE, int MODE, bool ASYNC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cl<LOGN, LOGE>::TPC, 512 / Cl<LOGN, LOGE>::TPC)
ntt_cluster_kernel(const __grid_constant__ DevNtt<T> tb0, const DevNtt<T> *__restrict__ tables, int limbs, const T *a, const T *b, T *out,
                   size_t npolys, unsigned stagger_ns) {  // out may alias a (in place): no __restrict__
    using C = Cl<LOGN, LOGE>;
    using Core = typename C::Core;
    constexpr int E = C::E, TPC = C::TPC, FB0 = C::FB0, N = C::N, HALF = C::HALF;
    extern __shared__ __align__(16) unsigned char smem_c[];
    double *buf = reinterpret_cast<double *>(smem_c);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_c + sizeof(double) * C::HALF);  // ASYNC: full, ready
    const uint32_t rank = cluster_rank();
    const int t = threadIdx.x, Tg = (int)rank * TPC + t;
    const size_t poly = blockIdx.x >> 1;
    const DevNtt<T> tb = limbs > 1 ? tables[poly % (size_t)limbs] : tb0;  // by value: only the fields in use occupy registers
    const F::Ctx c = F::ctx(tb);
    double x[E];
    // The two CTAs that share an SM start together and take equally long, so without help they stay in phase (both loading, both
    // computing, both storing) for the whole grid.  The clusters of the first wave therefore start with a pseudo-random delay of up to half a
    // CTA lifetime; later waves inherit the offset.  (stagger_ns = 0 switches it off.)
    if (stagger_ns && poly < 2 * 148) {
        const unsigned steps = (unsigned)(((unsigned)poly * 2654435761u) >> 28);  // 0..15
        for (unsigned i = 0; i < steps; i++) __nanosleep(stagger_ns);
    }
    if (ASYNC && t == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
    }
    cluster_sync();  // the peer CTA is resident (and its mbarriers initialised) before its shared memory is addressed
    typename C::Xchg xs = C::xchg_make(buf, bars, rank);
    auto forward = [&](const T *g, double *bf, int tg) {
        if constexpr (ASYNC) C::forward_g2r_async(g, x, bf, xs, tb, c, tg, rank);
        else C::forward_g2r(g, x, bf, tb, c, tg, rank);
    };
    auto inverse = [&](double *bf, int tg) {
        if constexpr (ASYNC) C::inverse_r2r_async(x, bf, xs, tb, c, tg, rank);
        else C::inverse_r2r(x, bf, tb, c, tg, rank);
    };
    // Exit needs no further cluster barrier in the ASYNC form: every remote store INTO this CTA had landed when its last wait returned, and
    // the peer -- the destination of this CTA's remote stores and remote arrivals -- cannot leave before they land because it waits for them.
    if (MODE == 0) {
        forward(a + poly * N, buf, Tg);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::fwd_bits(x[j], c);
        C::template store_pass<C::NPASS - 1>(x, buf, Tg);
        __syncthreads();
        C::copy_half_s2g(buf, out + poly * N + (size_t)rank * HALF, t);
    } else if (MODE == 1) {
        C::copy_half_g2s(a + poly * N + (size_t)rank * HALF, buf, t);
        __syncthreads();
        C::template load_pass<C::NPASS - 1>(x, buf, Tg);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load_bits(x[j], c);
        inverse(buf, Tg);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(out + poly * N + Core::elem_index(FB0, Tg, j), F::inv_word(x[j], c));
    } else {
        T *g_c = out + poly * N;
        forward(a + poly * N, buf, Tg);
        // fwd(a), canonical, parked in this thread's own contiguous words of the output polynomial (read back by the same thread)
        T *mine_c = g_c + (size_t)Tg * E;
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            uint4 raw;
            const uint64_t w0 = F::fwd_word(x[v * CW], c), w1 = F::fwd_word(x[v * CW + 1], c);
            raw.x = (uint32_t)w0; raw.y = (uint32_t)(w0 >> 32); raw.z = (uint32_t)w1; raw.w = (uint32_t)(w1 >> 32);
            *reinterpret_cast<uint4 *>(mine_c + v * CW) = raw;
        }
        if constexpr (!ASYNC) cluster_sync();  // both CTAs are done with the last-pass reads of their buffers before the next exchange writes into them
        // the second transform uses the same thread id: hide that from the compiler, which would otherwise keep every shared-memory / peer
        // address of the first transform alive for reuse (common-subexpression elimination) and spill them -- recomputing is 2-3 ALU ops
        int Tg2 = Tg;
        double *buf2 = buf;
        asm volatile("" : "+r"(Tg2), "+l"(buf2));
        forward(b + poly * N, buf2, Tg2);
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            uint4 raw;  // written above by this thread (program order); volatile: a few 16-byte loads in flight at a time
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(mine_c + v * CW) : "memory");
            const uint64_t w0 = ((uint64_t)raw.y << 32) | raw.x, w1 = ((uint64_t)raw.w << 32) | raw.z;
            x[v * CW] = F::pointwise(F::load(w0, c), x[v * CW], c);
            x[v * CW + 1] = F::pointwise(F::load(w1, c), x[v * CW + 1], c);
        }
        int Tg3 = Tg;
        double *buf3 = buf;
        asm volatile("" : "+r"(Tg3), "+l"(buf3));
        inverse(buf3, Tg3);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(g_c + Core::elem_index(FB0, Tg3, j), F::inv_word(x[j], c));
    }
}

template <int LOGN, int LOGE>
cudaError_t run_cluster(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, int mode, const uint64_t *a, const uint64_t *b,
                        uint64_t *out, size_t npolys, cudaStream_t s) {
    constexpr size_t smem = sizeof(double) * Cl<LOGN, LOGE>::HALF + 16;  // + two mbarriers (st.async variant)
    const unsigned grid = (unsigned)(2 * npolys);
    cudaError_t e;
    auto go = [&](auto k) -> cudaError_t {
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        static const unsigned stagger = getenv("PFHE_NTT_CLUSTER_STAGGER_NS") ? (unsigned)atoi(getenv("PFHE_NTT_CLUSTER_STAGGER_NS")) : 1000u;
        k<<<grid, Cl<LOGN, LOGE>::TPC, smem, s>>>(tb0, tables, limbs, a, b, out, npolys, npolys > 2 * 148 ? stagger : 0u);  // one wave: nothing to de-phase
        count_launch();
        return cudaGetLastError();
    };
    static const bool async = !(getenv("PFHE_NTT_CLUSTER_ASYNC") && getenv("PFHE_NTT_CLUSTER_ASYNC")[0] == '0');
    if (async) {
        if (mode == 0) return go(ntt_cluster_kernel<LOGN, LOGE, 0, true>);
        if (mode == 1) return go(ntt_cluster_kernel<LOGN, LOGE, 1, true>);
        return go(ntt_cluster_kernel<LOGN, LOGE, 2, true>);
    }
    if (mode == 0) return go(ntt_cluster_kernel<LOGN, LOGE, 0, false>);
    if (mode == 1) return go(ntt_cluster_kernel<LOGN, LOGE, 1, false>);
    return go(ntt_cluster_kernel<LOGN, LOGE, 2, false>);
}

}  // namespace

// cudaErrorNotSupported unless the table is a u64 FP64 lazy-fold N = 16384 layout (16- or 32-word register tiles)
cudaError_t launch_ntt_cluster(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, int mode, const uint64_t *a,
                               const uint64_t *b, uint64_t *out, size_t npolys, cudaStream_t s) {
    if (!tb0.use_f64 || !((tb0.log_n == 14 && (tb0.loge == 4 || tb0.loge == 5)) || (tb0.log_n == 13 && tb0.loge == 5))) return cudaErrorNotSupported;
    if (npolys == 0) return cudaSuccess;
    if (npolys > 0x3fffffffu) return cudaErrorNotSupported;
    if (tb0.log_n == 13) return run_cluster<13, 5>(tb0, tables, limbs, mode, a, b, out, npolys, s);
    return tb0.loge == 4 ? run_cluster<14, 4>(tb0, tables, limbs, mode, a, b, out, npolys, s) : run_cluster<14, 5>(tb0, tables, limbs, mode, a, b, out, npolys, s);
}

}  // namespace pfhe
