// ntt_cluster.cu -- N = 16384 transforms (u64 words, q < 2^50, FP64 lazy-fold butterflies) on a 2-CTA thread-block cluster.
//
// Why: with one CTA per polynomial the register file (512 threads x 32 doubles) and shared memory (one 128 KiB exchange buffer) of an SM
// hold exactly ONE polynomial, so the load, FP64 and store phases of an SM never overlap (profiles/r02_ncu_ntt_fwd_n16384.txt: FP64 pipe
// 57 %, 24 % of the stall samples wait for the loads, 17 % sit at EXIT).  Here a polynomial is split over the two CTAs of a cluster:
// each CTA runs 256 of the 512 threads, keeps 32 doubles per thread and a 64 KiB buffer -- so TWO CTAs of different clusters are resident
// per SM and one's memory phases hide behind the other's arithmetic.  Measured: +2..7 % on the transforms (default), the fused product is
// faster cold and slower inside a long run (opt-in); the FP64 pipe stays at 57 % (profiles/r02_large_n_experiments.md).  The code is generic
// over the register tile (Cl<5>: 2 x 256 threads x 32 words; Cl<4>: 2 x 512 threads x 16 words, measured slower).
//
// Index algebra for the default tile (Plan<14, 5>: passes of 4 | 5 | 5 stages over index bits 13..10 | 9..5 | 4..0; T = cluster-wide thread id, 9 bits):
//   pass 0: thread T owns indices  j*512 + T            (j = index bits 13..9)
//   pass 1: thread T owns indices  (T>>5)*1024 + j*32 + (T&31)
//   pass 2: thread T owns indices  T*32 + j
// CTA c runs the threads with T>>8 == c.  From pass 1 on a thread's indices all have bit 13 == c, so CTA c's buffer holds the index range
// [c*8192, (c+1)*8192) and the exchange between passes 1 and 2 is CTA-local (and warp-local).  Only the exchange between passes 0 and 1
// crosses the pair: a thread stores its words with j < 16 into CTA 0's buffer and those with j >= 16 into CTA 1's (st.shared::cluster).
// The inverse transform mirrors this with a pass-0 buffer laid out as [k][T & 255] per CTA.
#include <cstdlib>

#include "internal.hpp"

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

using F = F64LazyField;
using T = uint64_t;
constexpr int LOGN = 14, N = 1 << LOGN, HALF = N / 2, CW = 2;

// LOGE = 5: 2 x 256 threads x 32 words (128 registers, 16 warps per SM); LOGE = 4: 2 x 512 threads x 16 words (64 registers, 32 warps per SM)
template <int LOGE> struct Cl {
    using Core = NttCore<F, LOGN, LOGE>;
    using P = typename Core::P;
    static constexpr int E = 1 << LOGE, TPC = (N / E) / 2, LOG_TPC = LOGN - LOGE - 1, FB0 = P::fb(0), NPASS = P::NPASS;
    static_assert(FB0 == LOGN - LOGE && P::fb(1) + LOGE <= LOGN - 1, "from pass 1 on a thread's indices must stay inside one half of the index space");

    // The exchange-buffer swizzle of NttCore never touches index bit 13, so a CTA-local buffer addressed with (index & 8191) keeps the
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
template <int LOGE, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cl<LOGE>::TPC, 2)
ntt_cluster_kernel(const __grid_constant__ DevNtt<T> tb0, const DevNtt<T> *__restrict__ tables, int limbs, const T *a, const T *b, T *out,
                   size_t npolys, unsigned stagger_ns) {  // out may alias a (in place): no __restrict__
    using C = Cl<LOGE>;
    using Core = typename C::Core;
    constexpr int E = C::E, TPC = C::TPC, FB0 = C::FB0;
    extern __shared__ __align__(16) unsigned char smem_c[];
    double *buf = reinterpret_cast<double *>(smem_c);
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
    cluster_sync();  // the peer CTA is resident before its shared memory is addressed
    if (MODE == 0) {
        C::forward_g2r(a + poly * N, x, buf, tb, c, Tg, rank);
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
        C::inverse_r2r(x, buf, tb, c, Tg, rank);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(out + poly * N + Core::elem_index(FB0, Tg, j), F::inv_word(x[j], c));
    } else {
        T *g_c = out + poly * N;
        C::forward_g2r(a + poly * N, x, buf, tb, c, Tg, rank);
        // fwd(a), canonical, parked in this thread's own contiguous words of the output polynomial (read back by the same thread)
        T *mine_c = g_c + (size_t)Tg * E;
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            uint4 raw;
            const uint64_t w0 = F::fwd_word(x[v * CW], c), w1 = F::fwd_word(x[v * CW + 1], c);
            raw.x = (uint32_t)w0; raw.y = (uint32_t)(w0 >> 32); raw.z = (uint32_t)w1; raw.w = (uint32_t)(w1 >> 32);
            *reinterpret_cast<uint4 *>(mine_c + v * CW) = raw;
        }
        cluster_sync();  // both CTAs are done with the last-pass reads of their buffers before the next transform's exchange writes into them
        // the second transform uses the same thread id: hide that from the compiler, which would otherwise keep every shared-memory / peer
        // address of the first transform alive for reuse (common-subexpression elimination) and spill them -- recomputing is 2-3 ALU ops
        int Tg2 = Tg;
        double *buf2 = buf;
        asm volatile("" : "+r"(Tg2), "+l"(buf2));
        C::forward_g2r(b + poly * N, x, buf2, tb, c, Tg2, rank);
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
        C::inverse_r2r(x, buf3, tb, c, Tg3, rank);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(g_c + Core::elem_index(FB0, Tg3, j), F::inv_word(x[j], c));
    }
}

template <int LOGE>
cudaError_t run_cluster(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, int mode, const uint64_t *a, const uint64_t *b,
                        uint64_t *out, size_t npolys, cudaStream_t s) {
    constexpr size_t smem = sizeof(double) * HALF;
    const unsigned grid = (unsigned)(2 * npolys);
    cudaError_t e;
    auto go = [&](auto k) -> cudaError_t {
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        static const unsigned stagger = getenv("PFHE_NTT_CLUSTER_STAGGER_NS") ? (unsigned)atoi(getenv("PFHE_NTT_CLUSTER_STAGGER_NS")) : 1000u;
        k<<<grid, Cl<LOGE>::TPC, smem, s>>>(tb0, tables, limbs, a, b, out, npolys, npolys > 2 * 148 ? stagger : 0u);  // one wave: nothing to de-phase
        count_launch();
        return cudaGetLastError();
    };
    if (mode == 0) return go(ntt_cluster_kernel<LOGE, 0>);
    if (mode == 1) return go(ntt_cluster_kernel<LOGE, 1>);
    return go(ntt_cluster_kernel<LOGE, 2>);
}

}  // namespace

// cudaErrorNotSupported unless the table is a u64 FP64 lazy-fold N = 16384 layout (16- or 32-word register tiles)
cudaError_t launch_ntt_cluster(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, int mode, const uint64_t *a,
                               const uint64_t *b, uint64_t *out, size_t npolys, cudaStream_t s) {
    if (tb0.log_n != LOGN || (tb0.loge != 4 && tb0.loge != 5) || !tb0.use_f64) return cudaErrorNotSupported;
    if (npolys == 0) return cudaSuccess;
    if (npolys > 0x3fffffffu) return cudaErrorNotSupported;
    return tb0.loge == 4 ? run_cluster<4>(tb0, tables, limbs, mode, a, b, out, npolys, s) : run_cluster<5>(tb0, tables, limbs, mode, a, b, out, npolys, s);
}

}  // namespace pfhe
