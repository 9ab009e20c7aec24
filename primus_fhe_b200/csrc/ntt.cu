// ntt.cu -- batched negacyclic NTT / INTT / fused polymul / monomial-NTT kernels for sm_100a.
//
// Replaces, for batches resident in HBM:
//   NttTable::transform_slice / inverse_transform_slice   primus_ntt/src/ntt/prime64/table.rs:542-563
//   (scalar semantics primus_ntt/src/ntt/prime64/scalar/transform.rs:13-320, prime32/scalar/transform.rs:13-273)
//   DcrtTable::transform_slice (per-limb loop)            primus_ntt/src/dcrt/prime64.rs:106-111
//   monomial transforms                                   primus_ntt/src/ntt/prime64/table.rs:565-651
//   transform + NttPolynomial::mul_assign + inverse       primus_lattice/src/rlwe/coeff.rs:92-122
//
// One polynomial per thread group (TPP threads); a CTA carries PPB groups. Each polynomial is read from
// HBM once and written once; everything in between lives in registers / shared memory.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cuda.h>

#include "tma.cuh"

#include "internal.hpp"
#include "host_math.hpp"

namespace pfhe {

std::atomic<uint64_t> g_launches{0};

static int env_int_early(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// experiment hook: PFHE_BIGN_LOGE=4 lays the u64 N = 8192 / 16384 tables out for 16-word register tiles (twice the threads)
static int bign_loge() {
    static const int v = env_int_early("PFHE_BIGN_LOGE", 5);
    return v == 4 ? 4 : 5;
}
int choose_loge(int bits, int log_n) {
    if (bits == 64) {
        if (log_n == 13 || log_n == 14) return bign_loge();
        switch (log_n) {
            case 10: return 5;
            case 11: return 4;
            case 12: return 4;
            case 13: return 5;  // (LOGE = 4 with 512 / 1024 threads measured 10 % slower forward)
            case 14: return 5;
            default: return 0;
        }
    }
    switch (log_n) {
        case 10: return 5;
        case 11: return 6;
        case 12: return 6;
        case 13: return 5;
        case 14: return 5;
        case 15: return 5;
        default: return 0;
    }
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}


struct SyncBlock {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct SyncWarp {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
// thread group of one polynomial inside a CTA that carries several: named barrier (id 1 + group), so that the groups do
// not wait for each other at every exchange
template <int TPP> struct SyncGroup {
    int id;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TPP) : "memory"); }
};
template <int TPP> struct SyncFor {
    using type = SyncBlock;
};
template <> struct SyncFor<32> {
    using type = SyncWarp;
};

template <typename T> __device__ __forceinline__ DevNtt<T> pick_table(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, size_t poly) {
    if (limbs <= 1) return tb0;
    return tables[poly % (size_t)limbs];
}

// ------------------------------------------------------------------------------------------------
// register-pass kernels (F = IntField<T> or F64Field)
// ------------------------------------------------------------------------------------------------
// occupancy targets (CTAs per SM) for the register allocator, per configuration
template <typename F, int LOGN, int LOGE, int PPB> constexpr int ntt_min_blocks() {
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    if (sizeof(typename F::WordT) == 8 && LOGN == 12 && LOGE == 4 && PPB == 1)
        return sizeof(typename F::Elem) == 8 && !std::is_integral<typename F::Elem>::value ? 4 : 3;  // FP64: 64 regs x 4 CTAs; int: 80 x 3
    if (sizeof(typename F::WordT) == 8 && LOGN == 13 && LOGE == 4) return 2;  // 512 threads x 64 registers, two CTAs per SM
    return threads >= 512 ? 1 : 512 / threads;  // at least 16 warps per SM: cap the allocator at 128 registers
}

// polymul holds two polynomials' tiles in registers: only ask for 16 warps/SM when 2 tiles + scratch fit in 128 registers
template <typename F, int LOGN, int LOGE, int PPB> constexpr int polymul_min_blocks() {
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    constexpr int need = 2 * (1 << LOGE) * (int)(sizeof(typename F::WordT) / 4) + 48;
    return (need <= 128 && threads < 512) ? 512 / threads : 1;
}

// ---- TMA plumbing -------------------------------------------------------------------------------------------------
// Where a thread's row of the exchange buffer is 128 bytes (u64 with E = 16, u32 with E = 32) the buffer's XOR swizzle is
// exactly the SWIZZLE_128B pattern of a tensor map with box {E words, N/E rows}.  One elected thread then moves a whole
// polynomial between the swizzled buffer and its linear image in HBM with a single cp.async.bulk.tensor:
//   * store: replaces re-reading the buffer (LDS.128) + STG.128 and, above all, the store drain before the CTA can
//     retire -- the CTA only waits until the TMA unit has READ shared memory (wait_group.read);
//   * load (inverse input, arrives in bit-reversed = contiguous-per-thread order): replaces LDG.128 + STS.128 + a barrier.
// Measured on the headline kernel: 52.2 -> 61.4 M NTT/s (profiles/r01_ntt_kernel_experiments.md).
struct IoMaps {
    CUtensorMap in, out;
    const void *src;  // forward input (plain pointer: the forward transform reads with LDG)
};

// ---- LSU-path kernels (every configuration) ----------------------------------------------------------------------------
// PARAM_TB (forward, single modulus): table constants are read straight from the kernel-parameter bank instead of a
// by-value copy (measured +5..15 % on the forward LSU path; the inverse and the fused product are faster with the copy).
template <typename F, int LOGN, int LOGE, int PPB, bool FWD, bool PARAM_TB = false>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)) * PPB, ntt_min_blocks<F, LOGN, LOGE, PPB>())
ntt_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb0, const DevNtt<typename F::WordT> *__restrict__ tables, int limbs,
           const typename F::WordT *src, typename F::WordT *dst, size_t npolys) {  // src may equal dst (in place): no __restrict__
    using Core = NttCore<F, LOGN, LOGE>;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    constexpr int TPP = Core::TPP, N = Core::N, E = Core::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPP, t = threadIdx.x % TPP;
    Elem *sm = reinterpret_cast<Elem *>(smem_raw) + (size_t)grp * N;
    size_t poly = (size_t)blockIdx.x * PPB + grp;
    const bool active = poly < npolys;
    if (!active) {
        if (TPP == 32) return;  // warp-private group: nothing to synchronise with
        poly = npolys - 1;
    }
    DevNtt<T> tb_copy;
    if (!PARAM_TB) tb_copy = pick_table(tb0, tables, limbs, poly);
    const DevNtt<T> &tb = PARAM_TB ? tb0 : tb_copy;
    const typename F::Ctx c = F::ctx(tb);
    const T *g_in = src + poly * N;
    T *g_out = dst + poly * N;
    typename SyncFor<TPP>::type sync;
    Elem x[E];
    if (FWD) {
        Core::forward_g2r(g_in, x, sm, tb, c, t, sync);
        if (active) Core::fwd_regs_to_sm(x, sm, c, t);
        sync();
        if (active) Core::copy_s2g(sm, g_out, t);
    } else {
        Core::load_g2r(g_in, x, sm, c, t, sync);
        Core::template inv_from<Core::P::NPASS - 1>(x, sm, tb, c, t, sync);
        if (active) Core::inv_regs_to_global(x, g_out, c, t);
    }
}

// STASH: the register file cannot hold two tiles of 2^LOGE words (u64, E = 32: 128 registers of data alone), so fwd(a) is parked in
// the OUTPUT polynomial (canonical words, bit-reversed order; it stays in L2: one polynomial per SM is in flight) and read back by
// the thread that owns the same words after fwd(b).  Without it the N = 8192 / 16384 products spill (r01: 0.12 of the HBM peak on
// the 8-limb N = 16384 RNS product).  The host swaps a and b when c aliases b, and uses the two-tile kernel when a == b == c.
template <typename F, int LOGN, int LOGE, int PPB, bool STASH = false>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)) * PPB, STASH ? ntt_min_blocks<F, LOGN, LOGE, PPB>() : polymul_min_blocks<F, LOGN, LOGE, PPB>())
polymul_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb0, const DevNtt<typename F::WordT> *__restrict__ tables, int limbs,
               const typename F::WordT *a, const typename F::WordT *b, typename F::WordT *cc, size_t npolys) {
    using Core = NttCore<F, LOGN, LOGE>;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    constexpr int TPP = Core::TPP, N = Core::N, E = Core::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPP, t = threadIdx.x % TPP;
    Elem *sm = reinterpret_cast<Elem *>(smem_raw) + (size_t)grp * N;
    size_t poly = (size_t)blockIdx.x * PPB + grp;
    const bool active = poly < npolys;
    if (!active) {
        if (TPP == 32) return;
        poly = npolys - 1;
    }
    const DevNtt<T> tb = pick_table(tb0, tables, limbs, poly);
    const typename F::Ctx c = F::ctx(tb);
    typename SyncFor<TPP>::type sync;
    if constexpr (STASH) {
        static_assert(PPB == 1, "the stash variant runs one polynomial per CTA");
        T *g_c = cc + poly * N;
        Elem x[E];
        // (a bulk L2 prefetch of b at this point -- cp.async.bulk.prefetch.L2 -- was measured: 2.19 M against 2.78 M limb products/s)
        Core::forward_g2r(a + poly * N, x, sm, tb, c, t, sync);
        Core::fwd_regs_to_sm(x, sm, c, t);
        sync();
        Core::copy_s2g(sm, g_c, t);      // fwd(a), canonical, parked in the output polynomial
        sync();                          // buffer free again; the CTA's global writes are visible to the CTA after the barrier
        Core::forward_g2r(b + poly * N, x, sm, tb, c, t, sync);
        // this thread's words of fwd(a): the same contiguous run [t*E, (t+1)*E) it now holds of fwd(b) (coherent loads: written above)
        const T *ga = g_c + (size_t)t * E;
        constexpr int CW = Core::CW;  // 16-byte vectors, a few in flight at a time (the register file is full)
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            typename Core::WVec w;
            uint4 raw;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(ga + v * CW) : "memory");
            *reinterpret_cast<uint4 *>(&w) = raw;
#pragma unroll
            for (int k = 0; k < CW; k++) x[v * CW + k] = F::pointwise(F::load(w.v[k], c), x[v * CW + k], c);
        }
        sync();                          // every thread has read its stash before the output polynomial is overwritten
        Core::template inv_from<Core::P::NPASS - 1>(x, sm, tb, c, t, sync);
        Core::inv_regs_to_global(x, g_c, c, t);
    } else {
        Elem xa[E], xb[E];
        Core::forward_g2r(a + poly * N, xa, sm, tb, c, t, sync);
        sync();  // the exchange buffer is reused with the first pass's pattern
        Core::forward_g2r(b + poly * N, xb, sm, tb, c, t, sync);
        // pointwise product, exact mod q (BarrettModulus::reduce_mul, primus_modulus/src/barrett/ops.rs:276-283)
#pragma unroll
        for (int j = 0; j < E; j++) xa[j] = F::pointwise(xa[j], xb[j], c);
        Core::template inv_from<Core::P::NPASS - 1>(xa, sm, tb, c, t, sync);
        if (active) Core::inv_regs_to_global(xa, cc + poly * N, c, t);
    }
}


// ---- TMA variants (128-byte exchange-buffer rows only) -------------------------------------------------------------
// MULTI: per-limb tables (DCRT) are picked from the device array; otherwise every table constant is read straight from
// the kernel-parameter bank (no registers, no loads).
template <typename F, int LOGN, int LOGE, int PPB, bool FWD, bool MULTI>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)) * PPB, ntt_min_blocks<F, LOGN, LOGE, PPB>())
ntt_tma_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb0, const DevNtt<typename F::WordT> *__restrict__ tables, int limbs,
               size_t npolys, const __grid_constant__ IoMaps maps) {
    using Core = NttCore<F, LOGN, LOGE, true>;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    constexpr int TPP = Core::TPP, N = Core::N, E = Core::E;
    static_assert(Core::kTmaSwizzle, "tensor copies need the SWIZZLE_128B-compatible exchange-buffer swizzle");
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    const int grp = threadIdx.x / TPP, t = threadIdx.x % TPP;
    Elem *sm = reinterpret_cast<Elem *>(smem_tma) + (size_t)grp * N;
    size_t poly = (size_t)blockIdx.x * PPB + grp;
    const bool active = poly < npolys;
    if (!active) {
        if (TPP == 32) return;  // warp-private group: nothing to synchronise with
        poly = npolys - 1;
    }
    DevNtt<T> tb_limb;
    if (MULTI) tb_limb = tables[poly % (size_t)limbs];
    const DevNtt<T> &tb = MULTI ? tb_limb : tb0;
    const typename F::Ctx c = F::ctx(tb);
    const uint32_t row = (uint32_t)(poly * Core::kTmaRows);
    using SyncT = typename std::conditional<(PPB > 1 && TPP > 32), SyncGroup<TPP>, typename SyncFor<TPP>::type>::type;
    SyncT sync;
    if constexpr (PPB > 1 && TPP > 32) sync.id = 1 + grp;
    Elem x[E];
    if (FWD) {
        // input: strided coalesced LDG (a bulk-TMA copy-in measured no faster); output: one tensor store
        Core::forward_g2r(reinterpret_cast<const T *>(maps.src) + poly * N, x, sm, tb, c, t, sync);
        if (active) Core::fwd_regs_to_sm(x, sm, c, t);
    } else {
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem_tma + sizeof(T) * PPB * N) + grp;
        if (t == 0) {
            mbar_init(bar, 1);
            tma_load_poly(&maps.in, sm, row, bar, Core::kTmaBoxes, Core::kTmaBoxRows);
        }
        sync();  // barrier initialised before anyone polls it
        mbar_wait(bar, 0);
        Core::template sm_load<Core::P::NPASS - 1>(x, sm, t);
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load_bits(x[j], c);
        Core::template inv_from<Core::P::NPASS - 1>(x, sm, tb, c, t, sync);
        // canonical words back into the slots this thread just read (pass-0 pattern, no hazard)
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::inv_bits(x[j], c);
        if (active) Core::template sm_store<0>(x, sm, t);
    }
    fence_async_smem();
    sync();
    if (active && t == 0) tma_store_poly(&maps.out, sm, row, Core::kTmaBoxes, Core::kTmaBoxRows);
}

template <typename F, int LOGN, int LOGE, int PPB, bool MULTI>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)) * PPB, polymul_min_blocks<F, LOGN, LOGE, PPB>())
polymul_tma_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb0, const DevNtt<typename F::WordT> *__restrict__ tables, int limbs,
                   const typename F::WordT *a, const typename F::WordT *b, size_t npolys,  // the output map may alias a or b
                   const __grid_constant__ CUtensorMap out_map) {
    using Core = NttCore<F, LOGN, LOGE, true>;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    constexpr int TPP = Core::TPP, N = Core::N, E = Core::E;
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    const int grp = threadIdx.x / TPP, t = threadIdx.x % TPP;
    Elem *sm = reinterpret_cast<Elem *>(smem_tma) + (size_t)grp * N;
    size_t poly = (size_t)blockIdx.x * PPB + grp;
    const bool active = poly < npolys;
    if (!active) {
        if (TPP == 32) return;
        poly = npolys - 1;
    }
    DevNtt<T> tb_limb;
    if (MULTI) tb_limb = tables[poly % (size_t)limbs];
    const DevNtt<T> &tb = MULTI ? tb_limb : tb0;
    const typename F::Ctx c = F::ctx(tb);
    typename SyncFor<TPP>::type sync;
    Elem xa[E], xb[E];
    Core::forward_g2r(a + poly * N, xa, sm, tb, c, t, sync);
    sync();  // the exchange buffer is reused with the first pass's pattern
    Core::forward_g2r(b + poly * N, xb, sm, tb, c, t, sync);
#pragma unroll
    for (int j = 0; j < E; j++) xa[j] = F::pointwise(xa[j], xb[j], c);
    Core::template inv_from<Core::P::NPASS - 1>(xa, sm, tb, c, t, sync);
#pragma unroll
    for (int j = 0; j < E; j++) xa[j] = F::inv_bits(xa[j], c);
    if (active) Core::template sm_store<0>(xa, sm, t);
    fence_async_smem();
    sync();
    if (active && t == 0) tma_store_poly(&out_map, sm, (uint32_t)(poly * Core::kTmaRows), Core::kTmaBoxes, Core::kTmaBoxRows);
}

// ---- persistent forward kernel for the one-polynomial-per-SM sizes (u64, N = 8192 / 16384) ------------------------------------------
// At N = 16384 the register file holds exactly one polynomial (512 threads x 32 doubles) and shared memory one exchange buffer, so
// with one CTA per polynomial the load, FP64 and store/exit phases of an SM never overlap (ncu, profiles/r02_ncu_ntt_fwd_n16384.txt:
// FP64 pipe 57 % busy, 24 % of the stall samples wait for the loads, 17 % sit at EXIT).  Here a CTA is persistent and the NEXT
// polynomial is prefetched by bulk TMA copies into a staging buffer `in` while the current one is being transformed:
//   shared memory = in[N] (raw words of the next polynomial) + xb[N/2] -- 192 KiB at N = 16384;
//   both exchanges use xb for the lower half of the index space and the FIRST half of `in` for the upper half; the staging buffer is
//   free between the pass-0 loads and the prefetch, which therefore comes in two pieces: the second half right after the pass-0
//   loads, the first half after the loads of exchange 2 (it lands while pass 2 runs and the results are stored);
//   outputs go from registers to HBM as 128-bit streaming stores (each thread owns 256 contiguous bytes).
template <typename F, int LOGN, int LOGE>
__global__ void __launch_bounds__((1 << (LOGN - LOGE)), (LOGN <= 13 ? 2 : 1))
ntt_persist_fwd_kernel(const __grid_constant__ DevNtt<typename F::WordT> tb0, const DevNtt<typename F::WordT> *__restrict__ tables, int limbs,
                       const typename F::WordT *src, typename F::WordT *dst, size_t npolys) {  // src may equal dst
    using Core = NttCore<F, LOGN, LOGE>;
    using P = typename Core::P;
    using T = typename F::WordT;
    using Elem = typename F::Elem;
    static_assert(sizeof(T) == 8 && P::NPASS == 3 && LOGE == 5, "built for u64 words, three register passes of 2^5 words");
    constexpr int N = Core::N, E = Core::E, HALF = N / 2, CW = Core::CW;
    constexpr int FB0 = P::fb(0), FB1 = P::fb(1);
    constexpr uint32_t kHalfBytes = sizeof(T) * HALF;
    extern __shared__ __align__(1024) unsigned char smem_p[];
    T *in = reinterpret_cast<T *>(smem_p);                        // [N] raw words of the polynomial being (pre)fetched
    Elem *inx = reinterpret_cast<Elem *>(smem_p);                 // its first half doubles as the upper half of the exchange buffer
    Elem *xb = reinterpret_cast<Elem *>(smem_p + sizeof(T) * N);  // [N/2] lower half of the exchange buffer
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_p + sizeof(T) * (N + HALF));
    const int t = threadIdx.x;
    const int half = t >> (LOGN - LOGE - 1);                      // which half of the index space this thread owns from pass 1 on
    size_t poly = blockIdx.x;
    if (poly >= npolys) return;
    auto copy_half = [&](const T *g, int which) {  // elected thread: 64 KiB (N = 16384) from HBM into one half of the staging buffer
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(in + which * HALF)),
                     "l"(g + which * HALF), "r"(kHalfBytes), "r"(smem_addr(bar))
                     : "memory");
    };
    auto expect_poly = [&]() {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(2u * kHalfBytes) : "memory");
    };
    if (t == 0) {
        mbar_init(bar, 1);
        expect_poly();
        copy_half(src + poly * N, 0);
        copy_half(src + poly * N, 1);
    }
    __syncthreads();
    uint32_t phase = 0;
    for (; poly < npolys; poly += gridDim.x) {
        DevNtt<T> tb_copy;
        if (limbs > 1) tb_copy = tables[poly % (size_t)limbs];
        const DevNtt<T> &tb = limbs > 1 ? tb_copy : tb0;
        const typename F::Ctx c = F::ctx(tb);
        const size_t next = poly + gridDim.x;
        Elem x[E];
        mbar_wait(bar, phase);
        phase ^= 1u;
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = F::load(in[Core::elem_index(FB0, t, j)], c);
        __syncthreads();  // every thread holds its words: the staging buffer is free
        if (t == 0 && next < npolys) {
            fence_async_smem();
            expect_poly();
            copy_half(src + next * N, 1);  // the half the exchanges never touch
        }
        Core::template fwd_pass_regs<0>(x, tb, c, t);
        // exchange 1 (bit LOGN-1 of the index is bit LOGE-1 of j on the storing side, a thread bit on the loading side)
#pragma unroll
        for (int j = 0; j < E; j++) {
            Elem *buf = (j >> (LOGE - 1)) ? inx : xb;
            buf[Core::swz(Core::elem_index(FB0, t, j) & (HALF - 1))] = x[j];
        }
        __syncthreads();
        Elem *mine = half ? inx : xb;
#pragma unroll
        for (int j = 0; j < E; j++) x[j] = mine[Core::swz(Core::elem_index(FB1, t, j) & (HALF - 1))];
        Core::template fwd_pass_regs<1>(x, tb, c, t);
        // exchange 2: a thread stores to the slots it loaded from, so no barrier is needed in between
#pragma unroll
        for (int j = 0; j < E; j++) mine[Core::swz(Core::elem_index(FB1, t, j) & (HALF - 1))] = x[j];
        __syncthreads();
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            const typename Core::Vec w = *reinterpret_cast<const typename Core::Vec *>(mine + Core::swz(((t * E) & (HALF - 1)) + v * CW));
#pragma unroll
            for (int k = 0; k < CW; k++) x[v * CW + k] = w.v[k];
        }
        __syncthreads();  // exchange buffers are free: the first half of the next polynomial may land
        if (t == 0 && next < npolys) {
            fence_async_smem();
            copy_half(src + next * N, 0);
        }
        Core::template fwd_pass_regs<2>(x, tb, c, t);
        // canonical words straight to HBM: this thread's 2^LOGE contiguous output words
        T *o = dst + poly * N + (size_t)t * E;
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            typename Core::WVec w;
#pragma unroll
            for (int k = 0; k < CW; k++) w.v[k] = F::fwd_word(x[v * CW + k], c);
            stg_stream(reinterpret_cast<uint4 *>(o + v * CW), *reinterpret_cast<const uint4 *>(&w));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic radix-2 kernel: any 1 <= log_n that fits shared memory (sizes without a register-pass
// instantiation, e.g. the reference's small round-trip tests N = 8..512)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void ntt_generic_kernel(const __grid_constant__ DevNtt<T> tb0, const DevNtt<T> *__restrict__ tables, int limbs,
                                   const T *src, T *dst, size_t npolys, int forward) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const size_t poly = blockIdx.x;
    const DevNtt<T> tb = pick_table(tb0, tables, limbs, poly);
    const int logn = tb.log_n, n = 1 << logn, half = n >> 1;
    const T q = tb.q, two_q = tb.two_q;
    const T *g_in = src + poly * n;
    T *g_out = dst + poly * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = g_in[i];
    __syncthreads();
    if (forward) {
        for (int s = 0; s < logn; s++) {
            const int lg = logn - 1 - s, gap = 1 << lg;
            for (int bidx = threadIdx.x; bidx < half; bidx += blockDim.x) {
                const int blk = bidx >> lg, j = bidx & (gap - 1), i0 = (blk << (lg + 1)) | j;
                const auto w = ld_pair<T>(tb.fwd + ((1 << s) + blk));
                T x = sm[i0], y = sm[i0 + gap];
                fwd_bfly<T>(x, y, w.x, w.y, q, two_q);
                sm[i0] = x;
                sm[i0 + gap] = y;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) g_out[i] = csub(csub(sm[i], two_q), q);
    } else {
        for (int lg = 0; lg < logn; lg++) {
            const int gap = 1 << lg, base = 1 + n - (n >> lg);
            for (int bidx = threadIdx.x; bidx < half; bidx += blockDim.x) {
                const int blk = bidx >> lg, j = bidx & (gap - 1), i0 = (blk << (lg + 1)) | j;
                const auto w = ld_pair<T>(tb.inv + (base + blk));
                T x = sm[i0], y = sm[i0 + gap];
                if (lg == logn - 1) {
                    T tx = x + y, ty = x + two_q - y;
                    x = shoup<T>(tx, tb.inv_n, tb.inv_n_q, q);
                    y = shoup<T>(ty, w.x, w.y, q);
                } else {
                    inv_bfly<T>(x, y, w.x, w.y, q, two_q);
                }
                sm[i0] = x;
                sm[i0 + gap] = y;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) g_out[i] = sm[i];
    }
}

template <typename T>
__global__ void polymul_generic_kernel(const __grid_constant__ DevNtt<T> tb0, const DevNtt<T> *__restrict__ tables, int limbs,
                                       const T *a, const T *b, T *c, size_t npolys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t poly = blockIdx.x;
    const DevNtt<T> tb = pick_table(tb0, tables, limbs, poly);
    const int logn = tb.log_n, n = 1 << logn, half = n >> 1;
    T *sa = reinterpret_cast<T *>(smem_raw), *sb = sa + n;
    const T q = tb.q, two_q = tb.two_q;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sa[i] = a[poly * n + i];
        sb[i] = b[poly * n + i];
    }
    __syncthreads();
    for (int s = 0; s < logn; s++) {
        const int lg = logn - 1 - s, gap = 1 << lg;
        for (int bidx = threadIdx.x; bidx < half; bidx += blockDim.x) {
            const int blk = bidx >> lg, j = bidx & (gap - 1), i0 = (blk << (lg + 1)) | j;
            const auto w = ld_pair<T>(tb.fwd + ((1 << s) + blk));
            fwd_bfly<T>(sa[i0], sa[i0 + gap], w.x, w.y, q, two_q);
            fwd_bfly<T>(sb[i0], sb[i0 + gap], w.x, w.y, q, two_q);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T x = csub(csub(sa[i], two_q), q), y = csub(csub(sb[i], two_q), q);
        sa[i] = barrett_mul<T>(tb.br, x, y);
    }
    __syncthreads();
    for (int lg = 0; lg < logn; lg++) {
        const int gap = 1 << lg, base = 1 + n - (n >> lg);
        for (int bidx = threadIdx.x; bidx < half; bidx += blockDim.x) {
            const int blk = bidx >> lg, j = bidx & (gap - 1), i0 = (blk << (lg + 1)) | j;
            const auto w = ld_pair<T>(tb.inv + (base + blk));
            T x = sa[i0], y = sa[i0 + gap];
            if (lg == logn - 1) {
                T tx = x + y, ty = x + two_q - y;
                x = shoup<T>(tx, tb.inv_n, tb.inv_n_q, q);
                y = shoup<T>(ty, w.x, w.y, q);
            } else {
                inv_bfly<T>(x, y, w.x, w.y, q, two_q);
            }
            sa[i0] = x;
            sa[i0 + gap] = y;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) c[poly * n + i] = sa[i];
}

// NTT(coeff * X^degree)[i] = coeff * psi^(((2 brv(i)+1) degree) mod 2N)   (table.rs:565-609)
template <typename T>
__global__ void monomial_kernel(const __grid_constant__ DevNtt<T> tb, T coeff, T coeff_q, const uint32_t *__restrict__ degrees,
                                T *__restrict__ out, size_t batch) {
    const int logn = tb.log_n, n = 1 << logn;
    const size_t total = batch * (size_t)n;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t bidx = gid >> logn;
        const uint32_t i = (uint32_t)(gid & (n - 1));
        const uint32_t deg = degrees[bidx] & (2 * n - 1);
        T v;
        if (coeff == 0) {
            v = 0;
        } else {
            const uint32_t r = __brev(i) >> (32 - logn);
            const uint32_t e = ((2 * r + 1) * deg) & (2 * n - 1);
            const T w = tb.ordinal[e];
            v = (coeff == 1) ? w : shoup<T>(w, coeff, coeff_q, tb.q);
        }
        out[gid] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}


// tensor map over a batch of polynomials viewed as 128-byte rows: {128/w, npolys * N*w/128}, box {128/w, min(rows, 256)}, SWIZZLE_128B
template <typename T> bool make_poly_map(CUtensorMap *map, const T *base, size_t npolys, int log_n) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    const uint64_t e = 128 / sizeof(T);  // words per 128-byte row
    const uint64_t rows = ((uint64_t)1 << log_n) / e, total_rows = (uint64_t)npolys * rows;
    if (!enc || total_rows > 0xffffffffull || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)e, (cuuint64_t)total_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)128};
    const cuuint32_t box[2] = {(cuuint32_t)e, (cuuint32_t)(rows < 256 ? rows : 256)};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<T *>(base), gdim, gstride, box,
               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template bool make_poly_map<uint32_t>(CUtensorMap *, const uint32_t *, size_t, int);
template bool make_poly_map<uint64_t>(CUtensorMap *, const uint64_t *, size_t, int);

template <typename F, int LOGN, int LOGE, int PPB>
static cudaError_t run_ntt_f(const DevNtt<typename F::WordT> &tb0, const DevNtt<typename F::WordT> *tables, int limbs,
                             const typename F::WordT *src, typename F::WordT *dst, size_t npolys, bool fwd, cudaStream_t stream) {
    using T = typename F::WordT;
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    constexpr size_t smem = sizeof(T) * PPB * ((size_t)1 << LOGN);
    const unsigned grid = (unsigned)((npolys + PPB - 1) / PPB);
    cudaError_t e;
    // tensor copies: every direction where a thread row is 128 bytes; forward only for 256-byte rows (measured: +4..9 %
    // forward, but the 2-way conflicted row pattern costs the inverse and the fused product more than the copy-out saves)
    constexpr bool kRow128 = sizeof(T) * (1 << LOGE) == 128;
    if constexpr (sizeof(T) == 8 && LOGE == 5 && LOGN >= 13 && PPB == 1) {
        static const int persist = env_int("PFHE_NTT_PERSIST", 0);  // experiment, off: measured 8.2 M vs 9.2 M NTT/s at N = 16384 (profiles/r02_large_n_experiments.md)
        // in place is fine: a CTA only ever prefetches polynomials it will itself overwrite later
        if (fwd && persist && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            auto k = ntt_persist_fwd_kernel<F, LOGN, LOGE>;
            constexpr size_t bytes = sizeof(T) * (((size_t)1 << LOGN) + ((size_t)1 << (LOGN - 1))) + 64;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)) != cudaSuccess) return e;
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, bytes);
            if (per_sm < 1) per_sm = 1;
            const size_t slots = (size_t)sms * per_sm;
            const unsigned g = (unsigned)(npolys < slots ? npolys : slots);
            k<<<g, threads, bytes, stream>>>(tb0, tables, limbs, src, dst, npolys);
            count_launch();
            return cudaGetLastError();
        }
    }
    {
        static const bool use_tma = env_int("PFHE_NTT_TMA", 1) != 0;  // A/B tuning hook
        IoMaps maps;
        maps.src = src;
        if (use_tma && (kRow128 || fwd) && make_poly_map<T>(&maps.out, dst, npolys, LOGN) && (fwd || make_poly_map<T>(&maps.in, src, npolys, LOGN))) {
            if (fwd) memset(&maps.in, 0, sizeof(maps.in));
            auto launch = [&](auto k, size_t bytes) -> cudaError_t {
                if (bytes > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)) != cudaSuccess) return e;
                k<<<grid, threads, bytes, stream>>>(tb0, tables, limbs, npolys, maps);
                count_launch();
                return cudaGetLastError();
            };
            if (limbs > 1) {
                if (fwd) return launch(ntt_tma_kernel<F, LOGN, LOGE, PPB, true, true>, smem);
                return launch(ntt_tma_kernel<F, LOGN, LOGE, PPB, false, true>, smem + 8 * PPB);
            }
            if (fwd) return launch(ntt_tma_kernel<F, LOGN, LOGE, PPB, true, false>, smem);
            return launch(ntt_tma_kernel<F, LOGN, LOGE, PPB, false, false>, smem + 8 * PPB);
        }
    }
    if (fwd && limbs <= 1) {
        auto k = ntt_kernel<F, LOGN, LOGE, PPB, true, true>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<grid, threads, smem, stream>>>(tb0, tables, limbs, src, dst, npolys);
    } else if (fwd) {
        auto k = ntt_kernel<F, LOGN, LOGE, PPB, true>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<grid, threads, smem, stream>>>(tb0, tables, limbs, src, dst, npolys);
    } else {
        auto k = ntt_kernel<F, LOGN, LOGE, PPB, false>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<grid, threads, smem, stream>>>(tb0, tables, limbs, src, dst, npolys);
    }
    count_launch();
    return cudaGetLastError();
}
template <typename F, int LOGN, int LOGE, int PPB>
static cudaError_t run_polymul_f(const DevNtt<typename F::WordT> &tb0, const DevNtt<typename F::WordT> *tables, int limbs,
                                 const typename F::WordT *a, const typename F::WordT *b, typename F::WordT *c, size_t npolys,
                                 cudaStream_t stream) {
    using T = typename F::WordT;
    constexpr int threads = (1 << (LOGN - LOGE)) * PPB;
    constexpr size_t smem = sizeof(T) * PPB * ((size_t)1 << LOGN);
    const unsigned grid = (unsigned)((npolys + PPB - 1) / PPB);
    cudaError_t e;
    if constexpr (sizeof(T) == 8 && LOGN >= 13 && PPB == 1) {
        static const bool use_stash = env_int("PFHE_POLYMUL_STASH", 1) != 0;  // A/B tuning hook
        if (c == b) {  // the stash would overwrite b before it is read: the product commutes
            const T *tmp = a;
            a = b;
            b = tmp;
        }
        if (use_stash && c != b) {
            auto ks = polymul_kernel<F, LOGN, LOGE, PPB, true>;
            if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            ks<<<grid, threads, smem, stream>>>(tb0, tables, limbs, a, b, c, npolys);
            count_launch();
            return cudaGetLastError();
        }
    }
    if constexpr (sizeof(T) * (1 << LOGE) == 128 && !(sizeof(T) == 8 && LOGN >= 13)) {
        static const bool use_tma = env_int("PFHE_NTT_TMA", 1) != 0;
        CUtensorMap map;
        if (use_tma && make_poly_map<T>(&map, c, npolys, LOGN)) {
            auto launch = [&](auto k) -> cudaError_t {
                if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
                k<<<grid, threads, smem, stream>>>(tb0, tables, limbs, a, b, npolys, map);
                count_launch();
                return cudaGetLastError();
            };
            if (limbs > 1) return launch(polymul_tma_kernel<F, LOGN, LOGE, PPB, true>);
            return launch(polymul_tma_kernel<F, LOGN, LOGE, PPB, false>);
        }
    }
    auto k = polymul_kernel<F, LOGN, LOGE, PPB>;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, threads, smem, stream>>>(tb0, tables, limbs, a, b, c, npolys);
    count_launch();
    return cudaGetLastError();
}

// field selection: u64 tables with q < 2^50 run on the FP64 pipe, everything else on the integer pipe.
// PFHE_F64_LAZY=0 selects the per-stage-fold FP64 butterflies (A/B tuning hook).
template <typename T, int LOGN, int LOGE, int PPB>
static cudaError_t run_ntt(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, const T *src, T *dst, size_t npolys, bool fwd,
                           cudaStream_t stream) {
    if constexpr (sizeof(T) == 8) {
        static const bool lazy = env_int("PFHE_F64_LAZY", 1) != 0;
        if (tb0.use_f64 && lazy) return run_ntt_f<F64LazyField, LOGN, LOGE, PPB>(tb0, tables, limbs, src, dst, npolys, fwd, stream);
        if (tb0.use_f64) return run_ntt_f<F64Field, LOGN, LOGE, PPB>(tb0, tables, limbs, src, dst, npolys, fwd, stream);
    }
    return run_ntt_f<IntField<T>, LOGN, LOGE, PPB>(tb0, tables, limbs, src, dst, npolys, fwd, stream);
}
template <typename T, int LOGN, int LOGE, int PPB>
static cudaError_t run_polymul(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, const T *a, const T *b, T *c, size_t npolys,
                               cudaStream_t stream) {
    if constexpr (sizeof(T) == 8) {
        static const bool lazy = env_int("PFHE_F64_LAZY", 1) != 0;
        if (tb0.use_f64 && lazy) return run_polymul_f<F64LazyField, LOGN, LOGE, PPB>(tb0, tables, limbs, a, b, c, npolys, stream);
        if (tb0.use_f64) return run_polymul_f<F64Field, LOGN, LOGE, PPB>(tb0, tables, limbs, a, b, c, npolys, stream);
    }
    return run_polymul_f<IntField<T>, LOGN, LOGE, PPB>(tb0, tables, limbs, a, b, c, npolys, stream);
}

template <typename T> static int generic_threads(int log_n) {
    int half = 1 << (log_n - 1);
    int th = half < 32 ? 32 : half;
    return th > 512 ? 512 : th;
}

template <typename T>
static cudaError_t run_generic(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, const T *src, T *dst, size_t npolys, bool fwd,
                               cudaStream_t stream) {
    const size_t smem = sizeof(T) << tb0.log_n;
    auto k = ntt_generic_kernel<T>;
    cudaError_t e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)npolys, generic_threads<T>(tb0.log_n), smem, stream>>>(tb0, tables, limbs, src, dst, npolys, fwd ? 1 : 0);
    count_launch();
    return cudaGetLastError();
}

template <>
cudaError_t launch_ntt<uint64_t>(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, const uint64_t *src,
                                 uint64_t *dst, size_t npolys, bool fwd, cudaStream_t s) {
    using T = uint64_t;
    if (npolys == 0) return cudaSuccess;
    if (tb0.loge != 0) {
        switch (tb0.log_n) {
            case 10: return run_ntt<T, 10, 5, 4>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 11: return run_ntt<T, 11, 4, 2>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 12: return run_ntt<T, 12, 4, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 13: {
                static const bool cluster13 = env_int("PFHE_NTT_CLUSTER13", 0) != 0 && env_int("PFHE_F64_LAZY", 1) != 0;  // N = 8192 on the cluster kernels
                if (cluster13) {
                    const cudaError_t ce = launch_ntt_cluster(tb0, tables, limbs, fwd ? 0 : 1, src, nullptr, dst, npolys, s);
                    if (ce != cudaErrorNotSupported) return ce;
                }
                return tb0.loge == 4 ? run_ntt<T, 13, 4, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s)
                                     : run_ntt<T, 13, 5, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            }
            case 14: {
                // 2-CTA cluster per polynomial (ntt_cluster.cu): two half-polynomial CTAs resident per SM, cross-CTA exchange by st.async +
                // mbarriers.  Measured: forward 9.98 M against 9.32 M NTT/s, 8-limb fused product 353 K against 331 K (266 K inside the bench
                // sequence).  PFHE_NTT_CLUSTER: 0 = one CTA per polynomial, 1 = transforms only, 2 (default) = transforms and fused product
                static const bool cluster = env_int("PFHE_NTT_CLUSTER", 2) != 0 && env_int("PFHE_F64_LAZY", 1) != 0;
                if (cluster) {
                    const cudaError_t ce = launch_ntt_cluster(tb0, tables, limbs, fwd ? 0 : 1, src, nullptr, dst, npolys, s);
                    if (ce != cudaErrorNotSupported) return ce;
                }
                return tb0.loge == 4 ? run_ntt<T, 14, 4, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s)
                                     : run_ntt<T, 14, 5, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            }
        }
    }
    return run_generic<T>(tb0, tables, limbs, src, dst, npolys, fwd, s);
}
template <>
cudaError_t launch_ntt<uint32_t>(const DevNtt<uint32_t> &tb0, const DevNtt<uint32_t> *tables, int limbs, const uint32_t *src,
                                 uint32_t *dst, size_t npolys, bool fwd, cudaStream_t s) {
    using T = uint32_t;
    if (npolys == 0) return cudaSuccess;
    if (tb0.loge != 0) {
        switch (tb0.log_n) {
            case 10: return run_ntt<T, 10, 5, 4>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 11: return run_ntt<T, 11, 6, 4>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 12: return run_ntt<T, 12, 6, 2>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 13: return run_ntt<T, 13, 5, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 14: return run_ntt<T, 14, 5, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s);
            case 15: return run_ntt<T, 15, 5, 1>(tb0, tables, limbs, src, dst, npolys, fwd, s);
        }
    }
    return run_generic<T>(tb0, tables, limbs, src, dst, npolys, fwd, s);
}

template <typename T>
static cudaError_t run_polymul_generic(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, const T *a, const T *b, T *c,
                                       size_t npolys, cudaStream_t stream) {
    const size_t smem = 2 * (sizeof(T) << tb0.log_n);
    auto k = polymul_generic_kernel<T>;
    cudaError_t e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<(unsigned)npolys, generic_threads<T>(tb0.log_n), smem, stream>>>(tb0, tables, limbs, a, b, c, npolys);
    count_launch();
    return cudaGetLastError();
}

template <>
cudaError_t launch_polymul<uint64_t>(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, const uint64_t *a,
                                     const uint64_t *b, uint64_t *c, size_t npolys, cudaStream_t s) {
    using T = uint64_t;
    if (npolys == 0) return cudaSuccess;
    if (tb0.loge != 0) {
        switch (tb0.log_n) {
            case 10: return run_polymul<T, 10, 5, 4>(tb0, tables, limbs, a, b, c, npolys, s);
            case 11: return run_polymul<T, 11, 4, 2>(tb0, tables, limbs, a, b, c, npolys, s);
            case 12: return run_polymul<T, 12, 4, 1>(tb0, tables, limbs, a, b, c, npolys, s);
            case 13: {
                static const bool cluster13 = env_int("PFHE_NTT_CLUSTER13", 0) == 2 && env_int("PFHE_F64_LAZY", 1) != 0;
                if (cluster13 && !(c == b && c == a)) {
                    const bool swap = c == b;
                    const cudaError_t ce = launch_ntt_cluster(tb0, tables, limbs, 2, swap ? b : a, swap ? a : b, c, npolys, s);
                    if (ce != cudaErrorNotSupported) return ce;
                }
                return tb0.loge == 4 ? run_polymul<T, 13, 4, 1>(tb0, tables, limbs, a, b, c, npolys, s)
                                     : run_polymul<T, 13, 5, 1>(tb0, tables, limbs, a, b, c, npolys, s);
            }
            case 14: {
                static const bool cluster = env_int("PFHE_NTT_CLUSTER", 2) == 2 && env_int("PFHE_F64_LAZY", 1) != 0;
                if (cluster && !(c == b && c == a)) {  // the parked fwd(a) lives in c: c == b swaps the (commuting) operands, a == b == c cannot park
                    const bool swap = c == b;
                    const cudaError_t ce = launch_ntt_cluster(tb0, tables, limbs, 2, swap ? b : a, swap ? a : b, c, npolys, s);
                    if (ce != cudaErrorNotSupported) return ce;
                }
                return tb0.loge == 4 ? run_polymul<T, 14, 4, 1>(tb0, tables, limbs, a, b, c, npolys, s)
                                     : run_polymul<T, 14, 5, 1>(tb0, tables, limbs, a, b, c, npolys, s);
            }
        }
    }
    return run_polymul_generic<T>(tb0, tables, limbs, a, b, c, npolys, s);
}
template <>
cudaError_t launch_polymul<uint32_t>(const DevNtt<uint32_t> &tb0, const DevNtt<uint32_t> *tables, int limbs, const uint32_t *a,
                                     const uint32_t *b, uint32_t *c, size_t npolys, cudaStream_t s) {
    using T = uint32_t;
    if (npolys == 0) return cudaSuccess;
    if (tb0.loge != 0) {
        switch (tb0.log_n) {
            case 10: return run_polymul<T, 10, 5, 4>(tb0, tables, limbs, a, b, c, npolys, s);
            case 11: return run_polymul<T, 11, 6, 4>(tb0, tables, limbs, a, b, c, npolys, s);
            case 12: return run_polymul<T, 12, 6, 2>(tb0, tables, limbs, a, b, c, npolys, s);
            case 13: return run_polymul<T, 13, 5, 1>(tb0, tables, limbs, a, b, c, npolys, s);
            case 14: return run_polymul<T, 14, 5, 1>(tb0, tables, limbs, a, b, c, npolys, s);
            case 15: return run_polymul<T, 15, 5, 1>(tb0, tables, limbs, a, b, c, npolys, s);
        }
    }
    return run_polymul_generic<T>(tb0, tables, limbs, a, b, c, npolys, s);
}

template <typename T>
cudaError_t launch_monomial(const DevNtt<T> &tb0, T coeff, const uint32_t *degrees, T *out, size_t batch, cudaStream_t stream) {
    if (batch == 0) return cudaSuccess;
    const size_t total = batch << tb0.log_n;
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    const T cq = coeff ? host::shoup_quot<T>(coeff, tb0.q) : 0;
    monomial_kernel<T><<<grid, 256, 0, stream>>>(tb0, coeff, cq, degrees, out, batch);
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_monomial<uint32_t>(const DevNtt<uint32_t> &, uint32_t, const uint32_t *, uint32_t *, size_t, cudaStream_t);
template cudaError_t launch_monomial<uint64_t>(const DevNtt<uint64_t> &, uint64_t, const uint32_t *, uint64_t *, size_t, cudaStream_t);

}  // namespace pfhe
