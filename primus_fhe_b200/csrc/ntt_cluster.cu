// ntt_cluster.cu -- N = 16384 transforms (u64 words, q < 2^50, FP64 lazy-fold butterflies) on a 2-CTA thread-block cluster.
//
// Why: with one CTA per polynomial the register file (512 threads x 32 doubles) and shared memory (one 128 KiB exchange buffer) of an SM
// hold exactly ONE polynomial, so the load, FP64 and store phases of an SM never overlap (profiles/r02_ncu_ntt_fwd_n16384.txt: FP64 pipe
// 57 %, 24 % of the stall samples wait for the loads, 17 % sit at EXIT).  Here a polynomial is split over the two CTAs of a cluster:
// each CTA runs 256 of the 512 threads, keeps 32 doubles per thread and a 64 KiB buffer -- so TWO CTAs of different clusters are resident
// per SM and one's memory phases hide behind the other's arithmetic.
//
// Index algebra (Plan<14, 5>: passes of 4 | 5 | 5 stages over index bits 13..10 | 9..5 | 4..0; T = cluster-wide thread id, 9 bits):
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
constexpr int LOGN = 14, LOGE = 5;
using Core = NttCore<F, LOGN, LOGE>;
using P = Core::P;
constexpr int N = 1 << LOGN, E = 1 << LOGE, HALF = N / 2, TPC = 256, CW = 2;
static_assert(P::NPASS == 3 && P::fb(0) == 9 && P::fb(1) == 5 && P::fb(2) == 0, "index algebra above assumes Plan<14, 5>");

// The exchange-buffer swizzle of NttCore only uses index bits 1..3 and 5..7, so a CTA-local buffer addressed with (index & 8191) keeps
// the conflict-free patterns of the single-CTA kernels.
__device__ __forceinline__ int lswz(int idx) { return Core::swz(idx) & (HALF - 1); }

// forward: global (natural order) -> registers of the last pass (lazy form); T = cluster-wide thread id, buf = this CTA's 64 KiB buffer
__device__ __forceinline__ void forward_g2r(const T *g, double (&x)[E], double *buf, const DevNtt<T> &tb, const F::Ctx &c, int Tg, uint32_t rank) {
#pragma unroll
    for (int j = 0; j < E; j++) x[j] = F::load(ldg_stream(g + j * 512 + Tg), c);
    Core::template fwd_pass_regs<0>(x, tb, c, Tg);
    // exchange 0 -> 1 across the pair: word j goes to the CTA that owns index bit 13 == j >> 4
    const uint32_t base = smem_u32(buf);
    const uint32_t peer = map_to_cta(base, rank ^ 1u), mine = map_to_cta(base, rank);
#pragma unroll
    for (int j = 0; j < E; j++) {
        const uint32_t dst = ((uint32_t)(j >> 4) == rank) ? mine : peer;
        st_cluster(dst + 8u * (uint32_t)lswz(j * 512 + Tg), x[j]);
    }
    cluster_sync();
#pragma unroll
    for (int j = 0; j < E; j++) x[j] = buf[lswz(Core::elem_index(5, Tg, j))];
    Core::template fwd_pass_regs<1>(x, tb, c, Tg);
#pragma unroll
    for (int j = 0; j < E; j++) buf[lswz(Core::elem_index(5, Tg, j))] = x[j];   // the slots this thread just read
    __syncwarp();                                                                // exchange 1 -> 2 stays inside a warp (32 threads share T >> 5)
#pragma unroll
    for (int v = 0; v < E / CW; v++) {
        const double2 w = *reinterpret_cast<const double2 *>(buf + lswz(Tg * E + v * CW));
        x[v * CW] = w.x;
        x[v * CW + 1] = w.y;
    }
    Core::template fwd_pass_regs<2>(x, tb, c, Tg);
}

// inverse: registers of the last pass (inverse-transform inputs) -> registers of pass 0 in final form (F::inv_word gives the canonical word)
// buf must not be in use by the peer for anything else between the two cluster barriers inside.
__device__ __forceinline__ void inverse_r2r(double (&x)[E], double *buf, const DevNtt<T> &tb, const F::Ctx &c, int Tg, uint32_t rank) {
    Core::template inv_pass_regs<2>(x, tb, c, Tg);
#pragma unroll
    for (int v = 0; v < E / CW; v++) *reinterpret_cast<double2 *>(buf + lswz(Tg * E + v * CW)) = make_double2(x[v * CW], x[v * CW + 1]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < E; j++) x[j] = buf[lswz(Core::elem_index(5, Tg, j))];
    Core::template inv_pass_regs<1>(x, tb, c, Tg);
    // exchange 1 -> 0 across the pair.  Every thread of BOTH CTAs must have finished reading its pass-1 words before anyone overwrites a
    // buffer with the pass-0 layout [k][T & 255] (k = index bits 13..9, 32 x 256 doubles): index i belongs to pass-0 thread i & 511.
    cluster_sync();
    const uint32_t base = smem_u32(buf);
    const uint32_t peer = map_to_cta(base, rank ^ 1u), mine = map_to_cta(base, rank);
#pragma unroll
    for (int j = 0; j < E; j++) {
        const int i = Core::elem_index(5, Tg, j);
        const uint32_t owner = (uint32_t)(i >> 8) & 1u;
        const int slot = ((i >> 9) << 8) | (i & 255);
        st_cluster((owner == rank ? mine : peer) + 8u * (uint32_t)slot, x[j]);
    }
    cluster_sync();
#pragma unroll
    for (int j = 0; j < E; j++) x[j] = buf[(j << 8) | (Tg & 255)];
    Core::template inv_pass_regs<0>(x, tb, c, Tg);
}

// this CTA's half of the bit-reversed (= contiguous per thread) canonical words: buffer (pass-2 layout) -> global, coalesced 16-byte stores
__device__ __forceinline__ void copy_half_s2g(const double *buf, T *g_half, int t) {
#pragma unroll
    for (int v = t; v < HALF / CW; v += TPC) {
        const uint4 d = *reinterpret_cast<const uint4 *>(buf + lswz(v * CW));
        stg_stream(reinterpret_cast<uint4 *>(g_half + v * CW), d);
    }
}
__device__ __forceinline__ void copy_half_g2s(const T *g_half, double *buf, int t) {
#pragma unroll
    for (int v = t; v < HALF / CW; v += TPC) {
        const uint4 d = ldg_stream(reinterpret_cast<const uint4 *>(g_half + v * CW));
        *reinterpret_cast<uint4 *>(buf + lswz(v * CW)) = d;
    }
}

// MODE 0: forward, 1: inverse, 2: fused product c = a * b (fwd(a) parked in the output polynomial, as in polymul_kernel<STASH>)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TPC, 2)
ntt_cluster_kernel(const __grid_constant__ DevNtt<T> tb0, const DevNtt<T> *__restrict__ tables, int limbs, const T *a, const T *b, T *out,
                   size_t npolys) {  // out may alias a (in place): no __restrict__
    extern __shared__ __align__(16) unsigned char smem_c[];
    double *buf = reinterpret_cast<double *>(smem_c);
    const uint32_t rank = cluster_rank();
    const int t = threadIdx.x, Tg = (int)rank * TPC + t;
    const size_t poly = blockIdx.x >> 1;
    DevNtt<T> tb_copy;
    if (limbs > 1) tb_copy = tables[poly % (size_t)limbs];
    const DevNtt<T> &tb = limbs > 1 ? tb_copy : tb0;
    const F::Ctx c = F::ctx(tb);
    double x[E];
    cluster_sync();  // the peer CTA is resident before its shared memory is addressed
    if (MODE == 0) {
        forward_g2r(a + poly * N, x, buf, tb, c, Tg, rank);
#pragma unroll
        for (int v = 0; v < E / CW; v++)
            *reinterpret_cast<double2 *>(buf + lswz(Tg * E + v * CW)) = make_double2(F::fwd_bits(x[v * CW], c), F::fwd_bits(x[v * CW + 1], c));
        __syncthreads();
        copy_half_s2g(buf, out + poly * N + (size_t)rank * HALF, t);
    } else if (MODE == 1) {
        copy_half_g2s(a + poly * N + (size_t)rank * HALF, buf, t);
        __syncthreads();
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            const double2 w = *reinterpret_cast<const double2 *>(buf + lswz(Tg * E + v * CW));
            x[v * CW] = F::load_bits(w.x, c);
            x[v * CW + 1] = F::load_bits(w.y, c);
        }
        inverse_r2r(x, buf, tb, c, Tg, rank);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(out + poly * N + j * 512 + Tg, F::inv_word(x[j], c));
    } else {
        T *g_c = out + poly * N;
        forward_g2r(a + poly * N, x, buf, tb, c, Tg, rank);
        // fwd(a), canonical, parked in this thread's own 32 contiguous words of the output polynomial (read back by the same thread)
        T *mine_c = g_c + (size_t)Tg * E;
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            uint4 raw;
            const uint64_t w0 = F::fwd_word(x[v * CW], c), w1 = F::fwd_word(x[v * CW + 1], c);
            raw.x = (uint32_t)w0; raw.y = (uint32_t)(w0 >> 32); raw.z = (uint32_t)w1; raw.w = (uint32_t)(w1 >> 32);
            *reinterpret_cast<uint4 *>(mine_c + v * CW) = raw;
        }
        cluster_sync();  // both CTAs are done with the pass-2 reads of their buffers before the next transform's exchange writes into them
        forward_g2r(b + poly * N, x, buf, tb, c, Tg, rank);
#pragma unroll
        for (int v = 0; v < E / CW; v++) {
            uint4 raw;  // written above by this thread (program order); volatile: a few 16-byte loads in flight at a time, the register file is full
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(mine_c + v * CW) : "memory");
            const uint64_t w0 = ((uint64_t)raw.y << 32) | raw.x, w1 = ((uint64_t)raw.w << 32) | raw.z;
            x[v * CW] = F::pointwise(F::load(w0, c), x[v * CW], c);
            x[v * CW + 1] = F::pointwise(F::load(w1, c), x[v * CW + 1], c);
        }
        inverse_r2r(x, buf, tb, c, Tg, rank);
#pragma unroll
        for (int j = 0; j < E; j++) stg_stream(g_c + j * 512 + Tg, F::inv_word(x[j], c));
    }
}

}  // namespace

// cudaErrorNotSupported unless the table is the u64 FP64 lazy-fold N = 16384 layout with 32-word register tiles
cudaError_t launch_ntt_cluster(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, int mode, const uint64_t *a,
                               const uint64_t *b, uint64_t *out, size_t npolys, cudaStream_t s) {
    if (tb0.log_n != LOGN || tb0.loge != LOGE || !tb0.use_f64) return cudaErrorNotSupported;
    if (npolys == 0) return cudaSuccess;
    if (npolys > 0x3fffffffu) return cudaErrorNotSupported;
    constexpr size_t smem = sizeof(double) * HALF;
    const unsigned grid = (unsigned)(2 * npolys);
    cudaError_t e;
    auto go = [&](auto k) -> cudaError_t {
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<grid, TPC, smem, s>>>(tb0, tables, limbs, a, b, out, npolys);
        count_launch();
        return cudaGetLastError();
    };
    if (mode == 0) return go(ntt_cluster_kernel<0>);
    if (mode == 1) return go(ntt_cluster_kernel<1>);
    return go(ntt_cluster_kernel<2>);
}

}  // namespace pfhe
