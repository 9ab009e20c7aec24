// tma.cuh -- TMA (cp.async.bulk.tensor) plumbing shared by the NTT and lattice kernels.
//
// The XOR swizzle of the exchange buffer (16-byte chunk index ^ bits 7..9 of the byte address) is exactly the
// SWIZZLE_128B pattern of a tensor map with 128-byte rows, so one elected thread moves a whole polynomial between the
// swizzled buffer and its linear image in HBM with a single tensor copy (up to 256 rows per box).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace pfhe {

// host: tensor map over a batch of polynomials viewed as 128-byte rows (false: TMA unavailable / misaligned -> LSU path)
template <typename T> bool make_poly_map(CUtensorMap *map, const T *base, size_t npolys, int log_n);

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// elected thread: shared (swizzled) -> global tensor rows [row, row + boxes*box_rows); returns once shared memory has been read
__device__ __forceinline__ void tma_store_poly(const CUtensorMap *map, const void *sm, uint32_t row, int boxes, int box_rows) {
    for (int bx = 0; bx < boxes; bx++)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(0u), "r"(row + bx * box_rows),
                     "r"(smem_addr(sm) + (uint32_t)bx * box_rows * 128u)
                     : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// elected thread: global tensor rows -> shared (swizzled), completion on the mbarrier
__device__ __forceinline__ void tma_load_poly(const CUtensorMap *map, void *sm, uint32_t row, uint64_t *bar, int boxes, int box_rows) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((uint32_t)(boxes * box_rows) * 128u) : "memory");
    for (int bx = 0; bx < boxes; bx++)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_addr(sm) + (uint32_t)bx * box_rows * 128u),
                     "l"(map), "r"(0u), "r"(row + bx * box_rows), "r"(smem_addr(bar))
                     : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

}  // namespace pfhe
