// internal.hpp -- C++ interface between the C-ABI layer (capi.cu) and the kernel translation units.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "ntt_core.cuh"

namespace pfhe {

extern std::atomic<uint64_t> g_launches;  // every kernel launch issued by the library
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// LOGE of the register-pass kernel instantiated for (word bits, log_n); 0 = generic radix-2 kernel.
int choose_loge(int bits, int log_n);

// Gadget (ApproxSignedBasis) parameters, primus_decompose/src/primitive/basis.rs:47-176
template <typename T> struct GadgetParams {
    T q, basis_m1, q_minus_basis, carry_mask, threshold, add, init_mask;
    uint32_t log_basis, levels, drop_bits;
    uint32_t has_threshold, has_init_mask;
    // carry-free form of the same digits: digit_l = window_l(adjusted + offset) - half (balanced digits are unique), with
    // offset = 2^(drop-1) + sum_l half 2^(drop + l*beta), half = B/2 (0 for the binary basis)
    T offset, half;
};
// returns false when (q, log_basis, levels_in) is rejected by the reference constructor's asserts
template <typename T> bool make_gadget(T q, uint32_t log_basis, uint32_t levels_in, GadgetParams<T> &g);

// ---- NTT family (ntt.cu) ------------------------------------------------------------------
// polys are [batch*limbs][N]; polynomial p uses tables[p % limbs] (tb0 == tables[0] by value).
template <typename T>
cudaError_t launch_ntt(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, const T *src, T *dst, size_t npolys,
                       bool forward, cudaStream_t stream);
template <typename T>
cudaError_t launch_polymul(const DevNtt<T> &tb0, const DevNtt<T> *tables, int limbs, const T *a, const T *b, T *c,
                           size_t npolys, cudaStream_t stream);
template <typename T>
cudaError_t launch_monomial(const DevNtt<T> &tb0, T coeff, const uint32_t *degrees, T *out, size_t batch, cudaStream_t stream);

// ---- pointwise (pointwise.cu) -----------------------------------------------------------------
struct SliceOpArgs {
    int op;
    int limbs;
    size_t rows, n;
};
// per-limb constants passed in a small device-visible struct (<= 16 limbs)
constexpr int kMaxLimbs = 16;
template <typename T> struct LimbConsts {
    Barrett<T> br[kMaxLimbs];
    T scalar[kMaxLimbs], scalar_q[kMaxLimbs];
    double q_f[kMaxLimbs], qinv_f[kMaxLimbs];  // filled by the slice-op launcher when the FP64 product applies (u64 words, q < 2^50)
};
// N = 16384 u64 FP64 transforms on a 2-CTA cluster (ntt_cluster.cu); mode 0 forward, 1 inverse, 2 fused product out = a * b
cudaError_t launch_ntt_cluster(const DevNtt<uint64_t> &tb0, const DevNtt<uint64_t> *tables, int limbs, int mode, const uint64_t *a,
                               const uint64_t *b, uint64_t *out, size_t npolys, cudaStream_t s);
template <typename T>
cudaError_t launch_slice_op(int op, const LimbConsts<T> &lc, int limbs, const T *a, const T *b, const T *c, T *out,
                            size_t rows, size_t n, cudaStream_t stream, size_t b_group = 1);
template <typename T>
cudaError_t launch_butterfly_mul(const LimbConsts<T> &lc, int limbs, T *a, const T *s, const T *w, T *out, size_t rows, size_t n,
                                 cudaStream_t stream);
template <typename T>
cudaError_t launch_inv_slice(const Barrett<T> &br, const T *a, T *out, size_t count, unsigned long long *first_bad, cudaStream_t stream);
template <typename T>
cudaError_t launch_decompose(const GadgetParams<T> &g, const T *values, T *digits, size_t count, cudaStream_t stream);
template <typename T>
cudaError_t launch_rns_lift(const T *moduli_host, int limbs, T small_modulus, const T *small, T *out, size_t count,
                            cudaStream_t stream);
template <typename T>
cudaError_t launch_extract_lwe(T q, const T *rlwe, T *lwe, size_t n, size_t batch, size_t index, size_t count, cudaStream_t stream);

// ---- lattice (lattice.cu, lattice32.cu) -----------------------------------------------------------
// Host copy of the table entries the fast lattice kernels read from the kernel-parameter bank:
// fwd[0..7] (forward stages 0..2) and inv[N-8..N-1] (last three inverse stages; entry 7 = n^-1 * inv_roots[N-1]).
template <typename T> struct LatHead {
    typename Word<T>::Pair fwd_head[8], inv_tail[8];
};
// u32, N = 1024 bootstrapping shape (lattice32.cu); cudaErrorNotSupported when (table, gadget) does not qualify
cudaError_t launch_blind_rotate_fast32(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g,
                                       const uint32_t *bsk, uint32_t n_lwe, const uint32_t *lwe, const uint32_t *test_vector,
                                       uint32_t *acc_out, size_t batch, cudaStream_t stream);
// k = 1 external product, u32, N = 1024 / 2048 (lattice32_ep.cu); cudaErrorNotSupported when (table, gadget) does not qualify
cudaError_t launch_external_product_fast32(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g, uint32_t k,
                                           const uint32_t *key, const uint32_t *in, uint32_t *out, size_t batch, bool to_coeff,
                                           cudaStream_t stream);
// ternary-secret (monomial-combination) blind rotation, same shape restrictions
cudaError_t launch_blind_rotate_ternary32(const DevNtt<uint32_t> &tb, const LatHead<uint32_t> &head, const GadgetParams<uint32_t> &g,
                                          const uint32_t *bsk_plus, const uint32_t *bsk_minus, uint32_t n_lwe, const uint32_t *lwe,
                                          const uint32_t *test_vector, uint32_t *acc_out, size_t batch, cudaStream_t stream);
template <typename T>
cudaError_t launch_external_product(const DevNtt<T> &tb, const GadgetParams<T> &g, uint32_t k, const T *key, const T *in,
                                    T *out, size_t batch, bool to_coeff, cudaStream_t stream);
template <typename T>
cudaError_t launch_blind_rotate(const DevNtt<T> &tb, const GadgetParams<T> &g, const T *bsk, uint32_t n_lwe,
                                const uint32_t *lwe, const T *test_vector, T *acc_out, size_t batch, cudaStream_t stream);

// multi-limb external product from precomputed lifted digits, fused per (ciphertext, limb); `tables` = device array of the
// limbs' lattice-layout tables; policy: 0 integer pipe, 1 lazy FP64, 2 wide u32 forward butterflies
template <typename T>
cudaError_t launch_dcrt_external_product(int policy, const DevNtt<T> *tables, int limbs, uint32_t log_n, uint32_t k, uint32_t levels,
                                         const T *key, const T *digits, T *out, size_t batch, bool to_coeff, cudaStream_t stream);
bool dcrt_wide32_ok(uint64_t q, int logn);

cudaError_t run_modmul_microbench(int kind, uint32_t blocks, uint32_t iters, float *ms);

}  // namespace pfhe
