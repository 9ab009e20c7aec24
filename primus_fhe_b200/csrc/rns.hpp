// rns.hpp -- RNS base + multi-word gadget basis as one kernel-parameter struct, and the launchers of rns.cu.
#pragma once
#include "internal.hpp"

namespace pfhe {

constexpr int kRnsMaxLimbs = 8;  // limbs of an RNS base handled by the compose / gadget kernels
constexpr int kRnsMaxWords = 8;  // words of the composed value (8 x ~50-bit limbs = 400 bits = 7 u64 words)

// RNSBase (primus_rns/src/base.rs:26-122) + BigUintApproxSignedBasis (primus_decompose/src/big_integer/basis.rs:17-211)
template <typename T> struct RnsDev {
    int limbs, value_len;
    T q[kRnsMaxLimbs];
    T product[kRnsMaxWords];                 // Q = prod q_i, little endian
    T punct[kRnsMaxLimbs][kRnsMaxWords];     // Q / q_i
    T inv_punct[kRnsMaxLimbs], inv_punct_q[kRnsMaxLimbs];  // (Q/q_i)^-1 mod q_i and its Shoup quotient
    T pw[kRnsMaxLimbs][kRnsMaxWords], pw_q[kRnsMaxLimbs][kRnsMaxWords];  // 2^(BITS k) mod q_i and its Shoup quotient (decompose)
    // gadget part (log_basis == 0: RNS base only)
    uint32_t log_basis, levels, drop_bits;
    T basis_m1, carry_mask, init_mask;
    int has_threshold, has_init_mask, init_index;
    T threshold[kRnsMaxWords], add[kRnsMaxWords];
};

// BaseConverter (primus_rns/src/converter.rs:21-365): input base + output moduli + (Q/q_i) mod p_k matrix
template <typename T> struct BaseConvDev {
    int n_in, n_out;
    T inv[kRnsMaxLimbs], inv_q[kRnsMaxLimbs];      // (Q/q_i)^-1 mod q_i and its Shoup quotient
    Barrett<T> in_br[kRnsMaxLimbs], out_br[kRnsMaxLimbs];
    T matrix[kRnsMaxLimbs][kRnsMaxLimbs];          // [output k][input i]
    T matrix_q[kRnsMaxLimbs][kRnsMaxLimbs];        // Shoup quotient of matrix[k][i] for the modulus p_k (integer path with few input limbs)
    T q_mod_p[kRnsMaxLimbs];
    double q_f[kRnsMaxLimbs];
    // FP64-pipe formulation (u64 words, every modulus <= 2^50 - 2^10): the same constants as exact doubles
    int f64_ok;
    double in_qinv_f[kRnsMaxLimbs], inv_f[kRnsMaxLimbs], out_p_f[kRnsMaxLimbs], out_pinv_f[kRnsMaxLimbs], q_mod_p_f[kRnsMaxLimbs];
    double matrix_f[kRnsMaxLimbs][kRnsMaxLimbs];
};
template <typename T> int make_baseconv(const T *in_moduli, size_t n_in, const T *out_moduli, size_t n_out, BaseConvDev<T> &c);
template <typename T>
cudaError_t launch_baseconv(const BaseConvDev<T> &c, const T *in, T *out, size_t n, size_t polys, bool exact, cudaStream_t s);

// status codes mirror pfhe_status: 0 ok, 6 EmptyBase, 7 CoPrimeError, 9 invalid / unsupported size
template <typename T> int make_rns(const T *moduli, size_t limbs, uint32_t log_basis, uint32_t levels_in, RnsDev<T> &r);

template <typename T> cudaError_t launch_rns_compose(const RnsDev<T> &r, const T *residues, T *big, size_t count, cudaStream_t s);
template <typename T> cudaError_t launch_rns_decompose(const RnsDev<T> &r, const T *big, T *residues, size_t count, cudaStream_t s);
template <typename T>
cudaError_t launch_rns_gadget(const RnsDev<T> &r, const T *residues, T *digits, size_t count, size_t polys, size_t in_stride, size_t out_stride,
                              cudaStream_t s);
template <typename T>
cudaError_t launch_rns_key_mac(const LimbConsts<T> &lc, int limbs, int comps, uint32_t levels, const T *digits, const T *key, T *out, size_t n,
                               size_t batch, cudaStream_t s);
template <typename T>
cudaError_t launch_rns_lift_scaled_acc(const T *moduli, int limbs, T small_modulus, const T *scalars, const T *small, T *acc, size_t count,
                                       cudaStream_t s);
template <typename T>
cudaError_t launch_mul_monomial(const LimbConsts<T> &lc, int limbs, const uint32_t *degrees, const T *in, T *out, uint32_t log_n, size_t batch,
                                cudaStream_t s);
// single-kernel multi-limb external product (lattice.cu); cudaErrorNotSupported for composed values longer than two words
template <typename T>
cudaError_t launch_dcrt_external_product_fused(int policy, const DevNtt<T> *tables, const RnsDev<T> &r, uint32_t log_n, uint32_t k, const T *key,
                                               const T *in, T *out, size_t batch, bool to_coeff, cudaStream_t s);
template <typename T> cudaError_t launch_dot_product(const Barrett<T> &br, const T *a, const T *b, T *out, size_t rows, size_t n, cudaStream_t s);

}  // namespace pfhe
