/*
 * pfhe.h -- C-ABI of the B200-native polynomial-ring hot path (libpfhe_cuda.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every entry point names the reference interface it replaces (paths relative
 * to /root/reference/crates/).  All compute runs in hand-written sm_100a CUDA
 * kernels; there is no CPU fallback -- when no CUDA device is usable every
 * call returns PFHE_ERR_CUDA.
 *
 * Conventions
 *  - "host-slice shims" (`*_slice`, `*_slices`) take HOST memory exactly like
 *    the reference traits do (`&mut [T]`, primus_data/src/traits.rs:20) and do
 *    H2D -> kernel -> D2H internally, so a Rust `impl NttTable for CudaTable`
 *    is a 1:1 forwarder.
 *  - "device batch" entry points (`*_batch`, `*_dev`) take DEVICE pointers and a
 *    `cudaStream_t` passed as `void*` (NULL = default stream); they are
 *    stream-ordered, never synchronise and never allocate.
 *  - Tables/handles are immutable after `create` and safe to use from many
 *    host threads (NttTable: Send + Sync, primus_ntt/src/ntt/mod.rs:16).
 *  - Hot-path calls are infallible in the reference (length mismatches are
 *    debug_assert!, prime64/table.rs:543); here they return PFHE_ERR_INVALID_ARG
 *    for NULL/zero-size misuse and PFHE_ERR_CUDA for launch failures.
 *  - Layouts are the reference's flat element layouts: a batch of polynomials
 *    is `[batch][N]`; DCRT/CRT polynomials are limb-major `[L][N]`
 *    (primus_poly/src/dcrt/mod.rs:29,81-86); GGSW keys are
 *    `[row k+1][level l][component k+1][limb L][N]` (primus_lattice/src/ggsw/dcrt.rs:14-31).
 */
#ifndef PFHE_H
#define PFHE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  1..5 mirror NttError (primus_ntt/src/error.rs:7-49), 6..7 mirror
 * RNSError (primus_rns/src/error.rs:7-20). */
typedef enum {
    PFHE_OK = 0,
    PFHE_ERR_NO_PRIMITIVE_ROOT = 1, /* NttError::NoPrimitiveRoot   */
    PFHE_ERR_DEGREE_CONVERSION = 2, /* NttError::DegreeConversionErr */
    PFHE_ERR_DEGREE_TOO_LARGE = 3,  /* NttError::DegreeTooLarge    */
    PFHE_ERR_NTT_TABLE = 4,         /* NttError::NttTableErr       */
    PFHE_ERR_MODULUS_TOO_LARGE = 5, /* NttError::ModulusTooLarge   */
    PFHE_ERR_RNS_EMPTY = 6,         /* RNSError::EmptyBase         */
    PFHE_ERR_RNS_NOT_COPRIME = 7,   /* RNSError::CoPrimeError      */
    PFHE_ERR_CUDA = 8,              /* CUDA runtime / no device    */
    PFHE_ERR_INVALID_ARG = 9,
    PFHE_ERR_UNSUPPORTED = 10
} pfhe_status;

const char *pfhe_status_string(pfhe_status s);
/* Last CUDA error text seen by this thread (empty string when none). */
const char *pfhe_last_cuda_error(void);
/* Library/ABI version and the SM architecture the kernels were compiled for ("sm_100a"). */
const char *pfhe_version(void);
const char *pfhe_compiled_arch(void);
/* Number of kernel launches issued by this library since load (all threads). */
uint64_t pfhe_launch_count(void);

/* Opaque handles */
typedef struct pfhe_ntt32 pfhe_ntt32;   /* U32NttTable  primus_ntt/src/ntt/prime32/table.rs:37  */
typedef struct pfhe_ntt64 pfhe_ntt64;   /* U64NttTable  primus_ntt/src/ntt/prime64/table.rs:41  */
typedef struct pfhe_dcrt32 pfhe_dcrt32; /* U32DcrtTable primus_ntt/src/dcrt/prime32.rs          */
typedef struct pfhe_dcrt64 pfhe_dcrt64; /* U64DcrtTable primus_ntt/src/dcrt/prime64.rs:11-128   */

/* ===================================================================================== */
/* NttTable (primus_ntt/src/ntt/mod.rs:16-113)                                            */
/* ===================================================================================== */

/* NttTable::new(log_n, modulus) -- prime64/table.rs:308-516, prime32/table.rs:185-343.
 * Fails with NO_PRIMITIVE_ROOT when 2N does not divide q-1 (root.rs:72-81) and with
 * MODULUS_TOO_LARGE when q >= 2^62 (u64, table.rs:318) / q >= 2^30 (u32, table.rs:195).
 * Supported degrees: 1 <= log_n <= 14 (u64) / 15 (u32) (one polynomial per CTA in shared memory). */
pfhe_status pfhe_ntt64_create(int device, uint32_t log_n, uint64_t q, pfhe_ntt64 **out);
pfhe_status pfhe_ntt32_create(int device, uint32_t log_n, uint32_t q, pfhe_ntt32 **out);
void pfhe_ntt64_destroy(pfhe_ntt64 *t);
void pfhe_ntt32_destroy(pfhe_ntt32 *t);

/* NttTable::poly_length and table constants (prime64/table.rs:41-113) */
size_t pfhe_ntt64_poly_length(const pfhe_ntt64 *t);
size_t pfhe_ntt32_poly_length(const pfhe_ntt32 *t);
uint64_t pfhe_ntt64_modulus(const pfhe_ntt64 *t);
uint32_t pfhe_ntt32_modulus(const pfhe_ntt32 *t);
uint64_t pfhe_ntt64_root(const pfhe_ntt64 *t);      /* minimal primitive 2N-th root, root.rs:103-125 */
uint32_t pfhe_ntt32_root(const pfhe_ntt32 *t);
uint64_t pfhe_ntt64_inv_root(const pfhe_ntt64 *t);
uint32_t pfhe_ntt32_inv_root(const pfhe_ntt32 *t);
uint64_t pfhe_ntt64_inv_n(const pfhe_ntt64 *t);
uint32_t pfhe_ntt32_inv_n(const pfhe_ntt32 *t);
int pfhe_ntt64_device(const pfhe_ntt64 *t);
int pfhe_ntt32_device(const pfhe_ntt32 *t);

/* Host-slice shims: NttTable::{transform_slice, lazy_transform_slice, inverse_transform_slice,
 * lazy_inverse_transform_slice} (ntt/mod.rs:53-91; prime64/table.rs:542-563).
 * `poly` is HOST memory of N words, transformed in place (normal -> bit-reversed order forward,
 * bit-reversed -> normal inverse).  lazy != 0 selects the lazy contract: forward accepts inputs
 * in [0,4q), inverse in [0,2q); outputs are congruent mod q and inside the reference's lazy range
 * (this implementation always returns canonical values, which satisfy both contracts). */
pfhe_status pfhe_ntt64_transform_slice(const pfhe_ntt64 *t, uint64_t *poly, int lazy);
pfhe_status pfhe_ntt32_transform_slice(const pfhe_ntt32 *t, uint32_t *poly, int lazy);
pfhe_status pfhe_ntt64_inverse_transform_slice(const pfhe_ntt64 *t, uint64_t *values, int lazy);
pfhe_status pfhe_ntt32_inverse_transform_slice(const pfhe_ntt32 *t, uint32_t *values, int lazy);
/* Many polynomials in one call ([batch][N] HOST memory, in place); what the whole-ciphertext
 * macros loop over (primus_lattice/src/macros/mod.rs:537-674). */
pfhe_status pfhe_ntt64_transform_slices(const pfhe_ntt64 *t, uint64_t *polys, size_t batch, int lazy);
pfhe_status pfhe_ntt32_transform_slices(const pfhe_ntt32 *t, uint32_t *polys, size_t batch, int lazy);
pfhe_status pfhe_ntt64_inverse_transform_slices(const pfhe_ntt64 *t, uint64_t *polys, size_t batch, int lazy);
pfhe_status pfhe_ntt32_inverse_transform_slices(const pfhe_ntt32 *t, uint32_t *polys, size_t batch, int lazy);

/* NttTable::{transform_monomial, transform_coeff_one_monomial, transform_coeff_minus_one_monomial}
 * (ntt/mod.rs:93-112; prime64/table.rs:565-651): NTT(coeff * X^degree) into `values` (HOST, N words). */
pfhe_status pfhe_ntt64_transform_monomial(const pfhe_ntt64 *t, uint64_t coeff, size_t degree, uint64_t *values);
pfhe_status pfhe_ntt32_transform_monomial(const pfhe_ntt32 *t, uint32_t coeff, size_t degree, uint32_t *values);
pfhe_status pfhe_ntt64_transform_coeff_one_monomial(const pfhe_ntt64 *t, size_t degree, uint64_t *values);
pfhe_status pfhe_ntt32_transform_coeff_one_monomial(const pfhe_ntt32 *t, size_t degree, uint32_t *values);
pfhe_status pfhe_ntt64_transform_coeff_minus_one_monomial(const pfhe_ntt64 *t, size_t degree, uint64_t *values);
pfhe_status pfhe_ntt32_transform_coeff_minus_one_monomial(const pfhe_ntt32 *t, size_t degree, uint32_t *values);

/* Device batch API (stream ordered, in place, `dev` = [batch][N] DEVICE words, canonical values in [0,q);
 * lazy-range data must first go through PFHE_OP_REDUCE_LAZY -- the host-slice shims do that themselves). */
pfhe_status pfhe_ntt64_forward_batch(const pfhe_ntt64 *t, uint64_t *dev, size_t batch, void *stream);
pfhe_status pfhe_ntt32_forward_batch(const pfhe_ntt32 *t, uint32_t *dev, size_t batch, void *stream);
pfhe_status pfhe_ntt64_inverse_batch(const pfhe_ntt64 *t, uint64_t *dev, size_t batch, void *stream);
pfhe_status pfhe_ntt32_inverse_batch(const pfhe_ntt32 *t, uint32_t *dev, size_t batch, void *stream);
/* Out-of-place variants (src -> dst, src untouched). */
pfhe_status pfhe_ntt64_forward_batch_to(const pfhe_ntt64 *t, const uint64_t *src, uint64_t *dst, size_t batch, void *stream);
pfhe_status pfhe_ntt32_forward_batch_to(const pfhe_ntt32 *t, const uint32_t *src, uint32_t *dst, size_t batch, void *stream);
pfhe_status pfhe_ntt64_inverse_batch_to(const pfhe_ntt64 *t, const uint64_t *src, uint64_t *dst, size_t batch, void *stream);
pfhe_status pfhe_ntt32_inverse_batch_to(const pfhe_ntt32 *t, const uint32_t *src, uint32_t *dst, size_t batch, void *stream);
/* Monomial transforms on the device: degrees[batch] (device, each < 2N) -> out [batch][N];
 * coeff as in transform_monomial. */
pfhe_status pfhe_ntt64_monomial_batch(const pfhe_ntt64 *t, uint64_t coeff, const uint32_t *degrees, uint64_t *out, size_t batch, void *stream);
pfhe_status pfhe_ntt32_monomial_batch(const pfhe_ntt32 *t, uint32_t coeff, const uint32_t *degrees, uint32_t *out, size_t batch, void *stream);

/* Fused negacyclic product  c = inv(fwd(a) .* fwd(b))  in Z_q[X]/(X^N+1): the composition callers
 * make from transform_slice + NttPolynomial::mul_assign + inverse_transform_slice
 * (primus_lattice/src/rlwe/coeff.rs:92-122; primus_poly/src/ntt/mul.rs:84-90).  a,b,c = [batch][N]
 * device; c may alias a or b.  Host variant takes HOST pointers. */
pfhe_status pfhe_ntt64_polymul_batch(const pfhe_ntt64 *t, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, void *stream);
pfhe_status pfhe_ntt32_polymul_batch(const pfhe_ntt32 *t, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch, void *stream);
pfhe_status pfhe_ntt64_polymul_slices(const pfhe_ntt64 *t, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch);
pfhe_status pfhe_ntt32_polymul_slices(const pfhe_ntt32 *t, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch);

/* ===================================================================================== */
/* DcrtTable (primus_ntt/src/dcrt/mod.rs:19-135): L independent limb tables, layout [L][N] */
/* ===================================================================================== */
pfhe_status pfhe_dcrt64_create(int device, uint32_t log_n, const uint64_t *moduli, size_t count, pfhe_dcrt64 **out);
pfhe_status pfhe_dcrt32_create(int device, uint32_t log_n, const uint32_t *moduli, size_t count, pfhe_dcrt32 **out);
void pfhe_dcrt64_destroy(pfhe_dcrt64 *t);
void pfhe_dcrt32_destroy(pfhe_dcrt32 *t);
size_t pfhe_dcrt64_poly_length(const pfhe_dcrt64 *t);
size_t pfhe_dcrt32_poly_length(const pfhe_dcrt32 *t);
size_t pfhe_dcrt64_moduli_count(const pfhe_dcrt64 *t);
size_t pfhe_dcrt32_moduli_count(const pfhe_dcrt32 *t);
size_t pfhe_dcrt64_crt_poly_length(const pfhe_dcrt64 *t);
size_t pfhe_dcrt32_crt_poly_length(const pfhe_dcrt32 *t);
/* ntt_tables()/iter(): borrow the i-th limb table (owned by the DCRT handle). */
const pfhe_ntt64 *pfhe_dcrt64_ntt_table(const pfhe_dcrt64 *t, size_t limb);
const pfhe_ntt32 *pfhe_dcrt32_ntt_table(const pfhe_dcrt32 *t, size_t limb);
/* host-slice shims over [batch][L][N] HOST words (batch = 1 is the trait method) */
pfhe_status pfhe_dcrt64_transform_slices(const pfhe_dcrt64 *t, uint64_t *polys, size_t batch, int lazy);
pfhe_status pfhe_dcrt32_transform_slices(const pfhe_dcrt32 *t, uint32_t *polys, size_t batch, int lazy);
pfhe_status pfhe_dcrt64_inverse_transform_slices(const pfhe_dcrt64 *t, uint64_t *polys, size_t batch, int lazy);
pfhe_status pfhe_dcrt32_inverse_transform_slices(const pfhe_dcrt32 *t, uint32_t *polys, size_t batch, int lazy);
/* device batch: dev = [batch][L][N] */
pfhe_status pfhe_dcrt64_forward_batch(const pfhe_dcrt64 *t, uint64_t *dev, size_t batch, void *stream);
pfhe_status pfhe_dcrt32_forward_batch(const pfhe_dcrt32 *t, uint32_t *dev, size_t batch, void *stream);
pfhe_status pfhe_dcrt64_inverse_batch(const pfhe_dcrt64 *t, uint64_t *dev, size_t batch, void *stream);
pfhe_status pfhe_dcrt32_inverse_batch(const pfhe_dcrt32 *t, uint32_t *dev, size_t batch, void *stream);
/* RNS polynomial product per limb: DcrtTable::transform + DcrtPolynomial::mul_assign
 * (primus_poly/src/dcrt/mul.rs:176-187) + inverse, fused. */
pfhe_status pfhe_dcrt64_polymul_batch(const pfhe_dcrt64 *t, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, void *stream);
pfhe_status pfhe_dcrt32_polymul_batch(const pfhe_dcrt32 *t, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch, void *stream);

/* ===================================================================================== */
/* Pointwise modular slice operators on DEVICE memory.                                    */
/* Traits: primus_reduce/src/slice_ops.rs:63-230; bodies primus_modulus/src/barrett/slice.rs:185-295 */
/* -> common/compact/slice.rs:106-365.  `q` must satisfy 1 < q < 2^(BITS-2)                */
/* (BarrettModulus::new, primus_modulus/src/barrett/mod.rs:39-43).  Inputs in [0,q).       */
/* `limbs`/`n`: the slices are [rows][limbs][n] with modulus moduli[limb] per limb         */
/* (DcrtPolynomial ops, primus_poly/src/dcrt/mul.rs:176-187, dcrt/mod.rs:105-123);         */
/* single-modulus slices use limbs = 1, rows*n = length.                                   */
/* ===================================================================================== */
typedef enum {
    PFHE_OP_MUL = 0,         /* out = a*b          reduce_mul_slice_to / _assign (out==a)        */
    PFHE_OP_ADD_MUL = 1,     /* out = out + a*b    reduce_add_mul_slice_assign                   */
    PFHE_OP_SUB_MUL = 2,     /* out = out - a*b    reduce_sub_mul_slice_assign                   */
    PFHE_OP_MUL_ADD = 3,     /* out = a*b + c      reduce_mul_add_slice_to                       */
    PFHE_OP_ADD = 4,         /* out = a + b        reduce_add_slice_to                           */
    PFHE_OP_SUB = 5,         /* out = a - b        reduce_sub_slice_to                           */
    PFHE_OP_NEG = 6,         /* out = -a           reduce_neg_slice_to                           */
    PFHE_OP_MUL_SCALAR = 7,  /* out = a*s          reduce_mul_scalar_slice_to                    */
    PFHE_OP_ADD_MUL_SCALAR = 8, /* out = out + a*s reduce_add_mul_scalar_slice_assign            */
    PFHE_OP_FACTOR_MUL = 9,  /* out = f*a  (Shoup) FactorSliceOps::factor_mul_slice_to (primus_factor/src/ops.rs:58-118) */
    PFHE_OP_ADD_FACTOR_MUL = 10, /* out += f*a     add_factor_mul_slice_assign (common/slice.rs:61-70)   */
    PFHE_OP_SUB_FACTOR_MUL = 11, /* out -= f*a     sub_factor_mul_slice_assign                   */
    PFHE_OP_REDUCE_LAZY = 12,    /* out = a mod q for a in [0,4q): canonicalises lazy-range inputs (reduce_once twice) */
    PFHE_OP_DOUBLE = 13,         /* out = 2a           reduce_double_slice_to (slice_ops.rs:91-100)                */
    PFHE_OP_MUL_SCALAR_ADD = 14, /* out = a*s + c      reduce_mul_scalar_add_slice_to (slice_ops.rs:229)           */
    PFHE_OP_FACTOR_MUL_ADD = 15  /* out = f*a + c (Shoup) FactorSliceOps::factor_mul_add_slice_to (ops.rs:117)     */
} pfhe_slice_op;

/* One entry point per word size; `scalars` (HOST, `limbs` words) holds s / f per limb for the
 * scalar and factor ops (ignored otherwise); `b`, `c` may be NULL when the op does not read them. */
pfhe_status pfhe_mod64_slice_op(pfhe_slice_op op, const uint64_t *moduli, size_t limbs, const uint64_t *scalars,
                                const uint64_t *a, const uint64_t *b, const uint64_t *c, uint64_t *out,
                                size_t rows, size_t n, void *stream);
pfhe_status pfhe_mod32_slice_op(pfhe_slice_op op, const uint32_t *moduli, size_t limbs, const uint32_t *scalars,
                                const uint32_t *a, const uint32_t *b, const uint32_t *c, uint32_t *out,
                                size_t rows, size_t n, void *stream);
/* NTT-domain ciphertext x polynomial (primus_lattice/src/rlwe/ntt.rs:78-152: mul_ntt_polynomial_assign / _to,
 * add_ntt_rlwe_mul_ntt_polynomial_assign; the GLWE forms loop the same way): op in {PFHE_OP_MUL, PFHE_OP_ADD_MUL, PFHE_OP_SUB_MUL}
 * with `b` broadcast -- row r of `a`/`out` ([rows][limbs][n]) pairs with row r / group of `b` ([rows/group][limbs][n]), group =
 * components per ciphertext (2 for RLWE). */
pfhe_status pfhe_mod64_slice_op_bcast(pfhe_slice_op op, const uint64_t *moduli, size_t limbs, const uint64_t *a, const uint64_t *b,
                                      uint64_t *out, size_t rows, size_t n, size_t group, void *stream);
pfhe_status pfhe_mod32_slice_op_bcast(pfhe_slice_op op, const uint32_t *moduli, size_t limbs, const uint32_t *a, const uint32_t *b,
                                      uint32_t *out, size_t rows, size_t n, size_t group, void *stream);
/* DcrtPolynomial::butterfly_mul_factor_to / DcrtGlwe::butterfly_mul_factor_to (primus_poly/src/dcrt/mul.rs:189-222,
 * primus_lattice/src/glwe/dcrt.rs:150-175): (a, out) = (a + s, (a_orig - s) * w) per limb; a, s, out: device [rows][limbs][n]
 * in [0,q), a updated in place; w: device [limbs][n] factor polynomial (values; the Shoup quotients of the reference's
 * ShoupFactor array are not needed, the product is the same residue). */
pfhe_status pfhe_mod64_butterfly_mul_factor(const uint64_t *moduli, size_t limbs, uint64_t *a, const uint64_t *s, const uint64_t *w,
                                            uint64_t *out, size_t rows, size_t n, void *stream);
pfhe_status pfhe_mod32_butterfly_mul_factor(const uint32_t *moduli, size_t limbs, uint32_t *a, const uint32_t *s, const uint32_t *w,
                                            uint32_t *out, size_t rows, size_t n, void *stream);
/* ReduceInvSlice::reduce_inv_slice_to / NttPolynomial::inv_to (primus_poly/src/ntt/inv.rs:1-58, primus_reduce/src/slice_ops.rs:255-300):
 * out[i] = a[i]^-1 mod q for a PRIME q (every NTT modulus is).  Zero has no inverse: out[i] = 0 and, when `first_bad` (device
 * uint64, initialised by the caller to ~0) is not NULL, the smallest such index is recorded -- try_reduce_inv_slice_to's
 * ReduceError::NoInverseAtIndex (primus_reduce/src/error.rs:23). */
pfhe_status pfhe_mod64_inv_slice(uint64_t q, const uint64_t *a, uint64_t *out, size_t count, uint64_t *first_bad, void *stream);
pfhe_status pfhe_mod32_inv_slice(uint32_t q, const uint32_t *a, uint32_t *out, size_t count, uint64_t *first_bad, void *stream);
/* Host-slice shims of the same operators (HOST pointers; H2D -> kernel -> D2H). */
pfhe_status pfhe_mod64_slice_op_host(pfhe_slice_op op, const uint64_t *moduli, size_t limbs, const uint64_t *scalars,
                                     const uint64_t *a, const uint64_t *b, const uint64_t *c, uint64_t *out,
                                     size_t rows, size_t n);
pfhe_status pfhe_mod32_slice_op_host(pfhe_slice_op op, const uint32_t *moduli, size_t limbs, const uint32_t *scalars,
                                     const uint32_t *a, const uint32_t *b, const uint32_t *c, uint32_t *out,
                                     size_t rows, size_t n);

/* MultiplyFactor::new(operand, bit_shift, modulus) (primus_factor/src/mul_factor/mod.rs:4-43): the HEXL-style precomputed
 * quotient floor(operand * 2^bit_shift / modulus) (low 64 bits), bit_shift in {32, 52, 64}; setup-only host arithmetic (the
 * device tables use the same quotient with bit_shift = word size).  INVALID_ARG mirrors the constructor's asserts. */
pfhe_status pfhe_multiply_factor64(uint64_t operand, uint32_t bit_shift, uint64_t modulus, uint64_t *quotient);
/* MultiplyFactor::mul_modulo::<BIT_SHIFT> (mul_factor/mod.rs:45-70) for one value, host side (table-construction helper). */
pfhe_status pfhe_multiply_factor64_mul(uint64_t operand, uint64_t quotient, uint32_t bit_shift, uint64_t b, uint64_t modulus, uint64_t *out);

/* ===================================================================================== */
/* Gadget decomposition and RNS limb handling                                             */
/* ===================================================================================== */

/* ApproxSignedBasis::new(Some(q), log_basis, reverse_length) geometry
 * (primus_decompose/src/primitive/basis.rs:47-176).  levels_in = 0 means the full length
 * floor(bitlen(q)/log_basis).  Outputs the decompose length and drop bits. */
pfhe_status pfhe_basis64_geometry(uint64_t q, uint32_t log_basis, uint32_t levels_in, uint32_t *levels, uint32_t *drop_bits);
pfhe_status pfhe_basis32_geometry(uint32_t q, uint32_t log_basis, uint32_t levels_in, uint32_t *levels, uint32_t *drop_bits);

/* init_value_carry_slice + OnceSignedDecomposer::decompose_slice_to for every level, fused
 * (basis.rs:254-406; primitive/common.rs:219-273).  values: [count] device words in [0,q);
 * digits: [levels][count] device words, LSB level first, each digit canonical in [0,q)
 * (negative digits are q - |d|). */
pfhe_status pfhe_decompose64_batch(uint64_t q, uint32_t log_basis, uint32_t levels_in, const uint64_t *values,
                                   uint64_t *digits, size_t count, void *stream);
pfhe_status pfhe_decompose32_batch(uint32_t q, uint32_t log_basis, uint32_t levels_in, const uint32_t *values,
                                   uint32_t *digits, size_t count, void *stream);

/* RNSBase::wrapping_decompose_small_values_to (primus_rns/src/base.rs:279-315): centred lift of
 * small values in [0, small_modulus) to every limb; out is modulus-major [limbs][count]. */
pfhe_status pfhe_rns64_lift_small_batch(const uint64_t *moduli, size_t limbs, uint64_t small_modulus,
                                        const uint64_t *small, uint64_t *out, size_t count, void *stream);
pfhe_status pfhe_rns32_lift_small_batch(const uint32_t *moduli, size_t limbs, uint32_t small_modulus,
                                        const uint32_t *small, uint32_t *out, size_t count, void *stream);

/* ===================================================================================== */
/* External product and blind rotation                                                    */
/* ===================================================================================== */

/* Single-modulus (L = 1) GGSW external product, fused per ciphertext:
 *   out_c = [inv]( sum_{r<=k} sum_{l<levels} fwd(digit_l(in_r)) .* key[r][l][c] )
 * = CrtGlwe::mul_dcrt_ggsw_to (primus_lattice/src/glwe/crt.rs:200-227) -> gadget product
 * (glwe/dcrt.rs:178-255) -> MAC (glwe/dcrt.rs:108-126) [+ into_coeff_form, macros/mod.rs:892-937
 * when to_coeff != 0].  key: device [k+1][levels][k+1][N] in NTT domain, shared by the batch;
 * in/out: device [batch][k+1][N]; in is coefficient domain, canonical. */
pfhe_status pfhe_ggsw64_external_product_batch(const pfhe_ntt64 *t, uint32_t k, uint32_t log_basis, uint32_t levels_in,
                                               const uint64_t *key, const uint64_t *in, uint64_t *out,
                                               size_t batch, int to_coeff, void *stream);
pfhe_status pfhe_ggsw32_external_product_batch(const pfhe_ntt32 *t, uint32_t k, uint32_t log_basis, uint32_t levels_in,
                                               const uint32_t *key, const uint32_t *in, uint32_t *out,
                                               size_t batch, int to_coeff, void *stream);

/* Host-slice shim of the same product (HOST pointers; key uploaded once, ciphertexts pipelined H2D -> kernel -> D2H):
 * the 1:1 forwarder for a Rust `mul_dcrt_ggsw_to` over host-resident ciphertexts. */
pfhe_status pfhe_ggsw64_external_product_slices(const pfhe_ntt64 *t, uint32_t k, uint32_t log_basis, uint32_t levels_in,
                                                const uint64_t *key, const uint64_t *in, uint64_t *out, size_t batch, int to_coeff);
pfhe_status pfhe_ggsw32_external_product_slices(const pfhe_ntt32 *t, uint32_t k, uint32_t log_basis, uint32_t levels_in,
                                                const uint32_t *key, const uint32_t *in, uint32_t *out, size_t batch, int to_coeff);

/* Blind rotation composed from the reference's primitives (the reference has
 * no blind rotation: mul_monomial_assign primus_poly/src/poly/mul.rs:74-99, external product as
 * above, RLWE add).  bsk: device [n_lwe][2][levels][2][N] NTT domain; lwe: device
 * [batch][n_lwe+1] uint32 (a_0..a_{n-1}, b) already in Z_{2N}; test_vector: device [N];
 * acc_out: device [batch][2][N] (a then b polynomial), canonical coefficients. */
pfhe_status pfhe_blind_rotate64_batch(const pfhe_ntt64 *t, uint32_t log_basis, uint32_t levels_in, const uint64_t *bsk,
                                      uint32_t n_lwe, const uint32_t *lwe, const uint64_t *test_vector,
                                      uint64_t *acc_out, size_t batch, void *stream);
pfhe_status pfhe_blind_rotate32_batch(const pfhe_ntt32 *t, uint32_t log_basis, uint32_t levels_in, const uint32_t *bsk,
                                      uint32_t n_lwe, const uint32_t *lwe, const uint32_t *test_vector,
                                      uint32_t *acc_out, size_t batch, void *stream);
/* Rlwe::extract_lwe (primus_lattice/src/rlwe/coeff.rs:264-288): rlwe [batch][2][N] -> lwe [batch][N+1] */
pfhe_status pfhe_extract_lwe64_batch(uint64_t q, const uint64_t *rlwe, uint64_t *lwe, size_t n, size_t batch, void *stream);
pfhe_status pfhe_extract_lwe32_batch(uint32_t q, const uint32_t *rlwe, uint32_t *lwe, size_t n, size_t batch, void *stream);
/* Rlwe::extract_lwe_with_index (coeff.rs:194-226; count = 1) and Rlwe::extract_first_few_lwe (coeff.rs:229-261; index = 0,
 * MultiMsgLwe with `count` bodies): rlwe [batch][2][N] -> lwe [batch][N + count]; index + count <= N. */
pfhe_status pfhe_extract_lwe64_ex_batch(uint64_t q, const uint64_t *rlwe, uint64_t *lwe, size_t n, size_t batch, size_t index, size_t count, void *stream);
pfhe_status pfhe_extract_lwe32_ex_batch(uint32_t q, const uint32_t *rlwe, uint32_t *lwe, size_t n, size_t batch, size_t index, size_t count, void *stream);


/* ===================================================================================== */
/* RNS base, multi-word gadget basis and the multi-limb (L > 1) external product           */
/* ===================================================================================== */
typedef struct pfhe_rns32 pfhe_rns32; /* RNSBase<u32> primus_rns/src/base.rs:26-122 */
typedef struct pfhe_rns64 pfhe_rns64; /* RNSBase<u64> */

/* RNSBase::new(moduli) (base.rs:79-122): pairwise-coprime check (RNS_NOT_COPRIME), punctured products Q/q_i and
 * their inverses.  Up to 8 limbs / 8 words of composed value.  Host-only handle (constants travel as kernel
 * parameters), usable from any thread and on any device. */
pfhe_status pfhe_rns64_create(const uint64_t *moduli, size_t count, pfhe_rns64 **out);
pfhe_status pfhe_rns32_create(const uint32_t *moduli, size_t count, pfhe_rns32 **out);
void pfhe_rns64_destroy(pfhe_rns64 *r);
void pfhe_rns32_destroy(pfhe_rns32 *r);
size_t pfhe_rns64_moduli_count(const pfhe_rns64 *r);
size_t pfhe_rns32_moduli_count(const pfhe_rns32 *r);
size_t pfhe_rns64_big_uint_value_len(const pfhe_rns64 *r); /* RNSBase::big_uint_value_len */
size_t pfhe_rns32_big_uint_value_len(const pfhe_rns32 *r);
/* moduli_product(): Q as value_len little-endian words into `out` (HOST) */
pfhe_status pfhe_rns64_moduli_product(const pfhe_rns64 *r, uint64_t *out);
pfhe_status pfhe_rns32_moduli_product(const pfhe_rns32 *r, uint32_t *out);

/* RNSBase::compose_multiple_values_to (base.rs:638-673): residues [limbs][count] (modulus-major, device) ->
 * big values [count][value_len] little-endian words (device). */
pfhe_status pfhe_rns64_compose_batch(const pfhe_rns64 *r, const uint64_t *residues, uint64_t *big, size_t count, void *stream);
pfhe_status pfhe_rns32_compose_batch(const pfhe_rns32 *r, const uint32_t *residues, uint32_t *big, size_t count, void *stream);
/* RNSBase::decompose_big_uint_values_to (base.rs:457-481): the inverse map (values need not be < Q). */
pfhe_status pfhe_rns64_decompose_batch(const pfhe_rns64 *r, const uint64_t *big, uint64_t *residues, size_t count, void *stream);
pfhe_status pfhe_rns32_decompose_batch(const pfhe_rns32 *r, const uint32_t *big, uint32_t *residues, size_t count, void *stream);
/* RNSBase::wrapping_decompose_small_values_scaled_add_to (base.rs:326-386, :739-756):
 * acc[l][i] += scalars[l] * centred_lift(small[i]) mod q_l; acc is [limbs][count] device, scalars HOST. */
pfhe_status pfhe_rns64_lift_small_scaled_add_batch(const uint64_t *moduli, size_t limbs, uint64_t small_modulus, const uint64_t *scalars,
                                                   const uint64_t *small, uint64_t *acc, size_t count, void *stream);
pfhe_status pfhe_rns32_lift_small_scaled_add_batch(const uint32_t *moduli, size_t limbs, uint32_t small_modulus, const uint32_t *scalars,
                                                   const uint32_t *small, uint32_t *acc, size_t count, void *stream);

/* BigUintApproxSignedBasis::new(Q, log_basis, reverse_length) geometry (primus_decompose/src/big_integer/basis.rs:40-211). */
pfhe_status pfhe_bigbasis64_geometry(const pfhe_rns64 *r, uint32_t log_basis, uint32_t levels_in, uint32_t *levels, uint32_t *drop_bits);
pfhe_status pfhe_bigbasis32_geometry(const pfhe_rns32 *r, uint32_t log_basis, uint32_t levels_in, uint32_t *levels, uint32_t *drop_bits);
/* Gadget decomposition of CRT polynomials, fused: compose (base.rs:609-636) -> init_value_carry_slice_inplace
 * (big_integer/basis.rs:326-367) -> unsigned_decompose_slice_to for every level (big_integer/common.rs:275-325) ->
 * centred lift to every limb (base.rs:279-315), i.e. the digit pipeline of add_dcrt_glev_mul_crt_poly_assign
 * (primus_lattice/src/glwe/dcrt.rs:219-236).  residues: [polys][limbs][n] device; digits: [polys][levels][limbs][n]. */
pfhe_status pfhe_rns64_gadget_decompose_batch(const pfhe_rns64 *r, uint32_t log_basis, uint32_t levels_in, const uint64_t *residues,
                                              uint64_t *digits, size_t n, size_t polys, void *stream);
pfhe_status pfhe_rns32_gadget_decompose_batch(const pfhe_rns32 *r, uint32_t log_basis, uint32_t levels_in, const uint32_t *residues,
                                              uint32_t *digits, size_t n, size_t polys, void *stream);

/* Multi-limb GGSW external product = CrtGlwe::mul_dcrt_ggsw_to (primus_lattice/src/glwe/crt.rs:200-227) [+ into_coeff_form]:
 * digits (above) -> DCRT forward NTT -> key multiply-accumulate with lazy double-word sums
 * (glwe/dcrt.rs:108-126; reduce_dot_product, primus_modulus/src/common/compact/slice.rs:371-401) -> optional inverse NTT.
 * key: device [k+1][levels][k+1][limbs][N] NTT domain; in/out: device [batch][k+1][limbs][N].
 * `scratch`: device memory for the digits; any size >= ..._scratch_bytes(batch = 1) works, the batch is processed in
 * chunks that fit (no allocation on the hot path).  When the composed value Q fits two words (e.g. two 50-bit limbs) the
 * product runs as ONE kernel with the digits in registers and `scratch` is not touched (it may be NULL). */
size_t pfhe_dcrt64_external_product_scratch_bytes(const pfhe_dcrt64 *t, const pfhe_rns64 *r, uint32_t k, uint32_t log_basis,
                                                  uint32_t levels_in, size_t batch);
size_t pfhe_dcrt32_external_product_scratch_bytes(const pfhe_dcrt32 *t, const pfhe_rns32 *r, uint32_t k, uint32_t log_basis,
                                                  uint32_t levels_in, size_t batch);
pfhe_status pfhe_dcrt64_external_product_batch(const pfhe_dcrt64 *t, const pfhe_rns64 *r, uint32_t k, uint32_t log_basis,
                                               uint32_t levels_in, const uint64_t *key, const uint64_t *in, uint64_t *out, size_t batch,
                                               int to_coeff, void *scratch, size_t scratch_bytes, void *stream);
pfhe_status pfhe_dcrt32_external_product_batch(const pfhe_dcrt32 *t, const pfhe_rns32 *r, uint32_t k, uint32_t log_basis,
                                               uint32_t levels_in, const uint32_t *key, const uint32_t *in, uint32_t *out, size_t batch,
                                               int to_coeff, void *scratch, size_t scratch_bytes, void *stream);

/* BaseConverter (primus_rns/src/converter.rs:21-365): RNS base change with the precomputed matrix (Q/q_i) mod p_k.
 * Host-only handle; up to 8 input and 8 output moduli.  Layout: [polys][moduli][n] (modulus-major per polynomial). */
typedef struct pfhe_baseconv32 pfhe_baseconv32;
typedef struct pfhe_baseconv64 pfhe_baseconv64;
pfhe_status pfhe_baseconv64_create(const uint64_t *in_moduli, size_t n_in, const uint64_t *out_moduli, size_t n_out, pfhe_baseconv64 **out);
pfhe_status pfhe_baseconv32_create(const uint32_t *in_moduli, size_t n_in, const uint32_t *out_moduli, size_t n_out, pfhe_baseconv32 **out);
void pfhe_baseconv64_destroy(pfhe_baseconv64 *c);
void pfhe_baseconv32_destroy(pfhe_baseconv32 *c);
/* BaseConverter::fast_convert_array (converter.rs:186-213): in [polys][n_in][n] -> out [polys][n_out][n] (device). */
pfhe_status pfhe_baseconv64_fast_convert_batch(const pfhe_baseconv64 *c, const uint64_t *in, uint64_t *out, size_t n, size_t polys, void *stream);
pfhe_status pfhe_baseconv32_fast_convert_batch(const pfhe_baseconv32 *c, const uint32_t *in, uint32_t *out, size_t n, size_t polys, void *stream);
/* BaseConverter::exact_convert_array (converter.rs:257-365; exactly one output modulus, else INVALID_ARG):
 * in [polys][n_in][n] -> out [polys][n]; the floating correction term follows the reference's operation order. */
pfhe_status pfhe_baseconv64_exact_convert_batch(const pfhe_baseconv64 *c, const uint64_t *in, uint64_t *out, size_t n, size_t polys, void *stream);
pfhe_status pfhe_baseconv32_exact_convert_batch(const pfhe_baseconv32 *c, const uint32_t *in, uint32_t *out, size_t n, size_t polys, void *stream);

/* Polynomial::mul_monomial_assign / CrtGlwe::mul_monic_monomial_assign (primus_poly/src/poly/mul.rs:74-99,
 * primus_lattice/src/glwe/crt.rs:76-114): out = in * X^degree in Z_q[X]/(X^N+1) per limb; in/out [batch][limbs][N]
 * device (out != in), degrees[batch] device, each taken mod 2N. */
pfhe_status pfhe_poly64_mul_monomial_batch(const uint64_t *moduli, size_t limbs, const uint32_t *degrees, const uint64_t *in,
                                           uint64_t *out, uint32_t log_n, size_t batch, void *stream);
pfhe_status pfhe_poly32_mul_monomial_batch(const uint32_t *moduli, size_t limbs, const uint32_t *degrees, const uint32_t *in,
                                           uint32_t *out, uint32_t log_n, size_t batch, void *stream);
/* reduce_dot_product (primus_modulus/src/common/compact/slice.rs:371-438): out[row] = sum_i a[row][i]*b[row][i] mod q. */
pfhe_status pfhe_mod64_dot_product_batch(uint64_t q, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t rows, size_t n, void *stream);
pfhe_status pfhe_mod32_dot_product_batch(uint32_t q, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t rows, size_t n, void *stream);

/* ===================================================================================== */
/* Bootstrapping keys and the host-slice bootstrap (round 2)                               */
/* ===================================================================================== */
/* A bootstrapping key = n_lwe RGSW ciphertexts in NTT form, [n_lwe][2][levels][2][N] words (NttRgsw layout,
 * primus_lattice/src/ggsw/dcrt.rs:14-31 with L = 1), uploaded once to the table's device and kept resident like a table
 * (the reference keeps keys in host `Vec`s; a GPU caller holds this handle instead).  `bsk` is HOST memory.
 * `_create_from_bytes` takes the reference's serialised form (`to_bytes`, raw little-endian words, macros/mod.rs:39-97). */
typedef struct pfhe_bsk32 pfhe_bsk32;
typedef struct pfhe_bsk64 pfhe_bsk64;
pfhe_status pfhe_bsk64_create(const pfhe_ntt64 *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe, const uint64_t *bsk, pfhe_bsk64 **out);
pfhe_status pfhe_bsk32_create(const pfhe_ntt32 *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe, const uint32_t *bsk, pfhe_bsk32 **out);
pfhe_status pfhe_bsk64_create_from_bytes(const pfhe_ntt64 *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe, const uint8_t *bytes,
                                         size_t byte_count, pfhe_bsk64 **out);
pfhe_status pfhe_bsk32_create_from_bytes(const pfhe_ntt32 *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe, const uint8_t *bytes,
                                         size_t byte_count, pfhe_bsk32 **out);
void pfhe_bsk64_destroy(pfhe_bsk64 *b);
void pfhe_bsk32_destroy(pfhe_bsk32 *b);
uint32_t pfhe_bsk64_lwe_dimension(const pfhe_bsk64 *b);
uint32_t pfhe_bsk32_lwe_dimension(const pfhe_bsk32 *b);
uint32_t pfhe_bsk64_levels(const pfhe_bsk64 *b);
uint32_t pfhe_bsk32_levels(const pfhe_bsk32 *b);
const uint64_t *pfhe_bsk64_device_ptr(const pfhe_bsk64 *b); /* device pointer for the *_batch entry points */
const uint32_t *pfhe_bsk32_device_ptr(const pfhe_bsk32 *b);
/* Bootstrap on HOST slices: lwe [batch][n_lwe+1] uint32 in Z_2N (a_0..a_{n-1}, b), test_vector [N];
 * extract != 0: out [batch][N+1] = Rlwe::extract_lwe (rlwe/coeff.rs:264-288) of the rotated accumulator;
 * extract == 0: out [batch][2][N] = the accumulator itself.  H2D -> blind rotation -> D2H inside the call. */
pfhe_status pfhe_bootstrap64_slices(const pfhe_ntt64 *t, const pfhe_bsk64 *bsk, const uint32_t *lwe, const uint64_t *test_vector,
                                    uint64_t *out, size_t batch, int extract);
pfhe_status pfhe_bootstrap32_slices(const pfhe_ntt32 *t, const pfhe_bsk32 *bsk, const uint32_t *lwe, const uint32_t *test_vector,
                                    uint32_t *out, size_t batch, int extract);
/* Ternary-secret blind rotation by monomial combination (composed from the reference's primitives, parity unpinned
 * like the binary form): for every LWE coefficient a_i the accumulator is multiplied by RGSW(X^(a_i s_i)), s_i in {-1,0,1}, with ONE
 * external product against  K_i = (NTT(X^a_i) - 1) .* BSK+_i + (NTT(X^-a_i) - 1) .* BSK-_i, where NTT(X^d) is
 * NttTable::transform_coeff_one_monomial(d) (primus_ntt/src/ntt/prime64/table.rs:611-651) and BSK+-_i = RGSW([s_i = +-1]) in NTT form:
 *     ACC <- ACC + INTT( sum_{r,l} NTT(digit_l(ACC_r)) .* K_i[r][l][c] ).
 * bsk_plus / bsk_minus: device [n_lwe][2][levels][2][N]; other arguments as pfhe_blind_rotate32_batch.
 * Implemented for the bootstrapping shape of BASELINE config 5 (u32 words, N = 1024); PFHE_ERR_UNSUPPORTED otherwise. */
pfhe_status pfhe_blind_rotate_ternary32_batch(const pfhe_ntt32 *t, uint32_t log_basis, uint32_t levels_in, const uint32_t *bsk_plus,
                                              const uint32_t *bsk_minus, uint32_t n_lwe, const uint32_t *lwe, const uint32_t *test_vector,
                                              uint32_t *acc_out, size_t batch, void *stream);
/* LWE modulus switch to Z_2N before blind rotation (NOT in the reference -- it has no bootstrapping; convention fixed here and
 * in the oracle): out[i] = floor((lwe[i] * 2N + floor(q/2)) / q) mod 2N, exact integer arithmetic; 2N = 2^log_2n.
 * lwe: device, `count` canonical words (all a_i and b of a batch); out: device uint32. */
pfhe_status pfhe_lwe64_modulus_switch_batch(uint64_t q, uint32_t log_2n, const uint64_t *lwe, uint32_t *out, size_t count, void *stream);
pfhe_status pfhe_lwe32_modulus_switch_batch(uint32_t q, uint32_t log_2n, const uint32_t *lwe, uint32_t *out, size_t count, void *stream);

/* ===================================================================================== */
/* Host buffers of the *_slices entry points                                               */
/* ===================================================================================== */
/* The *_slices shims take plain host pointers -- the `&mut [T]` of NttTable::transform_slice (primus_data/src/traits.rs:20) lives in
 * pageable memory -- and stage pageable data through pinned bounce buffers (about half the PCIe rate: one extra pass over host DRAM per
 * direction).  A caller that reuses a long-lived buffer can page-lock it IN PLACE once; the shims recognise registered memory
 * (cudaPointerGetAttributes) and then copy straight from / to it at the full PCIe rate.  The registration is portable across devices
 * (the multi-device drivers use it too).  The caller must unregister before freeing or reallocating the buffer.
 * PFHE_ERR_INVALID_ARG: null pointer / zero length; PFHE_ERR_CUDA: the range cannot be locked (already registered, limits). */
pfhe_status pfhe_host_register(void *host_ptr, size_t bytes);
pfhe_status pfhe_host_unregister(void *host_ptr);
/* 1 if the shims would stage `host_ptr` through bounce buffers (pageable memory), 0 if it is page-locked / registered */
int pfhe_host_is_pageable(const void *host_ptr);

/* ===================================================================================== */
/* Multi-device drivers (round 2): one process, several GPUs, no torch                     */
/* ===================================================================================== */
/* NttTable is Send + Sync (primus_ntt/src/ntt/mod.rs:16): the reference lets a caller fan a batch out over threads.  Here the
 * fan-out is over devices: `tables[i]` is the same (log_n, q) table created on device i (pfhe_multi_ntt*_create replicates it),
 * the batch is split into contiguous shards (sizes differ by at most one: shard r of `total` starts at
 * r*floor(total/n) + min(r, total mod n)), one host thread with its own stream set drives each device, and the call returns
 * when every shard is back in host memory.  No collective on the data path.  The same device may appear
 * more than once (two handles on one GPU) -- the shards then overlap on that GPU. */
pfhe_status pfhe_multi_ntt64_create(const int *devices, size_t n_devices, uint32_t log_n, uint64_t q, pfhe_ntt64 **out_tables);
pfhe_status pfhe_multi_ntt32_create(const int *devices, size_t n_devices, uint32_t log_n, uint32_t q, pfhe_ntt32 **out_tables);
pfhe_status pfhe_multi_ntt64_transform_slices(const pfhe_ntt64 *const *tables, size_t n_devices, uint64_t *polys, size_t batch, int inverse, int lazy);
pfhe_status pfhe_multi_ntt32_transform_slices(const pfhe_ntt32 *const *tables, size_t n_devices, uint32_t *polys, size_t batch, int inverse, int lazy);
pfhe_status pfhe_multi_ntt64_polymul_slices(const pfhe_ntt64 *const *tables, size_t n_devices, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch);
pfhe_status pfhe_multi_ntt32_polymul_slices(const pfhe_ntt32 *const *tables, size_t n_devices, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch);
pfhe_status pfhe_multi_ggsw64_external_product_slices(const pfhe_ntt64 *const *tables, size_t n_devices, uint32_t k, uint32_t log_basis,
                                                      uint32_t levels_in, const uint64_t *key, const uint64_t *in, uint64_t *out, size_t batch, int to_coeff);
pfhe_status pfhe_multi_ggsw32_external_product_slices(const pfhe_ntt32 *const *tables, size_t n_devices, uint32_t k, uint32_t log_basis,
                                                      uint32_t levels_in, const uint32_t *key, const uint32_t *in, uint32_t *out, size_t batch, int to_coeff);
pfhe_status pfhe_multi_bootstrap64_slices(const pfhe_ntt64 *const *tables, const pfhe_bsk64 *const *bsks, size_t n_devices, const uint32_t *lwe,
                                          const uint64_t *test_vector, uint64_t *out, size_t batch, int extract);
pfhe_status pfhe_multi_bootstrap32_slices(const pfhe_ntt32 *const *tables, const pfhe_bsk32 *const *bsks, size_t n_devices, const uint32_t *lwe,
                                          const uint32_t *test_vector, uint32_t *out, size_t batch, int extract);

/* ===================================================================================== */
/* Whole-ciphertext transforms and the byte wire format (round 2)                          */
/* ===================================================================================== */
/* into_ntt_form / into_coeff_form / write_ntt_form / write_coeff_form of every ciphertext container
 * (primus_lattice/src/macros/mod.rs:537-621 impl_ntt / impl_intt; CRT forms :623-674, :892-937): transform every polynomial of
 * the flat HOST storage (`words` must be a multiple of N, resp. L*N) in one pipelined call.  `write_*` copies src to dst first,
 * exactly like `result.0.copy_from_slice(self.as_ref())`. */
pfhe_status pfhe_cipher64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data, size_t words);
pfhe_status pfhe_cipher32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data, size_t words);
pfhe_status pfhe_cipher64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data, size_t words);
pfhe_status pfhe_cipher32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data, size_t words);
pfhe_status pfhe_cipher64_write_ntt_form(const pfhe_ntt64 *t, const uint64_t *src, uint64_t *dst, size_t words);
pfhe_status pfhe_cipher32_write_ntt_form(const pfhe_ntt32 *t, const uint32_t *src, uint32_t *dst, size_t words);
pfhe_status pfhe_cipher64_write_coeff_form(const pfhe_ntt64 *t, const uint64_t *src, uint64_t *dst, size_t words);
pfhe_status pfhe_cipher32_write_coeff_form(const pfhe_ntt32 *t, const uint32_t *src, uint32_t *dst, size_t words);
pfhe_status pfhe_dcrt_cipher64_into_ntt_form(const pfhe_dcrt64 *t, uint64_t *data, size_t words);   /* CrtGlwe/CrtGlev/CrtGgsw -> Dcrt* */
pfhe_status pfhe_dcrt_cipher32_into_ntt_form(const pfhe_dcrt32 *t, uint32_t *data, size_t words);
pfhe_status pfhe_dcrt_cipher64_into_coeff_form(const pfhe_dcrt64 *t, uint64_t *data, size_t words); /* Dcrt* -> Crt* (macros/mod.rs:892-937) */
pfhe_status pfhe_dcrt_cipher32_into_coeff_form(const pfhe_dcrt32 *t, uint32_t *data, size_t words);
/* Named containers: word counts of the reference's flat layouts and the matching transforms.
 * Rlwe [2][N] (rlwe/coeff.rs), Rlev [levels][2][N], Rgsw [2][levels][2][N], Glwe [k+1][N] (glwe/), Glev [levels][k+1][N],
 * Ggsw [k+1][levels][k+1][N] (ggsw/dcrt.rs:14-31). */
size_t pfhe_rlwe64_words(const pfhe_ntt64 *t);
size_t pfhe_rlwe32_words(const pfhe_ntt32 *t);
size_t pfhe_rlev64_words(const pfhe_ntt64 *t, uint32_t levels);
size_t pfhe_rlev32_words(const pfhe_ntt32 *t, uint32_t levels);
size_t pfhe_rgsw64_words(const pfhe_ntt64 *t, uint32_t levels);
size_t pfhe_rgsw32_words(const pfhe_ntt32 *t, uint32_t levels);
size_t pfhe_glwe64_words(const pfhe_ntt64 *t, uint32_t k);
size_t pfhe_glwe32_words(const pfhe_ntt32 *t, uint32_t k);
size_t pfhe_glev64_words(const pfhe_ntt64 *t, uint32_t k, uint32_t levels);
size_t pfhe_glev32_words(const pfhe_ntt32 *t, uint32_t k, uint32_t levels);
size_t pfhe_ggsw64_words(const pfhe_ntt64 *t, uint32_t k, uint32_t levels);
size_t pfhe_ggsw32_words(const pfhe_ntt32 *t, uint32_t k, uint32_t levels);
pfhe_status pfhe_rlwe64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data);
pfhe_status pfhe_rlwe32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data);
pfhe_status pfhe_rlwe64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data);
pfhe_status pfhe_rlwe32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data);
pfhe_status pfhe_rlev64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t levels);
pfhe_status pfhe_rlev32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t levels);
pfhe_status pfhe_rlev64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t levels);
pfhe_status pfhe_rlev32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t levels);
pfhe_status pfhe_rgsw64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t levels);
pfhe_status pfhe_rgsw32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t levels);
pfhe_status pfhe_rgsw64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t levels);
pfhe_status pfhe_rgsw32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t levels);
pfhe_status pfhe_glwe64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t k);
pfhe_status pfhe_glwe32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t k);
pfhe_status pfhe_glwe64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t k);
pfhe_status pfhe_glwe32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t k);
pfhe_status pfhe_glev64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_glev32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_glev64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_glev32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_ggsw64_into_ntt_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_ggsw32_into_ntt_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_ggsw64_into_coeff_form(const pfhe_ntt64 *t, uint64_t *data, uint32_t k, uint32_t levels);
pfhe_status pfhe_ggsw32_into_coeff_form(const pfhe_ntt32 *t, uint32_t *data, uint32_t k, uint32_t levels);
/* from_bytes / read_bytes / to_bytes / write_bytes / byte_count (macros/mod.rs:39-97: `bytemuck::cast_slice`, i.e. the raw
 * little-endian bytes of the flat word array -- also the host<->device wire format, so a serialised key can be uploaded as is).
 * Unlike bytemuck, unaligned byte buffers are accepted.  byte_count must equal word_count * sizeof(word). */
pfhe_status pfhe_cipher64_read_bytes(uint64_t *words, size_t word_count, const uint8_t *bytes, size_t byte_count);
pfhe_status pfhe_cipher32_read_bytes(uint32_t *words, size_t word_count, const uint8_t *bytes, size_t byte_count);
pfhe_status pfhe_cipher64_write_bytes(const uint64_t *words, size_t word_count, uint8_t *bytes, size_t byte_count);
pfhe_status pfhe_cipher32_write_bytes(const uint32_t *words, size_t word_count, uint8_t *bytes, size_t byte_count);
size_t pfhe_cipher64_byte_count(size_t word_count);
size_t pfhe_cipher32_byte_count(size_t word_count);

/* ===================================================================================== */
/* UintNttTable<T> (primus_ntt/src/ntt/primitive.rs:37-396) -- the generic table (round 2)   */
/* ===================================================================================== */
/* A distinct table type with the reference's constructor rules: NoPrimitiveRoot (root.rs:72-81), DegreeConversionErr when
 * N does not fit the word type (primitive.rs:160-161), DegreeTooLarge when N >= q (:163-168); ModulusTooLarge when 4q does not
 * fit the word (the lazy butterflies keep values in [0,4q), :219-236).  Words: u16 / u32 / u64 (FheUint,
 * primus_integer/src/unsigned_integer.rs:27-112).  Always runs the plain radix-2 kernel (the reference's cross-check
 * implementation: canonical results equal U32/U64NttTable, prime64/tests.rs:78-237). */
typedef struct pfhe_uintntt16 pfhe_uintntt16;
typedef struct pfhe_uintntt32 pfhe_uintntt32;
typedef struct pfhe_uintntt64 pfhe_uintntt64;
pfhe_status pfhe_uintntt16_create(int device, uint32_t log_n, uint16_t q, pfhe_uintntt16 **out);
pfhe_status pfhe_uintntt32_create(int device, uint32_t log_n, uint32_t q, pfhe_uintntt32 **out);
pfhe_status pfhe_uintntt64_create(int device, uint32_t log_n, uint64_t q, pfhe_uintntt64 **out);
void pfhe_uintntt16_destroy(pfhe_uintntt16 *t);
void pfhe_uintntt32_destroy(pfhe_uintntt32 *t);
void pfhe_uintntt64_destroy(pfhe_uintntt64 *t);
size_t pfhe_uintntt16_poly_length(const pfhe_uintntt16 *t);
size_t pfhe_uintntt32_poly_length(const pfhe_uintntt32 *t);
size_t pfhe_uintntt64_poly_length(const pfhe_uintntt64 *t);
uint16_t pfhe_uintntt16_root(const pfhe_uintntt16 *t);
uint32_t pfhe_uintntt32_root(const pfhe_uintntt32 *t);
uint64_t pfhe_uintntt64_root(const pfhe_uintntt64 *t);
uint16_t pfhe_uintntt16_inv_root(const pfhe_uintntt16 *t);
uint32_t pfhe_uintntt32_inv_root(const pfhe_uintntt32 *t);
uint64_t pfhe_uintntt64_inv_root(const pfhe_uintntt64 *t);
/* the same table viewed through the NttTable entry points (monomial transforms, device batches) */
const pfhe_ntt32 *pfhe_uintntt32_as_table(const pfhe_uintntt32 *t);
const pfhe_ntt64 *pfhe_uintntt64_as_table(const pfhe_uintntt64 *t);
pfhe_status pfhe_uintntt16_transform_slices(const pfhe_uintntt16 *t, uint16_t *polys, size_t batch, int lazy);
pfhe_status pfhe_uintntt32_transform_slices(const pfhe_uintntt32 *t, uint32_t *polys, size_t batch, int lazy);
pfhe_status pfhe_uintntt64_transform_slices(const pfhe_uintntt64 *t, uint64_t *polys, size_t batch, int lazy);
pfhe_status pfhe_uintntt16_inverse_transform_slices(const pfhe_uintntt16 *t, uint16_t *polys, size_t batch, int lazy);
pfhe_status pfhe_uintntt32_inverse_transform_slices(const pfhe_uintntt32 *t, uint32_t *polys, size_t batch, int lazy);
pfhe_status pfhe_uintntt64_inverse_transform_slices(const pfhe_uintntt64 *t, uint64_t *polys, size_t batch, int lazy);

/* ===================================================================================== */
/* Plumbing for hosts without a CUDA binding of their own (the Rust FFI crate, tests)      */
/* ===================================================================================== */
pfhe_status pfhe_device_count(int *count);
pfhe_status pfhe_malloc(int device, size_t bytes, void **dev_ptr);
pfhe_status pfhe_free(int device, void *dev_ptr);
pfhe_status pfhe_memcpy_h2d(int device, void *dev_dst, const void *host_src, size_t bytes, void *stream);
pfhe_status pfhe_memcpy_d2h(int device, void *host_dst, const void *dev_src, size_t bytes, void *stream);
pfhe_status pfhe_stream_synchronize(int device, void *stream);

/* Integer-pipe microbenchmark used to measure the modmul roofline denominator
 * (SURVEY.md 8d): runs `iters` dependent Shoup modmuls per thread in registers on
 * `blocks` x 256 threads and returns elapsed milliseconds via *ms. kind: 0 = u32 Harvey butterflies, 1 = u64 Harvey
 * butterflies, 2 = bare u32 Shoup products (1 high + 2 low multiplies), 3 = bare u64 Shoup products; 8 per thread per iteration. */
pfhe_status pfhe_modmul_microbench(int device, int kind, uint32_t blocks, uint32_t iters, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* PFHE_H */
