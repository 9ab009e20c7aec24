// pointwise.cu -- HBM-streaming element-wise kernels: modular slice operators, gadget decomposition,
// RNS centred lift, LWE sample extraction, and the integer-pipe microbenchmark.
//
// Replaces:
//   ReduceMulSlice / ReduceMulAddSlice / Reduce{Add,Sub,Neg}Slice     primus_reduce/src/slice_ops.rs:43-230
//     (bodies primus_modulus/src/common/compact/slice.rs:106-365, Barrett primus_modulus/src/barrett/ops.rs:276-322)
//   FactorSliceOps with a ShoupFactor                                  primus_factor/src/ops.rs:58-118, common/slice.rs:7-107
//   per-limb DcrtPolynomial mul / add_mul / factor ops                 primus_poly/src/dcrt/mul.rs:142-187, dcrt/mod.rs:105-123
//   ApproxSignedBasis init + OnceSignedDecomposer levels              primus_decompose/src/primitive/basis.rs:254-406, common.rs:219-273
//   RNSBase::wrapping_decompose_small_values_to                        primus_rns/src/base.rs:279-315, :721-731
//   Rlwe::extract_lwe                                                  primus_lattice/src/rlwe/coeff.rs:264-288
// All kernels are pure streaming: 128-bit vector accesses, grid sized in multiples of the SM count.
#include "internal.hpp"
#include "host_math.hpp"
#include "pfhe.h"

namespace pfhe {

constexpr int kSMs = 148;

template <typename T> struct VecOf {
    static constexpr int W = 16 / sizeof(T);
    struct alignas(16) type {
        T v[W];
    };
};

template <typename T, int OP>
__device__ __forceinline__ T apply_op(const Barrett<T> &br, T s, T sq, T a, T b, T c, T o) {
    const T q = br.q;
    switch (OP) {
        case PFHE_OP_MUL: return barrett_mul<T>(br, a, b);
        case PFHE_OP_ADD_MUL: return barrett_mul_add<T>(br, a, b, o);
        case PFHE_OP_SUB_MUL: return mod_sub<T>(o, barrett_mul<T>(br, a, b), q);
        case PFHE_OP_MUL_ADD: return barrett_mul_add<T>(br, a, b, c);
        case PFHE_OP_ADD: return mod_add<T>(a, b, q);
        case PFHE_OP_SUB: return mod_sub<T>(a, b, q);
        case PFHE_OP_NEG: return mod_neg<T>(a, q);
        case PFHE_OP_MUL_SCALAR: return barrett_mul<T>(br, a, s);
        case PFHE_OP_ADD_MUL_SCALAR: return barrett_mul_add<T>(br, a, s, o);
        case PFHE_OP_FACTOR_MUL: return shoup<T>(a, s, sq, q);
        case PFHE_OP_ADD_FACTOR_MUL: return mod_add<T>(o, shoup<T>(a, s, sq, q), q);
        case PFHE_OP_SUB_FACTOR_MUL: return mod_sub<T>(o, shoup<T>(a, s, sq, q), q);
        case PFHE_OP_REDUCE_LAZY: return csub<T>(csub<T>(a, q + q), q);
        case PFHE_OP_DOUBLE: return mod_add<T>(a, a, q);
        case PFHE_OP_MUL_SCALAR_ADD: return barrett_mul_add<T>(br, a, s, c);
        case PFHE_OP_FACTOR_MUL_ADD: return mod_add<T>(shoup<T>(a, s, sq, q), c, q);
    }
    return 0;
}

constexpr bool op_reads_b(int op) { return op == PFHE_OP_MUL || op == PFHE_OP_ADD_MUL || op == PFHE_OP_SUB_MUL || op == PFHE_OP_MUL_ADD || op == PFHE_OP_ADD || op == PFHE_OP_SUB; }
constexpr bool op_reads_c(int op) { return op == PFHE_OP_MUL_ADD || op == PFHE_OP_MUL_SCALAR_ADD || op == PFHE_OP_FACTOR_MUL_ADD; }
constexpr bool op_reads_out(int op) { return op == PFHE_OP_ADD_MUL || op == PFHE_OP_SUB_MUL || op == PFHE_OP_ADD_MUL_SCALAR || op == PFHE_OP_ADD_FACTOR_MUL || op == PFHE_OP_SUB_FACTOR_MUL; }

// reduce_mul on the FP64 pipe (u64 words, q <= 2^50 - 2^10, canonical operands): 9 FP64 instructions per product instead of ~9 64-bit integer
// multiplies -- the integer Barrett product is what kept reduce_mul_slice at 0.80 of the HBM copy peak while the additive operators reach 0.93-0.95.
// A vector with a non-canonical word (>= q) takes the integer product.
__device__ __forceinline__ uint64_t f64_mul_canonical(uint64_t a, uint64_t b, double qf, double qinvf, uint64_t q) {
    using F = F64LazyField;
    const F::Ctx cx{qf, qinvf, 0.0, F::kTwo52 + qf, q, 0};
    const double r = F::mulmod(F::from_u64(a), F::from_u64(b), cx, 0);          // |r| <= 0.75 q
    return csub<uint64_t>(F::mant(__dadd_rn(r, cx.off1)), q);                    // r + q in (0.25 q, 1.75 q)
}

// slices are [rows][limbs][n]; vectorised when n % W == 0 (always true for polynomial lengths >= 4)
template <typename T, int OP, bool VEC, bool F64 = false>
__global__ void __launch_bounds__(256) slice_op_kernel(const __grid_constant__ LimbConsts<T> lc, int limbs, const T *a, const T *b, const T *c,
                                                       T *out, size_t rows, size_t n, size_t b_group) {
    // no __restrict__: the *_assign operators of the reference run in place (out aliases a, b or c)
    // b_group > 1: row r of `a` pairs with row r / b_group of `b` (one polynomial against every component of a ciphertext)
    constexpr int W = VEC ? VecOf<T>::W : 1;
    using V = typename VecOf<T>::type;
    const size_t per_row = n / W, total = rows * (size_t)limbs * per_row;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const int limb = limbs == 1 ? 0 : (int)((gid / per_row) % (size_t)limbs);
        const Barrett<T> br = lc.br[limb];
        const T s = lc.scalar[limb], sq = lc.scalar_q[limb];
        size_t bgid = gid;
        if (op_reads_b(OP) && b_group > 1) {
            const size_t row_len = (size_t)limbs * per_row, row = gid / row_len;
            bgid = (row / b_group) * row_len + gid % row_len;
        }
        if (VEC) {
            V va = reinterpret_cast<const V *>(a)[gid], vb, vc, vo;
            if (op_reads_b(OP)) vb = reinterpret_cast<const V *>(b)[bgid];
            if (op_reads_c(OP)) vc = reinterpret_cast<const V *>(c)[gid];
            if (op_reads_out(OP)) vo = reinterpret_cast<const V *>(out)[gid];
            bool done = false;
            if constexpr (F64 && OP == PFHE_OP_MUL && sizeof(T) == 8) {
                bool canonical = true;
#pragma unroll
                for (int k = 0; k < W; k++) canonical = canonical && va.v[k] < br.q && vb.v[k] < br.q;
                if (canonical) {
#pragma unroll
                    for (int k = 0; k < W; k++) vo.v[k] = f64_mul_canonical(va.v[k], vb.v[k], lc.q_f[limb], lc.qinv_f[limb], br.q);
                    done = true;
                }
            }
            if (!done) {
#pragma unroll
                for (int k = 0; k < W; k++)
                    vo.v[k] = apply_op<T, OP>(br, s, sq, va.v[k], op_reads_b(OP) ? vb.v[k] : T(0), op_reads_c(OP) ? vc.v[k] : T(0),
                                              op_reads_out(OP) ? vo.v[k] : T(0));
            }
            reinterpret_cast<V *>(out)[gid] = vo;
        } else {
            out[gid] = apply_op<T, OP>(br, s, sq, a[gid], op_reads_b(OP) ? b[bgid] : T(0), op_reads_c(OP) ? c[gid] : T(0),
                                       op_reads_out(OP) ? out[gid] : T(0));
        }
    }
}

static unsigned stream_grid(size_t work_items, int threads) {
    size_t blocks = (work_items + threads - 1) / threads;
    const size_t cap = (size_t)kSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return (unsigned)blocks;
}

template <typename T, int OP>
static cudaError_t run_slice_op(const LimbConsts<T> &lc, int limbs, const T *a, const T *b, const T *c, T *out, size_t rows, size_t n,
                                size_t b_group, cudaStream_t stream) {
    constexpr int W = VecOf<T>::W;
    auto aligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = (n % W == 0) && aligned(a) && aligned(out) && (!b || aligned(b)) && (!c || aligned(c));
    const size_t total = rows * (size_t)limbs * (vec ? n / W : n);
    if (total == 0) return cudaSuccess;
    if constexpr (OP == PFHE_OP_MUL && sizeof(T) == 8) {
        static const bool f64_off = getenv("PFHE_DISABLE_F64") && getenv("PFHE_DISABLE_F64")[0] == '1';
        bool f64 = vec && !f64_off;
        for (int i = 0; i < limbs && f64; i++) f64 = (uint64_t)lc.br[i].q <= (((uint64_t)1 << 50) - 1024) && lc.br[i].q > 1;
        if (f64) {
            LimbConsts<T> lf = lc;
            for (int i = 0; i < limbs; i++) {
                lf.q_f[i] = (double)lc.br[i].q;
                lf.qinv_f[i] = 1.0 / (double)lc.br[i].q;
            }
            slice_op_kernel<T, OP, true, true><<<stream_grid(total, 256), 256, 0, stream>>>(lf, limbs, a, b, c, out, rows, n, b_group);
            count_launch();
            return cudaGetLastError();
        }
    }
    if (vec)
        slice_op_kernel<T, OP, true><<<stream_grid(total, 256), 256, 0, stream>>>(lc, limbs, a, b, c, out, rows, n, b_group);
    else
        slice_op_kernel<T, OP, false><<<stream_grid(total, 256), 256, 0, stream>>>(lc, limbs, a, b, c, out, rows, n, b_group);
    count_launch();
    return cudaGetLastError();
}

// Inverse butterfly with a factor polynomial, per limb: (a, out) = (a + s, (a - s) * w)   (all in [0, q))
// DcrtPolynomial::butterfly_mul_factor_to  primus_poly/src/dcrt/mul.rs:189-222 (DcrtGlwe form: primus_lattice/src/glwe/dcrt.rs:150-175).
// The reference multiplies with a precomputed ShoupFactor array; the product is the same residue, computed here by Barrett.
template <typename T>
__global__ void __launch_bounds__(256) butterfly_mul_kernel(const __grid_constant__ LimbConsts<T> lc, int limbs, T *__restrict__ a,
                                                            const T *__restrict__ s, const T *__restrict__ w, T *__restrict__ out, size_t rows,
                                                            size_t n) {
    const size_t total = rows * (size_t)limbs * n;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const int limb = limbs == 1 ? 0 : (int)((gid / n) % (size_t)limbs);
        const Barrett<T> br = lc.br[limb];
        const T x = a[gid], y = s[gid];
        // w is one factor polynomial per limb, shared by every row (factor_poly: [limbs][n])
        const T f = w[(size_t)limb * n + gid % n];
        a[gid] = mod_add<T>(x, y, br.q);
        out[gid] = barrett_mul<T>(br, mod_sub<T>(x, y, br.q), f);
    }
}
template <typename T>
cudaError_t launch_butterfly_mul(const LimbConsts<T> &lc, int limbs, T *a, const T *s, const T *w, T *out, size_t rows, size_t n,
                                 cudaStream_t stream) {
    const size_t total = rows * (size_t)limbs * n;
    if (!total) return cudaSuccess;
    butterfly_mul_kernel<T><<<stream_grid(total, 256), 256, 0, stream>>>(lc, limbs, a, s, w, out, rows, n);
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_butterfly_mul<uint32_t>(const LimbConsts<uint32_t> &, int, uint32_t *, const uint32_t *, const uint32_t *, uint32_t *,
                                                    size_t, size_t, cudaStream_t);
template cudaError_t launch_butterfly_mul<uint64_t>(const LimbConsts<uint64_t> &, int, uint64_t *, const uint64_t *, const uint64_t *, uint64_t *,
                                                    size_t, size_t, cudaStream_t);

// Element-wise modular inverse over a prime modulus (every NTT modulus is prime): a^(q-2) by square-and-multiply with
// Barrett products.  Replaces ReduceInvSlice::reduce_inv_slice_to / NttPolynomial::inv_to (primus_poly/src/ntt/inv.rs:1-58;
// the CPU uses Montgomery's batch-inversion trick, primus_modulus/src/barrett/slice.rs:499-625 -- a serial prefix product
// that does not map to one-thread-per-element; the inverse is unique, so the results are the same bits).
// Non-invertible (zero) inputs give 0 and, like try_reduce_inv_slice_to -> ReduceError::NoInverseAtIndex, report the
// smallest offending index through *first_bad (atomicMin; initialise to ~0ull).
template <typename T>
__global__ void __launch_bounds__(256) inv_slice_kernel(const Barrett<T> br, const T *__restrict__ a, T *__restrict__ out, size_t count,
                                                        unsigned long long *first_bad) {
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < count; gid += (size_t)gridDim.x * blockDim.x) {
        const T x = a[gid];
        T e = br.q - 2, base = x, r = 1;
        while (e) {
            if (e & 1) r = barrett_mul<T>(br, r, base);
            base = barrett_mul<T>(br, base, base);
            e >>= 1;
        }
        if (x == 0) {
            r = 0;
            if (first_bad) atomicMin(first_bad, (unsigned long long)gid);
        }
        out[gid] = r;
    }
}
template <typename T>
cudaError_t launch_inv_slice(const Barrett<T> &br, const T *a, T *out, size_t count, unsigned long long *first_bad, cudaStream_t stream) {
    if (!count) return cudaSuccess;
    inv_slice_kernel<T><<<stream_grid(count, 256), 256, 0, stream>>>(br, a, out, count, first_bad);
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_inv_slice<uint32_t>(const Barrett<uint32_t> &, const uint32_t *, uint32_t *, size_t, unsigned long long *, cudaStream_t);
template cudaError_t launch_inv_slice<uint64_t>(const Barrett<uint64_t> &, const uint64_t *, uint64_t *, size_t, unsigned long long *, cudaStream_t);

template <typename T>
cudaError_t launch_slice_op(int op, const LimbConsts<T> &lc, int limbs, const T *a, const T *b, const T *c, T *out, size_t rows, size_t n,
                            cudaStream_t s, size_t b_group) {
    switch (op) {
#define PFHE_CASE(OPC) \
    case OPC: return run_slice_op<T, OPC>(lc, limbs, a, b, c, out, rows, n, b_group, s);
        PFHE_CASE(PFHE_OP_MUL)
        PFHE_CASE(PFHE_OP_ADD_MUL)
        PFHE_CASE(PFHE_OP_SUB_MUL)
        PFHE_CASE(PFHE_OP_MUL_ADD)
        PFHE_CASE(PFHE_OP_ADD)
        PFHE_CASE(PFHE_OP_SUB)
        PFHE_CASE(PFHE_OP_NEG)
        PFHE_CASE(PFHE_OP_MUL_SCALAR)
        PFHE_CASE(PFHE_OP_ADD_MUL_SCALAR)
        PFHE_CASE(PFHE_OP_FACTOR_MUL)
        PFHE_CASE(PFHE_OP_ADD_FACTOR_MUL)
        PFHE_CASE(PFHE_OP_SUB_FACTOR_MUL)
        PFHE_CASE(PFHE_OP_REDUCE_LAZY)
        PFHE_CASE(PFHE_OP_DOUBLE)
        PFHE_CASE(PFHE_OP_MUL_SCALAR_ADD)
        PFHE_CASE(PFHE_OP_FACTOR_MUL_ADD)
#undef PFHE_CASE
    }
    return cudaErrorInvalidValue;
}
template cudaError_t launch_slice_op<uint32_t>(int, const LimbConsts<uint32_t> &, int, const uint32_t *, const uint32_t *, const uint32_t *,
                                               uint32_t *, size_t, size_t, cudaStream_t, size_t);
template cudaError_t launch_slice_op<uint64_t>(int, const LimbConsts<uint64_t> &, int, const uint64_t *, const uint64_t *, const uint64_t *,
                                               uint64_t *, size_t, size_t, cudaStream_t, size_t);

// ---- gadget parameters (host) ---------------------------------------------------------------------
template <typename T> bool make_gadget(T q, uint32_t log_basis, uint32_t levels_in, GadgetParams<T> &g) {
    constexpr int BITS = sizeof(T) * 8;
    if (log_basis == 0 || (int)log_basis >= BITS) return false;
    if (q < 3 || (q & (q - 1)) == 0) return false;  // NTT primes only (power-of-two moduli are the torus side)
    const uint32_t value_bits = (uint32_t)host::bit_length<T>(q);
    if (value_bits < log_basis) return false;
    uint32_t levels = value_bits / log_basis, drop = value_bits - levels * log_basis;
    if (levels_in) {
        if (levels < levels_in) return false;
        levels = levels_in;
        drop = value_bits - levels * log_basis;
    }
    if (levels == 0) return false;
    g.q = q;
    g.log_basis = log_basis;
    g.levels = levels;
    g.drop_bits = drop;
    const T basis = (T)1 << log_basis;
    g.basis_m1 = basis - 1;
    g.q_minus_basis = q - basis;
    g.carry_mask = log_basis == 1 ? (T)2 : (T)(basis | (basis >> 1));
    g.has_init_mask = drop > 0;
    g.init_mask = drop > 0 ? (T)1 << (drop - 1) : 0;
    T value = 0;
    bool have = false;
    if (log_basis == 1) {
        if (drop != 0) {
            for (uint32_t i = 0; i < levels; i++) value = (T)((value << 1) | 1);
            value = (T)((value << 1) | 1);
            value = (T)(value << (drop - 1));
            have = value < q;
        }
    } else {
        for (uint32_t i = 0; i < levels; i++) value = (T)((value << log_basis) | (g.basis_m1 >> 1));
        if (drop > 0) {
            value = (T)((value << 1) | 1);
            value = (T)(value << (drop - 1));
        } else {
            value = (T)(value + 1);
        }
        have = value < q;
    }
    g.has_threshold = have;
    g.threshold = value;
    const T all = value_bits == (uint32_t)BITS ? (T)~(T)0 : (T)(((T)1 << value_bits) - 1);
    g.add = (T)(all - (q - 1));
    g.half = log_basis == 1 ? (T)0 : (T)(basis >> 1);
    g.offset = drop > 0 ? (T)((T)1 << (drop - 1)) : (T)0;
    for (uint32_t l = 0; l < levels; l++) g.offset = (T)(g.offset + (T)(g.half << (drop + l * log_basis)));
    return true;
}
template bool make_gadget<uint32_t>(uint32_t, uint32_t, uint32_t, GadgetParams<uint32_t> &);
template bool make_gadget<uint64_t>(uint64_t, uint32_t, uint32_t, GadgetParams<uint64_t> &);

// digits[l][i], LSB level first; each digit canonical mod q
template <typename T>
__global__ void __launch_bounds__(256) decompose_kernel(const __grid_constant__ GadgetParams<T> g, const T *__restrict__ values,
                                                        T *__restrict__ digits, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T v = values[i];
        if (g.has_threshold && v >= g.threshold) v += g.add;
        uint32_t carry = g.has_init_mask ? (uint32_t)((v & g.init_mask) != 0) : 0u;
        for (uint32_t l = 0; l < g.levels; l++) {
            T t = ((v >> (g.drop_bits + l * g.log_basis)) & g.basis_m1) + carry;
            carry = (t & g.carry_mask) != 0;
            T d = carry ? (t > g.basis_m1 ? T(0) : t + g.q_minus_basis) : t;
            digits[(size_t)l * count + i] = d;
        }
    }
}
// 16 bytes per access (the kernel writes `levels` words per word read: it is bound by its stores), streaming hints on both sides
template <typename T>
__global__ void __launch_bounds__(256) decompose_vec_kernel(const __grid_constant__ GadgetParams<T> g, const T *__restrict__ values,
                                                            T *__restrict__ digits, size_t count) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V {
        T v[VEC];
    };
    const size_t cv = count / VEC;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cv; i += (size_t)gridDim.x * blockDim.x) {
        V in;
        *reinterpret_cast<uint4 *>(&in) = ldg_stream(reinterpret_cast<const uint4 *>(values) + i);
        uint32_t carry[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            if (g.has_threshold && in.v[k] >= g.threshold) in.v[k] += g.add;
            carry[k] = g.has_init_mask ? (uint32_t)((in.v[k] & g.init_mask) != 0) : 0u;
        }
        for (uint32_t l = 0; l < g.levels; l++) {
            V o;
            const uint32_t sh = g.drop_bits + l * g.log_basis;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const T t = ((in.v[k] >> sh) & g.basis_m1) + carry[k];
                carry[k] = (t & g.carry_mask) != 0;
                o.v[k] = carry[k] ? (t > g.basis_m1 ? T(0) : t + g.q_minus_basis) : t;
            }
            stg_stream(reinterpret_cast<uint4 *>(digits + (size_t)l * count) + i, *reinterpret_cast<const uint4 *>(&o));
        }
    }
}
template <typename T> cudaError_t launch_decompose(const GadgetParams<T> &g, const T *values, T *digits, size_t count, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    constexpr size_t VEC = 16 / sizeof(T);
    if (count % VEC == 0 && ((reinterpret_cast<uintptr_t>(values) | reinterpret_cast<uintptr_t>(digits)) & 15) == 0)
        decompose_vec_kernel<T><<<stream_grid(count / VEC, 256), 256, 0, stream>>>(g, values, digits, count);
    else
        decompose_kernel<T><<<stream_grid(count, 256), 256, 0, stream>>>(g, values, digits, count);
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_decompose<uint32_t>(const GadgetParams<uint32_t> &, const uint32_t *, uint32_t *, size_t, cudaStream_t);
template cudaError_t launch_decompose<uint64_t>(const GadgetParams<uint64_t> &, const uint64_t *, uint64_t *, size_t, cudaStream_t);

template <typename T> struct LiftConsts {
    T temp[kMaxLimbs];  // q_i - small_modulus
    T half;
    int limbs, unsigned_mode;
};
template <typename T>
__global__ void __launch_bounds__(256) rns_lift_kernel(const __grid_constant__ LiftConsts<T> lc, const T *__restrict__ small,
                                                       T *__restrict__ out, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const T v = small[i];
        for (int l = 0; l < lc.limbs; l++) out[(size_t)l * count + i] = (lc.unsigned_mode || v < lc.half) ? v : lc.temp[l] + v;
    }
}
template <typename T>
cudaError_t launch_rns_lift(const T *moduli_host, int limbs, T small_modulus, const T *small, T *out, size_t count, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    LiftConsts<T> lc;
    lc.limbs = limbs;
    lc.unsigned_mode = small_modulus == 2;
    lc.half = (T)((small_modulus + 1) / 2);
    for (int l = 0; l < limbs; l++) lc.temp[l] = moduli_host[l] - small_modulus;
    rns_lift_kernel<T><<<stream_grid(count, 256), 256, 0, stream>>>(lc, small, out, count);
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_rns_lift<uint32_t>(const uint32_t *, int, uint32_t, const uint32_t *, uint32_t *, size_t, cudaStream_t);
template cudaError_t launch_rns_lift<uint64_t>(const uint64_t *, int, uint64_t, const uint64_t *, uint64_t *, size_t, cudaStream_t);

// lwe = [a_0, -a_{N-1}, ..., -a_1, b_0]
// Rlwe::extract_lwe_with_index / extract_first_few_lwe / extract_lwe (primus_lattice/src/rlwe/coeff.rs:194-288) in one kernel:
// out[i] = a[index - i] for i <= index, -a[N + index - i] for index < i < N, then `count` body coefficients b[index ..].
// (index = 0, count = 1: extract_lwe; count > 1: the MultiMsgLwe of extract_first_few_lwe.)
template <typename T>
__global__ void __launch_bounds__(256) extract_lwe_kernel(T q, const T *__restrict__ rlwe, T *__restrict__ lwe, size_t n, size_t batch,
                                                          size_t index, size_t count) {
    const size_t per = n + count, total = batch * per;
    for (size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t bidx = gid / per, i = gid % per;
        const T *src = rlwe + bidx * 2 * n;
        T v;
        if (i >= n) v = src[n + index + (i - n)];
        else if (i <= index) v = src[index - i];
        else v = mod_neg<T>(src[n + index - i], q);
        lwe[gid] = v;
    }
}
template <typename T>
cudaError_t launch_extract_lwe(T q, const T *rlwe, T *lwe, size_t n, size_t batch, size_t index, size_t count, cudaStream_t stream) {
    if (batch == 0) return cudaSuccess;
    extract_lwe_kernel<T><<<stream_grid(batch * (n + count), 256), 256, 0, stream>>>(q, rlwe, lwe, n, batch, index, count);
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_extract_lwe<uint32_t>(uint32_t, const uint32_t *, uint32_t *, size_t, size_t, size_t, size_t, cudaStream_t);
template cudaError_t launch_extract_lwe<uint64_t>(uint64_t, const uint64_t *, uint64_t *, size_t, size_t, size_t, size_t, cudaStream_t);

// ---- integer-pipe microbenchmark --------------------------------------------------------------------
// 8 independent chains of Harvey butterflies per thread, registers only: measures the achievable
// butterfly rate (the "modmul ops at integer-pipe peak" roofline denominator).
template <typename T> __global__ void __launch_bounds__(256) modmul_bench_kernel(T q, T w, T wq, uint32_t iters, T *sink) {
    T x[8], y[8];
    const T two_q = q * 2;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x[k] = (T)(threadIdx.x * 7 + k + blockIdx.x) % q;
        y[k] = (T)(threadIdx.x * 13 + 5 * k + 1) % q;
    }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) fwd_bfly<T>(x[k], y[k], w, wq, q, two_q);
    }
    T acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= x[k] ^ y[k];
    if (acc == (T)0x1234567) sink[0] = acc;
}

// 8 independent chains of bare Shoup products per thread (1 high + 2 low multiplies, nothing else): the "modmul ops at
// integer-pipe peak" denominator of the lattice-kernel rooflines.
template <typename T> __global__ void __launch_bounds__(256) shoup_bench_kernel(T q, T w, T wq, uint32_t iters, T *sink) {
    T x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = (T)(threadIdx.x * 7 + k + blockIdx.x) % q;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = shoup_lazy<T>(x[k], w, wq, q);
    }
    T acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= x[k];
    if (acc == (T)0x1234567) sink[0] = acc;
}

cudaError_t run_modmul_microbench(int kind, uint32_t blocks, uint32_t iters, float *ms) {
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    void *sink = nullptr;
    if ((e = cudaMalloc(&sink, 64)) != cudaSuccess) return e;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (kind == 0) {
            const uint32_t q = 132120577u, w = 73993u;
            modmul_bench_kernel<uint32_t><<<blocks, 256>>>(q, w, host::shoup_quot<uint32_t>(w, q), iters, (uint32_t *)sink);
        } else if (kind == 2) {
            const uint32_t q = 132120577u, w = 73993u;
            shoup_bench_kernel<uint32_t><<<blocks, 256>>>(q, w, host::shoup_quot<uint32_t>(w, q), iters, (uint32_t *)sink);
        } else if (kind == 3) {
            const uint64_t q = 1152921504606830593ull, w = 459811883340678ull;
            shoup_bench_kernel<uint64_t><<<blocks, 256>>>(q, w, host::shoup_quot<uint64_t>(w, q), iters, (uint64_t *)sink);
        } else {
            const uint64_t q = 1125899906826241ull, w = 46909545429ull;
            modmul_bench_kernel<uint64_t><<<blocks, 256>>>(q, w, host::shoup_quot<uint64_t>(w, q), iters, (uint64_t *)sink);
        }
        count_launch();
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        cudaEventElapsedTime(ms, e0, e1);
    }
    cudaFree(sink);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return e;
}

}  // namespace pfhe
