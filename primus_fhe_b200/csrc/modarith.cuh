// modarith.cuh -- integer-pipe modular arithmetic for sm_100a (u32 and u64 words).
//
// Semantics follow the reference (canonical results are the exact residues):
//   Shoup lazy/canonical product   primus_factor/src/shoup_factor/mod.rs:124-142
//   Harvey butterflies             primus_ntt/src/ntt/prime64/scalar/arithmetic.rs:43-79
//   Barrett reduce of a 2-word value primus_modulus/src/barrett/mod.rs:99-139
//   add/sub/neg (wrapping-min trick) primus_modulus/src/common/compact/primitive.rs:8-58
// No tensor cores: an exact modular product is not a floating-point contraction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pfhe {

template <typename T> struct Word;
template <> struct Word<uint32_t> {
    static constexpr int BITS = 32;
    using Pair = uint2;  // (w, w') twiddle pair, 8 bytes
};
template <> struct Word<uint64_t> {
    static constexpr int BITS = 64;
    using Pair = ulonglong2;  // (w, w') twiddle pair, 16 bytes
};

__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ uint64_t mulhi(uint64_t a, uint64_t b) { return __umul64hi(a, b); }

// x mod m for x < 2m.  u32: the reference's wrapping-min trick (one IADD + one VIMNMX,
// primus_ntt/src/ntt/prime32/scalar/arithmetic.rs:3-6); u64: compare/select (a 64-bit min is no cheaper).
template <typename T> __device__ __forceinline__ T csub(T x, T m) { return x >= m ? x - m : x; }
template <> __device__ __forceinline__ uint32_t csub<uint32_t>(uint32_t x, uint32_t m) { return min(x, x - m); }

// Shoup lazy product: w*y - q*floor(w'*y / 2^BITS)  in [0, 2q), any word-sized y.
template <typename T> __device__ __forceinline__ T shoup_lazy(T y, T w, T wq, T q) {
    T h = mulhi(y, wq);
    return w * y - h * q;
}
// (Measured and dropped in round 2: a u64 quotient estimate from the three upper 32x32 partial products only -- the trick of the
//  reference's AVX-512 DQ back-end, primus_ntt/src/ntt/prime64/avx512/utils/arithmetic.rs:82-121 -- plus one extra conditional
//  subtraction ran the 60-bit N = 4096 transform at 25.5 M/s against 26.9 M/s with the compiler's own __umul64hi.)
template <typename T> __device__ __forceinline__ T shoup(T y, T w, T wq, T q) { return csub(shoup_lazy(y, w, wq, q), q); }
// exact quotient: valid for every q < 2^(BITS-1) (ShoupFactor's own bound, primus_factor/src/shoup_factor/mod.rs:35-43)
template <typename T> __device__ __forceinline__ T shoup_exact(T y, T w, T wq, T q) { return csub<T>((T)(w * y - mulhi(y, wq) * q), q); }

template <typename T> __device__ __forceinline__ T mod_add(T a, T b, T q) { return csub<T>(a + b, q); }
template <typename T> __device__ __forceinline__ T mod_sub(T a, T b, T q) { return a >= b ? a - b : a + q - b; }
template <typename T> __device__ __forceinline__ T mod_neg(T a, T q) { return a == 0 ? T(0) : q - a; }

// Barrett constants for one modulus: ratio = floor(2^(2*BITS) / q) as two words.
template <typename T> struct Barrett {
    T q, r0, r1;
};

// Reduce the double word (hi, lo) < q * 2^BITS ... to [0, q). Mirrors lazy_reduce_wide + one
// conditional subtract (primus_modulus/src/barrett/mod.rs:99-139).
__device__ __forceinline__ uint32_t barrett_reduce_wide(const Barrett<uint32_t> &m, uint32_t lo, uint32_t hi) {
    uint32_t ah = __umulhi(lo, m.r0);
    uint64_t b = (uint64_t)lo * m.r1 + ah;
    uint64_t c = (uint64_t)hi * m.r0;
    uint64_t bc = (b >> 32) + (c >> 32) + (((b & 0xffffffffull) + (c & 0xffffffffull)) >> 32);
    uint32_t q3 = hi * m.r1 + (uint32_t)bc;
    uint32_t r = lo - q3 * m.q;
    return csub(r, m.q);
}
__device__ __forceinline__ uint64_t barrett_reduce_wide(const Barrett<uint64_t> &m, uint64_t lo, uint64_t hi) {
    uint64_t ah = __umul64hi(lo, m.r0);
    uint64_t b0 = lo * m.r1, b1 = __umul64hi(lo, m.r1);
    b0 += ah;
    b1 += (b0 < ah);
    uint64_t c0 = hi * m.r0, c1 = __umul64hi(hi, m.r0);
    uint64_t s = b0 + c0;
    uint64_t bch = b1 + c1 + (s < b0);
    uint64_t q3 = hi * m.r1 + bch;
    uint64_t r = lo - q3 * m.q;
    return csub(r, m.q);
}

__device__ __forceinline__ void mul_wide(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi) {
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}
__device__ __forceinline__ void mul_wide(uint64_t a, uint64_t b, uint64_t &lo, uint64_t &hi) {
    lo = a * b;
    hi = __umul64hi(a, b);
}

// reduce_mul(a,b) = a*b mod q     (primus_modulus/src/barrett/ops.rs:276-283)
template <typename T> __device__ __forceinline__ T barrett_mul(const Barrett<T> &m, T a, T b) {
    T lo, hi;
    mul_wide(a, b, lo, hi);
    return barrett_reduce_wide(m, lo, hi);
}
// reduce_mul_add(a,b,c) = (a*b + c) mod q   (ops.rs:308-315)
template <typename T> __device__ __forceinline__ T barrett_mul_add(const Barrett<T> &m, T a, T b, T c) {
    T lo, hi;
    mul_wide(a, b, lo, hi);
    lo += c;
    hi += (lo < c);
    return barrett_reduce_wide(m, lo, hi);
}

// Harvey forward butterfly: X,Y in [0,4q) -> X',Y' in [0,4q)
template <typename T> __device__ __forceinline__ void fwd_bfly(T &x, T &y, T w, T wq, T q, T two_q) {
    T tx = csub(x, two_q);
    T t = shoup_lazy(y, w, wq, q);
    x = tx + t;
    y = tx + two_q - t;
}
// Harvey inverse (Gentleman-Sande) butterfly: X,Y in [0,2q) -> X',Y' in [0,2q)
template <typename T> __device__ __forceinline__ void inv_bfly(T &x, T &y, T w, T wq, T q, T two_q) {
    T tx = x + y;
    T ty = x + two_q - y;
    x = csub(tx, two_q);
    y = shoup_lazy(ty, w, wq, q);
}

}  // namespace pfhe
