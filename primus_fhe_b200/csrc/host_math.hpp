// host_math.hpp -- host-side number theory for table construction (product code; independent of oracle/).
// Follows the reference's constructors:
//   minimal primitive root        primus_ntt/src/root.rs:41-125
//   table contents                primus_ntt/src/ntt/prime64/table.rs:308-405, prime32/table.rs:185-257
//   Shoup quotient                primus_factor/src/shoup_factor/mod.rs:35-43
//   Barrett ratio                 primus_modulus/src/barrett/mod.rs:52-60
#pragma once
#include <cstdint>
#include <vector>

namespace pfhe {
namespace host {

template <typename T> struct Wide;
template <> struct Wide<uint32_t> {
    using type = uint64_t;
    static constexpr int BITS = 32;
};
template <> struct Wide<uint64_t> {
    using type = unsigned __int128;
    static constexpr int BITS = 64;
};

template <typename T> inline T mulmod(T a, T b, T q) {
    using W = typename Wide<T>::type;
    return (T)(((W)a * b) % q);
}
template <typename T> inline T powmod(T a, T e, T q) {
    T r = (T)(1 % q);
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
template <typename T> inline T shoup_quot(T w, T q) {
    using W = typename Wide<T>::type;
    return (T)((((W)w) << Wide<T>::BITS) / q);
}
// floor(2^(2*BITS) / q) as (lo, hi) words
template <typename T> inline void barrett_ratio(T q, T &r0, T &r1) {
    using W = typename Wide<T>::type;
    W rem = 1;
    W t = rem << Wide<T>::BITS;
    r1 = (T)(t / q);
    rem = t % q;
    t = rem << Wide<T>::BITS;
    r0 = (T)(t / q);
}
template <typename T> inline int bit_length(T v) {
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

// Smallest primitive 2^log_degree-th root of unity mod q; false when 2^log_degree does not divide q-1
// or no generator is found (composite q).
template <typename T> inline bool min_primitive_root(unsigned log_degree, T q, T &out) {
    if (log_degree == 0 || log_degree >= (unsigned)Wide<T>::BITS || q < 3) return false;
    const T qm1 = q - 1;
    const T quotient = qm1 >> log_degree;
    if ((T)(quotient << log_degree) != qm1) return false;
    T w = 0;
    bool found = false;
    for (T r = 2; r < q && r < 5000; r++) {
        w = powmod<T>(r, quotient, q);
        T t = w;
        for (unsigned i = 0; i + 1 < log_degree; i++) t = mulmod(t, t, q);
        if (w != 0 && t == qm1) {
            found = true;
            break;
        }
    }
    if (!found) return false;
    const T sq = mulmod(w, w, q);
    T cur = w, best = w;
    const uint64_t half = 1ull << (log_degree - 1);  // the primitive roots are the `half` odd powers
    for (uint64_t j = 0; j < half; j++) {
        if (cur < best) best = cur;
        cur = mulmod(cur, sq, q);
    }
    out = best;
    return true;
}

inline uint32_t bitrev(uint32_t i, unsigned bits) {
    uint32_t r = 0;
    for (unsigned b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}

template <typename T> struct HostTables {
    unsigned log_n = 0;
    size_t n = 0;
    T q = 0, root = 0, inv_root = 0, inv_n = 0, inv_n_q = 0, inv_n_w = 0, inv_n_w_q = 0;
    std::vector<T> roots, roots_q, inv_roots, inv_roots_q, ordinal;
};

template <typename T> inline void build_tables(unsigned log_n, T q, T root, HostTables<T> &h) {
    const size_t n = (size_t)1 << log_n;
    h.log_n = log_n;
    h.n = n;
    h.q = q;
    h.root = root;
    h.ordinal.resize(2 * n);
    h.ordinal[0] = 1;
    for (size_t k = 1; k < 2 * n; k++) h.ordinal[k] = mulmod(h.ordinal[k - 1], root, q);
    h.inv_root = h.ordinal[2 * n - 1];
    h.roots.assign(n, 0);
    h.inv_roots.assign(n, 0);
    h.roots[0] = 1;
    h.inv_roots[0] = 1;
    for (size_t k = 0; k < n; k++) h.roots[bitrev((uint32_t)k, log_n)] = h.ordinal[k];
    for (size_t k = 0; k + 1 < n; k++) h.inv_roots[bitrev((uint32_t)k, log_n) + 1] = h.ordinal[2 * n - 1 - k];
    h.roots_q.resize(n);
    h.inv_roots_q.resize(n);
    for (size_t k = 0; k < n; k++) {
        h.roots_q[k] = shoup_quot(h.roots[k], q);
        h.inv_roots_q[k] = shoup_quot(h.inv_roots[k], q);
    }
    h.inv_n = powmod<T>((T)(n % q), (T)(q - 2), q);
    h.inv_n_q = shoup_quot(h.inv_n, q);
    h.inv_n_w = mulmod(h.inv_n, h.inv_roots[n - 1], q);
    h.inv_n_w_q = shoup_quot(h.inv_n_w, q);
}

}  // namespace host
}  // namespace pfhe
