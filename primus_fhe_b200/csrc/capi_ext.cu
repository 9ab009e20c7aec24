// capi_ext.cu -- second half of the C-ABI (include/pfhe.h, "round 2" sections):
//   * bootstrapping-key handles and the host-slice bootstrap shim (LWE in -> blind rotation -> sample extraction -> LWE out);
//   * multi-device drivers: one host thread + stream set per device, tables / keys replicated, contiguous batch shards,
//     no data-path collective (NttTable is Send + Sync, primus_ntt/src/ntt/mod.rs:16);
//   * named whole-ciphertext transforms (into_ntt_form / write_ntt_form / into_coeff_form / write_coeff_form,
//     primus_lattice/src/macros/mod.rs:537-674, :892-937) and the raw little-endian byte layout (macros/mod.rs:39-97);
//   * LWE modulus switch to Z_2N (not in the reference: convention stated in pfhe.h);
//   * UintNttTable<T> handles (primus_ntt/src/ntt/primitive.rs:37-396) with their own constructor rules.
// No CPU compute path: every transform / product below is a kernel launch.
#include <cstring>
#include <new>
#include <thread>

#include "handles.hpp"

namespace pfhe {

constexpr int kSMsExt = 148;
static unsigned ext_grid(size_t work, int threads) {
    size_t blocks = (work + threads - 1) / threads;
    const size_t cap = (size_t)kSMsExt * 8;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

// ---- LWE modulus switch q -> 2N:  a' = floor((a * 2N + floor(q/2)) / q) mod 2N  (exact: shift-subtract long division) -------
template <typename T>
__global__ void __launch_bounds__(256) modswitch_kernel(T q, uint32_t log_2n, const T *__restrict__ in, uint32_t *__restrict__ out, size_t count) {
    const T half_up = (q >> 1) + 1;  // ceil(q/2) for odd q
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T r = in[i];  // canonical, < q < 2^(BITS-2): 2r never wraps
        uint32_t quo = 0;
        for (uint32_t b = 0; b < log_2n; b++) {
            r <<= 1;
            const bool ge = r >= q;
            r -= ge ? q : (T)0;
            quo = (quo << 1) | (uint32_t)ge;
        }
        quo += (r >= half_up) ? 1u : 0u;
        out[i] = quo & ((1u << log_2n) - 1u);
    }
}
template <typename T> static cudaError_t launch_modswitch(T q, uint32_t log_2n, const T *in, uint32_t *out, size_t count, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    modswitch_kernel<T><<<ext_grid(count, 256), 256, 0, s>>>(q, log_2n, in, out, count);
    count_launch();
    return cudaGetLastError();
}

// ---- u16 <-> u32 widening for UintNttTable<u16> --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) widen16_kernel(const uint16_t *__restrict__ in, uint32_t *__restrict__ out, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void __launch_bounds__(256) narrow16_kernel(const uint32_t *__restrict__ in, uint16_t *__restrict__ out, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) out[i] = (uint16_t)in[i];
}

// ---- bootstrapping key --------------------------------------------------------------------------------------------------------
template <typename T> struct BskHandle {
    int device = 0;
    uint32_t n_lwe = 0, log_basis = 0, levels_in = 0, levels = 0, log_n = 0;
    T q = 0;
    T *dev = nullptr;  // [n_lwe][2][levels][2][N], NTT domain (NttRgsw layout, primus_lattice/src/ggsw/dcrt.rs:14-31 with L = 1)
};

template <typename T, typename H, typename B>
static pfhe_status bsk_create(const H *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe, const T *bsk_host, size_t words, B **out) {
    if (!out) return PFHE_ERR_INVALID_ARG;
    *out = nullptr;
    if (!t || !bsk_host || n_lwe == 0) return PFHE_ERR_INVALID_ARG;
    GadgetParams<T> g;
    if (!make_gadget<T>(t->h.q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;
    const size_t need = (size_t)n_lwe * 2 * g.levels * 2 * t->h.n;
    if (words != need) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    auto *b = new (std::nothrow) B();
    if (!b) return PFHE_ERR_INVALID_ARG;
    b->device = t->device;
    b->n_lwe = n_lwe;
    b->log_basis = log_basis;
    b->levels_in = levels_in;
    b->levels = g.levels;
    b->log_n = t->h.log_n;
    b->q = t->h.q;
    cudaError_t e = cudaMalloc(&b->dev, need * sizeof(T));
    if (e == cudaSuccess) e = cudaMemcpy(b->dev, bsk_host, need * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (b->dev) cudaFree(b->dev);
        delete b;
        return cuda_fail(e);
    }
    *out = b;
    return PFHE_OK;
}
template <typename B> static void bsk_destroy(B *b) {
    if (!b) return;
    DeviceGuard guard(b->device);
    if (b->dev) cudaFree(b->dev);
    delete b;
}

// LWE (mod 2N) in -> blind rotation -> [extract_lwe] -> out, all through the pipelined host <-> device path.
template <typename T, typename H, typename B>
static pfhe_status bootstrap_slices(const H *t, const B *bsk, const uint32_t *lwe, const T *tv, T *out, size_t batch, int extract) {
    if (!t || !bsk || ((!lwe || !tv || !out) && batch)) return PFHE_ERR_INVALID_ARG;
    if (bsk->device != t->device || bsk->q != t->h.q || bsk->log_n != t->h.log_n) return PFHE_ERR_INVALID_ARG;
    if (t->dev_lat.loge == 0) return PFHE_ERR_UNSUPPORTED;
    if (batch == 0) return PFHE_OK;
    GadgetParams<T> g;
    if (!make_gadget<T>(t->h.q, bsk->log_basis, bsk->levels_in, g)) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    const size_t n = t->h.n;
    T *dtv = nullptr;
    PFHE_CUDA(cudaMalloc(&dtv, n * sizeof(T)));
    cudaError_t e = cudaMemcpy(dtv, tv, n * sizeof(T), cudaMemcpyHostToDevice);
    pfhe_status status = e == cudaSuccess ? PFHE_OK : cuda_fail(e);
    if (status == PFHE_OK) {
        const void *ins[1] = {lwe};
        const size_t inb[1] = {(size_t)(bsk->n_lwe + 1) * sizeof(uint32_t)};
        const size_t acc_bytes = 2 * n * sizeof(T), out_bytes = extract ? (n + 1) * sizeof(T) : acc_bytes;
        status = pipelined(
            t->device, ins, 1, inb, out, out_bytes, batch,
            [&](const void *const *din, void *dout, size_t nu, cudaStream_t s) -> cudaError_t {
                T *acc = extract ? const_cast<T *>(static_cast<const T *>(din[1])) : static_cast<T *>(dout);
                cudaError_t err = cudaErrorNotSupported;
                if constexpr (sizeof(T) == 4)
                    err = launch_blind_rotate_fast32(t->dev_lat, t->head, g, bsk->dev, bsk->n_lwe, static_cast<const uint32_t *>(din[0]), dtv, acc, nu, s);
                if (err == cudaErrorNotSupported)
                    err = launch_blind_rotate<T>(t->dev_lat, g, bsk->dev, bsk->n_lwe, static_cast<const uint32_t *>(din[0]), dtv, acc, nu, s);
                if (err != cudaSuccess || !extract) return err;
                return launch_extract_lwe<T>(t->h.q, acc, static_cast<T *>(dout), n, nu, 0, 1, s);
            },
            -1, extract ? acc_bytes : 0);
    }
    cudaFree(dtv);
    return status;
}

// ---- multi-device drivers --------------------------------------------------------------------------------------------------------
// contiguous shard of `total` units owned by part `idx` of `parts` (sizes differ by at most one; same rule as shard.py)
static void shard_range(size_t total, size_t parts, size_t idx, size_t &begin, size_t &end) {
    const size_t base = total / parts, rem = total % parts;
    begin = idx * base + (idx < rem ? idx : rem);
    end = begin + base + (idx < rem ? 1 : 0);
}
// run fn(part, begin, end) on one host thread per device; first non-OK status wins
template <typename F> static pfhe_status run_sharded(size_t parts, size_t total, F fn) {
    if (parts == 0) return PFHE_ERR_INVALID_ARG;
    std::vector<pfhe_status> st(parts, PFHE_OK);
    std::vector<std::string> err(parts);
    std::vector<std::thread> th;
    th.reserve(parts);
    for (size_t p = 0; p < parts; p++) {
        size_t b, e;
        shard_range(total, parts, p, b, e);
        th.emplace_back([&, p, b, e] {
            st[p] = b < e ? fn(p, b, e) : PFHE_OK;
            if (st[p] != PFHE_OK) err[p] = t_last_cuda_error;
        });
    }
    for (auto &t : th) t.join();
    for (size_t p = 0; p < parts; p++)
        if (st[p] != PFHE_OK) {
            t_last_cuda_error = err[p];
            return st[p];
        }
    return PFHE_OK;
}
template <typename H> static bool same_tables(const H *const *tables, size_t n) {
    if (!tables || n == 0) return false;
    for (size_t i = 0; i < n; i++)
        if (!tables[i] || tables[i]->h.q != tables[0]->h.q || tables[i]->h.log_n != tables[0]->h.log_n) return false;
    return true;
}

// ---- UintNttTable<T> -------------------------------------------------------------------------------------------------------------
template <typename T> struct UintHandle {
    NttHandle<T> *inner = nullptr;  // same tables, generic radix-2 kernel only
};
// UintNttTable::new (primitive.rs:114-181): NoPrimitiveRoot, DegreeConversionErr (n does not fit the word), DegreeTooLarge (n >= q);
// the lazy butterflies additionally need 4q to fit the word (values live in [0,4q), primitive.rs:219-236).
template <typename T, typename W, typename H, typename U>
static pfhe_status uint_create(int device, uint32_t log_n, T q, U **out) {
    constexpr int BITS = sizeof(T) * 8;
    if (!out) return PFHE_ERR_INVALID_ARG;
    *out = nullptr;
    if (log_n == 0) return PFHE_ERR_DEGREE_TOO_LARGE;
    W root;  // same order as the reference: root first (primitive.rs:121), then the degree checks (:160-168)
    if (log_n + 1 >= (uint32_t)BITS || !host::min_primitive_root<W>(log_n + 1, (W)q, root)) return PFHE_ERR_NO_PRIMITIVE_ROOT;
    if (log_n >= (uint32_t)BITS) return PFHE_ERR_DEGREE_CONVERSION;  // T::try_from(n) fails
    if ((((uint64_t)1) << log_n) >= (uint64_t)q) return PFHE_ERR_DEGREE_TOO_LARGE;
    if ((q >> (BITS - 2)) != 0) return PFHE_ERR_MODULUS_TOO_LARGE;
    H *inner = nullptr;
    const pfhe_status s = create_handle<W, H>(device, log_n, (W)q, &inner, true);
    if (s != PFHE_OK) return s;
    auto *u = new (std::nothrow) U();
    if (!u) {
        destroy_handle(inner);
        return PFHE_ERR_NTT_TABLE;
    }
    u->inner = inner;
    *out = u;
    return PFHE_OK;
}
// u16 words: widened to the u32 kernels on the device
static pfhe_status uint16_transform(const NttHandle<uint32_t> *t, uint16_t *polys, size_t batch, bool fwd) {
    if (!t || (!polys && batch)) return PFHE_ERR_INVALID_ARG;
    const size_t n = t->h.n;
    const void *ins[1] = {polys};
    const size_t inb[1] = {n * sizeof(uint16_t)};
    return pipelined(
        t->device, ins, 1, inb, polys, n * sizeof(uint16_t), batch,
        [&](const void *const *din, void *dout, size_t nu, cudaStream_t s) -> cudaError_t {
            uint32_t *wide = const_cast<uint32_t *>(static_cast<const uint32_t *>(din[1]));
            widen16_kernel<<<ext_grid(nu * n, 256), 256, 0, s>>>(static_cast<const uint16_t *>(din[0]), wide, nu * n);
            count_launch();
            cudaError_t e = launch_ntt<uint32_t>(t->dev, nullptr, 1, wide, wide, nu, fwd, s);
            if (e != cudaSuccess) return e;
            narrow16_kernel<<<ext_grid(nu * n, 256), 256, 0, s>>>(wide, static_cast<uint16_t *>(dout), nu * n);
            count_launch();
            return cudaGetLastError();
        },
        0, n * sizeof(uint32_t));
}

}  // namespace pfhe

using namespace pfhe;

struct pfhe_bsk32 : BskHandle<uint32_t> {};
struct pfhe_bsk64 : BskHandle<uint64_t> {};
struct pfhe_uintntt16 : UintHandle<uint32_t> {};
struct pfhe_uintntt32 : UintHandle<uint32_t> {};
struct pfhe_uintntt64 : UintHandle<uint64_t> {};

extern "C" {

#define PFHE_DEFINE_EXT(B, T)                                                                                                          \
    pfhe_status pfhe_bsk##B##_create(const pfhe_ntt##B *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe, const T *bsk,       \
                                     pfhe_bsk##B **out) {                                                                              \
        GadgetParams<T> g;                                                                                                             \
        if (!t || !make_gadget<T>(t->h.q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;                                       \
        return bsk_create<T>(t, log_basis, levels_in, n_lwe, bsk, (size_t)n_lwe * 2 * g.levels * 2 * t->h.n, out);                     \
    }                                                                                                                                  \
    pfhe_status pfhe_bsk##B##_create_from_bytes(const pfhe_ntt##B *t, uint32_t log_basis, uint32_t levels_in, uint32_t n_lwe,          \
                                                const uint8_t *bytes, size_t byte_count, pfhe_bsk##B **out) {                          \
        if (!bytes || byte_count % sizeof(T)) return PFHE_ERR_INVALID_ARG;                                                             \
        if (reinterpret_cast<uintptr_t>(bytes) % alignof(T) == 0)                                                                      \
            return bsk_create<T>(t, log_basis, levels_in, n_lwe, reinterpret_cast<const T *>(bytes), byte_count / sizeof(T), out);     \
        std::vector<T> tmp(byte_count / sizeof(T));                                                                                    \
        memcpy(tmp.data(), bytes, byte_count);                                                                                         \
        return bsk_create<T>(t, log_basis, levels_in, n_lwe, tmp.data(), tmp.size(), out);                                             \
    }                                                                                                                                  \
    void pfhe_bsk##B##_destroy(pfhe_bsk##B *b) { bsk_destroy(b); }                                                                     \
    uint32_t pfhe_bsk##B##_lwe_dimension(const pfhe_bsk##B *b) { return b ? b->n_lwe : 0; }                                            \
    uint32_t pfhe_bsk##B##_levels(const pfhe_bsk##B *b) { return b ? b->levels : 0; }                                                  \
    const T *pfhe_bsk##B##_device_ptr(const pfhe_bsk##B *b) { return b ? b->dev : nullptr; }                                           \
    pfhe_status pfhe_bootstrap##B##_slices(const pfhe_ntt##B *t, const pfhe_bsk##B *bsk, const uint32_t *lwe, const T *test_vector,    \
                                           T *out, size_t batch, int extract) {                                                        \
        return bootstrap_slices<T>(t, bsk, lwe, test_vector, out, batch, extract);                                                     \
    }                                                                                                                                  \
    pfhe_status pfhe_lwe##B##_modulus_switch_batch(T q, uint32_t log_2n, const T *lwe, uint32_t *out, size_t count, void *stream) {    \
        if (q < 3 || (q & 1) == 0 || (q >> (sizeof(T) * 8 - 2)) != 0 || log_2n == 0 || log_2n > 24) return PFHE_ERR_INVALID_ARG;       \
        if (count == 0) return PFHE_OK;                                                                                                \
        if (!lwe || !out) return PFHE_ERR_INVALID_ARG;                                                                                 \
        const int dev = device_of(lwe);                                                                                                \
        if (dev < 0) return PFHE_ERR_INVALID_ARG;                                                                                      \
        DeviceGuard guard(dev);                                                                                                        \
        if (!guard.ok) return PFHE_ERR_CUDA;                                                                                           \
        PFHE_CUDA(launch_modswitch<T>(q, log_2n, lwe, out, count, static_cast<cudaStream_t>(stream)));                                 \
        return PFHE_OK;                                                                                                                \
    }                                                                                                                                  \
    /* multi-device drivers */                                                                                                         \
    pfhe_status pfhe_multi_ntt##B##_create(const int *devices, size_t n_devices, uint32_t log_n, T q, pfhe_ntt##B **out_tables) {      \
        if (!devices || !out_tables || n_devices == 0) return PFHE_ERR_INVALID_ARG;                                                    \
        for (size_t i = 0; i < n_devices; i++) out_tables[i] = nullptr;                                                                \
        for (size_t i = 0; i < n_devices; i++) {                                                                                       \
            const pfhe_status s = pfhe_ntt##B##_create(devices[i], log_n, q, &out_tables[i]);                                          \
            if (s != PFHE_OK) {                                                                                                        \
                for (size_t j = 0; j < i; j++) {                                                                                       \
                    pfhe_ntt##B##_destroy(out_tables[j]);                                                                              \
                    out_tables[j] = nullptr;                                                                                           \
                }                                                                                                                      \
                return s;                                                                                                              \
            }                                                                                                                          \
        }                                                                                                                              \
        return PFHE_OK;                                                                                                                \
    }                                                                                                                                  \
    pfhe_status pfhe_multi_ntt##B##_transform_slices(const pfhe_ntt##B *const *tables, size_t n_devices, T *polys, size_t batch,       \
                                                     int inverse, int lazy) {                                                          \
        if (!same_tables(tables, n_devices) || (!polys && batch)) return PFHE_ERR_INVALID_ARG;                                         \
        const size_t n = tables[0]->h.n;                                                                                               \
        return run_sharded(n_devices, batch, [&](size_t p, size_t b, size_t e) {                                                       \
            return host_transform<T>(tables[p], polys + b * n, e - b, inverse == 0, lazy != 0);                                        \
        });                                                                                                                            \
    }                                                                                                                                  \
    pfhe_status pfhe_multi_ntt##B##_polymul_slices(const pfhe_ntt##B *const *tables, size_t n_devices, const T *a, const T *b_,        \
                                                   T *c, size_t batch) {                                                               \
        if (!same_tables(tables, n_devices) || ((!a || !b_ || !c) && batch)) return PFHE_ERR_INVALID_ARG;                              \
        const size_t n = tables[0]->h.n;                                                                                               \
        return run_sharded(n_devices, batch, [&](size_t p, size_t b, size_t e) {                                                       \
            return host_polymul<T>(tables[p], a + b * n, b_ + b * n, c + b * n, e - b);                                                \
        });                                                                                                                            \
    }                                                                                                                                  \
    pfhe_status pfhe_multi_ggsw##B##_external_product_slices(const pfhe_ntt##B *const *tables, size_t n_devices, uint32_t k,           \
                                                             uint32_t log_basis, uint32_t levels_in, const T *key, const T *in,        \
                                                             T *out, size_t batch, int to_coeff) {                                     \
        if (!same_tables(tables, n_devices) || ((!key || !in || !out) && batch)) return PFHE_ERR_INVALID_ARG;                          \
        const size_t len = (size_t)(k + 1) * tables[0]->h.n;                                                                           \
        return run_sharded(n_devices, batch, [&](size_t p, size_t b, size_t e) {                                                       \
            return ext_prod_host<T, pfhe_ntt##B>(tables[p], k, log_basis, levels_in, key, in + b * len, out + b * len, e - b,          \
                                                 to_coeff);                                                                            \
        });                                                                                                                            \
    }                                                                                                                                  \
    pfhe_status pfhe_multi_bootstrap##B##_slices(const pfhe_ntt##B *const *tables, const pfhe_bsk##B *const *bsks, size_t n_devices,   \
                                                 const uint32_t *lwe, const T *test_vector, T *out, size_t batch, int extract) {       \
        if (!same_tables(tables, n_devices) || !bsks || ((!lwe || !test_vector || !out) && batch)) return PFHE_ERR_INVALID_ARG;        \
        for (size_t i = 0; i < n_devices; i++)                                                                                         \
            if (!bsks[i] || bsks[i]->n_lwe != bsks[0]->n_lwe) return PFHE_ERR_INVALID_ARG;                                             \
        const size_t n = tables[0]->h.n, in_len = (size_t)bsks[0]->n_lwe + 1, out_len = extract ? n + 1 : 2 * n;                       \
        return run_sharded(n_devices, batch, [&](size_t p, size_t b, size_t e) {                                                       \
            return bootstrap_slices<T>(tables[p], bsks[p], lwe + b * in_len, test_vector, out + b * out_len, e - b, extract);          \
        });                                                                                                                            \
    }                                                                                                                                  \
    /* whole-ciphertext transforms: every polynomial of the flat storage (macros/mod.rs:537-674) */                                    \
    pfhe_status pfhe_cipher##B##_into_ntt_form(const pfhe_ntt##B *t, T *data, size_t words) {                                          \
        if (!t || t->h.n == 0 || words % t->h.n) return PFHE_ERR_INVALID_ARG;                                                          \
        return host_transform<T>(t, data, words / t->h.n, true, false);                                                                \
    }                                                                                                                                  \
    pfhe_status pfhe_cipher##B##_into_coeff_form(const pfhe_ntt##B *t, T *data, size_t words) {                                        \
        if (!t || t->h.n == 0 || words % t->h.n) return PFHE_ERR_INVALID_ARG;                                                          \
        return host_transform<T>(t, data, words / t->h.n, false, false);                                                               \
    }                                                                                                                                  \
    pfhe_status pfhe_cipher##B##_write_ntt_form(const pfhe_ntt##B *t, const T *src, T *dst, size_t words) {                            \
        if (!t || t->h.n == 0 || words % t->h.n || ((!src || !dst) && words)) return PFHE_ERR_INVALID_ARG;                             \
        if (src != dst) memmove(dst, src, words * sizeof(T)); /* result.copy_from_slice(self), macros/mod.rs:568 */                    \
        return host_transform<T>(t, dst, words / t->h.n, true, false);                                                                 \
    }                                                                                                                                  \
    pfhe_status pfhe_cipher##B##_write_coeff_form(const pfhe_ntt##B *t, const T *src, T *dst, size_t words) {                          \
        if (!t || t->h.n == 0 || words % t->h.n || ((!src || !dst) && words)) return PFHE_ERR_INVALID_ARG;                             \
        if (src != dst) memmove(dst, src, words * sizeof(T));                                                                          \
        return host_transform<T>(t, dst, words / t->h.n, false, false);                                                                \
    }                                                                                                                                  \
    pfhe_status pfhe_dcrt_cipher##B##_into_ntt_form(const pfhe_dcrt##B *t, T *data, size_t words) {                                    \
        const size_t unit = t ? t->limbs.size() * t->limbs[0]->h.n : 0;                                                                \
        if (!unit || words % unit) return PFHE_ERR_INVALID_ARG;                                                                        \
        return dcrt_host_transform<T>(t, data, words / unit, true, false);                                                             \
    }                                                                                                                                  \
    pfhe_status pfhe_dcrt_cipher##B##_into_coeff_form(const pfhe_dcrt##B *t, T *data, size_t words) {                                  \
        const size_t unit = t ? t->limbs.size() * t->limbs[0]->h.n : 0;                                                                \
        if (!unit || words % unit) return PFHE_ERR_INVALID_ARG;                                                                        \
        return dcrt_host_transform<T>(t, data, words / unit, false, false);                                                            \
    }                                                                                                                                  \
    /* named shapes: word counts of the reference's flat containers */                                                                 \
    size_t pfhe_rlwe##B##_words(const pfhe_ntt##B *t) { return t ? 2 * t->h.n : 0; }                                                   \
    size_t pfhe_rlev##B##_words(const pfhe_ntt##B *t, uint32_t levels) { return t ? (size_t)levels * 2 * t->h.n : 0; }                 \
    size_t pfhe_rgsw##B##_words(const pfhe_ntt##B *t, uint32_t levels) { return t ? (size_t)2 * levels * 2 * t->h.n : 0; }             \
    size_t pfhe_glwe##B##_words(const pfhe_ntt##B *t, uint32_t k) { return t ? (size_t)(k + 1) * t->h.n : 0; }                         \
    size_t pfhe_glev##B##_words(const pfhe_ntt##B *t, uint32_t k, uint32_t levels) { return t ? (size_t)levels * (k + 1) * t->h.n : 0; } \
    size_t pfhe_ggsw##B##_words(const pfhe_ntt##B *t, uint32_t k, uint32_t levels) {                                                   \
        return t ? (size_t)(k + 1) * levels * (k + 1) * t->h.n : 0;                                                                    \
    }                                                                                                                                  \
    pfhe_status pfhe_rlwe##B##_into_ntt_form(const pfhe_ntt##B *t, T *d) { return pfhe_cipher##B##_into_ntt_form(t, d, pfhe_rlwe##B##_words(t)); }          \
    pfhe_status pfhe_rlwe##B##_into_coeff_form(const pfhe_ntt##B *t, T *d) { return pfhe_cipher##B##_into_coeff_form(t, d, pfhe_rlwe##B##_words(t)); }      \
    pfhe_status pfhe_rlev##B##_into_ntt_form(const pfhe_ntt##B *t, T *d, uint32_t l) { return pfhe_cipher##B##_into_ntt_form(t, d, pfhe_rlev##B##_words(t, l)); }     \
    pfhe_status pfhe_rlev##B##_into_coeff_form(const pfhe_ntt##B *t, T *d, uint32_t l) { return pfhe_cipher##B##_into_coeff_form(t, d, pfhe_rlev##B##_words(t, l)); } \
    pfhe_status pfhe_rgsw##B##_into_ntt_form(const pfhe_ntt##B *t, T *d, uint32_t l) { return pfhe_cipher##B##_into_ntt_form(t, d, pfhe_rgsw##B##_words(t, l)); }     \
    pfhe_status pfhe_rgsw##B##_into_coeff_form(const pfhe_ntt##B *t, T *d, uint32_t l) { return pfhe_cipher##B##_into_coeff_form(t, d, pfhe_rgsw##B##_words(t, l)); } \
    pfhe_status pfhe_glwe##B##_into_ntt_form(const pfhe_ntt##B *t, T *d, uint32_t k) { return pfhe_cipher##B##_into_ntt_form(t, d, pfhe_glwe##B##_words(t, k)); }     \
    pfhe_status pfhe_glwe##B##_into_coeff_form(const pfhe_ntt##B *t, T *d, uint32_t k) { return pfhe_cipher##B##_into_coeff_form(t, d, pfhe_glwe##B##_words(t, k)); } \
    pfhe_status pfhe_glev##B##_into_ntt_form(const pfhe_ntt##B *t, T *d, uint32_t k, uint32_t l) {                                     \
        return pfhe_cipher##B##_into_ntt_form(t, d, pfhe_glev##B##_words(t, k, l));                                                    \
    }                                                                                                                                  \
    pfhe_status pfhe_glev##B##_into_coeff_form(const pfhe_ntt##B *t, T *d, uint32_t k, uint32_t l) {                                   \
        return pfhe_cipher##B##_into_coeff_form(t, d, pfhe_glev##B##_words(t, k, l));                                                  \
    }                                                                                                                                  \
    pfhe_status pfhe_ggsw##B##_into_ntt_form(const pfhe_ntt##B *t, T *d, uint32_t k, uint32_t l) {                                     \
        return pfhe_cipher##B##_into_ntt_form(t, d, pfhe_ggsw##B##_words(t, k, l));                                                    \
    }                                                                                                                                  \
    pfhe_status pfhe_ggsw##B##_into_coeff_form(const pfhe_ntt##B *t, T *d, uint32_t k, uint32_t l) {                                   \
        return pfhe_cipher##B##_into_coeff_form(t, d, pfhe_ggsw##B##_words(t, k, l));                                                  \
    }                                                                                                                                  \
    /* from_bytes / read_bytes / to_bytes / write_bytes (macros/mod.rs:39-97): raw little-endian element bytes */                      \
    pfhe_status pfhe_cipher##B##_read_bytes(T *words, size_t word_count, const uint8_t *bytes, size_t byte_count) {                    \
        if (byte_count != word_count * sizeof(T) || ((!words || !bytes) && byte_count)) return PFHE_ERR_INVALID_ARG;                   \
        if (byte_count) memcpy(words, bytes, byte_count); /* x86-64 / aarch64 hosts are little endian, like the reference's targets */ \
        return PFHE_OK;                                                                                                                \
    }                                                                                                                                  \
    pfhe_status pfhe_cipher##B##_write_bytes(const T *words, size_t word_count, uint8_t *bytes, size_t byte_count) {                   \
        if (byte_count != word_count * sizeof(T) || ((!words || !bytes) && byte_count)) return PFHE_ERR_INVALID_ARG;                   \
        if (byte_count) memcpy(bytes, words, byte_count);                                                                              \
        return PFHE_OK;                                                                                                                \
    }                                                                                                                                  \
    size_t pfhe_cipher##B##_byte_count(size_t word_count) { return word_count * sizeof(T); }

PFHE_DEFINE_EXT(32, uint32_t)
PFHE_DEFINE_EXT(64, uint64_t)

pfhe_status pfhe_blind_rotate_ternary32_batch(const pfhe_ntt32 *t, uint32_t log_basis, uint32_t levels_in, const uint32_t *bsk_plus,
                                              const uint32_t *bsk_minus, uint32_t n_lwe, const uint32_t *lwe, const uint32_t *test_vector,
                                              uint32_t *acc_out, size_t batch, void *stream) {
    if (!t || ((!bsk_plus || !bsk_minus || !lwe || !test_vector || !acc_out) && batch)) return PFHE_ERR_INVALID_ARG;
    GadgetParams<uint32_t> g;
    if (!make_gadget<uint32_t>(t->h.q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(launch_blind_rotate_ternary32(t->dev_lat, t->head, g, bsk_plus, bsk_minus, n_lwe, lwe, test_vector, acc_out, batch,
                                            static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}

/* ---- UintNttTable<T> ---- */
#define PFHE_DEFINE_UINT(B, T)                                                                                                   \
    pfhe_status pfhe_uintntt##B##_create(int device, uint32_t log_n, T q, pfhe_uintntt##B **out) {                               \
        return uint_create<T, T, pfhe_ntt##B, pfhe_uintntt##B>(device, log_n, q, out);                                           \
    }                                                                                                                            \
    void pfhe_uintntt##B##_destroy(pfhe_uintntt##B *t) {                                                                         \
        if (!t) return;                                                                                                          \
        destroy_handle(static_cast<pfhe_ntt##B *>(t->inner));                                                                    \
        delete t;                                                                                                                \
    }                                                                                                                            \
    size_t pfhe_uintntt##B##_poly_length(const pfhe_uintntt##B *t) { return t ? t->inner->h.n : 0; }                             \
    T pfhe_uintntt##B##_root(const pfhe_uintntt##B *t) { return t ? t->inner->h.root : 0; }                                      \
    T pfhe_uintntt##B##_inv_root(const pfhe_uintntt##B *t) { return t ? t->inner->h.inv_root : 0; }                              \
    const pfhe_ntt##B *pfhe_uintntt##B##_as_table(const pfhe_uintntt##B *t) {                                                    \
        return t ? static_cast<const pfhe_ntt##B *>(t->inner) : nullptr;                                                         \
    }                                                                                                                            \
    pfhe_status pfhe_uintntt##B##_transform_slices(const pfhe_uintntt##B *t, T *polys, size_t batch, int lazy) {                 \
        return t ? host_transform<T>(t->inner, polys, batch, true, lazy != 0) : PFHE_ERR_INVALID_ARG;                            \
    }                                                                                                                            \
    pfhe_status pfhe_uintntt##B##_inverse_transform_slices(const pfhe_uintntt##B *t, T *polys, size_t batch, int lazy) {         \
        return t ? host_transform<T>(t->inner, polys, batch, false, lazy != 0) : PFHE_ERR_INVALID_ARG;                           \
    }
PFHE_DEFINE_UINT(32, uint32_t)
PFHE_DEFINE_UINT(64, uint64_t)

pfhe_status pfhe_uintntt16_create(int device, uint32_t log_n, uint16_t q, pfhe_uintntt16 **out) {
    if (!out) return PFHE_ERR_INVALID_ARG;
    *out = nullptr;
    if (log_n == 0) return PFHE_ERR_DEGREE_TOO_LARGE;
    uint32_t root;
    if (log_n + 1 >= 16 || !host::min_primitive_root<uint32_t>(log_n + 1, (uint32_t)q, root)) return PFHE_ERR_NO_PRIMITIVE_ROOT;
    if (log_n >= 16) return PFHE_ERR_DEGREE_CONVERSION;  // u16::try_from(n) fails (primitive.rs:160-161)
    if ((1u << log_n) >= (uint32_t)q) return PFHE_ERR_DEGREE_TOO_LARGE;
    if ((q >> 14) != 0) return PFHE_ERR_MODULUS_TOO_LARGE;
    pfhe_ntt32 *inner = nullptr;
    const pfhe_status s = create_handle<uint32_t, pfhe_ntt32>(device, log_n, (uint32_t)q, &inner, true);
    if (s != PFHE_OK) return s;
    auto *u = new (std::nothrow) pfhe_uintntt16();
    if (!u) {
        destroy_handle(inner);
        return PFHE_ERR_NTT_TABLE;
    }
    u->inner = inner;
    *out = u;
    return PFHE_OK;
}
void pfhe_uintntt16_destroy(pfhe_uintntt16 *t) {
    if (!t) return;
    destroy_handle(static_cast<pfhe_ntt32 *>(t->inner));
    delete t;
}
size_t pfhe_uintntt16_poly_length(const pfhe_uintntt16 *t) { return t ? t->inner->h.n : 0; }
uint16_t pfhe_uintntt16_root(const pfhe_uintntt16 *t) { return t ? (uint16_t)t->inner->h.root : 0; }
uint16_t pfhe_uintntt16_inv_root(const pfhe_uintntt16 *t) { return t ? (uint16_t)t->inner->h.inv_root : 0; }
pfhe_status pfhe_uintntt16_transform_slices(const pfhe_uintntt16 *t, uint16_t *polys, size_t batch, int) {
    return t ? uint16_transform(t->inner, polys, batch, true) : PFHE_ERR_INVALID_ARG;
}
pfhe_status pfhe_uintntt16_inverse_transform_slices(const pfhe_uintntt16 *t, uint16_t *polys, size_t batch, int) {
    return t ? uint16_transform(t->inner, polys, batch, false) : PFHE_ERR_INVALID_ARG;
}

// ---- page-locking of caller-owned host buffers --------------------------------------------------------------------------------------
pfhe_status pfhe_host_register(void *host_ptr, size_t bytes) {
    if (!host_ptr || !bytes) return PFHE_ERR_INVALID_ARG;
    const cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) return cuda_fail(e);
    return PFHE_OK;
}
pfhe_status pfhe_host_unregister(void *host_ptr) {
    if (!host_ptr) return PFHE_ERR_INVALID_ARG;
    const cudaError_t e = cudaHostUnregister(host_ptr);
    if (e != cudaSuccess) return cuda_fail(e);
    return PFHE_OK;
}
int pfhe_host_is_pageable(const void *host_ptr) { return host_is_pageable(host_ptr) ? 1 : 0; }

}  // extern "C"
