// handles.hpp -- shared by the C-ABI translation units (capi.cu, capi_ext.cu): handle layouts, device guard, per-thread
// copy streams and the pipelined host <-> device driver of the host-slice shims.
#pragma once
#include <cstdlib>
#include <string>
#include <vector>

#include "host_math.hpp"
#include "internal.hpp"
#include "pfhe.h"

namespace pfhe {
int lattice_loge(int bits, int log_n);

extern thread_local std::string t_last_cuda_error;
inline pfhe_status cuda_fail(cudaError_t e) {
    t_last_cuda_error = cudaGetErrorString(e);
    cudaGetLastError();  // clear sticky-free errors
    if (e == cudaErrorNotSupported) return PFHE_ERR_UNSUPPORTED;
    return PFHE_ERR_CUDA;
}
#define PFHE_CUDA(expr)                                    \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return pfhe::cuda_fail(_e); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = prev == dev || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (ok && prev >= 0) {
            int cur = -1;
            cudaGetDevice(&cur);
            if (cur != prev) cudaSetDevice(prev);
        }
    }
};
// device that owns a device pointer (entry points without a handle); -1 when it is not a device allocation
inline int device_of(const void *p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

// per-thread copy streams (thread safe by construction) and a PRIVATE stream-ordered pool per device for the staging
// buffers of the host-slice shims (freed staging memory stays in the pool; the process-wide default pool is untouched)
constexpr int kPipe = 4;  // pipeline depth of the host-slice shims (H2D / kernel / D2H + one slack stage)
struct ThreadStreams {
    struct PerDevice {
        std::vector<cudaStream_t> streams;
        cudaMemPool_t pool = nullptr;
    };
    std::vector<PerDevice> per_device;
    cudaError_t get(int device, cudaStream_t *out, cudaMemPool_t *pool = nullptr) {
        if ((int)per_device.size() <= device) per_device.resize(device + 1);
        auto &d = per_device[device];
        if (d.streams.empty()) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            cudaError_t e = cudaMemPoolCreate(&d.pool, &props);
            if (e != cudaSuccess) return e;
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep);
            d.streams.resize(kPipe);
            for (auto &s : d.streams) {
                e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                if (e != cudaSuccess) {
                    d.streams.clear();
                    return e;
                }
            }
        }
        for (int i = 0; i < kPipe; i++) out[i] = d.streams[i];
        if (pool) *pool = d.pool;
        return cudaSuccess;
    }
};
extern thread_local ThreadStreams t_streams;

template <typename T> struct NttHandle {
    int device = 0;
    host::HostTables<T> h;
    DevNtt<T> dev{};      // register-pass layout for the standalone NTT kernels
    DevNtt<T> dev_lat{};  // same table, per-pass layout for the lattice kernels (loge == 0: unsupported size)
    LatHead<T> head{};    // first forward / last inverse twiddles by value (kernel-parameter operands of lattice32.cu)
    void *blob = nullptr;
};

template <typename T> struct DcrtHandle {
    int device = 0;
    std::vector<NttHandle<T> *> limbs;
    DevNtt<T> *d_tables = nullptr;  // device array of the limb tables
    DevNtt<T> *d_tables_lat = nullptr;  // same limbs, lattice-kernel pass layout (nullptr: unsupported degree)
    int lat_policy = 0;                 // field policy of the fused multi-limb external product (internal.hpp)
    DevNtt<T> tb0{};                // limb 0 by value (same field-policy flag as the device array)
};

// Pipelined host <-> device processing of `units` independent work items (`in_bytes`/`out_bytes` each).
// launch(dev_in_chunks[], dev_out, n_units, stream).
template <typename LaunchF>
inline pfhe_status pipelined(int device, const void *const *host_in, int n_in, const size_t *in_bytes, void *host_out, size_t out_bytes,
                             size_t units, LaunchF launch, int out_alias = -1, size_t scratch_bytes = 0) {
    if (units == 0) return PFHE_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    cudaStream_t st[kPipe];
    cudaMemPool_t pool = nullptr;
    PFHE_CUDA(t_streams.get(device, st, &pool));
    size_t per_unit = (out_alias >= 0 ? 0 : out_bytes) + scratch_bytes;  // out_alias: the kernel updates input region #out_alias in place
    for (int i = 0; i < n_in; i++) per_unit += in_bytes[i];  // scratch_bytes: device-only work area per unit, handed to launch as din[n_in]
    // chunk so that each stage moves ~64 MiB (measured best on PCIe Gen5; PFHE_PIPE_CHUNK_MB overrides); at least one unit
    size_t chunk_bytes = (size_t)64 << 20;
    if (const char *e = getenv("PFHE_PIPE_CHUNK_MB")) {
        const long mb = atol(e);
        if (mb > 0 && mb <= 1024) chunk_bytes = (size_t)mb << 20;
    }
    size_t chunk = chunk_bytes / (per_unit ? per_unit : 1);
    if (chunk == 0) chunk = 1;
    if (chunk > units) chunk = units;
    const int nbuf = (int)((units + chunk - 1) / chunk < (size_t)kPipe ? (units + chunk - 1) / chunk : kPipe);
    void *dbuf[kPipe] = {};
    pfhe_status status = PFHE_OK;
    for (int i = 0; i < nbuf; i++) {
        cudaError_t e = cudaMallocFromPoolAsync(&dbuf[i], chunk * per_unit + 6 * 256, pool, st[i]);  // + alignment slack of the regions
        if (e != cudaSuccess) {
            status = cuda_fail(e);
            break;
        }
    }
    if (status == PFHE_OK) {
        size_t done = 0;
        for (int c = 0; done < units; c++) {
            const int b = c % nbuf;
            const size_t nu = units - done < chunk ? units - done : chunk;
            unsigned char *base = static_cast<unsigned char *>(dbuf[b]);
            const void *din[4] = {nullptr, nullptr, nullptr, nullptr};
            size_t off = 0;
            cudaError_t e = cudaSuccess;
            for (int i = 0; i < n_in && e == cudaSuccess; i++) {
                din[i] = base + off;
                e = cudaMemcpyAsync(base + off, static_cast<const unsigned char *>(host_in[i]) + done * in_bytes[i], nu * in_bytes[i],
                                    cudaMemcpyHostToDevice, st[b]);
                off += (chunk * in_bytes[i] + 255) & ~(size_t)255;  // every region starts 256-byte aligned
            }
            void *dout = out_alias >= 0 ? const_cast<void *>(din[out_alias]) : static_cast<void *>(base + off);
            if (scratch_bytes && n_in < 4) din[n_in] = base + off + (out_alias >= 0 ? 0 : ((chunk * out_bytes + 255) & ~(size_t)255));
            if (e == cudaSuccess) e = launch(din, dout, nu, st[b]);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(static_cast<unsigned char *>(host_out) + done * out_bytes, dout, nu * out_bytes, cudaMemcpyDeviceToHost,
                                    st[b]);
            if (e != cudaSuccess) {
                status = cuda_fail(e);
                break;
            }
            done += nu;
        }
    }
    for (int i = 0; i < nbuf; i++) {
        if (dbuf[i]) cudaFreeAsync(dbuf[i], st[i]);
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && status == PFHE_OK) status = cuda_fail(e);
    }
    return status;
}

}  // namespace pfhe

struct pfhe_ntt32 : pfhe::NttHandle<uint32_t> {};
struct pfhe_ntt64 : pfhe::NttHandle<uint64_t> {};
struct pfhe_dcrt32 : pfhe::DcrtHandle<uint32_t> {};
struct pfhe_dcrt64 : pfhe::DcrtHandle<uint64_t> {};

namespace pfhe {
// defined in capi.cu (explicitly instantiated for the u32 / u64 handle types)
template <typename T, typename H> pfhe_status create_handle(int device, uint32_t log_n, T q, H **out, bool generic_only);
template <typename H> void destroy_handle(H *h);
template <typename T> pfhe_status host_transform(const NttHandle<T> *t, T *polys, size_t batch, bool fwd, bool lazy);
template <typename T> pfhe_status host_polymul(const NttHandle<T> *t, const T *a, const T *b, T *c, size_t batch);
template <typename T, typename H>
pfhe_status ext_prod_host(const H *t, uint32_t k, uint32_t log_basis, uint32_t levels_in, const T *key, const T *in, T *out, size_t batch, int to_coeff);
template <typename T, typename D> pfhe_status dcrt_host_transform(const D *t, T *polys, size_t batch, bool fwd, bool lazy);
}  // namespace pfhe
