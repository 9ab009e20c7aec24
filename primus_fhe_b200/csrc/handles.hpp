// handles.hpp -- shared by the C-ABI translation units (capi.cu, capi_ext.cu): handle layouts, device guard, per-thread
// copy streams and the pipelined host <-> device driver of the host-slice shims.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
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

// Entry points that take device pointers but no handle run on the device that owns the pointer (cudaPointerGetAttributes), not on
// whatever device happens to be current: a process that drives several GPUs gets INVALID_ARG for a host / foreign pointer instead of
// an illegal-address fault.  Handle-based entry points switch to the handle's device and report PFHE_ERR_CUDA when that fails.
#define PFHE_PTR_GUARD(ptr)                                   \
    const int _pdev = device_of(ptr);                         \
    if (_pdev < 0) return PFHE_ERR_INVALID_ARG;               \
    DeviceGuard _pguard(_pdev);                               \
    if (!_pguard.ok) return PFHE_ERR_CUDA
#define PFHE_DEV_GUARD(dev)       \
    DeviceGuard guard(dev);       \
    if (!guard.ok) return PFHE_ERR_CUDA

// Stream contexts of the host-slice shims: kPipe copy/compute streams, a PRIVATE stream-ordered memory pool for the device staging
// buffers (the process-wide default pool is untouched), pinned bounce buffers for the pageable-memory path and their events.
// Contexts live in a process-wide pool keyed by device and are leased for the duration of one call, so concurrent callers (any host
// thread, including the short-lived worker threads of the multi-device drivers) never share streams and nothing is re-created or
// leaked per call.
constexpr int kPipe = 4;  // pipeline depth of the host-slice shims (H2D / kernel / D2H + one slack stage)
struct StreamCtx {
    int device = -1;
    cudaStream_t streams[kPipe] = {};
    cudaMemPool_t pool = nullptr;
    void *pin[kPipe] = {nullptr, nullptr, nullptr, nullptr};
    size_t pin_bytes[kPipe] = {0, 0, 0, 0};
    cudaEvent_t ev[kPipe] = {nullptr, nullptr, nullptr, nullptr};
    cudaError_t init(int dev) {
        device = dev;
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaError_t e = cudaMemPoolCreate(&pool, &props);
        if (e != cudaSuccess) return e;
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        for (int i = 0; i < kPipe; i++) {
            if ((e = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    // pinned bounce buffer of slot `i` with at least `bytes` (grown on demand, kept with the context)
    cudaError_t staging(int i, size_t bytes, void **out) {
        if (pin_bytes[i] < bytes) {
            if (pin[i]) cudaFreeHost(pin[i]);
            pin[i] = nullptr;
            pin_bytes[i] = 0;
            cudaError_t e = cudaHostAlloc(&pin[i], bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) return e;
            pin_bytes[i] = bytes;
        }
        *out = pin[i];
        return cudaSuccess;
    }
};
struct StreamPool {
    std::mutex mu;
    std::vector<std::vector<StreamCtx *>> idle;  // per device
    cudaError_t acquire(int device, StreamCtx **out) {
        {
            std::lock_guard<std::mutex> lock(mu);
            if ((int)idle.size() <= device) idle.resize(device + 1);
            if (!idle[device].empty()) {
                *out = idle[device].back();
                idle[device].pop_back();
                return cudaSuccess;
            }
        }
        auto *c = new (std::nothrow) StreamCtx();
        if (!c) return cudaErrorMemoryAllocation;
        const cudaError_t e = c->init(device);
        if (e != cudaSuccess) {
            delete c;  // (streams created so far are abandoned: init only fails when the device is unusable)
            return e;
        }
        *out = c;
        return cudaSuccess;
    }
    void release(StreamCtx *c) {
        std::lock_guard<std::mutex> lock(mu);
        idle[c->device].push_back(c);
    }
};
extern StreamPool g_stream_pool;
struct StreamLease {
    StreamCtx *ctx = nullptr;
    cudaError_t err;
    explicit StreamLease(int device) { err = g_stream_pool.acquire(device, &ctx); }
    ~StreamLease() {
        if (ctx) g_stream_pool.release(ctx);
    }
};

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

inline bool host_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (!p) return false;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;  // old runtimes report plain malloc memory as an error
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Staged variant of pipelined() for pageable host memory; same chunk / region layout on the device.
template <typename LaunchF>
inline pfhe_status pipelined_staged(int device, const void *const *host_in, int n_in, const size_t *in_bytes, void *host_out, size_t out_bytes,
                                    size_t units, LaunchF launch, int out_alias, size_t scratch_bytes, size_t chunk, int nbuf, size_t per_unit,
                                    StreamCtx *ctx) {
    cudaStream_t *st = ctx->streams;
    cudaMemPool_t pool = ctx->pool;
    const size_t nchunks = (units + chunk - 1) / chunk;
    int nthreads = 4;
    if (const char *e = getenv("PFHE_STAGE_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 32) nthreads = v;
    }
    // pinned layout per slot: [input 0][input 1]...[output] (the output aliases input `out_alias` when the kernel works in place)
    size_t pin_off[5] = {0, 0, 0, 0, 0}, pin_total = 0;
    for (int i = 0; i < n_in; i++) {
        pin_off[i] = pin_total;
        pin_total += (chunk * in_bytes[i] + 255) & ~(size_t)255;
    }
    const size_t pin_out_off = out_alias >= 0 ? pin_off[out_alias] : pin_total;
    if (out_alias < 0) pin_total += (chunk * out_bytes + 255) & ~(size_t)255;
    void *pin[kPipe] = {};
    cudaEvent_t ev[kPipe] = {};
    void *dbuf[kPipe] = {};
    pfhe_status status = PFHE_OK;
    for (int i = 0; i < nbuf && status == PFHE_OK; i++) {
        cudaError_t e = ctx->staging(i, pin_total, &pin[i]);
        ev[i] = ctx->ev[i];
        if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&dbuf[i], chunk * per_unit + 6 * 256, pool, st[i]);
        if (e != cudaSuccess) status = cuda_fail(e);
    }
    std::unique_ptr<std::atomic<int>[]> in_ready(new std::atomic<int>[nchunks]), out_done(new std::atomic<int>[nchunks]),
        launched(new std::atomic<int>[nchunks]);
    for (size_t c = 0; c < nchunks; c++) {
        in_ready[c].store(0);
        out_done[c].store(0);
        launched[c].store(0);
    }
    std::atomic<int> abort_flag{0};
    auto slice = [&](size_t bytes, int k, size_t &b, size_t &e) {  // 64-byte aligned share k of nthreads
        const size_t per = ((bytes / nthreads) + 63) & ~(size_t)63;
        b = per * k < bytes ? per * k : bytes;
        e = (k == nthreads - 1) ? bytes : (b + per < bytes ? b + per : bytes);
    };
    auto units_of = [&](size_t c) { return c + 1 < nchunks ? chunk : units - c * chunk; };
    std::vector<std::thread> workers;
    if (status == PFHE_OK) {
        for (int k = 0; k < nthreads; k++) {
            workers.emplace_back([&, k] {  // caller memory -> pinned
                for (size_t c = 0; c < nchunks && !abort_flag.load(); c++) {
                    const int b = (int)(c % nbuf);
                    if (c >= (size_t)nbuf)
                        while (out_done[c - nbuf].load(std::memory_order_acquire) < nthreads && !abort_flag.load()) std::this_thread::yield();
                    const size_t nu = units_of(c);
                    for (int i = 0; i < n_in; i++) {
                        size_t sb, se;
                        slice(nu * in_bytes[i], k, sb, se);
                        if (se > sb)
                            memcpy(static_cast<unsigned char *>(pin[b]) + pin_off[i] + sb,
                                   static_cast<const unsigned char *>(host_in[i]) + c * chunk * in_bytes[i] + sb, se - sb);
                    }
                    in_ready[c].fetch_add(1, std::memory_order_release);
                }
            });
            workers.emplace_back([&, k] {  // pinned -> caller memory
                cudaSetDevice(device);
                for (size_t c = 0; c < nchunks && !abort_flag.load(); c++) {
                    const int b = (int)(c % nbuf);
                    while (!launched[c].load(std::memory_order_acquire) && !abort_flag.load()) std::this_thread::yield();
                    if (abort_flag.load()) break;
                    if (cudaEventSynchronize(ev[b]) != cudaSuccess) {
                        abort_flag.store(1);
                        break;
                    }
                    size_t sb, se;
                    slice(units_of(c) * out_bytes, k, sb, se);
                    if (se > sb)
                        memcpy(static_cast<unsigned char *>(host_out) + c * chunk * out_bytes + sb,
                               static_cast<const unsigned char *>(pin[b]) + pin_out_off + sb, se - sb);
                    out_done[c].fetch_add(1, std::memory_order_release);
                }
            });
        }
        for (size_t c = 0; c < nchunks; c++) {
            const int b = (int)(c % nbuf);
            const size_t nu = units_of(c);
            while (in_ready[c].load(std::memory_order_acquire) < nthreads && !abort_flag.load()) std::this_thread::yield();
            if (abort_flag.load()) break;
            unsigned char *base = static_cast<unsigned char *>(dbuf[b]);
            const void *din[4] = {nullptr, nullptr, nullptr, nullptr};
            size_t off = 0;
            cudaError_t e = cudaSuccess;
            for (int i = 0; i < n_in && e == cudaSuccess; i++) {
                din[i] = base + off;
                e = cudaMemcpyAsync(base + off, static_cast<const unsigned char *>(pin[b]) + pin_off[i], nu * in_bytes[i], cudaMemcpyHostToDevice, st[b]);
                off += (chunk * in_bytes[i] + 255) & ~(size_t)255;
            }
            void *dout = out_alias >= 0 ? const_cast<void *>(din[out_alias]) : static_cast<void *>(base + off);
            if (scratch_bytes && n_in < 4) din[n_in] = base + off + (out_alias >= 0 ? 0 : ((chunk * out_bytes + 255) & ~(size_t)255));
            if (e == cudaSuccess) e = launch(din, dout, nu, st[b]);
            if (e == cudaSuccess) e = cudaMemcpyAsync(static_cast<unsigned char *>(pin[b]) + pin_out_off, dout, nu * out_bytes, cudaMemcpyDeviceToHost, st[b]);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], st[b]);
            if (e != cudaSuccess) {
                status = cuda_fail(e);
                abort_flag.store(1);
                break;
            }
            launched[c].store(1, std::memory_order_release);
        }
        for (auto &w : workers) w.join();
        if (abort_flag.load() && status == PFHE_OK) status = PFHE_ERR_CUDA;
    }
    for (int i = 0; i < nbuf; i++) {
        if (dbuf[i]) cudaFreeAsync(dbuf[i], st[i]);
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && status == PFHE_OK) status = cuda_fail(e);
    }
    return status;
}

// Pipelined host <-> device processing of `units` independent work items (`in_bytes`/`out_bytes` each).
// launch(dev_in_chunks[], dev_out, n_units, stream).
template <typename LaunchF>
inline pfhe_status pipelined(int device, const void *const *host_in, int n_in, const size_t *in_bytes, void *host_out, size_t out_bytes,
                             size_t units, LaunchF launch, int out_alias = -1, size_t scratch_bytes = 0) {
    if (units == 0) return PFHE_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    StreamLease lease(device);
    PFHE_CUDA(lease.err);
    cudaStream_t *st = lease.ctx->streams;
    cudaMemPool_t pool = lease.ctx->pool;
    size_t per_unit = (out_alias >= 0 ? 0 : out_bytes) + scratch_bytes;  // out_alias: the kernel updates input region #out_alias in place
    for (int i = 0; i < n_in; i++) per_unit += in_bytes[i];  // scratch_bytes: device-only work area per unit, handed to launch as din[n_in]
    // chunk so that each stage moves ~64 MiB (measured best on PCIe Gen5; PFHE_PIPE_CHUNK_MB overrides); at least one unit
    size_t chunk_bytes = (size_t)64 << 20;
    if (const char *e = getenv("PFHE_PIPE_CHUNK_MB")) {
        const long mb = atol(e);
        if (mb > 0 && mb <= 1024) chunk_bytes = (size_t)mb << 20;
    }
    // Pageable host memory (what a Rust `&mut [T]` / Vec is): cudaMemcpyAsync would stage it synchronously through the driver's own
    // bounce buffer (measured 14 GB/s per direction against 46 GB/s pinned).  Large pageable calls are staged here instead: copy
    // threads move slices between the caller's memory and pinned bounce buffers while the DMA engines and the kernels work on other
    // chunks.  PFHE_STAGE=0 disables it, PFHE_STAGE_THREADS sets the copy threads per direction (default 4).
    size_t host_total = out_bytes * units;
    for (int i = 0; i < n_in; i++) host_total += in_bytes[i] * units;
    bool staged = false;
    if (host_total >= ((size_t)8 << 20)) {
        static const bool stage_on = !(getenv("PFHE_STAGE") && getenv("PFHE_STAGE")[0] == '0');
        bool pageable = host_is_pageable(host_out);
        for (int i = 0; i < n_in; i++) pageable = pageable || host_is_pageable(host_in[i]);
        staged = stage_on && pageable;
        if (staged && !getenv("PFHE_PIPE_CHUNK_MB")) chunk_bytes = (size_t)32 << 20;
    }
    size_t chunk = chunk_bytes / (per_unit ? per_unit : 1);
    if (chunk == 0) chunk = 1;
    if (chunk > units) chunk = units;
    const int nbuf = (int)((units + chunk - 1) / chunk < (size_t)kPipe ? (units + chunk - 1) / chunk : kPipe);
    void *dbuf[kPipe] = {};
    pfhe_status status = PFHE_OK;
    if (staged) return pipelined_staged(device, host_in, n_in, in_bytes, host_out, out_bytes, units, launch, out_alias, scratch_bytes, chunk, nbuf, per_unit, lease.ctx);
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
