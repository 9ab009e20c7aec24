// capi.cu -- the C-ABI (include/pfhe.h): handle construction, table upload, host-slice shims with a
// pipelined H2D -> kernel -> D2H path, and the thin device-batch forwarders.
// No CPU compute path exists here: every transform / product is a kernel launch (no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "handles.hpp"
#include "rns.hpp"

namespace pfhe {

thread_local std::string t_last_cuda_error;
StreamPool g_stream_pool;

template <typename T>
static void fill_pass_tables(const host::HostTables<T> &h, int loge, std::vector<typename Word<T>::Pair> &fwd,
                             std::vector<typename Word<T>::Pair> &inv) {
    using Pair = typename Word<T>::Pair;
    const int logn = (int)h.log_n;
    const size_t n = h.n;
    PlanRt pl(logn, loge);
    fwd.assign(pl.total(), Pair{});
    inv.assign(pl.total(), Pair{});
    for (int p = 0; p < pl.npass; p++) {
        const int ns = pl.nstages(p), fb = pl.fb(p), nh = pl.nh(p), off = pl.pass_offset(p), s0 = pl.s0(p);
        for (int ls = 0; ls < ns; ls++) {
            const int jb = loge - 1 - ls;
            const int b = fb + jb;  // index bit of this stage
            const size_t inv_base = 1 + n - (n >> b);
            for (int jp = 0; jp < (1 << ls); jp++) {
                for (int high = 0; high < nh; high++) {
                    const size_t blk = ((size_t)high << ls) | (size_t)jp;
                    const size_t slot = (size_t)off + (size_t)((1 << ls) - 1 + jp) * nh + high;
                    const size_t fi = ((size_t)1 << (s0 + ls)) + blk;
                    fwd[slot].x = h.roots[fi];
                    fwd[slot].y = h.roots_q[fi];
                    const size_t ii = inv_base + blk;
                    if (b == logn - 1) {  // final inverse stage: inv_n * inv_roots[n-1]
                        inv[slot].x = h.inv_n_w;
                        inv[slot].y = h.inv_n_w_q;
                    } else {
                        inv[slot].x = h.inv_roots[ii];
                        inv[slot].y = h.inv_roots_q[ii];
                    }
                }
            }
        }
    }
}

template <typename T, typename H> pfhe_status create_handle(int device, uint32_t log_n, T q, H **out, bool generic_only) {
    using Pair = typename Word<T>::Pair;
    constexpr int BITS = sizeof(T) * 8;
    if (!out) return PFHE_ERR_INVALID_ARG;
    *out = nullptr;
    if (log_n == 0 || log_n + 1 >= (uint32_t)BITS) return PFHE_ERR_DEGREE_TOO_LARGE;
    T root;
    if (!host::min_primitive_root<T>(log_n + 1, q, root)) return PFHE_ERR_NO_PRIMITIVE_ROOT;  // root.rs:72-81
    if ((q >> (BITS - 2)) != 0) return PFHE_ERR_MODULUS_TOO_LARGE;                           // table.rs:318 / :195
    const uint32_t max_log_n = BITS == 64 ? 14 : 15;  // one polynomial per CTA in shared memory
    if (log_n > max_log_n) return PFHE_ERR_DEGREE_TOO_LARGE;
    int count = 0;
    PFHE_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;

    auto *hd = new (std::nothrow) H();
    if (!hd) return PFHE_ERR_NTT_TABLE;
    hd->device = device;
    host::build_tables<T>(log_n, q, root, hd->h);
    const auto &h = hd->h;
    const size_t n = h.n;
    std::vector<Pair> fwd(n), inv(n), fp, ip, fpl, ipl;
    std::vector<double> fpf, ipf, fplf, iplf;
    for (size_t k = 0; k < n; k++) {
        fwd[k].x = h.roots[k];
        fwd[k].y = h.roots_q[k];
        inv[k].x = h.inv_roots[k];
        inv[k].y = h.inv_roots_q[k];
    }
    inv[n - 1].x = h.inv_n_w;
    inv[n - 1].y = h.inv_n_w_q;
    for (size_t k = 0; k < 8 && n >= 8; k++) {
        hd->head.fwd_head[k] = fwd[k];
        hd->head.inv_tail[k] = inv[n - 8 + k];
    }
    // generic_only: UintNttTable handles always run the generic radix-2 kernel (no register-pass / FP64 layouts)
    const int loge = generic_only ? 0 : choose_loge(BITS, (int)log_n), loge_lat = generic_only ? 0 : lattice_loge(BITS, (int)log_n);
    if (loge) fill_pass_tables<T>(h, loge, fp, ip);
    if (loge_lat) fill_pass_tables<T>(h, loge_lat, fpl, ipl);
    // FP64-pipe path: u64 words and q < 2^50 (every table value is then an exact double)
    const char *no_f64 = getenv("PFHE_DISABLE_F64");
    // lazy-fold exactness budget (tools/f64_bounds.py): q <= 2^50 - 2^10
    const bool f64_ok = BITS == 64 && (uint64_t)q <= ((uint64_t)1 << 50) - 1024 && !(no_f64 && no_f64[0] == '1');
    const bool use_f64 = f64_ok && loge != 0;
    if (use_f64) {
        fpf.resize(fp.size());
        ipf.resize(ip.size());
        for (size_t i = 0; i < fp.size(); i++) {
            fpf[i] = (double)fp[i].x;
            ipf[i] = (double)ip[i].x;
        }
    }
    const bool use_f64_lat = f64_ok && loge_lat != 0;
    if (use_f64_lat) {
        fplf.resize(fpl.size());
        iplf.resize(ipl.size());
        for (size_t i = 0; i < fpl.size(); i++) {
            fplf[i] = (double)fpl[i].x;
            iplf[i] = (double)ipl[i].x;
        }
    }
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t sz_pair = sizeof(Pair);
    size_t off_fwd = 0, off_inv = align(off_fwd + n * sz_pair), off_fp = align(off_inv + n * sz_pair),
           off_ip = align(off_fp + fp.size() * sz_pair), off_fpl = align(off_ip + ip.size() * sz_pair),
           off_ipl = align(off_fpl + fpl.size() * sz_pair), off_ord = align(off_ipl + ipl.size() * sz_pair),
           off_fpf = align(off_ord + 2 * n * sizeof(T)), off_ipf = align(off_fpf + fpf.size() * sizeof(double)),
           off_fplf = align(off_ipf + ipf.size() * sizeof(double)), off_iplf = align(off_fplf + fplf.size() * sizeof(double)),
           total = align(off_iplf + iplf.size() * sizeof(double));
    std::vector<unsigned char> stage(total, 0);
    memcpy(stage.data() + off_fwd, fwd.data(), n * sz_pair);
    memcpy(stage.data() + off_inv, inv.data(), n * sz_pair);
    if (!fp.empty()) memcpy(stage.data() + off_fp, fp.data(), fp.size() * sz_pair);
    if (!ip.empty()) memcpy(stage.data() + off_ip, ip.data(), ip.size() * sz_pair);
    if (!fpl.empty()) memcpy(stage.data() + off_fpl, fpl.data(), fpl.size() * sz_pair);
    if (!ipl.empty()) memcpy(stage.data() + off_ipl, ipl.data(), ipl.size() * sz_pair);
    memcpy(stage.data() + off_ord, h.ordinal.data(), 2 * n * sizeof(T));
    if (!fpf.empty()) memcpy(stage.data() + off_fpf, fpf.data(), fpf.size() * sizeof(double));
    if (!ipf.empty()) memcpy(stage.data() + off_ipf, ipf.data(), ipf.size() * sizeof(double));
    if (!fplf.empty()) memcpy(stage.data() + off_fplf, fplf.data(), fplf.size() * sizeof(double));
    if (!iplf.empty()) memcpy(stage.data() + off_iplf, iplf.data(), iplf.size() * sizeof(double));
    cudaError_t e = cudaMalloc(&hd->blob, total);
    if (e == cudaSuccess) e = cudaMemcpy(hd->blob, stage.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (hd->blob) cudaFree(hd->blob);
        delete hd;
        return cuda_fail(e);
    }
    unsigned char *base = static_cast<unsigned char *>(hd->blob);
    DevNtt<T> d{};
    d.q = q;
    d.two_q = (T)(q << 1);
    d.inv_n = h.inv_n;
    d.inv_n_q = h.inv_n_q;
    d.br.q = q;
    host::barrett_ratio<T>(q, d.br.r0, d.br.r1);
    d.log_n = log_n;
    d.loge = (uint32_t)loge;
    d.fwd = reinterpret_cast<const Pair *>(base + off_fwd);
    d.inv = reinterpret_cast<const Pair *>(base + off_inv);
    d.fwd_pass = reinterpret_cast<const Pair *>(base + off_fp);
    d.inv_pass = reinterpret_cast<const Pair *>(base + off_ip);
    d.ordinal = reinterpret_cast<const T *>(base + off_ord);
    d.fwd_pass_f = reinterpret_cast<const double *>(base + off_fpf);
    d.inv_pass_f = reinterpret_cast<const double *>(base + off_ipf);
    d.q_f = (double)q;
    d.qinv_f = 1.0 / (double)q;
    d.inv_n_f = (double)h.inv_n;
    d.use_f64 = use_f64 ? 1u : 0u;
    hd->dev = d;
    hd->dev_lat = d;
    hd->dev_lat.loge = (uint32_t)loge_lat;
    hd->dev_lat.use_f64 = use_f64_lat ? 1u : 0u;
    hd->dev_lat.fwd_pass_f = reinterpret_cast<const double *>(base + off_fplf);
    hd->dev_lat.inv_pass_f = reinterpret_cast<const double *>(base + off_iplf);
    hd->dev_lat.fwd_pass = reinterpret_cast<const Pair *>(base + off_fpl);
    hd->dev_lat.inv_pass = reinterpret_cast<const Pair *>(base + off_ipl);
    *out = hd;
    return PFHE_OK;
}

template <typename H> void destroy_handle(H *h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    if (h->blob) cudaFree(h->blob);
    delete h;
}

template <typename T> pfhe_status host_transform(const NttHandle<T> *t, T *polys, size_t batch, bool fwd, bool lazy) {
    if (!t || (!polys && batch)) return PFHE_ERR_INVALID_ARG;
    const size_t bytes = sizeof(T) << t->h.log_n;
    const void *ins[1] = {polys};
    const size_t inb[1] = {bytes};
    // in place on the device: the "out" region doubles as the input region (n_in = 0 inputs + copy in manually)
    return pipelined(t->device, ins, 1, inb, polys, bytes, batch, [&](const void *const *din, void *dout, size_t nu, cudaStream_t s) {
        const T *src = static_cast<const T *>(din[0]);
        if (lazy) {  // lazy trait contract: inputs in [0,4q) / [0,2q) -> canonicalise on the device first
            LimbConsts<T> lc{};
            lc.br[0] = t->dev.br;
            cudaError_t e = launch_slice_op<T>(PFHE_OP_REDUCE_LAZY, lc, 1, src, nullptr, nullptr, const_cast<T *>(src), nu,
                                               (size_t)1 << t->h.log_n, s);
            if (e != cudaSuccess) return e;
        }
        return launch_ntt<T>(t->dev, nullptr, 1, src, static_cast<T *>(dout), nu, fwd, s);
    });
}

template <typename T> pfhe_status host_polymul(const NttHandle<T> *t, const T *a, const T *b, T *c, size_t batch) {
    if (!t || ((!a || !b || !c) && batch)) return PFHE_ERR_INVALID_ARG;
    const size_t bytes = sizeof(T) << t->h.log_n;
    const void *ins[2] = {a, b};
    const size_t inb[2] = {bytes, bytes};
    return pipelined(t->device, ins, 2, inb, c, bytes, batch, [&](const void *const *din, void *dout, size_t nu, cudaStream_t s) {
        return launch_polymul<T>(t->dev, nullptr, 1, static_cast<const T *>(din[0]), static_cast<const T *>(din[1]), static_cast<T *>(dout),
                                 nu, s);
    });
}

template <typename T> static pfhe_status host_monomial(const NttHandle<T> *t, T coeff, size_t degree, T *values) {
    if (!t || !values) return PFHE_ERR_INVALID_ARG;
    if (degree >= 2 * t->h.n || coeff >= t->h.q) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    StreamLease lease(t->device);
    PFHE_CUDA(lease.err);
    cudaStream_t *st = lease.ctx->streams;
    const size_t bytes = sizeof(T) << t->h.log_n;
    void *d = nullptr;
    PFHE_CUDA(cudaMallocFromPoolAsync(&d, bytes + 256, lease.ctx->pool, st[0]));
    uint32_t *ddeg = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(d) + bytes);
    const uint32_t deg32 = (uint32_t)degree;
    cudaError_t e = cudaMemcpyAsync(ddeg, &deg32, sizeof(deg32), cudaMemcpyHostToDevice, st[0]);
    if (e == cudaSuccess) e = launch_monomial<T>(t->dev, coeff, ddeg, static_cast<T *>(d), 1, st[0]);
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, d, bytes, cudaMemcpyDeviceToHost, st[0]);
    cudaFreeAsync(d, st[0]);
    cudaError_t e2 = cudaStreamSynchronize(st[0]);
    if (e != cudaSuccess) return cuda_fail(e);
    if (e2 != cudaSuccess) return cuda_fail(e2);
    return PFHE_OK;
}

template <typename T> static pfhe_status make_limb_consts(const T *moduli, size_t limbs, const T *scalars, int op, LimbConsts<T> &lc) {
    constexpr int BITS = sizeof(T) * 8;
    if (!moduli || limbs == 0 || limbs > (size_t)kMaxLimbs) return PFHE_ERR_INVALID_ARG;
    const bool needs_scalar = op == PFHE_OP_MUL_SCALAR || op == PFHE_OP_ADD_MUL_SCALAR || op == PFHE_OP_FACTOR_MUL ||
                              op == PFHE_OP_ADD_FACTOR_MUL || op == PFHE_OP_SUB_FACTOR_MUL || op == PFHE_OP_MUL_SCALAR_ADD ||
                              op == PFHE_OP_FACTOR_MUL_ADD;
    if (needs_scalar && !scalars) return PFHE_ERR_INVALID_ARG;
    for (size_t i = 0; i < limbs; i++) {
        const T q = moduli[i];
        if (q <= 1) return PFHE_ERR_INVALID_ARG;
        if ((q >> (BITS - 2)) != 0) return PFHE_ERR_MODULUS_TOO_LARGE;  // BarrettModulus::new, barrett/mod.rs:39-43
        lc.br[i].q = q;
        host::barrett_ratio<T>(q, lc.br[i].r0, lc.br[i].r1);
        lc.scalar[i] = needs_scalar ? scalars[i] : 0;
        if (needs_scalar && scalars[i] >= q) return PFHE_ERR_INVALID_ARG;
        lc.scalar_q[i] = needs_scalar ? host::shoup_quot<T>(scalars[i], q) : 0;
    }
    return PFHE_OK;
}

template <typename T>
static pfhe_status slice_op_dev(int op, const T *moduli, size_t limbs, const T *scalars, const T *a, const T *b, const T *c, T *out, size_t rows,
                                size_t n, void *stream) {
    if (op < 0 || op > PFHE_OP_FACTOR_MUL_ADD) return PFHE_ERR_INVALID_ARG;
    if (rows * n == 0) return PFHE_OK;
    if (!a || !out) return PFHE_ERR_INVALID_ARG;
    if (op <= PFHE_OP_SUB && !b) return PFHE_ERR_INVALID_ARG;
    if ((op == PFHE_OP_MUL_ADD || op == PFHE_OP_MUL_SCALAR_ADD || op == PFHE_OP_FACTOR_MUL_ADD) && !c) return PFHE_ERR_INVALID_ARG;
    LimbConsts<T> lc;
    pfhe_status s = make_limb_consts<T>(moduli, limbs, scalars, op, lc);
    if (s != PFHE_OK) return s;
    PFHE_PTR_GUARD(out);
    PFHE_CUDA(launch_slice_op<T>(op, lc, (int)limbs, a, b, c, out, rows, n, static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}

}  // namespace pfhe

using namespace pfhe;


template <typename T> struct RnsHandle {
    std::vector<T> moduli;
    RnsDev<T> base{};  // RNS part only (log_basis = 0); gadget variants are derived per call (a few hundred host cycles)
};
template <typename T> struct BaseConvHandle {
    BaseConvDev<T> dev{};
};
struct pfhe_baseconv32 : BaseConvHandle<uint32_t> {};
struct pfhe_baseconv64 : BaseConvHandle<uint64_t> {};
struct pfhe_rns32 : RnsHandle<uint32_t> {};
struct pfhe_rns64 : RnsHandle<uint64_t> {};

namespace pfhe {

template <typename T, typename H, typename D>
static pfhe_status create_dcrt(int device, uint32_t log_n, const T *moduli, size_t count, D **out) {
    if (!out) return PFHE_ERR_INVALID_ARG;
    *out = nullptr;
    if (!moduli || count == 0) return PFHE_ERR_RNS_EMPTY;
    if (count > (size_t)kMaxLimbs) return PFHE_ERR_INVALID_ARG;
    auto *d = new (std::nothrow) D();
    if (!d) return PFHE_ERR_NTT_TABLE;
    d->device = device;
    pfhe_status s = PFHE_OK;
    for (size_t i = 0; i < count && s == PFHE_OK; i++) {
        H *h = nullptr;
        s = create_handle<T, H>(device, log_n, moduli[i], &h, false);
        if (s == PFHE_OK) d->limbs.push_back(h);
    }
    if (s == PFHE_OK) {
        DeviceGuard guard(device);
        std::vector<DevNtt<T>> tabs;
        bool all_f64 = true;
        for (auto *h : d->limbs) all_f64 = all_f64 && h->dev.use_f64;
        for (auto *h : d->limbs) {
            tabs.push_back(h->dev);
            tabs.back().use_f64 = all_f64 ? 1u : 0u;  // one field policy per launch
        }
        d->tb0 = tabs[0];
        cudaError_t e = cudaMalloc(&d->d_tables, tabs.size() * sizeof(DevNtt<T>));
        if (e == cudaSuccess) e = cudaMemcpy(d->d_tables, tabs.data(), tabs.size() * sizeof(DevNtt<T>), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && d->limbs[0]->dev_lat.loge != 0) {
            std::vector<DevNtt<T>> lat;
            bool all_f64_lat = true, all_wide = sizeof(T) == 4;
            for (auto *h : d->limbs) {
                all_f64_lat = all_f64_lat && h->dev_lat.use_f64;
                all_wide = all_wide && dcrt_wide32_ok((uint64_t)h->h.q, (int)log_n);
            }
            for (auto *h : d->limbs) {
                lat.push_back(h->dev_lat);
                lat.back().use_f64 = all_f64_lat ? 1u : 0u;
            }
            d->lat_policy = all_f64_lat ? 1 : (all_wide ? 2 : 0);
            e = cudaMalloc(&d->d_tables_lat, lat.size() * sizeof(DevNtt<T>));
            if (e == cudaSuccess) e = cudaMemcpy(d->d_tables_lat, lat.data(), lat.size() * sizeof(DevNtt<T>), cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) s = cuda_fail(e);
    }
    if (s != PFHE_OK) {
        for (auto *h : d->limbs) destroy_handle(static_cast<H *>(h));
        if (d->d_tables) cudaFree(d->d_tables);
        if (d->d_tables_lat) cudaFree(d->d_tables_lat);
        delete d;
        return s;
    }
    *out = d;
    return PFHE_OK;
}

template <typename T, typename D> pfhe_status dcrt_host_transform(const D *t, T *polys, size_t batch, bool fwd, bool lazy) {
    if (!t || (!polys && batch)) return PFHE_ERR_INVALID_ARG;
    const size_t L = t->limbs.size();
    const size_t bytes = (sizeof(T) << t->limbs[0]->h.log_n) * L;
    const void *ins[1] = {polys};
    const size_t inb[1] = {bytes};
    LimbConsts<T> lc{};
    for (size_t i = 0; i < L; i++) lc.br[i] = t->limbs[i]->dev.br;
    return pipelined(t->device, ins, 1, inb, polys, bytes, batch, [&](const void *const *din, void *dout, size_t nu, cudaStream_t s) {
        if (lazy) {  // lazy trait contract (dcrt/mod.rs:77-103): inputs in [0,4q_i) / [0,2q_i) -> canonicalise per limb on the device first
            T *src = const_cast<T *>(static_cast<const T *>(din[0]));
            cudaError_t e = launch_slice_op<T>(PFHE_OP_REDUCE_LAZY, lc, (int)L, src, nullptr, nullptr, src, nu, (size_t)1 << t->limbs[0]->h.log_n, s);
            if (e != cudaSuccess) return e;
        }
        return launch_ntt<T>(t->tb0, t->d_tables, (int)L, static_cast<const T *>(din[0]), static_cast<T *>(dout), nu * L, fwd, s);
    });
}

template <typename T>
static pfhe_status slice_op_host(int op, const T *moduli, size_t limbs, const T *scalars, const T *a, const T *b, const T *c, T *out, size_t rows,
                                 size_t n) {
    if (op < 0 || op > PFHE_OP_FACTOR_MUL_ADD) return PFHE_ERR_INVALID_ARG;
    if (rows * n == 0) return PFHE_OK;
    if (!a || !out) return PFHE_ERR_INVALID_ARG;
    if (op <= PFHE_OP_SUB && !b) return PFHE_ERR_INVALID_ARG;
    if ((op == PFHE_OP_MUL_ADD || op == PFHE_OP_MUL_SCALAR_ADD || op == PFHE_OP_FACTOR_MUL_ADD) && !c) return PFHE_ERR_INVALID_ARG;
    LimbConsts<T> lc;
    pfhe_status s = make_limb_consts<T>(moduli, limbs, scalars, op, lc);
    if (s != PFHE_OK) return s;
    const bool acc = op == PFHE_OP_ADD_MUL || op == PFHE_OP_SUB_MUL || op == PFHE_OP_ADD_MUL_SCALAR || op == PFHE_OP_ADD_FACTOR_MUL ||
                     op == PFHE_OP_SUB_FACTOR_MUL;
    const bool rb = op <= PFHE_OP_SUB, rc = op == PFHE_OP_MUL_ADD || op == PFHE_OP_MUL_SCALAR_ADD || op == PFHE_OP_FACTOR_MUL_ADD;
    const size_t bytes = sizeof(T) * limbs * n;
    const void *ins[4];
    size_t inb[4];
    int n_in = 0, ia = 0, ib = -1, ic = -1, io = -1;
    ins[n_in] = a; inb[n_in] = bytes; ia = n_in++;
    if (rb) { ins[n_in] = b; inb[n_in] = bytes; ib = n_in++; }
    if (rc) { ins[n_in] = c; inb[n_in] = bytes; ic = n_in++; }
    if (acc) { ins[n_in] = out; inb[n_in] = bytes; io = n_in++; }
    int dev = 0;
    cudaGetDevice(&dev);
    return pipelined(dev, ins, n_in, inb, out, bytes, rows, [&](const void *const *din, void *dout, size_t nu, cudaStream_t st) {
        return launch_slice_op<T>(op, lc, (int)limbs, static_cast<const T *>(din[ia]), ib >= 0 ? static_cast<const T *>(din[ib]) : nullptr,
                                  ic >= 0 ? static_cast<const T *>(din[ic]) : nullptr, static_cast<T *>(dout), nu, n, st);
    }, io);
}

// single-modulus external product: the re-scheduled u32 kernels (lattice32_ep.cu) where they apply, the generic kernel otherwise
template <typename T, typename H>
static cudaError_t run_external_product(const H *t, const GadgetParams<T> &g, uint32_t k, const T *key, const T *in, T *out, size_t batch, bool to_coeff,
                                        cudaStream_t s) {
    if constexpr (sizeof(T) == 4) {
        const cudaError_t e = launch_external_product_fast32(t->dev_lat, t->head, g, k, key, in, out, batch, to_coeff, s);
        if (e != cudaErrorNotSupported) return e;
    }
    return launch_external_product<T>(t->dev_lat, g, k, key, in, out, batch, to_coeff, s);
}

template <typename T, typename H>
static pfhe_status ext_prod(const H *t, uint32_t k, uint32_t log_basis, uint32_t levels_in, const T *key, const T *in, T *out, size_t batch,
                            int to_coeff, void *stream) {
    if (!t || ((!key || !in || !out) && batch)) return PFHE_ERR_INVALID_ARG;
    GadgetParams<T> g;
    if (!make_gadget<T>(t->h.q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;
    if (t->dev_lat.loge == 0 || k < 1 || k > 2) return PFHE_ERR_UNSUPPORTED;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(run_external_product<T>(t, g, k, key, in, out, batch, to_coeff != 0, static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}
// Host-slice shim of the single-modulus external product: key uploaded once, ciphertexts streamed through the pipelined
// H2D -> kernel -> D2H path (the reference's CrtGlwe::mul_dcrt_ggsw_to works on host slices, primus_data/src/traits.rs:20).
template <typename T, typename H>
pfhe_status ext_prod_host(const H *t, uint32_t k, uint32_t log_basis, uint32_t levels_in, const T *key, const T *in, T *out, size_t batch,
                                 int to_coeff) {
    if (!t || ((!key || !in || !out) && batch)) return PFHE_ERR_INVALID_ARG;
    GadgetParams<T> g;
    if (!make_gadget<T>(t->h.q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;
    if (t->dev_lat.loge == 0 || k < 1 || k > 2) return PFHE_ERR_UNSUPPORTED;
    if (batch == 0) return PFHE_OK;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    const size_t n = t->h.n, comps = (size_t)k + 1;
    const size_t key_bytes = comps * g.levels * comps * n * sizeof(T), ct_bytes = comps * n * sizeof(T);
    void *dkey = nullptr;
    PFHE_CUDA(cudaMalloc(&dkey, key_bytes));
    cudaError_t e = cudaMemcpy(dkey, key, key_bytes, cudaMemcpyHostToDevice);
    pfhe_status status = e == cudaSuccess ? PFHE_OK : cuda_fail(e);
    if (status == PFHE_OK) {
        const void *ins[1] = {in};
        const size_t inb[1] = {ct_bytes};
        status = pipelined(t->device, ins, 1, inb, out, ct_bytes, batch, [&](const void *const *din, void *dout, size_t nu, cudaStream_t s) {
            return run_external_product<T>(t, g, k, static_cast<const T *>(dkey), static_cast<const T *>(din[0]), static_cast<T *>(dout), nu,
                                           to_coeff != 0, s);
        });
    }
    cudaFree(dkey);
    return status;
}
template <typename T, typename H>
static pfhe_status blind_rot(const H *t, uint32_t log_basis, uint32_t levels_in, const T *bsk, uint32_t n_lwe, const uint32_t *lwe,
                             const T *tv, T *acc_out, size_t batch, void *stream) {
    if (!t || ((!bsk || !lwe || !tv || !acc_out) && batch)) return PFHE_ERR_INVALID_ARG;
    GadgetParams<T> g;
    if (!make_gadget<T>(t->h.q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;
    if (t->dev_lat.loge == 0) return PFHE_ERR_UNSUPPORTED;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    if constexpr (sizeof(T) == 4) {  // bootstrapping shape of BASELINE config 5: the re-scheduled kernel of lattice32.cu
        const cudaError_t e = launch_blind_rotate_fast32(t->dev_lat, t->head, g, bsk, n_lwe, lwe, tv, acc_out, batch, static_cast<cudaStream_t>(stream));
        if (e != cudaErrorNotSupported) {
            PFHE_CUDA(e);
            return PFHE_OK;
        }
    }
    PFHE_CUDA(launch_blind_rotate<T>(t->dev_lat, g, bsk, n_lwe, lwe, tv, acc_out, batch, static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}


template <typename T, typename R> static pfhe_status rns_create(const T *moduli, size_t count, R **out) {
    if (!out) return PFHE_ERR_INVALID_ARG;
    *out = nullptr;
    if (!moduli || count == 0) return PFHE_ERR_RNS_EMPTY;
    auto *h = new (std::nothrow) R();
    if (!h) return PFHE_ERR_INVALID_ARG;
    const int rc = make_rns<T>(moduli, count, 0, 0, h->base);
    if (rc != 0) {
        delete h;
        return (pfhe_status)rc;
    }
    h->moduli.assign(moduli, moduli + count);
    *out = h;
    return PFHE_OK;
}
// RnsDev with the gadget part filled in (BigUintApproxSignedBasis::new)
template <typename T, typename R> static pfhe_status rns_with_gadget(const R *r, uint32_t log_basis, uint32_t levels_in, RnsDev<T> &out) {
    if (!r || log_basis == 0) return PFHE_ERR_INVALID_ARG;
    const int rc = make_rns<T>(r->moduli.data(), r->moduli.size(), log_basis, levels_in, out);
    return (pfhe_status)rc;
}
template <typename T> static pfhe_status limb_consts_plain(const T *moduli, size_t limbs, LimbConsts<T> &lc) {
    return make_limb_consts<T>(moduli, limbs, nullptr, PFHE_OP_MUL, lc);
}

template <typename T, typename D, typename R>
static size_t ext_scratch_bytes(const D *t, const R *r, uint32_t k, uint32_t log_basis, uint32_t levels_in, size_t batch) {
    RnsDev<T> g;
    if (!t || rns_with_gadget<T>(r, log_basis, levels_in, g) != PFHE_OK) return 0;
    const size_t n = t->limbs[0]->h.n, L = t->limbs.size();
    return batch * (size_t)(k + 1) * g.levels * L * n * sizeof(T);
}

template <typename T, typename D, typename R>
static pfhe_status dcrt_ext_prod(const D *t, const R *r, uint32_t k, uint32_t log_basis, uint32_t levels_in, const T *key, const T *in, T *out,
                                 size_t batch, int to_coeff, void *scratch, size_t scratch_bytes, void *stream) {
    if (!t || !r || ((!key || !in || !out) && batch)) return PFHE_ERR_INVALID_ARG;
    if (batch == 0) return PFHE_OK;
    const size_t L = t->limbs.size(), n = t->limbs[0]->h.n, comps = (size_t)k + 1;
    if (L != r->moduli.size() || k < 1) return PFHE_ERR_INVALID_ARG;
    if (k > 2) return PFHE_ERR_UNSUPPORTED;  // GLWE dimension 1 (RLWE) and 2, like the fused single-modulus kernel
    for (size_t i = 0; i < L; i++)
        if (t->limbs[i]->h.q != r->moduli[i]) return PFHE_ERR_INVALID_ARG;
    RnsDev<T> g;
    pfhe_status st = rns_with_gadget<T>(r, log_basis, levels_in, g);
    if (st != PFHE_OK) return st;
    LimbConsts<T> lc;
    if ((st = limb_consts_plain<T>(r->moduli.data(), L, lc)) != PFHE_OK) return st;
    static const bool unfused = getenv("PFHE_DCRT_EP_UNFUSED") != nullptr;     // A/B tuning hooks
    static const bool two_kernel = getenv("PFHE_DCRT_EP_TWO_KERNEL") != nullptr;
    if (t->d_tables_lat && !unfused && !two_kernel) {  // composed values of at most two words: ONE kernel, no digit round trip, no scratch
        DeviceGuard fguard(t->device);
        if (!fguard.ok) return PFHE_ERR_CUDA;
        const cudaError_t fe = launch_dcrt_external_product_fused<T>(t->lat_policy, t->d_tables_lat, g, t->limbs[0]->h.log_n, k, key, in, out, batch,
                                                                     to_coeff != 0, static_cast<cudaStream_t>(stream));
        if (fe != cudaErrorNotSupported) {
            PFHE_CUDA(fe);
            return PFHE_OK;
        }
    }
    const size_t per_ct = comps * g.levels * L * n * sizeof(T);
    if (!scratch || scratch_bytes < per_ct) return PFHE_ERR_INVALID_ARG;
    const size_t chunk = scratch_bytes / per_ct;
    DeviceGuard guard(t->device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    T *digits = static_cast<T *>(scratch);
    const size_t glwe_len = comps * L * n;
    for (size_t done = 0; done < batch; done += chunk) {
        const size_t nb = batch - done < chunk ? batch - done : chunk;
        const T *cin = in + done * glwe_len;
        T *cout = out + done * glwe_len;
        // digits of every input component: [ct][r][level][limb][n]
        PFHE_CUDA(launch_rns_gadget<T>(g, cin, digits, n, nb * comps, L * n, (size_t)g.levels * L * n, s));
        if (t->d_tables_lat && !unfused) {  // one fused kernel per (ciphertext, limb): digits read once, nothing written back
            PFHE_CUDA(launch_dcrt_external_product<T>(t->lat_policy, t->d_tables_lat, (int)L, t->limbs[0]->h.log_n, k, g.levels, key, digits,
                                                      cout, nb, to_coeff != 0, s));
            continue;
        }
        PFHE_CUDA(launch_ntt<T>(t->tb0, t->d_tables, (int)L, digits, digits, nb * comps * g.levels * L, true, s));
        PFHE_CUDA(launch_rns_key_mac<T>(lc, (int)L, (int)comps, g.levels, digits, key, cout, n, nb, s));
        if (to_coeff) PFHE_CUDA(launch_ntt<T>(t->tb0, t->d_tables, (int)L, cout, cout, nb * comps * L, false, s));
    }
    return PFHE_OK;
}

template pfhe_status create_handle<uint32_t, pfhe_ntt32>(int, uint32_t, uint32_t, pfhe_ntt32 **, bool);
template pfhe_status create_handle<uint64_t, pfhe_ntt64>(int, uint32_t, uint64_t, pfhe_ntt64 **, bool);
template void destroy_handle<pfhe_ntt32>(pfhe_ntt32 *);
template void destroy_handle<pfhe_ntt64>(pfhe_ntt64 *);
template pfhe_status host_transform<uint32_t>(const NttHandle<uint32_t> *, uint32_t *, size_t, bool, bool);
template pfhe_status host_transform<uint64_t>(const NttHandle<uint64_t> *, uint64_t *, size_t, bool, bool);
template pfhe_status host_polymul<uint32_t>(const NttHandle<uint32_t> *, const uint32_t *, const uint32_t *, uint32_t *, size_t);
template pfhe_status host_polymul<uint64_t>(const NttHandle<uint64_t> *, const uint64_t *, const uint64_t *, uint64_t *, size_t);
template pfhe_status ext_prod_host<uint32_t, pfhe_ntt32>(const pfhe_ntt32 *, uint32_t, uint32_t, uint32_t, const uint32_t *, const uint32_t *, uint32_t *, size_t, int);
template pfhe_status ext_prod_host<uint64_t, pfhe_ntt64>(const pfhe_ntt64 *, uint32_t, uint32_t, uint32_t, const uint64_t *, const uint64_t *, uint64_t *, size_t, int);
template pfhe_status dcrt_host_transform<uint32_t, pfhe_dcrt32>(const pfhe_dcrt32 *, uint32_t *, size_t, bool, bool);
template pfhe_status dcrt_host_transform<uint64_t, pfhe_dcrt64>(const pfhe_dcrt64 *, uint64_t *, size_t, bool, bool);

}  // namespace pfhe

extern "C" {

const char *pfhe_status_string(pfhe_status s) {
    switch (s) {
        case PFHE_OK: return "Ok";
        case PFHE_ERR_NO_PRIMITIVE_ROOT: return "NoPrimitiveRoot";
        case PFHE_ERR_DEGREE_CONVERSION: return "DegreeConversionErr";
        case PFHE_ERR_DEGREE_TOO_LARGE: return "DegreeTooLarge";
        case PFHE_ERR_NTT_TABLE: return "NttTableErr";
        case PFHE_ERR_MODULUS_TOO_LARGE: return "ModulusTooLarge";
        case PFHE_ERR_RNS_EMPTY: return "EmptyBase";
        case PFHE_ERR_RNS_NOT_COPRIME: return "CoPrimeError";
        case PFHE_ERR_CUDA: return "CudaError";
        case PFHE_ERR_INVALID_ARG: return "InvalidArgument";
        case PFHE_ERR_UNSUPPORTED: return "Unsupported";
    }
    return "Unknown";
}
const char *pfhe_last_cuda_error(void) { return t_last_cuda_error.c_str(); }
const char *pfhe_version(void) { return "primus_fhe_b200 0.2.0 (abi 2)"; }
const char *pfhe_compiled_arch(void) { return "sm_100a"; }
uint64_t pfhe_launch_count(void) { return g_launches.load(); }

#define PFHE_DEFINE_WORD(B, T)                                                                                                        \
    pfhe_status pfhe_ntt##B##_create(int device, uint32_t log_n, T q, pfhe_ntt##B **out) {                                            \
        return create_handle<T, pfhe_ntt##B>(device, log_n, q, out, false);                                                                  \
    }                                                                                                                                 \
    void pfhe_ntt##B##_destroy(pfhe_ntt##B *t) { destroy_handle(t); }                                                                 \
    size_t pfhe_ntt##B##_poly_length(const pfhe_ntt##B *t) { return t ? t->h.n : 0; }                                                 \
    T pfhe_ntt##B##_modulus(const pfhe_ntt##B *t) { return t ? t->h.q : 0; }                                                          \
    T pfhe_ntt##B##_root(const pfhe_ntt##B *t) { return t ? t->h.root : 0; }                                                          \
    T pfhe_ntt##B##_inv_root(const pfhe_ntt##B *t) { return t ? t->h.inv_root : 0; }                                                  \
    T pfhe_ntt##B##_inv_n(const pfhe_ntt##B *t) { return t ? t->h.inv_n : 0; }                                                        \
    int pfhe_ntt##B##_device(const pfhe_ntt##B *t) { return t ? t->device : -1; }                                                     \
    pfhe_status pfhe_ntt##B##_transform_slice(const pfhe_ntt##B *t, T *poly, int lazy) { return host_transform<T>(t, poly, 1, true, lazy != 0); }     \
    pfhe_status pfhe_ntt##B##_inverse_transform_slice(const pfhe_ntt##B *t, T *v, int lazy) { return host_transform<T>(t, v, 1, false, lazy != 0); }  \
    pfhe_status pfhe_ntt##B##_transform_slices(const pfhe_ntt##B *t, T *p, size_t batch, int lazy) {                                  \
        return host_transform<T>(t, p, batch, true, lazy != 0);                                                                                \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_inverse_transform_slices(const pfhe_ntt##B *t, T *p, size_t batch, int lazy) {                          \
        return host_transform<T>(t, p, batch, false, lazy != 0);                                                                                \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_transform_monomial(const pfhe_ntt##B *t, T coeff, size_t degree, T *values) {                           \
        return host_monomial<T>(t, coeff, degree, values);                                                                            \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_transform_coeff_one_monomial(const pfhe_ntt##B *t, size_t degree, T *values) {                          \
        return host_monomial<T>(t, (T)1, degree, values);                                                                             \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_transform_coeff_minus_one_monomial(const pfhe_ntt##B *t, size_t degree, T *values) {                    \
        return t ? host_monomial<T>(t, (T)(t->h.q - 1), degree, values) : PFHE_ERR_INVALID_ARG;                                       \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_forward_batch(const pfhe_ntt##B *t, T *dev, size_t batch, void *stream) {                               \
        if (!t || (!dev && batch)) return PFHE_ERR_INVALID_ARG;                                                                       \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_ntt<T>(t->dev, nullptr, 1, dev, dev, batch, true, static_cast<cudaStream_t>(stream)));                       \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_inverse_batch(const pfhe_ntt##B *t, T *dev, size_t batch, void *stream) {                               \
        if (!t || (!dev && batch)) return PFHE_ERR_INVALID_ARG;                                                                       \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_ntt<T>(t->dev, nullptr, 1, dev, dev, batch, false, static_cast<cudaStream_t>(stream)));                      \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_forward_batch_to(const pfhe_ntt##B *t, const T *src, T *dst, size_t batch, void *stream) {              \
        if (!t || ((!src || !dst) && batch)) return PFHE_ERR_INVALID_ARG;                                                             \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_ntt<T>(t->dev, nullptr, 1, src, dst, batch, true, static_cast<cudaStream_t>(stream)));                       \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_inverse_batch_to(const pfhe_ntt##B *t, const T *src, T *dst, size_t batch, void *stream) {              \
        if (!t || ((!src || !dst) && batch)) return PFHE_ERR_INVALID_ARG;                                                             \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_ntt<T>(t->dev, nullptr, 1, src, dst, batch, false, static_cast<cudaStream_t>(stream)));                      \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_monomial_batch(const pfhe_ntt##B *t, T coeff, const uint32_t *degrees, T *out, size_t batch,            \
                                             void *stream) {                                                                          \
        if (!t || ((!degrees || !out) && batch) || coeff >= t->h.q) return PFHE_ERR_INVALID_ARG;                                      \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_monomial<T>(t->dev, coeff, degrees, out, batch, static_cast<cudaStream_t>(stream)));                         \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_polymul_batch(const pfhe_ntt##B *t, const T *a, const T *b, T *c, size_t batch, void *stream) {         \
        if (!t || ((!a || !b || !c) && batch)) return PFHE_ERR_INVALID_ARG;                                                           \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_polymul<T>(t->dev, nullptr, 1, a, b, c, batch, static_cast<cudaStream_t>(stream)));                          \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ntt##B##_polymul_slices(const pfhe_ntt##B *t, const T *a, const T *b, T *c, size_t batch) {                      \
        return host_polymul<T>(t, a, b, c, batch);                                                                                    \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_create(int device, uint32_t log_n, const T *moduli, size_t count, pfhe_dcrt##B **out) {                \
        return create_dcrt<T, pfhe_ntt##B, pfhe_dcrt##B>(device, log_n, moduli, count, out);                                          \
    }                                                                                                                                 \
    void pfhe_dcrt##B##_destroy(pfhe_dcrt##B *t) {                                                                                    \
        if (!t) return;                                                                                                               \
        for (auto *h : t->limbs) destroy_handle(static_cast<pfhe_ntt##B *>(h));                                                       \
        if (t->d_tables) {                                                                                                            \
            DeviceGuard guard(t->device);                                                                                             \
            cudaFree(t->d_tables);                                                                                                    \
            if (t->d_tables_lat) cudaFree(t->d_tables_lat);                                                                           \
        }                                                                                                                             \
        delete t;                                                                                                                     \
    }                                                                                                                                 \
    size_t pfhe_dcrt##B##_poly_length(const pfhe_dcrt##B *t) { return t ? t->limbs[0]->h.n : 0; }                                     \
    size_t pfhe_dcrt##B##_moduli_count(const pfhe_dcrt##B *t) { return t ? t->limbs.size() : 0; }                                     \
    size_t pfhe_dcrt##B##_crt_poly_length(const pfhe_dcrt##B *t) { return t ? t->limbs.size() * t->limbs[0]->h.n : 0; }               \
    const pfhe_ntt##B *pfhe_dcrt##B##_ntt_table(const pfhe_dcrt##B *t, size_t limb) {                                                 \
        return (t && limb < t->limbs.size()) ? static_cast<const pfhe_ntt##B *>(t->limbs[limb]) : nullptr;                            \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_transform_slices(const pfhe_dcrt##B *t, T *p, size_t batch, int lazy) {                                \
        return dcrt_host_transform<T>(t, p, batch, true, lazy != 0);                                                                  \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_inverse_transform_slices(const pfhe_dcrt##B *t, T *p, size_t batch, int lazy) {                        \
        return dcrt_host_transform<T>(t, p, batch, false, lazy != 0);                                                                 \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_forward_batch(const pfhe_dcrt##B *t, T *dev, size_t batch, void *stream) {                             \
        if (!t || (!dev && batch)) return PFHE_ERR_INVALID_ARG;                                                                       \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_ntt<T>(t->tb0, t->d_tables, (int)t->limbs.size(), dev, dev, batch * t->limbs.size(), true,         \
                                static_cast<cudaStream_t>(stream)));                                                                  \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_inverse_batch(const pfhe_dcrt##B *t, T *dev, size_t batch, void *stream) {                             \
        if (!t || (!dev && batch)) return PFHE_ERR_INVALID_ARG;                                                                       \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_ntt<T>(t->tb0, t->d_tables, (int)t->limbs.size(), dev, dev, batch * t->limbs.size(), false,        \
                                static_cast<cudaStream_t>(stream)));                                                                  \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_polymul_batch(const pfhe_dcrt##B *t, const T *a, const T *b, T *c, size_t batch, void *stream) {       \
        if (!t || ((!a || !b || !c) && batch)) return PFHE_ERR_INVALID_ARG;                                                           \
        PFHE_DEV_GUARD(t->device);                                                                                                 \
        PFHE_CUDA(launch_polymul<T>(t->tb0, t->d_tables, (int)t->limbs.size(), a, b, c, batch * t->limbs.size(),            \
                                    static_cast<cudaStream_t>(stream)));                                                              \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_mod##B##_slice_op(pfhe_slice_op op, const T *moduli, size_t limbs, const T *scalars, const T *a, const T *b,     \
                                       const T *c, T *out, size_t rows, size_t n, void *stream) {                                     \
        return slice_op_dev<T>((int)op, moduli, limbs, scalars, a, b, c, out, rows, n, stream);                                       \
    }                                                                                                                                 \
    pfhe_status pfhe_mod##B##_slice_op_host(pfhe_slice_op op, const T *moduli, size_t limbs, const T *scalars, const T *a,            \
                                            const T *b, const T *c, T *out, size_t rows, size_t n) {                                  \
        return slice_op_host<T>((int)op, moduli, limbs, scalars, a, b, c, out, rows, n);                                              \
    }                                                                                                                                 \
    pfhe_status pfhe_mod##B##_slice_op_bcast(pfhe_slice_op op, const T *moduli, size_t limbs, const T *a, const T *b, T *out,         \
                                             size_t rows, size_t n, size_t group, void *stream) {                                     \
        if (op != PFHE_OP_MUL && op != PFHE_OP_ADD_MUL && op != PFHE_OP_SUB_MUL) return PFHE_ERR_INVALID_ARG;                         \
        if (group == 0 || rows % group) return PFHE_ERR_INVALID_ARG;                                                                  \
        if (rows * n == 0) return PFHE_OK;                                                                                            \
        if (!a || !b || !out) return PFHE_ERR_INVALID_ARG;                                                                            \
        LimbConsts<T> lc;                                                                                                             \
        pfhe_status st = make_limb_consts<T>(moduli, limbs, nullptr, (int)op, lc);                                                    \
        if (st != PFHE_OK) return st;                                                                                                 \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_slice_op<T>((int)op, lc, (int)limbs, a, b, nullptr, out, rows, n, static_cast<cudaStream_t>(stream), group)); \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_mod##B##_butterfly_mul_factor(const T *moduli, size_t limbs, T *a, const T *s, const T *w, T *out, size_t rows,    \
                                                   size_t n, void *stream) {                                                          \
        LimbConsts<T> lc;                                                                                                             \
        pfhe_status st = make_limb_consts<T>(moduli, limbs, nullptr, PFHE_OP_MUL, lc);                                                \
        if (st != PFHE_OK) return st;                                                                                                 \
        if ((!a || !s || !w || !out) && rows * n) return PFHE_ERR_INVALID_ARG;                                                        \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_butterfly_mul<T>(lc, (int)limbs, a, s, w, out, rows, n, static_cast<cudaStream_t>(stream)));                 \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_mod##B##_inv_slice(T q, const T *a, T *out, size_t count, uint64_t *first_bad, void *stream) {                   \
        LimbConsts<T> lc;                                                                                                             \
        pfhe_status st = make_limb_consts<T>(&q, 1, nullptr, PFHE_OP_MUL, lc);                                                        \
        if (st != PFHE_OK) return st;                                                                                                 \
        if ((!a || !out) && count) return PFHE_ERR_INVALID_ARG;                                                                       \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_inv_slice<T>(lc.br[0], a, out, count, reinterpret_cast<unsigned long long *>(first_bad),                     \
                                      static_cast<cudaStream_t>(stream)));                                                            \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_basis##B##_geometry(T q, uint32_t log_basis, uint32_t levels_in, uint32_t *levels, uint32_t *drop_bits) {        \
        GadgetParams<T> g;                                                                                                            \
        if (!make_gadget<T>(q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;                                                 \
        if (levels) *levels = g.levels;                                                                                               \
        if (drop_bits) *drop_bits = g.drop_bits;                                                                                      \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_decompose##B##_batch(T q, uint32_t log_basis, uint32_t levels_in, const T *values, T *digits, size_t count,      \
                                          void *stream) {                                                                             \
        GadgetParams<T> g;                                                                                                            \
        if (!make_gadget<T>(q, log_basis, levels_in, g)) return PFHE_ERR_INVALID_ARG;                                                 \
        if ((!values || !digits) && count) return PFHE_ERR_INVALID_ARG;                                                               \
        PFHE_PTR_GUARD(digits);                                                                                                       \
        PFHE_CUDA(launch_decompose<T>(g, values, digits, count, static_cast<cudaStream_t>(stream)));                                  \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_rns##B##_lift_small_batch(const T *moduli, size_t limbs, T small_modulus, const T *small, T *out, size_t count,  \
                                               void *stream) {                                                                        \
        if (!moduli || limbs == 0) return PFHE_ERR_RNS_EMPTY;                                                                         \
        if (limbs > (size_t)kMaxLimbs || ((!small || !out) && count)) return PFHE_ERR_INVALID_ARG;                                    \
        for (size_t i = 0; i < limbs; i++)                                                                                            \
            if (moduli[i] <= small_modulus) return PFHE_ERR_INVALID_ARG;                                                              \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_rns_lift<T>(moduli, (int)limbs, small_modulus, small, out, count, static_cast<cudaStream_t>(stream)));       \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_ggsw##B##_external_product_batch(const pfhe_ntt##B *t, uint32_t k, uint32_t log_basis, uint32_t levels_in,       \
                                                      const T *key, const T *in, T *out, size_t batch, int to_coeff, void *stream) {  \
        return ext_prod<T>(t, k, log_basis, levels_in, key, in, out, batch, to_coeff, stream);                                        \
    }                                                                                                                                 \
    pfhe_status pfhe_ggsw##B##_external_product_slices(const pfhe_ntt##B *t, uint32_t k, uint32_t log_basis, uint32_t levels_in,      \
                                                       const T *key, const T *in, T *out, size_t batch, int to_coeff) {               \
        return ext_prod_host<T>(t, k, log_basis, levels_in, key, in, out, batch, to_coeff);                                           \
    }                                                                                                                                 \
    pfhe_status pfhe_blind_rotate##B##_batch(const pfhe_ntt##B *t, uint32_t log_basis, uint32_t levels_in, const T *bsk,              \
                                             uint32_t n_lwe, const uint32_t *lwe, const T *test_vector, T *acc_out, size_t batch,     \
                                             void *stream) {                                                                          \
        return blind_rot<T>(t, log_basis, levels_in, bsk, n_lwe, lwe, test_vector, acc_out, batch, stream);                           \
    }                                                                                                                                 \
    pfhe_status pfhe_extract_lwe##B##_batch(T q, const T *rlwe, T *lwe, size_t n, size_t batch, void *stream) {                       \
        if ((!rlwe || !lwe) && batch) return PFHE_ERR_INVALID_ARG;                                                                    \
        PFHE_PTR_GUARD(lwe);                                                                                                          \
        PFHE_CUDA(launch_extract_lwe<T>(q, rlwe, lwe, n, batch, 0, 1, static_cast<cudaStream_t>(stream)));                            \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_extract_lwe##B##_ex_batch(T q, const T *rlwe, T *lwe, size_t n, size_t batch, size_t index, size_t count,        \
                                               void *stream) {                                                                        \
        if (((!rlwe || !lwe) && batch) || count == 0 || index + count > n) return PFHE_ERR_INVALID_ARG;                               \
        PFHE_PTR_GUARD(lwe);                                                                                                          \
        PFHE_CUDA(launch_extract_lwe<T>(q, rlwe, lwe, n, batch, index, count, static_cast<cudaStream_t>(stream)));                    \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_rns##B##_create(const T *moduli, size_t count, pfhe_rns##B **out) { return rns_create<T>(moduli, count, out); }  \
    void pfhe_rns##B##_destroy(pfhe_rns##B *r) { delete r; }                                                                          \
    size_t pfhe_rns##B##_moduli_count(const pfhe_rns##B *r) { return r ? r->moduli.size() : 0; }                                      \
    size_t pfhe_rns##B##_big_uint_value_len(const pfhe_rns##B *r) { return r ? (size_t)r->base.value_len : 0; }                       \
    pfhe_status pfhe_rns##B##_moduli_product(const pfhe_rns##B *r, T *out) {                                                          \
        if (!r || !out) return PFHE_ERR_INVALID_ARG;                                                                                  \
        for (int i = 0; i < r->base.value_len; i++) out[i] = r->base.product[i];                                                      \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_rns##B##_compose_batch(const pfhe_rns##B *r, const T *residues, T *big, size_t count, void *stream) {            \
        if (!r || ((!residues || !big) && count)) return PFHE_ERR_INVALID_ARG;                                                        \
        PFHE_PTR_GUARD(big);                                                                                                          \
        PFHE_CUDA(launch_rns_compose<T>(r->base, residues, big, count, static_cast<cudaStream_t>(stream)));                           \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_rns##B##_decompose_batch(const pfhe_rns##B *r, const T *big, T *residues, size_t count, void *stream) {          \
        if (!r || ((!residues || !big) && count)) return PFHE_ERR_INVALID_ARG;                                                        \
        PFHE_PTR_GUARD(residues);                                                                                                     \
        PFHE_CUDA(launch_rns_decompose<T>(r->base, big, residues, count, static_cast<cudaStream_t>(stream)));                         \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_rns##B##_lift_small_scaled_add_batch(const T *moduli, size_t limbs, T small_modulus, const T *scalars,           \
                                                          const T *small, T *acc, size_t count, void *stream) {                       \
        if (!moduli || limbs == 0) return PFHE_ERR_RNS_EMPTY;                                                                         \
        if (limbs > (size_t)kMaxLimbs || !scalars || ((!small || !acc) && count)) return PFHE_ERR_INVALID_ARG;                        \
        for (size_t i = 0; i < limbs; i++)                                                                                            \
            if (moduli[i] <= small_modulus || scalars[i] >= moduli[i] || (moduli[i] >> (sizeof(T) * 8 - 1)) != 0)                     \
                return PFHE_ERR_INVALID_ARG;                                                                                          \
        PFHE_PTR_GUARD(acc);                                                                                                          \
        PFHE_CUDA(launch_rns_lift_scaled_acc<T>(moduli, (int)limbs, small_modulus, scalars, small, acc, count,                        \
                                                static_cast<cudaStream_t>(stream)));                                                  \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_bigbasis##B##_geometry(const pfhe_rns##B *r, uint32_t log_basis, uint32_t levels_in, uint32_t *levels,           \
                                            uint32_t *drop_bits) {                                                                    \
        RnsDev<T> g;                                                                                                                  \
        pfhe_status st = rns_with_gadget<T>(r, log_basis, levels_in, g);                                                              \
        if (st != PFHE_OK) return st;                                                                                                 \
        if (levels) *levels = g.levels;                                                                                               \
        if (drop_bits) *drop_bits = g.drop_bits;                                                                                      \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_rns##B##_gadget_decompose_batch(const pfhe_rns##B *r, uint32_t log_basis, uint32_t levels_in, const T *residues, \
                                                     T *digits, size_t n, size_t polys, void *stream) {                               \
        RnsDev<T> g;                                                                                                                  \
        pfhe_status st = rns_with_gadget<T>(r, log_basis, levels_in, g);                                                              \
        if (st != PFHE_OK) return st;                                                                                                 \
        if ((!residues || !digits) && n * polys) return PFHE_ERR_INVALID_ARG;                                                         \
        const size_t L = r->moduli.size();                                                                                            \
        PFHE_PTR_GUARD(digits);                                                                                                       \
        PFHE_CUDA(launch_rns_gadget<T>(g, residues, digits, n, polys, L * n, (size_t)g.levels * L * n,                                \
                                       static_cast<cudaStream_t>(stream)));                                                           \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    size_t pfhe_dcrt##B##_external_product_scratch_bytes(const pfhe_dcrt##B *t, const pfhe_rns##B *r, uint32_t k, uint32_t log_basis, \
                                                         uint32_t levels_in, size_t batch) {                                          \
        return ext_scratch_bytes<T>(t, r, k, log_basis, levels_in, batch);                                                            \
    }                                                                                                                                 \
    pfhe_status pfhe_dcrt##B##_external_product_batch(const pfhe_dcrt##B *t, const pfhe_rns##B *r, uint32_t k, uint32_t log_basis,    \
                                                      uint32_t levels_in, const T *key, const T *in, T *out, size_t batch,            \
                                                      int to_coeff, void *scratch, size_t scratch_bytes, void *stream) {              \
        return dcrt_ext_prod<T>(t, r, k, log_basis, levels_in, key, in, out, batch, to_coeff, scratch, scratch_bytes, stream);        \
    }                                                                                                                                 \
    pfhe_status pfhe_baseconv##B##_create(const T *in_moduli, size_t n_in, const T *out_moduli, size_t n_out, pfhe_baseconv##B **out) { \
        if (!out) return PFHE_ERR_INVALID_ARG;                                                                                        \
        *out = nullptr;                                                                                                               \
        if (!in_moduli || !out_moduli || n_in == 0 || n_out == 0) return PFHE_ERR_RNS_EMPTY;                                          \
        auto *h = new (std::nothrow) pfhe_baseconv##B();                                                                              \
        if (!h) return PFHE_ERR_INVALID_ARG;                                                                                          \
        const int rc = make_baseconv<T>(in_moduli, n_in, out_moduli, n_out, h->dev);                                                  \
        if (rc != 0) {                                                                                                                \
            delete h;                                                                                                                 \
            return (pfhe_status)rc;                                                                                                   \
        }                                                                                                                             \
        *out = h;                                                                                                                     \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    void pfhe_baseconv##B##_destroy(pfhe_baseconv##B *c) { delete c; }                                                                \
    pfhe_status pfhe_baseconv##B##_fast_convert_batch(const pfhe_baseconv##B *c, const T *in, T *out, size_t n, size_t polys,         \
                                                      void *stream) {                                                                 \
        if (!c || ((!in || !out) && n * polys)) return PFHE_ERR_INVALID_ARG;                                                          \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_baseconv<T>(c->dev, in, out, n, polys, false, static_cast<cudaStream_t>(stream)));                           \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_baseconv##B##_exact_convert_batch(const pfhe_baseconv##B *c, const T *in, T *out, size_t n, size_t polys,        \
                                                       void *stream) {                                                                \
        if (!c || c->dev.n_out != 1 || ((!in || !out) && n * polys)) return PFHE_ERR_INVALID_ARG;                                     \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_baseconv<T>(c->dev, in, out, n, polys, true, static_cast<cudaStream_t>(stream)));                            \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_poly##B##_mul_monomial_batch(const T *moduli, size_t limbs, const uint32_t *degrees, const T *in, T *out,        \
                                                  uint32_t log_n, size_t batch, void *stream) {                                       \
        LimbConsts<T> lc;                                                                                                             \
        pfhe_status st = limb_consts_plain<T>(moduli, limbs, lc);                                                                     \
        if (st != PFHE_OK) return st;                                                                                                 \
        if (((!degrees || !in || !out) && batch) || in == out || log_n == 0 || log_n > 20) return PFHE_ERR_INVALID_ARG;               \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_mul_monomial<T>(lc, (int)limbs, degrees, in, out, log_n, batch, static_cast<cudaStream_t>(stream)));         \
        return PFHE_OK;                                                                                                               \
    }                                                                                                                                 \
    pfhe_status pfhe_mod##B##_dot_product_batch(T q, const T *a, const T *b, T *out, size_t rows, size_t n, void *stream) {           \
        LimbConsts<T> lc;                                                                                                             \
        pfhe_status st = limb_consts_plain<T>(&q, 1, lc);                                                                             \
        if (st != PFHE_OK) return st;                                                                                                 \
        if ((!a || !b || !out) && rows) return PFHE_ERR_INVALID_ARG;                                                                  \
        PFHE_PTR_GUARD(out);                                                                                                          \
        PFHE_CUDA(launch_dot_product<T>(lc.br[0], a, b, out, rows, n, static_cast<cudaStream_t>(stream)));                            \
        return PFHE_OK;                                                                                                               \
    }

PFHE_DEFINE_WORD(32, uint32_t)
PFHE_DEFINE_WORD(64, uint64_t)

pfhe_status pfhe_multiply_factor64(uint64_t operand, uint32_t bit_shift, uint64_t modulus, uint64_t *quotient) {
    if (!quotient || modulus == 0 || operand >= modulus || (bit_shift != 32 && bit_shift != 52 && bit_shift != 64)) return PFHE_ERR_INVALID_ARG;
    const unsigned __int128 n = (unsigned __int128)operand << bit_shift;  // (op_hi, op_lo) of mul_factor/mod.rs:21-28
    *quotient = (uint64_t)(n / modulus);
    return PFHE_OK;
}
pfhe_status pfhe_multiply_factor64_mul(uint64_t operand, uint64_t quotient, uint32_t bit_shift, uint64_t b, uint64_t modulus, uint64_t *out) {
    if (!out || (bit_shift != 32 && bit_shift != 52 && bit_shift != 64)) return PFHE_ERR_INVALID_ARG;
    const uint64_t hw = bit_shift == 32 ? (quotient * b) >> 32 : (uint64_t)(((unsigned __int128)quotient * b) >> bit_shift);
    const uint64_t r = operand * b - modulus * hw;  // lazy_mul_modulo, wrapping
    const uint64_t r2 = r - modulus;
    *out = r < r2 ? r : r2;                          // r.min(r.wrapping_sub(modulus))
    return PFHE_OK;
}
pfhe_status pfhe_device_count(int *count) {
    if (!count) return PFHE_ERR_INVALID_ARG;
    PFHE_CUDA(cudaGetDeviceCount(count));
    return PFHE_OK;
}
pfhe_status pfhe_malloc(int device, size_t bytes, void **dev_ptr) {
    if (!dev_ptr) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(cudaMalloc(dev_ptr, bytes));
    return PFHE_OK;
}
pfhe_status pfhe_free(int device, void *dev_ptr) {
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(cudaFree(dev_ptr));
    return PFHE_OK;
}
pfhe_status pfhe_memcpy_h2d(int device, void *dev_dst, const void *host_src, size_t bytes, void *stream) {
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}
pfhe_status pfhe_memcpy_d2h(int device, void *host_dst, const void *dev_src, size_t bytes, void *stream) {
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}
pfhe_status pfhe_stream_synchronize(int device, void *stream) {
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return PFHE_OK;
}
pfhe_status pfhe_modmul_microbench(int device, int kind, uint32_t blocks, uint32_t iters, float *ms) {
    if (!ms) return PFHE_ERR_INVALID_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) return PFHE_ERR_CUDA;
    PFHE_CUDA(run_modmul_microbench(kind, blocks, iters, ms));
    return PFHE_OK;
}

}  // extern "C"
