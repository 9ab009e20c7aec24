//! `primus_cuda`: drop-in CUDA (B200, sm_100a) implementations of primus-fhe's plug-in traits for the polynomial-ring hot path.
//!
//! * [`CudaU64NttTable`] / [`CudaU32NttTable`] implement `primus_ntt::NttTable` (crates/primus_ntt/src/ntt/mod.rs:16-113);
//! * [`CudaU64DcrtTable`] / [`CudaU32DcrtTable`] implement `primus_ntt::DcrtTable` (crates/primus_ntt/src/dcrt/mod.rs:19-135);
//! * [`BootstrappingKey`] + [`bootstrap`] keep a bootstrapping key on the device and run blind rotation + sample extraction;
//! * [`multi`] fans host slices out over several GPUs (the traits are `Send + Sync`; no collective on the data path);
//! * [`Pinned`] page-locks a caller-owned buffer in place so the host-slice calls run at the full PCIe rate.
//!
//! Every method is a 1:1 forwarder to a C symbol of `include/pfhe.h` (raw bindings in [`sys`], generated from the header).
//! There is no CPU fallback: a failing CUDA call panics, exactly like an infallible trait method has to.
pub mod bootstrap;
pub mod dcrt;
pub mod multi;
pub mod ntt;
pub mod pinned;
pub mod sys;

pub use bootstrap::{bootstrap, BootstrappingKey};
pub use dcrt::{CudaU32DcrtTable, CudaU64DcrtTable};
pub use ntt::{CudaU32NttTable, CudaU64NttTable};
pub use pinned::Pinned;

/// Panics with the library's message unless `rc == 0` (hot-path trait methods are infallible in the reference).
#[inline]
pub(crate) fn check(rc: sys::pfhe_status, what: &str) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::pfhe_status_string(rc)) }.to_string_lossy().into_owned();
        let cuda = unsafe { std::ffi::CStr::from_ptr(sys::pfhe_last_cuda_error()) }.to_string_lossy().into_owned();
        panic!("{what} failed: {msg} {cuda}");
    }
}
