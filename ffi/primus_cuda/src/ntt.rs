//! `impl NttTable` for the CUDA tables (drop-in for `U64NttTable` / `U32NttTable`,
//! crates/primus_ntt/src/ntt/prime64/table.rs:41, prime32/table.rs:37).
use primus_data::{DataMut, RawData};
use primus_ntt::{NttError, NttTable};
use primus_poly::{NttPolynomial, Polynomial};
use primus_reduce::FieldContext;

use crate::{check, sys::*};

macro_rules! cuda_ntt_table {
    ($name:ident, $t:ty, $h:ident, $max_bits:expr, $create:ident, $destroy:ident, $fwd:ident, $inv:ident, $fwds:ident, $invs:ident,
     $mono:ident, $one:ident, $minus:ident) => {
        /// Device-resident table; immutable after `new`, so it is `Send + Sync` like the reference's tables (ntt/mod.rs:16).
        pub struct $name {
            pub(crate) h: *mut $h,
            n: usize,
        }
        unsafe impl Send for $name {}
        unsafe impl Sync for $name {}
        impl Drop for $name {
            fn drop(&mut self) {
                unsafe { $destroy(self.h) }
            }
        }
        impl $name {
            /// `new` on a chosen device (the trait constructor uses device 0).
            pub fn new_on(device: i32, log_n: u32, q: $t) -> Result<Self, NttError<$t>> {
                let mut h = core::ptr::null_mut();
                match unsafe { $create(device, log_n, q, &mut h) } {
                    0 => Ok(Self { h, n: 1usize << log_n }),
                    1 => Err(NttError::NoPrimitiveRoot { degree: (1 as $t) << (log_n + 1), modulus: q }), // root.rs:72-81
                    2 => Err(NttError::DegreeConversionErr { degree: 1usize << log_n, modulus: q }),
                    3 => Err(NttError::DegreeTooLarge { degree: 1usize << log_n, modulus: q }),
                    5 => Err(NttError::ModulusTooLarge { modulus: q, max_bits: $max_bits }),              // table.rs:318-323 / :195
                    _ => Err(NttError::NttTableErr),                                                      // incl. CUDA failures
                }
            }
            /// Raw handle for the device-batch entry points of [`crate::sys`].
            pub fn raw(&self) -> *const $h {
                self.h
            }
            /// Every polynomial of a flat ciphertext storage in ONE pipelined call: what `into_ntt_form` loops over
            /// (crates/primus_lattice/src/macros/mod.rs:551-553).
            pub fn transform_slices(&self, polys: &mut [$t]) {
                assert_eq!(polys.len() % self.n, 0);
                check(unsafe { $fwds(self.h, polys.as_mut_ptr(), polys.len() / self.n, 0) }, stringify!($fwds));
            }
            pub fn inverse_transform_slices(&self, polys: &mut [$t]) {
                assert_eq!(polys.len() % self.n, 0);
                check(unsafe { $invs(self.h, polys.as_mut_ptr(), polys.len() / self.n, 0) }, stringify!($invs));
            }
        }
        impl NttTable for $name {
            type ValueT = $t;
            fn new<M: FieldContext<$t>>(log_n: u32, modulus: M) -> Result<Self, NttError<$t>> {
                let q = modulus.value().ok_or(NttError::NttTableErr)?;
                Self::new_on(0, log_n, q)
            }
            #[inline]
            fn poly_length(&self) -> usize {
                self.n
            }
            fn transform_slice(&self, poly: &mut [$t]) {
                debug_assert_eq!(poly.len(), self.n); // prime64/table.rs:549
                check(unsafe { $fwd(self.h, poly.as_mut_ptr(), 0) }, stringify!($fwd));
            }
            fn lazy_transform_slice(&self, poly: &mut [$t]) {
                debug_assert_eq!(poly.len(), self.n); // inputs in [0, 4q): canonicalised on the device, outputs canonical
                check(unsafe { $fwd(self.h, poly.as_mut_ptr(), 1) }, stringify!($fwd));
            }
            fn inverse_transform_slice(&self, values: &mut [$t]) {
                debug_assert_eq!(values.len(), self.n);
                check(unsafe { $inv(self.h, values.as_mut_ptr(), 0) }, stringify!($inv));
            }
            fn lazy_inverse_transform_slice(&self, values: &mut [$t]) {
                debug_assert_eq!(values.len(), self.n);
                check(unsafe { $inv(self.h, values.as_mut_ptr(), 1) }, stringify!($inv));
            }
            fn transform_inplace<S: RawData<Elem = $t> + DataMut>(&self, mut poly: Polynomial<S>) -> NttPolynomial<S> {
                self.transform_slice(poly.as_mut_slice()); // prime64/table.rs:524-531
                NttPolynomial::new(poly.0)
            }
            fn inverse_transform_inplace<S: RawData<Elem = $t> + DataMut>(&self, mut values: NttPolynomial<S>) -> Polynomial<S> {
                self.inverse_transform_slice(values.as_mut_slice());
                Polynomial::new(values.0)
            }
            fn transform_monomial(&self, coeff: $t, degree: usize, values: &mut [$t]) {
                debug_assert_eq!(values.len(), self.n);
                check(unsafe { $mono(self.h, coeff, degree, values.as_mut_ptr()) }, stringify!($mono));
            }
            fn transform_coeff_one_monomial(&self, degree: usize, values: &mut [$t]) {
                check(unsafe { $one(self.h, degree, values.as_mut_ptr()) }, stringify!($one));
            }
            fn transform_coeff_minus_one_monomial(&self, degree: usize, values: &mut [$t]) {
                check(unsafe { $minus(self.h, degree, values.as_mut_ptr()) }, stringify!($minus));
            }
        }
    };
}

cuda_ntt_table!(CudaU64NttTable, u64, pfhe_ntt64, 62, pfhe_ntt64_create, pfhe_ntt64_destroy, pfhe_ntt64_transform_slice,
                pfhe_ntt64_inverse_transform_slice, pfhe_ntt64_transform_slices, pfhe_ntt64_inverse_transform_slices,
                pfhe_ntt64_transform_monomial, pfhe_ntt64_transform_coeff_one_monomial, pfhe_ntt64_transform_coeff_minus_one_monomial);
cuda_ntt_table!(CudaU32NttTable, u32, pfhe_ntt32, 30, pfhe_ntt32_create, pfhe_ntt32_destroy, pfhe_ntt32_transform_slice,
                pfhe_ntt32_inverse_transform_slice, pfhe_ntt32_transform_slices, pfhe_ntt32_inverse_transform_slices,
                pfhe_ntt32_transform_monomial, pfhe_ntt32_transform_coeff_one_monomial, pfhe_ntt32_transform_coeff_minus_one_monomial);
