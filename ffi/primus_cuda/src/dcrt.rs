//! `impl DcrtTable` for the multi-limb CUDA tables (drop-in for `U64DcrtTable` / `U32DcrtTable`,
//! crates/primus_ntt/src/dcrt/prime64.rs:11-128): L independent transforms over limb-major `[L][N]` storage in one launch.
use primus_ntt::{DcrtTable, NttError};
use primus_reduce::FieldContext;

use crate::{check, sys::*};

macro_rules! cuda_dcrt_table {
    ($name:ident, $t:ty, $h:ident, $create:ident, $destroy:ident, $fwds:ident, $invs:ident) => {
        pub struct $name {
            pub(crate) h: *mut $h,
            n: usize,
            limbs: usize,
        }
        unsafe impl Send for $name {}
        unsafe impl Sync for $name {}
        impl Drop for $name {
            fn drop(&mut self) {
                unsafe { $destroy(self.h) }
            }
        }
        impl $name {
            pub fn new_on(device: i32, log_n: u32, moduli: &[$t]) -> Result<Self, NttError<$t>> {
                let mut h = core::ptr::null_mut();
                match unsafe { $create(device, log_n, moduli.as_ptr(), moduli.len(), &mut h) } {
                    0 => Ok(Self { h, n: 1usize << log_n, limbs: moduli.len() }),
                    1 => Err(NttError::NoPrimitiveRoot { degree: (1 as $t) << (log_n + 1), modulus: moduli[0] }),
                    3 => Err(NttError::DegreeTooLarge { degree: 1usize << log_n, modulus: moduli[0] }),
                    5 => Err(NttError::ModulusTooLarge { modulus: moduli[0], max_bits: (<$t>::BITS - 2) }),
                    _ => Err(NttError::NttTableErr),
                }
            }
            pub fn raw(&self) -> *const $h {
                self.h
            }
        }
        impl DcrtTable for $name {
            type ValueT = $t;
            fn new<M: FieldContext<$t>>(log_n: u32, moduli: &[M]) -> Result<Self, NttError<$t>> {
                let qs: Vec<$t> = moduli.iter().map(|m| m.value().ok_or(NttError::NttTableErr)).collect::<Result<_, _>>()?;
                Self::new_on(0, log_n, &qs)
            }
            #[inline]
            fn poly_length(&self) -> usize {
                self.n
            }
            #[inline]
            fn moduli_count(&self) -> usize {
                self.limbs
            }
            #[inline]
            fn crt_poly_length(&self) -> usize {
                self.n * self.limbs
            }
            fn transform_slice(&self, poly: &mut [$t]) {
                debug_assert_eq!(poly.len(), self.n * self.limbs); // dcrt/prime64.rs:106-111
                check(unsafe { $fwds(self.h, poly.as_mut_ptr(), 1, 0) }, stringify!($fwds));
            }
            fn lazy_transform_slice(&self, poly: &mut [$t]) {
                check(unsafe { $fwds(self.h, poly.as_mut_ptr(), 1, 1) }, stringify!($fwds));
            }
            fn inverse_transform_slice(&self, values: &mut [$t]) {
                debug_assert_eq!(values.len(), self.n * self.limbs);
                check(unsafe { $invs(self.h, values.as_mut_ptr(), 1, 0) }, stringify!($invs));
            }
            fn lazy_inverse_transform_slice(&self, values: &mut [$t]) {
                check(unsafe { $invs(self.h, values.as_mut_ptr(), 1, 1) }, stringify!($invs));
            }
        }
    };
}

cuda_dcrt_table!(CudaU64DcrtTable, u64, pfhe_dcrt64, pfhe_dcrt64_create, pfhe_dcrt64_destroy, pfhe_dcrt64_transform_slices,
                 pfhe_dcrt64_inverse_transform_slices);
cuda_dcrt_table!(CudaU32DcrtTable, u32, pfhe_dcrt32, pfhe_dcrt32_create, pfhe_dcrt32_destroy, pfhe_dcrt32_transform_slices,
                 pfhe_dcrt32_inverse_transform_slices);
