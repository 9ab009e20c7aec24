//! Device-resident bootstrapping key and the host-slice bootstrap (blind rotation composed from the reference's primitives,
//! SURVEY.md App. A.6: `mul_monomial_assign` crates/primus_poly/src/poly/mul.rs:74-99, `mul_dcrt_ggsw_to`
//! crates/primus_lattice/src/glwe/crt.rs:200-227, `extract_lwe` crates/primus_lattice/src/rlwe/coeff.rs:264-288).
use crate::{check, ntt::CudaU32NttTable, sys::*};

/// `n_lwe` RGSW ciphertexts in NTT form (`[n_lwe][2][levels][2][N]`, the NttRgsw layout), uploaded once.
pub struct BootstrappingKey<'t> {
    pub(crate) h: *mut pfhe_bsk32,
    pub(crate) table: &'t CudaU32NttTable,
}
unsafe impl Send for BootstrappingKey<'_> {}
unsafe impl Sync for BootstrappingKey<'_> {}
impl Drop for BootstrappingKey<'_> {
    fn drop(&mut self) {
        unsafe { pfhe_bsk32_destroy(self.h) }
    }
}
impl<'t> BootstrappingKey<'t> {
    /// `levels = 0` selects the full decomposition length of `ApproxSignedBasis::new(q, log_basis, None)`.
    pub fn new(table: &'t CudaU32NttTable, log_basis: u32, levels: u32, n_lwe: u32, key: &[u32]) -> Self {
        let mut h = core::ptr::null_mut();
        check(unsafe { pfhe_bsk32_create(table.raw(), log_basis, levels, n_lwe, key.as_ptr(), &mut h) }, "pfhe_bsk32_create");
        Self { h, table }
    }
    /// From the serialised form (`to_bytes()` of the key container: raw little-endian words, macros/mod.rs:74-80).
    pub fn from_bytes(table: &'t CudaU32NttTable, log_basis: u32, levels: u32, n_lwe: u32, bytes: &[u8]) -> Self {
        let mut h = core::ptr::null_mut();
        check(unsafe { pfhe_bsk32_create_from_bytes(table.raw(), log_basis, levels, n_lwe, bytes.as_ptr(), bytes.len(), &mut h) },
              "pfhe_bsk32_create_from_bytes");
        Self { h, table }
    }
    pub fn lwe_dimension(&self) -> usize {
        unsafe { pfhe_bsk32_lwe_dimension(self.h) as usize }
    }
}

/// Bootstraps a batch of LWE samples already switched to `Z_2N` (`lwe`: `[batch][n_lwe + 1]`); returns `[batch][N + 1]`
/// extracted LWE samples (`Rlwe::extract_lwe` of the rotated accumulator).
pub fn bootstrap(key: &BootstrappingKey<'_>, lwe: &[u32], test_vector: &[u32]) -> Vec<u32> {
    use primus_ntt::NttTable;
    let n = key.table.poly_length();
    let per = key.lwe_dimension() + 1;
    assert_eq!(lwe.len() % per, 0);
    assert_eq!(test_vector.len(), n);
    let batch = lwe.len() / per;
    let mut out = vec![0u32; batch * (n + 1)];
    check(unsafe { pfhe_bootstrap32_slices(key.table.raw(), key.h, lwe.as_ptr(), test_vector.as_ptr(), out.as_mut_ptr(), batch, 1) },
          "pfhe_bootstrap32_slices");
    out
}
