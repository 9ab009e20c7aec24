//! Several GPUs from one process: the same table replicated per device, host slices split into contiguous shards, one host
//! thread + stream set per device inside the library, no collective on the data path (pfhe_multi_*).
use primus_ntt::NttError;

use crate::{check, ntt::CudaU64NttTable, sys::*};

pub struct MultiU64NttTable {
    tables: Vec<CudaU64NttTable>,
    raw: Vec<*const pfhe_ntt64>,
    n: usize,
}
unsafe impl Send for MultiU64NttTable {}
unsafe impl Sync for MultiU64NttTable {}

impl MultiU64NttTable {
    pub fn new(devices: &[i32], log_n: u32, q: u64) -> Result<Self, NttError<u64>> {
        let tables = devices.iter().map(|&d| CudaU64NttTable::new_on(d, log_n, q)).collect::<Result<Vec<_>, _>>()?;
        let raw = tables.iter().map(|t| t.raw()).collect();
        Ok(Self { tables, raw, n: 1usize << log_n })
    }
    pub fn device_count(&self) -> usize {
        self.tables.len()
    }
    /// `[batch][N]` host polynomials, transformed in place; shard r of `batch` starts at `r*floor(batch/n) + min(r, batch mod n)`.
    pub fn transform_slices(&self, polys: &mut [u64], inverse: bool) {
        assert_eq!(polys.len() % self.n, 0);
        check(unsafe { pfhe_multi_ntt64_transform_slices(self.raw.as_ptr(), self.raw.len(), polys.as_mut_ptr(), polys.len() / self.n,
                                                         inverse as i32, 0) }, "pfhe_multi_ntt64_transform_slices");
    }
    pub fn polymul_slices(&self, a: &[u64], b: &[u64], c: &mut [u64]) {
        assert!(a.len() == b.len() && a.len() == c.len() && a.len() % self.n == 0);
        check(unsafe { pfhe_multi_ntt64_polymul_slices(self.raw.as_ptr(), self.raw.len(), a.as_ptr(), b.as_ptr(), c.as_mut_ptr(),
                                                       a.len() / self.n) }, "pfhe_multi_ntt64_polymul_slices");
    }
}
