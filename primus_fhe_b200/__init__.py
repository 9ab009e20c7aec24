"""primus_fhe_b200 -- B200-native (sm_100a) implementation of primus-fhe's polynomial-ring hot path.

The package holds only what the path needs: `csrc/` (CUDA kernels + the C-ABI of include/pfhe.h),
`build.py` (in-tree nvcc build) and `api.py` (host-side mirror of the reference's trait surface).
"""
from ._cabi import PfheError, LIB_PATH, declared_symbols, launch_count  # noqa: F401
from .api import (  # noqa: F401
    ApproxSignedBasis, BarrettModulus, BaseConverter, BigUintApproxSignedBasis, MultiplyFactor, RNSBase, U32DcrtTable, U32NttTable, U64DcrtTable, U64NttTable,
    BootstrappingKey, MultiNttTable, UintNttTable, cipher_words, dcrt_into_coeff_form, dcrt_into_ntt_form, from_bytes, into_coeff_form, into_ntt_form,
    modulus_switch_batch, to_bytes, write_coeff_form, write_ntt_form, registered_host_buffer, host_is_pageable,
    butterfly_mul_factor_batch, dcrt_external_product_batch, device_count, dot_product_batch, extract_lwe_batch, extract_lwe_ex_batch, inv_slice_batch, modmul_microbench, mul_monomial_batch, slice_op_bcast,
)
