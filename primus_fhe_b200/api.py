"""Host-side mirror of the reference's operator surface for the polynomial-ring hot path.

Python stands in for the Rust FFI crate here (no Rust toolchain in this image):
every class forwards 1:1 to the C-ABI in include/pfhe.h, exactly as `impl NttTable for CudaU64NttTable`
would.  Names, argument meaning and error behaviour follow the reference:

  U32NttTable / U64NttTable     trait NttTable            primus_ntt/src/ntt/mod.rs:16-113
  U32DcrtTable / U64DcrtTable   trait DcrtTable           primus_ntt/src/dcrt/mod.rs:19-135
  BarrettModulus (slice ops)    ReduceMulSlice & co       primus_reduce/src/slice_ops.rs:63-230
  ShoupFactor (slice ops)       FactorSliceOps            primus_factor/src/ops.rs:58-118
  ApproxSignedBasis             decompose_slice_to        primus_decompose/src/primitive/basis.rs:12-407
  RNSBase.wrapping_decompose_small_values_to              primus_rns/src/base.rs:279-315
  external_product / blind_rotate                         primus_lattice/src/glwe/crt.rs:200-227, SURVEY App. A.6

Host-slice methods take numpy arrays / CPU torch tensors (in place, like `&mut [T]`); `*_batch`
methods take CUDA torch tensors and run on torch's current stream.  PyTorch is plumbing only
(device memory, streams); all arithmetic happens in libpfhe_cuda.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import PfheError, check, lib

OP_MUL, OP_ADD_MUL, OP_SUB_MUL, OP_MUL_ADD, OP_ADD, OP_SUB, OP_NEG = range(7)
OP_MUL_SCALAR, OP_ADD_MUL_SCALAR, OP_FACTOR_MUL, OP_ADD_FACTOR_MUL, OP_SUB_FACTOR_MUL = range(7, 12)
OP_REDUCE_LAZY, OP_DOUBLE, OP_MUL_SCALAR_ADD, OP_FACTOR_MUL_ADD = range(12, 16)


def _ct(bits):
    return C.c_uint32 if bits == 32 else C.c_uint64


def _np_dtype(bits):
    return np.uint32 if bits == 32 else np.uint64


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _host_ptr(x, bits, n_words=None):
    """Pointer to a contiguous HOST buffer of `bits`-wide words (numpy array or CPU torch tensor)."""
    if _is_torch(x):
        if x.is_cuda:
            raise TypeError("host-slice methods take host memory; use the *_batch methods for CUDA tensors")
        if not x.is_contiguous() or x.element_size() * 8 != bits:
            raise TypeError("expected a contiguous tensor of %d-bit words" % bits)
        if n_words is not None and x.numel() != n_words:
            raise ValueError(f"length mismatch: {x.numel()} != {n_words}")
        return C.c_void_p(x.data_ptr())
    if not isinstance(x, np.ndarray) or x.dtype.itemsize * 8 != bits or not x.flags.c_contiguous:
        raise TypeError("expected a C-contiguous numpy array of %d-bit words" % bits)
    if n_words is not None and x.size != n_words:
        raise ValueError(f"length mismatch: {x.size} != {n_words}")
    return x.ctypes.data_as(C.c_void_p)


def _dev_ptr(x, bits, n_words=None):
    if not _is_torch(x) or not x.is_cuda:
        raise TypeError("*_batch methods take CUDA torch tensors")
    if not x.is_contiguous() or x.element_size() * 8 != bits:
        raise TypeError("expected a contiguous CUDA tensor of %d-bit words" % bits)
    if n_words is not None and x.numel() != n_words:
        raise ValueError(f"length mismatch: {x.numel()} != {n_words}")
    return C.c_void_p(x.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_count() -> int:
    n = C.c_int(0)
    check(lib().pfhe_device_count(C.byref(n)))
    return n.value


class _NttTable:
    """NttTable (primus_ntt/src/ntt/mod.rs:16-113) backed by device-resident tables."""
    bits = 64

    def __init__(self, log_n: int, modulus: int, device: int = 0):
        self._h = C.c_void_p()
        self._p = f"pfhe_ntt{self.bits}_"
        f = getattr(lib(), self._p + "create")
        f.argtypes = [C.c_int, C.c_uint32, _ct(self.bits), C.c_void_p]
        check(f(device, log_n, int(modulus), C.byref(self._h)))
        self.log_n, self.n, self.q, self.device = log_n, 1 << log_n, int(modulus), device

    @classmethod
    def new(cls, log_n, modulus, device=0):
        return cls(log_n, modulus, device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                f = getattr(lib(), self._p + "destroy"); f.argtypes = [C.c_void_p]; f.restype = None
                f(h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def _scalar(self, name):
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p]; f.restype = _ct(self.bits)
        return int(f(self._h))

    def poly_length(self) -> int:
        f = getattr(lib(), self._p + "poly_length"); f.argtypes = [C.c_void_p]; f.restype = C.c_size_t
        return int(f(self._h))

    def modulus(self): return self._scalar("modulus")
    def root(self): return self._scalar("root")
    def inv_root(self): return self._scalar("inv_root")
    def inv_n(self): return self._scalar("inv_n")

    # ---- host-slice trait methods (in place) ----
    def _host1(self, name, poly, lazy):
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        check(f(self._h, _host_ptr(poly, self.bits, self.n), int(lazy)))

    def transform_slice(self, poly): self._host1("transform_slice", poly, 0)
    def lazy_transform_slice(self, poly): self._host1("transform_slice", poly, 1)
    def inverse_transform_slice(self, values): self._host1("inverse_transform_slice", values, 0)
    def lazy_inverse_transform_slice(self, values): self._host1("inverse_transform_slice", values, 1)

    def transform_inplace(self, poly):
        self.transform_slice(poly); return poly

    def inverse_transform_inplace(self, values):
        self.inverse_transform_slice(values); return values

    def _hostn(self, name, polys, lazy=0):
        size = polys.numel() if _is_torch(polys) else polys.size
        if size % self.n:
            raise ValueError("batch buffer is not a multiple of the polynomial length")
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        check(f(self._h, _host_ptr(polys, self.bits), size // self.n, lazy))

    def transform_slices(self, polys): self._hostn("transform_slices", polys)
    def inverse_transform_slices(self, polys): self._hostn("inverse_transform_slices", polys)

    def transform_monomial(self, coeff, degree, values=None):
        values = np.empty(self.n, dtype=_np_dtype(self.bits)) if values is None else values
        f = getattr(lib(), self._p + "transform_monomial")
        f.argtypes = [C.c_void_p, _ct(self.bits), C.c_size_t, C.c_void_p]
        check(f(self._h, int(coeff), int(degree), _host_ptr(values, self.bits, self.n)))
        return values

    def transform_coeff_one_monomial(self, degree, values=None):
        values = np.empty(self.n, dtype=_np_dtype(self.bits)) if values is None else values
        f = getattr(lib(), self._p + "transform_coeff_one_monomial"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._h, int(degree), _host_ptr(values, self.bits, self.n))); return values

    def transform_coeff_minus_one_monomial(self, degree, values=None):
        values = np.empty(self.n, dtype=_np_dtype(self.bits)) if values is None else values
        f = getattr(lib(), self._p + "transform_coeff_minus_one_monomial"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._h, int(degree), _host_ptr(values, self.bits, self.n))); return values

    def polymul_slices(self, a, b, c=None):
        if c is None:
            c = np.empty_like(a)
        size = a.numel() if _is_torch(a) else a.size
        f = getattr(lib(), self._p + "polymul_slices"); f.argtypes = [C.c_void_p] * 4 + [C.c_size_t]
        check(f(self._h, _host_ptr(a, self.bits), _host_ptr(b, self.bits, size), _host_ptr(c, self.bits, size), size // self.n))
        return c

    # ---- device batch API (CUDA torch tensors, current stream) ----
    def _batch(self, name, dev):
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(dev, self.bits), dev.numel() // self.n, _stream()))

    def forward_batch(self, dev): self._batch("forward_batch", dev)
    def inverse_batch(self, dev): self._batch("inverse_batch", dev)

    def forward_batch_to(self, src, dst):
        f = getattr(lib(), self._p + "forward_batch_to"); f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(src, self.bits), _dev_ptr(dst, self.bits, src.numel()), src.numel() // self.n, _stream()))

    def inverse_batch_to(self, src, dst):
        f = getattr(lib(), self._p + "inverse_batch_to"); f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(src, self.bits), _dev_ptr(dst, self.bits, src.numel()), src.numel() // self.n, _stream()))

    def polymul_batch(self, a, b, c):
        f = getattr(lib(), self._p + "polymul_batch"); f.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(a, self.bits), _dev_ptr(b, self.bits, a.numel()), _dev_ptr(c, self.bits, a.numel()),
                a.numel() // self.n, _stream()))

    def monomial_batch(self, coeff, degrees, out):
        f = getattr(lib(), self._p + "monomial_batch")
        f.argtypes = [C.c_void_p, _ct(self.bits), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._h, int(coeff), _dev_ptr(degrees, 32), _dev_ptr(out, self.bits, degrees.numel() * self.n),
                degrees.numel(), _stream()))

    # ---- lattice ops on this table's ring (L = 1) ----
    def external_product_batch(self, k, log_basis, levels, key, glwe_in, out, to_coeff=True):
        """GGSW external product, fused (primus_lattice/src/glwe/crt.rs:200-227 [+ into_coeff_form])."""
        f = getattr(lib(), f"pfhe_ggsw{self.bits}_external_product_batch")
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        batch = glwe_in.numel() // ((k + 1) * self.n)
        check(f(self._h, k, log_basis, levels or 0, _dev_ptr(key, self.bits), _dev_ptr(glwe_in, self.bits),
                _dev_ptr(out, self.bits, glwe_in.numel()), batch, int(to_coeff), _stream()))

    def external_product_slices(self, k, log_basis, levels, key, glwe_in, out, to_coeff=True):
        """Host-slice form of external_product_batch (numpy / CPU tensors)."""
        f = getattr(lib(), f"pfhe_ggsw{self.bits}_external_product_slices")
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        size = glwe_in.numel() if _is_torch(glwe_in) else glwe_in.size
        batch = size // ((k + 1) * self.n)
        check(f(self._h, k, log_basis, levels or 0, _host_ptr(key, self.bits), _host_ptr(glwe_in, self.bits), _host_ptr(out, self.bits, size),
                batch, int(bool(to_coeff))))
        return out

    def blind_rotate_batch(self, log_basis, levels, bsk, n_lwe, lwe, test_vector, acc_out):
        """Composed blind rotation."""
        f = getattr(lib(), f"pfhe_blind_rotate{self.bits}_batch")
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        batch = lwe.numel() // (n_lwe + 1)
        check(f(self._h, log_basis, levels or 0, _dev_ptr(bsk, self.bits), n_lwe, _dev_ptr(lwe, 32),
                _dev_ptr(test_vector, self.bits, self.n), _dev_ptr(acc_out, self.bits, batch * 2 * self.n), batch, _stream()))


    def blind_rotate_ternary_batch(self, log_basis, levels, bsk_plus, bsk_minus, n_lwe, lwe, test_vector, acc_out):
        """Ternary-secret blind rotation by monomial combination (pfhe_blind_rotate_ternary32_batch; u32, N = 1024)."""
        if self.bits != 32:
            raise PfheError(10, "ternary blind rotation is built for u32 words")
        f = lib().pfhe_blind_rotate_ternary32_batch
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        batch = lwe.numel() // (n_lwe + 1)
        check(f(self._h, log_basis, levels or 0, _dev_ptr(bsk_plus, 32), _dev_ptr(bsk_minus, 32, bsk_plus.numel()), n_lwe, _dev_ptr(lwe, 32),
                _dev_ptr(test_vector, 32, self.n), _dev_ptr(acc_out, 32, batch * 2 * self.n), batch, _stream()))


class U64NttTable(_NttTable):
    bits = 64


class U32NttTable(_NttTable):
    bits = 32


class _DcrtTable:
    """DcrtTable (primus_ntt/src/dcrt/mod.rs:19-135): limb-major [L][N]."""
    bits = 64

    def __init__(self, log_n: int, moduli, device: int = 0):
        self._h = C.c_void_p()
        self._p = f"pfhe_dcrt{self.bits}_"
        self.moduli = [int(m) for m in moduli]
        arr = (_ct(self.bits) * max(1, len(self.moduli)))(*self.moduli)
        f = getattr(lib(), self._p + "create")
        f.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(device, log_n, arr, len(self.moduli), C.byref(self._h)))
        self.log_n, self.n = log_n, 1 << log_n

    @classmethod
    def new(cls, log_n, moduli, device=0):
        return cls(log_n, moduli, device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                f = getattr(lib(), self._p + "destroy"); f.argtypes = [C.c_void_p]; f.restype = None
                f(h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def _size(self, name):
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p]; f.restype = C.c_size_t
        return int(f(self._h))

    def poly_length(self): return self._size("poly_length")
    def moduli_count(self): return self._size("moduli_count")
    def crt_poly_length(self): return self._size("crt_poly_length")

    def _hostn(self, name, polys, lazy=0):
        size = polys.numel() if _is_torch(polys) else polys.size
        unit = self.n * len(self.moduli)
        if size % unit:
            raise ValueError("buffer is not a multiple of the CRT polynomial length")
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        check(f(self._h, _host_ptr(polys, self.bits), size // unit, int(lazy)))

    def transform_slice(self, poly): self._hostn("transform_slices", poly)
    def inverse_transform_slice(self, poly): self._hostn("inverse_transform_slices", poly)
    # lazy trait contract (primus_ntt/src/dcrt/mod.rs:77-103): inputs in [0,4q_i) / [0,2q_i); outputs canonical (congruent, in range)
    def lazy_transform_slice(self, poly): self._hostn("transform_slices", poly, 1)
    def lazy_inverse_transform_slice(self, poly): self._hostn("inverse_transform_slices", poly, 1)

    def _batch(self, name, dev):
        unit = self.n * len(self.moduli)
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(dev, self.bits), dev.numel() // unit, _stream()))

    def forward_batch(self, dev): self._batch("forward_batch", dev)
    def inverse_batch(self, dev): self._batch("inverse_batch", dev)

    def polymul_batch(self, a, b, c):
        unit = self.n * len(self.moduli)
        f = getattr(lib(), self._p + "polymul_batch"); f.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(a, self.bits), _dev_ptr(b, self.bits, a.numel()), _dev_ptr(c, self.bits, a.numel()),
                a.numel() // unit, _stream()))


class U64DcrtTable(_DcrtTable):
    bits = 64


class U32DcrtTable(_DcrtTable):
    bits = 32


def _slice_op(bits, op, moduli, scalars, a, b, c, out, rows, n, host):
    L = len(moduli)
    m = (_ct(bits) * L)(*[int(x) for x in moduli])
    s = (_ct(bits) * L)(*[int(x) for x in scalars]) if scalars is not None else None
    ptr = (lambda x: _host_ptr(x, bits)) if host else (lambda x: _dev_ptr(x, bits))
    args = [op, m, L, s, ptr(a), ptr(b) if b is not None else None, ptr(c) if c is not None else None, ptr(out), rows, n]
    f = getattr(lib(), f"pfhe_mod{bits}_slice_op" + ("_host" if host else ""))
    f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p] + [C.c_void_p] * 4 + [C.c_size_t, C.c_size_t] + ([] if host else [C.c_void_p])
    if not host:
        args.append(_stream())
    check(f(*args))


class BarrettModulus:
    """Slice operators of a BarrettModulus (primus_modulus/src/barrett/slice.rs:185-295) on device or host slices.
    With several moduli it is the per-limb DcrtPolynomial form (slices are [rows][L][n])."""

    def __init__(self, value, bits=64, n=None):
        self.moduli = [int(v) for v in (value if isinstance(value, (list, tuple)) else [value])]
        self.bits, self.n = bits, n
        for q in self.moduli:
            if q <= 1:
                raise ValueError("modulus can't be 0 or 1.")
            if q >> (bits - 2):
                raise ValueError("modulus is too large.")

    def value(self): return self.moduli[0]

    def _run(self, op, a, b=None, c=None, out=None, scalars=None):
        host = not (_is_torch(a) and a.is_cuda)
        size = a.numel() if _is_torch(a) else a.size
        L = len(self.moduli)
        n = self.n if self.n else size // L
        rows = size // (L * n)
        _slice_op(self.bits, op, self.moduli, scalars, a, b, c, out, rows, n, host)
        return out

    def reduce_mul_slice_to(self, a, b, out): return self._run(OP_MUL, a, b, None, out)
    def reduce_mul_slice_assign(self, a, b): return self._run(OP_MUL, a, b, None, a)
    def reduce_add_mul_slice_assign(self, acc, a, b): return self._run(OP_ADD_MUL, a, b, None, acc)
    def reduce_sub_mul_slice_assign(self, acc, a, b): return self._run(OP_SUB_MUL, a, b, None, acc)
    def reduce_mul_add_slice_to(self, a, b, c, out): return self._run(OP_MUL_ADD, a, b, c, out)
    def reduce_add_slice_to(self, a, b, out): return self._run(OP_ADD, a, b, None, out)
    def reduce_sub_slice_to(self, a, b, out): return self._run(OP_SUB, a, b, None, out)
    def reduce_neg_slice_to(self, a, out): return self._run(OP_NEG, a, None, None, out)

    def reduce_mul_scalar_slice_to(self, a, scalar, out):
        return self._run(OP_MUL_SCALAR, a, None, None, out, self._sc(scalar))

    def reduce_double_slice_to(self, a, out): return self._run(OP_DOUBLE, a, None, None, out)
    def reduce_sub_slice_rev_assign(self, a, b): return self._run(OP_SUB, a, b, None, b)          # b = a - b
    def reduce_mul_scalar_add_slice_to(self, a, scalar, c, out): return self._run(OP_MUL_SCALAR_ADD, a, None, c, out, self._sc(scalar))
    def factor_mul_add_slice_to(self, factor, rhs, addend, out): return self._run(OP_FACTOR_MUL_ADD, rhs, None, addend, out, self._sc(factor))

    def reduce_add_mul_scalar_slice_assign(self, acc, a, scalar):
        return self._run(OP_ADD_MUL_SCALAR, a, None, None, acc, self._sc(scalar))

    def _sc(self, s):
        return [int(x) for x in s] if isinstance(s, (list, tuple)) else [int(s)] * len(self.moduli)

    # FactorSliceOps with Shoup factors (primus_factor/src/ops.rs:58-118); factor value(s) per limb
    def factor_mul_slice_to(self, factor, rhs, out): return self._run(OP_FACTOR_MUL, rhs, None, None, out, self._sc(factor))
    def add_factor_mul_slice_assign(self, factor, acc, rhs): return self._run(OP_ADD_FACTOR_MUL, rhs, None, None, acc, self._sc(factor))
    def sub_factor_mul_slice_assign(self, factor, acc, rhs): return self._run(OP_SUB_FACTOR_MUL, rhs, None, None, acc, self._sc(factor))


class ApproxSignedBasis:
    """ApproxSignedBasis::new(Some(q), log_basis, reverse_length) + fused slice decomposition
    (primus_decompose/src/primitive/basis.rs:47-176, :254-406; primitive/common.rs:219-273)."""

    def __init__(self, modulus, log_basis, reverse_length=None, bits=64):
        self.q, self.bits, self._log_basis, self._rev = int(modulus), bits, log_basis, reverse_length or 0
        lv, dr = C.c_uint32(0), C.c_uint32(0)
        f = getattr(lib(), f"pfhe_basis{bits}_geometry")
        f.argtypes = [_ct(bits), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        check(f(self.q, log_basis, self._rev, C.byref(lv), C.byref(dr)))
        self._levels, self._drop = lv.value, dr.value

    def decompose_length(self): return self._levels
    def drop_bits(self): return self._drop
    def log_basis(self): return self._log_basis
    def basis_value(self): return 1 << self._log_basis
    def scalars(self): return [1 << (self._drop + l * self._log_basis) for l in range(self._levels)]

    def decompose_batch(self, values, digits):
        """values: CUDA [count]; digits: CUDA [levels][count], LSB level first, canonical mod q."""
        f = getattr(lib(), f"pfhe_decompose{self.bits}_batch")
        f.argtypes = [_ct(self.bits), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self.q, self._log_basis, self._rev, _dev_ptr(values, self.bits),
                _dev_ptr(digits, self.bits, values.numel() * self._levels), values.numel(), _stream()))


class RNSBase:
    """RNSBase (primus_rns/src/base.rs:26-122): constructor checks (EmptyBase / CoPrimeError), compose / decompose
    between residues [L][count] and little-endian big values [count][value_len], centred lifts of small values."""

    def __init__(self, moduli, bits=64):
        self.moduli, self.bits = [int(m) for m in moduli], bits
        self._h = C.c_void_p()
        self._p = f"pfhe_rns{bits}_"
        L = len(self.moduli)
        arr = (_ct(bits) * max(1, L))(*self.moduli)
        f = getattr(lib(), self._p + "create"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(arr, L, C.byref(self._h)))
        g = getattr(lib(), self._p + "big_uint_value_len"); g.argtypes = [C.c_void_p]; g.restype = C.c_size_t
        self.value_len = int(g(self._h))

    @classmethod
    def new(cls, moduli, bits=64):
        return cls(moduli, bits)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                f = getattr(lib(), self._p + "destroy"); f.argtypes = [C.c_void_p]; f.restype = None
                f(h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def moduli_count(self): return len(self.moduli)
    def big_uint_value_len(self): return self.value_len

    def moduli_product(self) -> int:
        out = (_ct(self.bits) * self.value_len)()
        f = getattr(lib(), self._p + "moduli_product"); f.argtypes = [C.c_void_p, C.c_void_p]
        check(f(self._h, out))
        return sum(int(w) << (self.bits * i) for i, w in enumerate(out))

    def _marr(self):
        return (_ct(self.bits) * len(self.moduli))(*self.moduli)

    def compose_multiple_values_to(self, multi_residues, big_values):
        """multi_residues: CUDA [L][count]; big_values: CUDA [count][value_len]."""
        count = multi_residues.numel() // len(self.moduli)
        f = getattr(lib(), self._p + "compose_batch"); f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(multi_residues, self.bits), _dev_ptr(big_values, self.bits, count * self.value_len), count, _stream()))

    def decompose_big_uint_values_to(self, big_values, multi_residues):
        count = big_values.numel() // self.value_len
        f = getattr(lib(), self._p + "decompose_batch"); f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(big_values, self.bits), _dev_ptr(multi_residues, self.bits, count * len(self.moduli)), count, _stream()))

    def wrapping_decompose_small_values_to(self, small, multi_residues, small_modulus):
        L = len(self.moduli)
        f = getattr(lib(), f"pfhe_rns{self.bits}_lift_small_batch")
        f.argtypes = [C.c_void_p, C.c_size_t, _ct(self.bits), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._marr(), L, int(small_modulus), _dev_ptr(small, self.bits), _dev_ptr(multi_residues, self.bits, small.numel() * L),
                small.numel(), _stream()))

    def wrapping_decompose_small_values_scaled_add_to(self, small, acc, small_modulus, scalars):
        """acc[l] += scalars[l] * centred_lift(small) mod q_l (base.rs:326-386)."""
        L = len(self.moduli)
        sc = (_ct(self.bits) * L)(*[int(x) for x in scalars])
        f = getattr(lib(), f"pfhe_rns{self.bits}_lift_small_scaled_add_batch")
        f.argtypes = [C.c_void_p, C.c_size_t, _ct(self.bits), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(self._marr(), L, int(small_modulus), sc, _dev_ptr(small, self.bits), _dev_ptr(acc, self.bits, small.numel() * L),
                small.numel(), _stream()))


class BaseConverter:
    """BaseConverter::new(input_base, output_base) (primus_rns/src/converter.rs:21-76) with the array conversions."""

    def __init__(self, in_moduli, out_moduli, bits=64):
        self.bits, self.in_moduli, self.out_moduli = bits, [int(m) for m in in_moduli], [int(m) for m in out_moduli]
        self._h = C.c_void_p()
        self._p = f"pfhe_baseconv{bits}_"
        a = (_ct(bits) * max(1, len(self.in_moduli)))(*self.in_moduli)
        b = (_ct(bits) * max(1, len(self.out_moduli)))(*self.out_moduli)
        f = getattr(lib(), self._p + "create"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        check(f(a, len(self.in_moduli), b, len(self.out_moduli), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                f = getattr(lib(), self._p + "destroy"); f.argtypes = [C.c_void_p]; f.restype = None
                f(h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def input_moduli_count(self): return len(self.in_moduli)
    def output_moduli_count(self): return len(self.out_moduli)

    def _run(self, name, crt_in, crt_out, n, out_limbs):
        polys = crt_in.numel() // (len(self.in_moduli) * n)
        f = getattr(lib(), self._p + name); f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_size_t, C.c_void_p]
        check(f(self._h, _dev_ptr(crt_in, self.bits), _dev_ptr(crt_out, self.bits, polys * out_limbs * n), n, polys, _stream()))

    def fast_convert_array(self, crt_in, crt_out, poly_length):
        """crt_in: CUDA [polys][L_in][n]; crt_out: CUDA [polys][L_out][n] (converter.rs:186-213)."""
        self._run("fast_convert_batch", crt_in, crt_out, poly_length, len(self.out_moduli))

    def exact_convert_array(self, crt_in, crt_out, poly_length):
        """crt_in: CUDA [polys][L_in][n]; crt_out: CUDA [polys][n]; one output modulus (converter.rs:257-365)."""
        self._run("exact_convert_batch", crt_in, crt_out, poly_length, 1)


class BigUintApproxSignedBasis:
    """BigUintApproxSignedBasis::new(Q, log_basis, reverse_length) over an RNS base
    (primus_decompose/src/big_integer/basis.rs:17-211) + the fused digit pipeline of the gadget product."""

    def __init__(self, rns: RNSBase, log_basis, reverse_length=None):
        self.rns, self.bits, self._log_basis, self._rev = rns, rns.bits, log_basis, reverse_length or 0
        lv, dr = C.c_uint32(0), C.c_uint32(0)
        f = getattr(lib(), f"pfhe_bigbasis{self.bits}_geometry")
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        check(f(rns._h, log_basis, self._rev, C.byref(lv), C.byref(dr)))
        self._levels, self._drop = lv.value, dr.value

    def decompose_length(self): return self._levels
    def drop_bits(self): return self._drop
    def log_basis(self): return self._log_basis
    def basis_value(self): return 1 << self._log_basis

    def gadget_decompose_batch(self, residues, digits, n):
        """residues: CUDA [polys][L][n] (CRT polynomials); digits: CUDA [polys][levels][L][n] lifted digits."""
        L = self.rns.moduli_count()
        polys = residues.numel() // (L * n)
        f = getattr(lib(), f"pfhe_rns{self.bits}_gadget_decompose_batch")
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        check(f(self.rns._h, self._log_basis, self._rev, _dev_ptr(residues, self.bits),
                _dev_ptr(digits, self.bits, polys * self._levels * L * n), n, polys, _stream()))


def dcrt_external_product_batch(table: _DcrtTable, basis: BigUintApproxSignedBasis, k, key, glwe_in, out, to_coeff=True, scratch=None):
    """CrtGlwe::mul_dcrt_ggsw_to (primus_lattice/src/glwe/crt.rs:200-227) [+ into_coeff_form] over a batch.
    key: CUDA [k+1][levels][k+1][L][N]; glwe_in/out: CUDA [batch][k+1][L][N]; scratch: CUDA byte tensor (allocated when None)."""
    import torch
    bits, L, n = table.bits, len(table.moduli), table.n
    batch = glwe_in.numel() // ((k + 1) * L * n)
    fs = getattr(lib(), f"pfhe_dcrt{bits}_external_product_scratch_bytes")
    fs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t]; fs.restype = C.c_size_t
    if scratch is None:
        need = int(fs(table._h, basis.rns._h, k, basis._log_basis, basis._rev, min(batch, 256) or 1))
        scratch = torch.empty(max(need, 16), dtype=torch.uint8, device=glwe_in.device)
    f = getattr(lib(), f"pfhe_dcrt{bits}_external_product_batch")
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                  C.c_void_p, C.c_size_t, C.c_void_p]
    check(f(table._h, basis.rns._h, k, basis._log_basis, basis._rev, _dev_ptr(key, bits), _dev_ptr(glwe_in, bits),
            _dev_ptr(out, bits, glwe_in.numel()), batch, int(bool(to_coeff)), C.c_void_p(scratch.data_ptr()), scratch.numel(), _stream()))
    return scratch


def mul_monomial_batch(moduli, degrees, polys, out, log_n, bits=64):
    """out = polys * X^degree per limb (primus_poly/src/poly/mul.rs:74-99, primus_lattice/src/glwe/crt.rs:76-114).
    polys/out: CUDA [batch][L][N]; degrees: CUDA uint32/int32 [batch]."""
    L = len(moduli)
    m = (_ct(bits) * L)(*[int(x) for x in moduli])
    batch = polys.numel() // (L << log_n)
    f = getattr(lib(), f"pfhe_poly{bits}_mul_monomial_batch")
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p]
    check(f(m, L, _dev_ptr(degrees, 32, batch), _dev_ptr(polys, bits), _dev_ptr(out, bits, polys.numel()), log_n, batch, _stream()))


def slice_op_bcast(op, moduli, a, b, out, n, group, bits=64):
    """NTT-domain ciphertext x polynomial with `b` broadcast over the `group` components of each ciphertext
    (primus_lattice/src/rlwe/ntt.rs:78-152). a, out: CUDA [rows][L][n]; b: CUDA [rows/group][L][n]; op in OP_MUL/OP_ADD_MUL/OP_SUB_MUL."""
    L = len(moduli)
    m = (_ct(bits) * L)(*[int(x) for x in moduli])
    rows = a.numel() // (L * n)
    f = getattr(lib(), f"pfhe_mod{bits}_slice_op_bcast")
    f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    check(f(op, m, L, _dev_ptr(a, bits), _dev_ptr(b, bits, a.numel() // group), _dev_ptr(out, bits, a.numel()), rows, n, group, _stream()))


def butterfly_mul_factor_batch(moduli, a, s, w, out, n, bits=64):
    """(a, out) = (a + s, (a - s) * w) per limb (primus_poly/src/dcrt/mul.rs:189-222). a, s, out: CUDA [rows][L][n]; w: CUDA [L][n]."""
    L = len(moduli)
    m = (_ct(bits) * L)(*[int(x) for x in moduli])
    rows = a.numel() // (L * n)
    f = getattr(lib(), f"pfhe_mod{bits}_butterfly_mul_factor")
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    check(f(m, L, _dev_ptr(a, bits), _dev_ptr(s, bits, a.numel()), _dev_ptr(w, bits, L * n), _dev_ptr(out, bits, a.numel()), rows, n, _stream()))


def inv_slice_batch(q, a, out, bits=64):
    """out = a^-1 mod q element-wise (prime q). Returns the smallest index without an inverse (zero element) or None
    (try_reduce_inv_slice_to, primus_reduce/src/slice_ops.rs:293-300)."""
    import torch
    bad = torch.full((1,), -1, dtype=torch.int64, device=a.device)
    f = getattr(lib(), f"pfhe_mod{bits}_inv_slice")
    f.argtypes = [_ct(bits), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    check(f(int(q), _dev_ptr(a, bits), _dev_ptr(out, bits, a.numel()), a.numel(), C.c_void_p(bad.data_ptr()), _stream()))
    v = int(bad.item())
    return None if v == -1 else v


def dot_product_batch(q, a, b, out, n, bits=64):
    """reduce_dot_product per row (primus_modulus/src/common/compact/slice.rs:371-438). a, b: CUDA [rows][n]; out: CUDA [rows]."""
    rows = a.numel() // n
    f = getattr(lib(), f"pfhe_mod{bits}_dot_product_batch")
    f.argtypes = [_ct(bits), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    check(f(int(q), _dev_ptr(a, bits), _dev_ptr(b, bits, a.numel()), _dev_ptr(out, bits, rows), rows, n, _stream()))


class MultiplyFactor:
    """MultiplyFactor (primus_factor/src/mul_factor/mod.rs:4-88): operand with its precomputed quotient for bit shift 32/52/64."""

    def __init__(self, operand, bit_shift, modulus):
        q = C.c_uint64(0)
        f = lib().pfhe_multiply_factor64; f.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]
        check(f(int(operand), int(bit_shift), int(modulus), C.byref(q)))
        self._operand, self._quotient, self.bit_shift, self.modulus = int(operand), int(q.value), int(bit_shift), int(modulus)

    def operand(self): return self._operand
    def quotient(self): return self._quotient

    def mul_modulo(self, b):
        out = C.c_uint64(0)
        f = lib().pfhe_multiply_factor64_mul; f.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        check(f(self._operand, self._quotient, self.bit_shift, int(b), self.modulus, C.byref(out)))
        return int(out.value)


def extract_lwe_batch(q, rlwe, lwe, n, bits=64):
    """Rlwe::extract_lwe (primus_lattice/src/rlwe/coeff.rs:264-288) over a batch."""
    f = getattr(lib(), f"pfhe_extract_lwe{bits}_batch")
    f.argtypes = [_ct(bits), C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    batch = rlwe.numel() // (2 * n)
    check(f(int(q), _dev_ptr(rlwe, bits), _dev_ptr(lwe, bits, batch * (n + 1)), n, batch, _stream()))


def extract_lwe_ex_batch(q, rlwe, lwe, n, index=0, count=1, bits=64):
    """Rlwe::extract_lwe_with_index (count = 1) / extract_first_few_lwe (index = 0) (primus_lattice/src/rlwe/coeff.rs:194-261)."""
    f = getattr(lib(), f"pfhe_extract_lwe{bits}_ex_batch")
    f.argtypes = [_ct(bits), C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    batch = rlwe.numel() // (2 * n)
    check(f(int(q), _dev_ptr(rlwe, bits), _dev_ptr(lwe, bits, batch * (n + count)), n, batch, index, count, _stream()))


def modmul_microbench(kind: int, blocks: int, iters: int, device: int = 0) -> float:
    ms = C.c_float(0)
    f = lib().pfhe_modmul_microbench
    f.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    check(f(device, kind, blocks, iters, C.byref(ms)))
    return ms.value


# ======================================================================================================
# Round 2: bootstrapping keys, multi-device drivers, whole-ciphertext transforms, bytes, UintNttTable
# ======================================================================================================
def _size(x):
    return x.numel() if _is_torch(x) else x.size


class BootstrappingKey:
    """n_lwe RGSW ciphertexts in NTT form, resident on the table's device (pfhe_bsk*_create).
    `bsk`: host array [n_lwe][2][levels][2][N] (NttRgsw layout) or its `to_bytes()` image (bytes / bytearray / memoryview)."""

    def __init__(self, table: _NttTable, log_basis, levels, n_lwe, bsk):
        self.table, self.bits, self.n_lwe = table, table.bits, int(n_lwe)
        self._h = C.c_void_p()
        if isinstance(bsk, (bytes, bytearray, memoryview)):
            buf = (C.c_char * len(bsk)).from_buffer_copy(bytes(bsk))
            f = getattr(lib(), f"pfhe_bsk{self.bits}_create_from_bytes")
            f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
            check(f(table._h, log_basis, levels or 0, n_lwe, buf, len(bsk), C.byref(self._h)))
        else:
            f = getattr(lib(), f"pfhe_bsk{self.bits}_create")
            f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
            lv = ApproxSignedBasis(table.q, log_basis, levels, self.bits).decompose_length()
            check(f(table._h, log_basis, levels or 0, n_lwe, _host_ptr(bsk, self.bits, n_lwe * 2 * lv * 2 * table.n), C.byref(self._h)))
        g = getattr(lib(), f"pfhe_bsk{self.bits}_levels"); g.argtypes = [C.c_void_p]; g.restype = C.c_uint32
        self.levels = int(g(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                f = getattr(lib(), f"pfhe_bsk{self.bits}_destroy"); f.argtypes = [C.c_void_p]; f.restype = None
                f(self._h); self._h = C.c_void_p()
        except Exception:
            pass

    def device_ptr(self) -> int:
        f = getattr(lib(), f"pfhe_bsk{self.bits}_device_ptr"); f.argtypes = [C.c_void_p]; f.restype = C.c_void_p
        return int(f(self._h) or 0)

    def bootstrap_slices(self, lwe, test_vector, out=None, extract=True):
        """Host LWE samples (uint32 [batch][n_lwe+1], already in Z_2N) -> host LWE [batch][N+1] (extract) or accumulators [batch][2][N]."""
        t = self.table
        batch = _size(lwe) // (self.n_lwe + 1)
        width = (t.n + 1) if extract else 2 * t.n
        if out is None:
            out = np.empty((batch, width), dtype=_np_dtype(self.bits))
        f = getattr(lib(), f"pfhe_bootstrap{self.bits}_slices")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        check(f(t._h, self._h, _host_ptr(lwe, 32, batch * (self.n_lwe + 1)), _host_ptr(test_vector, self.bits, t.n),
                _host_ptr(out, self.bits, batch * width), batch, int(bool(extract))))
        return out


def modulus_switch_batch(q, log_2n, lwe, out, bits=64):
    """LWE modulus switch q -> 2N = 2^log_2n (round to nearest; not in the reference): CUDA words in, CUDA uint32 out."""
    f = getattr(lib(), f"pfhe_lwe{bits}_modulus_switch_batch")
    f.argtypes = [_ct(bits), C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    check(f(int(q), int(log_2n), _dev_ptr(lwe, bits), _dev_ptr(out, 32, lwe.numel()), lwe.numel(), _stream()))


class registered_host_buffer:
    """Context manager: page-lock a caller-owned numpy buffer in place (pfhe_host_register) so that the *_slices shims copy straight
    from / to it instead of staging through bounce buffers; unregisters on exit.  Mirrors the RAII guard of the Rust FFI crate."""

    def __init__(self, array):
        self.array = array
        self._ptr = array.ctypes.data

    def __enter__(self):
        f = lib().pfhe_host_register
        f.argtypes = [C.c_void_p, C.c_size_t]
        check(f(self._ptr, self.array.nbytes))
        return self.array

    def __exit__(self, *exc):
        f = lib().pfhe_host_unregister
        f.argtypes = [C.c_void_p]
        check(f(self._ptr))
        return False


def host_is_pageable(array) -> bool:
    f = lib().pfhe_host_is_pageable
    f.argtypes = [C.c_void_p]
    f.restype = C.c_int
    return bool(f(array.ctypes.data))


class MultiNttTable:
    """The same (log_n, q) table replicated on several devices of one process; host-slice calls are split into contiguous
    shards, one host thread per device (pfhe_multi_*).  `devices` may repeat a device index."""

    def __init__(self, log_n, modulus, devices, bits=64):
        self.bits, self.devices = bits, list(devices)
        cls = U64NttTable if bits == 64 else U32NttTable
        self.tables = [cls(log_n, modulus, device=d) for d in self.devices]
        self.n, self.q = self.tables[0].n, self.tables[0].q
        self._arr = (C.c_void_p * len(self.tables))(*[t._h for t in self.tables])

    def _p(self, name):
        return getattr(lib(), name.replace("#", str(self.bits)))

    def transform_slices(self, polys, inverse=False, lazy=False):
        f = self._p("pfhe_multi_ntt#_transform_slices"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        check(f(self._arr, len(self.tables), _host_ptr(polys, self.bits), _size(polys) // self.n, int(inverse), int(lazy)))

    def inverse_transform_slices(self, polys): self.transform_slices(polys, inverse=True)

    def polymul_slices(self, a, b, c):
        f = self._p("pfhe_multi_ntt#_polymul_slices"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        check(f(self._arr, len(self.tables), _host_ptr(a, self.bits), _host_ptr(b, self.bits, _size(a)), _host_ptr(c, self.bits, _size(a)),
                _size(a) // self.n))
        return c

    def external_product_slices(self, k, log_basis, levels, key, glwe_in, out, to_coeff=True):
        f = self._p("pfhe_multi_ggsw#_external_product_slices")
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        check(f(self._arr, len(self.tables), k, log_basis, levels or 0, _host_ptr(key, self.bits), _host_ptr(glwe_in, self.bits),
                _host_ptr(out, self.bits, _size(glwe_in)), _size(glwe_in) // ((k + 1) * self.n), int(bool(to_coeff))))
        return out

    def bootstrapping_keys(self, log_basis, levels, n_lwe, bsk):
        """Replicate one bootstrapping key on every device."""
        return [BootstrappingKey(t, log_basis, levels, n_lwe, bsk) for t in self.tables]

    def bootstrap_slices(self, keys, lwe, test_vector, out=None, extract=True):
        n_lwe = keys[0].n_lwe
        batch = _size(lwe) // (n_lwe + 1)
        width = (self.n + 1) if extract else 2 * self.n
        if out is None:
            out = np.empty((batch, width), dtype=_np_dtype(self.bits))
        karr = (C.c_void_p * len(keys))(*[k._h for k in keys])
        f = self._p("pfhe_multi_bootstrap#_slices")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        check(f(self._arr, karr, len(self.tables), _host_ptr(lwe, 32, batch * (n_lwe + 1)), _host_ptr(test_vector, self.bits, self.n),
                _host_ptr(out, self.bits, batch * width), batch, int(bool(extract))))
        return out


# ---- whole-ciphertext transforms (primus_lattice/src/macros/mod.rs:537-674) and the byte layout (:39-97) ----
_SHAPES = {"rlwe": 0, "rlev": 1, "rgsw": 1, "glwe": 1, "glev": 2, "ggsw": 2}


def cipher_words(table: _NttTable, shape: str, *dims) -> int:
    """Word count of a flat ciphertext container: rlwe() / rlev(levels) / rgsw(levels) / glwe(k) / glev(k, levels) / ggsw(k, levels)."""
    f = getattr(lib(), f"pfhe_{shape}{table.bits}_words"); f.argtypes = [C.c_void_p] + [C.c_uint32] * _SHAPES[shape]; f.restype = C.c_size_t
    return int(f(table._h, *[int(d) for d in dims]))


def into_ntt_form(table: _NttTable, shape: str, data, *dims):
    """`cipher.into_ntt_form(table)` for the named container: in place on host storage, returns it."""
    f = getattr(lib(), f"pfhe_{shape}{table.bits}_into_ntt_form"); f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_uint32] * _SHAPES[shape]
    check(f(table._h, _host_ptr(data, table.bits, cipher_words(table, shape, *dims)), *[int(d) for d in dims]))
    return data


def into_coeff_form(table: _NttTable, shape: str, data, *dims):
    f = getattr(lib(), f"pfhe_{shape}{table.bits}_into_coeff_form"); f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_uint32] * _SHAPES[shape]
    check(f(table._h, _host_ptr(data, table.bits, cipher_words(table, shape, *dims)), *[int(d) for d in dims]))
    return data


def write_ntt_form(table: _NttTable, src, dst):
    f = getattr(lib(), f"pfhe_cipher{table.bits}_write_ntt_form"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    check(f(table._h, _host_ptr(src, table.bits), _host_ptr(dst, table.bits, _size(src)), _size(src)))
    return dst


def write_coeff_form(table: _NttTable, src, dst):
    f = getattr(lib(), f"pfhe_cipher{table.bits}_write_coeff_form"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    check(f(table._h, _host_ptr(src, table.bits), _host_ptr(dst, table.bits, _size(src)), _size(src)))
    return dst


def dcrt_into_ntt_form(table: _DcrtTable, data):
    f = getattr(lib(), f"pfhe_dcrt_cipher{table.bits}_into_ntt_form"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    check(f(table._h, _host_ptr(data, table.bits), _size(data)))
    return data


def dcrt_into_coeff_form(table: _DcrtTable, data):
    f = getattr(lib(), f"pfhe_dcrt_cipher{table.bits}_into_coeff_form"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    check(f(table._h, _host_ptr(data, table.bits), _size(data)))
    return data


def to_bytes(words, bits=64) -> bytes:
    """`cipher.to_bytes()`: raw little-endian bytes of the flat word array (macros/mod.rs:74-80)."""
    n = _size(words)
    out = (C.c_uint8 * (n * bits // 8))()
    f = getattr(lib(), f"pfhe_cipher{bits}_write_bytes"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    check(f(_host_ptr(words, bits), n, out, len(out)))
    return bytes(out)


def from_bytes(data: bytes, bits=64):
    """`Cipher::from_bytes(data)`: the flat word array of a serialised container (macros/mod.rs:47-52)."""
    if len(data) % (bits // 8):
        raise PfheError(9, "byte length is not a multiple of the word size")
    out = np.empty(len(data) // (bits // 8), dtype=_np_dtype(bits))
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    f = getattr(lib(), f"pfhe_cipher{bits}_read_bytes"); f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    check(f(_host_ptr(out, bits), out.size, buf, len(data)))
    return out


class UintNttTable:
    """UintNttTable<T> (primus_ntt/src/ntt/primitive.rs:37-396), T = u16 / u32 / u64: the generic table with its own constructor
    rules; transforms run the plain radix-2 kernel."""

    def __init__(self, log_n: int, modulus: int, bits: int = 64, device: int = 0):
        if bits not in (16, 32, 64):
            raise ValueError("word size must be 16, 32 or 64")
        self.bits, self.q, self.log_n, self.n = bits, int(modulus), int(log_n), 1 << int(log_n)
        self._h = C.c_void_p()
        ct = {16: C.c_uint16, 32: C.c_uint32, 64: C.c_uint64}[bits]
        if self.q >> bits:
            raise PfheError(5, "modulus does not fit the word type")
        f = getattr(lib(), f"pfhe_uintntt{bits}_create"); f.argtypes = [C.c_int, C.c_uint32, ct, C.c_void_p]
        check(f(device, log_n, self.q, C.byref(self._h)))
        self._ct = ct

    new = classmethod(lambda cls, log_n, modulus, bits=64, device=0: cls(log_n, modulus, bits, device))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                f = getattr(lib(), f"pfhe_uintntt{self.bits}_destroy"); f.argtypes = [C.c_void_p]; f.restype = None
                f(self._h); self._h = C.c_void_p()
        except Exception:
            pass

    def poly_length(self): return self.n
    def modulus(self): return self.q

    def root(self):
        f = getattr(lib(), f"pfhe_uintntt{self.bits}_root"); f.argtypes = [C.c_void_p]; f.restype = self._ct
        return int(f(self._h))

    def inv_root(self):
        f = getattr(lib(), f"pfhe_uintntt{self.bits}_inv_root"); f.argtypes = [C.c_void_p]; f.restype = self._ct
        return int(f(self._h))

    def _run(self, name, polys, lazy):
        if not isinstance(polys, np.ndarray) or polys.dtype.itemsize * 8 != self.bits or not polys.flags.c_contiguous or polys.size % self.n:
            raise TypeError("expected a C-contiguous numpy array of whole polynomials in the table's word type")
        f = getattr(lib(), f"pfhe_uintntt{self.bits}_{name}"); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        check(f(self._h, polys.ctypes.data_as(C.c_void_p), polys.size // self.n, int(lazy)))

    def transform_slice(self, poly): self._run("transform_slices", poly, 0)
    def lazy_transform_slice(self, poly): self._run("transform_slices", poly, 1)
    def inverse_transform_slice(self, values): self._run("inverse_transform_slices", values, 0)
    def lazy_inverse_transform_slice(self, values): self._run("inverse_transform_slices", values, 1)
    transform_slices = transform_slice
    inverse_transform_slices = inverse_transform_slice
