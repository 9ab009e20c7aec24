"""ctypes binding of the CPU oracle (oracle/libpfhe_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(primus_fhe_b200) never imports this module.

The class names mirror the reference's types so the parity tests read like the
reference's own tests: U32NttTable / U64NttTable (primus_ntt/src/ntt/mod.rs:16-113),
BarrettModulus (primus_modulus/src/barrett/mod.rs:25), ShoupFactor
(primus_factor/src/shoup_factor/mod.rs:22), ApproxSignedBasis
(primus_decompose/src/primitive/basis.rs:12), RNSBase (primus_rns/src/base.rs:26),
BigUintApproxSignedBasis (primus_decompose/src/big_integer/basis.rs:17).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpfhe_oracle.so")

ERR_NAMES = {
    0: "Ok", 1: "NoPrimitiveRoot", 2: "DegreeConversionErr", 3: "DegreeTooLarge",
    4: "NttTableErr", 5: "ModulusTooLarge", 6: "EmptyBase", 7: "CoPrimeError",
}


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        super().__init__(ERR_NAMES.get(code, str(code)))


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("pfhe_oracle.c", "oracle_impl.inc", "pfhe_oracle.h", "pfhe_oracle_avx512.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _np(bits):
    return np.uint32 if bits == 32 else np.uint64


def _ct(bits):
    return C.c_uint32 if bits == 32 else C.c_uint64


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(x, bits):
    a = np.ascontiguousarray(x, dtype=_np(bits))
    return a


def max_threads() -> int:
    return int(lib().o_max_threads())


class _Sized:
    bits = 64

    @classmethod
    def fn(cls, name, restype=None, argtypes=None):
        f = getattr(lib(), f"{name}{cls.bits}")
        f.restype = restype
        if argtypes is not None:
            f.argtypes = argtypes
        return f


# --------------------------------------------------------------------------- scalar helpers
def shoup_quot(value, q, bits=64):
    f = getattr(lib(), f"o_shoup_quot{bits}"); f.restype = _ct(bits); f.argtypes = [_ct(bits)] * 2
    return int(f(value, q))


def shoup_mul(w, wq, y, q, bits=64, lazy=False):
    f = getattr(lib(), f"o_shoup_mul{'_lazy' if lazy else ''}{bits}")
    f.restype = _ct(bits); f.argtypes = [_ct(bits)] * 4
    return int(f(w, wq, y, q))


def min_primitive_root(log_degree, q, bits=64):
    out = _ct(bits)()
    f = getattr(lib(), f"o_min_primitive_root{bits}"); f.restype = C.c_int
    f.argtypes = [C.c_uint, _ct(bits), C.c_void_p]
    e = f(log_degree, q, C.byref(out))
    if e:
        raise OracleError(e)
    return int(out.value)


class BarrettModulus:
    """primus_modulus/src/barrett/mod.rs:25-139 + slice ops (barrett/slice.rs:185-295)."""

    def __init__(self, q, bits=64):
        self.q, self.bits = int(q), bits
        ct = _ct(bits)

        class S(C.Structure):
            _fields_ = [("q", ct), ("ratio", ct * 2)]
        self._s = S()
        f = getattr(lib(), f"o_barrett_new{bits}"); f.restype = C.c_int; f.argtypes = [ct, C.c_void_p]
        e = f(self.q, C.byref(self._s))
        if e:
            raise ValueError("modulus is too large." if e == -2 else "modulus can't be 0 or 1.")

    @property
    def ratio(self):
        return [int(self._s.ratio[0]), int(self._s.ratio[1])]

    def value(self):
        return self.q

    def _f(self, name, restype, argtypes):
        f = getattr(lib(), f"{name}{self.bits}"); f.restype = restype; f.argtypes = argtypes
        return f

    def reduce(self, v):
        ct = _ct(self.bits)
        return int(self._f("o_barrett_reduce", ct, [C.c_void_p, ct])(C.byref(self._s), v))

    def reduce_wide(self, lo, hi):
        ct = _ct(self.bits)
        return int(self._f("o_barrett_reduce_wide", ct, [C.c_void_p, ct, ct])(C.byref(self._s), lo, hi))

    def reduce_mul(self, a, b):
        ct = _ct(self.bits)
        return int(self._f("o_barrett_mul", ct, [C.c_void_p, ct, ct])(C.byref(self._s), a, b))

    def reduce_mul_add(self, a, b, c):
        ct = _ct(self.bits)
        return int(self._f("o_barrett_mul_add", ct, [C.c_void_p, ct, ct, ct])(C.byref(self._s), a, b, c))

    # slice operators (primus_reduce/src/slice_ops.rs:137-230)
    def _slice(self, name, *arrays_and_scalars):
        ct = _ct(self.bits)
        args, types = [self.q], [ct]
        n = None
        for x in arrays_and_scalars:
            if isinstance(x, np.ndarray):
                args.append(_ptr(x)); types.append(C.c_void_p); n = x.size
            else:
                args.append(int(x)); types.append(ct)
        args.append(n); types.append(C.c_size_t)
        self._f(name, None, types)(*args)

    def reduce_mul_slice_to(self, a, b):
        a, b = _arr(a, self.bits), _arr(b, self.bits); out = np.empty_like(a)
        self._slice("o_mod_mul_slice", a, b, out); return out

    def reduce_add_mul_slice_assign(self, acc, a, b):
        self._slice("o_mod_add_mul_slice", acc, _arr(a, self.bits), _arr(b, self.bits)); return acc

    def reduce_sub_mul_slice_assign(self, acc, a, b):
        self._slice("o_mod_sub_mul_slice", acc, _arr(a, self.bits), _arr(b, self.bits)); return acc

    def reduce_mul_add_slice_to(self, a, b, c):
        a = _arr(a, self.bits); out = np.empty_like(a)
        self._slice("o_mod_mul_add_slice", a, _arr(b, self.bits), _arr(c, self.bits), out); return out

    def reduce_mul_scalar_slice_to(self, a, s):
        a = _arr(a, self.bits); out = np.empty_like(a)
        ct = _ct(self.bits)
        self._f("o_mod_mul_scalar_slice", None, [ct, C.c_void_p, ct, C.c_void_p, C.c_size_t])(
            self.q, _ptr(a), int(s), _ptr(out), a.size)
        return out

    def reduce_add_mul_scalar_slice_assign(self, acc, a, s):
        a = _arr(a, self.bits); ct = _ct(self.bits)
        self._f("o_mod_add_mul_scalar_slice", None, [ct, C.c_void_p, C.c_void_p, ct, C.c_size_t])(
            self.q, _ptr(acc), _ptr(a), int(s), a.size)
        return acc

    def reduce_add_slice_to(self, a, b):
        a = _arr(a, self.bits); out = np.empty_like(a)
        self._slice("o_mod_add_slice", a, _arr(b, self.bits), out); return out

    def reduce_sub_slice_to(self, a, b):
        a = _arr(a, self.bits); out = np.empty_like(a)
        self._slice("o_mod_sub_slice", a, _arr(b, self.bits), out); return out

    def reduce_neg_slice_to(self, a):
        a = _arr(a, self.bits); out = np.empty_like(a)
        self._slice("o_mod_neg_slice", a, out); return out

    def reduce_dot_product(self, a, b):
        a, b = _arr(a, self.bits), _arr(b, self.bits); ct = _ct(self.bits)
        return int(self._f("o_mod_dot_product", ct, [ct, C.c_void_p, C.c_void_p, C.c_size_t])(
            self.q, _ptr(a), _ptr(b), a.size))


class ShoupFactor:
    """primus_factor/src/shoup_factor/mod.rs:22-143 + FactorSliceOps (ops.rs:58-118)."""

    def __init__(self, value, q, bits=64):
        self.value, self.q, self.bits = int(value), int(q), bits
        self.quotient = shoup_quot(value, q, bits)

    def factor_mul_modulo(self, y):
        return shoup_mul(self.value, self.quotient, y, self.q, self.bits)

    def lazy_factor_mul_modulo(self, y):
        return shoup_mul(self.value, self.quotient, y, self.q, self.bits, lazy=True)

    def _f(self, name, types):
        f = getattr(lib(), f"{name}{self.bits}"); f.restype = None; f.argtypes = types; return f

    def factor_mul_slice_to(self, rhs):
        rhs = _arr(rhs, self.bits); out = np.empty_like(rhs); ct = _ct(self.bits)
        self._f("o_factor_mul_slice", [ct, ct, C.c_void_p, C.c_void_p, C.c_size_t])(
            self.value, self.q, _ptr(rhs), _ptr(out), rhs.size)
        return out

    def add_factor_mul_slice_assign(self, acc, rhs):
        rhs = _arr(rhs, self.bits); ct = _ct(self.bits)
        self._f("o_add_factor_mul_slice", [ct, ct, C.c_void_p, C.c_void_p, C.c_size_t])(
            self.value, self.q, _ptr(acc), _ptr(rhs), rhs.size)
        return acc

    def sub_factor_mul_slice_assign(self, acc, rhs):
        rhs = _arr(rhs, self.bits); ct = _ct(self.bits)
        self._f("o_sub_factor_mul_slice", [ct, ct, C.c_void_p, C.c_void_p, C.c_size_t])(
            self.value, self.q, _ptr(acc), _ptr(rhs), rhs.size)
        return acc


# --------------------------------------------------------------------------- NTT tables
class _NttTable:
    bits = 64

    def __init__(self, log_n, q):
        ct = _ct(self.bits)
        err = C.c_int(0)
        f = getattr(lib(), f"o_ntt_create{self.bits}"); f.restype = C.c_void_p
        f.argtypes = [C.c_uint, ct, C.c_void_p]
        self._h = f(log_n, int(q), C.byref(err))
        if not self._h:
            raise OracleError(err.value)
        self.log_n, self.q, self.n = log_n, int(q), 1 << log_n

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            f = getattr(lib(), f"o_ntt_destroy{self.bits}"); f.restype = None; f.argtypes = [C.c_void_p]
            f(h); self._h = None

    def _get(self, name):
        f = getattr(lib(), f"{name}{self.bits}"); f.restype = _ct(self.bits); f.argtypes = [C.c_void_p]
        return int(f(self._h))

    def poly_length(self): return self.n
    def root(self): return self._get("o_ntt_root")
    def inv_root(self): return self._get("o_ntt_inv_root")
    def inv_n(self): return self._get("o_ntt_inv_n")
    def modulus(self): return self.q

    def roots(self):
        f = getattr(lib(), f"o_ntt_roots{self.bits}"); f.restype = C.POINTER(_ct(self.bits)); f.argtypes = [C.c_void_p]
        return np.ctypeslib.as_array(f(self._h), shape=(self.n,)).copy()

    def inv_roots(self):
        f = getattr(lib(), f"o_ntt_inv_roots{self.bits}"); f.restype = C.POINTER(_ct(self.bits)); f.argtypes = [C.c_void_p]
        return np.ctypeslib.as_array(f(self._h), shape=(self.n,)).copy()

    def _tf(self, name, v, *extra):
        assert v.dtype == _np(self.bits) and v.flags.c_contiguous
        f = getattr(lib(), f"{name}{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * len(extra)
        f(self._h, _ptr(v), *extra)

    # trait methods (in place on numpy arrays of length n)
    def transform_slice(self, v): self._tf("o_ntt_forward", v, 1)
    def lazy_transform_slice(self, v): self._tf("o_ntt_forward", v, 4)
    def inverse_transform_slice(self, v): self._tf("o_ntt_inverse", v, 1)
    def lazy_inverse_transform_slice(self, v): self._tf("o_ntt_inverse", v, 2)
    def generic_transform_slice(self, v): self._tf("o_ntt_forward_generic", v)
    def generic_inverse_transform_slice(self, v): self._tf("o_ntt_inverse_generic", v)

    def direct_transform(self, x):
        x = _arr(x, self.bits); out = np.empty_like(x)
        f = getattr(lib(), f"o_ntt_forward_direct{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f(self._h, _ptr(x), _ptr(out)); return out

    def transform_monomial(self, coeff, degree):
        out = np.empty(self.n, dtype=_np(self.bits))
        f = getattr(lib(), f"o_ntt_monomial{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, _ct(self.bits), C.c_size_t, C.c_void_p]
        f(self._h, int(coeff), int(degree), _ptr(out)); return out

    def transform_coeff_one_monomial(self, degree): return self.transform_monomial(1, degree)
    def transform_coeff_minus_one_monomial(self, degree): return self.transform_monomial(self.q - 1, degree)

    # batch helpers ([batch, n] arrays, OpenMP)
    def forward_batch(self, v, threads=0):
        assert v.dtype == _np(self.bits) and v.flags.c_contiguous
        f = getattr(lib(), f"o_ntt_forward_batch{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        f(self._h, _ptr(v), v.size // self.n, threads or max_threads())

    def inverse_batch(self, v, threads=0):
        assert v.dtype == _np(self.bits) and v.flags.c_contiguous
        f = getattr(lib(), f"o_ntt_inverse_batch{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        f(self._h, _ptr(v), v.size // self.n, threads or max_threads())

    def polymul_batch(self, a, b, threads=0):
        a, b = _arr(a, self.bits), _arr(b, self.bits); c = np.empty_like(a)
        f = getattr(lib(), f"o_ntt_polymul_batch{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int]
        f(self._h, _ptr(a), _ptr(b), _ptr(c), a.size // self.n, threads or max_threads())
        return c


class U64NttTable(_NttTable):
    bits = 64

    # ---- AVX-512 IFMA restatement of the reference's fast forward path (pfhe_oracle_avx512.c): the CPU baseline of bench.py ----
    def simd_supported(self) -> bool:
        """True when this CPU has AVX-512 IFMA and the table qualifies for the reference's BIT_SHIFT = 52 back-end (q < 2^50, N >= 16)."""
        f = lib().o_ifma_supported; f.restype = C.c_int
        return bool(f()) and self.q < (1 << 50) and self.n >= 16

    def _simd(self):
        if getattr(self, "_ifma", None) is None:
            f = lib().o_ifma_create; f.restype = C.c_void_p; f.argtypes = [C.c_uint, C.c_uint64, C.c_void_p]
            roots = np.ascontiguousarray(self.roots(), dtype=np.uint64)
            self._ifma = f(self.log_n, self.q, _ptr(roots))
            if not self._ifma:
                raise OracleError(4)
        return self._ifma

    def forward_batch_simd(self, v, threads=0):
        """forward_batch through the IFMA path (canonical outputs, bit-identical to forward_batch)."""
        assert v.dtype == np.uint64 and v.flags.c_contiguous and self.simd_supported()
        f = lib().o_ifma_forward_batch; f.restype = None; f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        f(self._simd(), _ptr(v), v.size // self.n, threads or max_threads())


class U32NttTable(_NttTable):
    bits = 32


def naive_mul(a, b, q, bits=64):
    """Schoolbook negacyclic product (primus_poly/src/poly/mul.rs:107-134)."""
    a, b = _arr(a, bits), _arr(b, bits); c = np.empty_like(a)
    f = getattr(lib(), f"o_poly_naive_mul{bits}"); f.restype = None
    f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, _ct(bits)]
    f(_ptr(a), _ptr(b), _ptr(c), a.size, int(q)); return c


def mul_monomial(p, r, q, bits=64):
    """p * X^r (primus_poly/src/poly/mul.rs:74-99)."""
    p = _arr(p, bits); out = np.empty_like(p)
    f = getattr(lib(), f"o_poly_mul_monomial{bits}"); f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, _ct(bits)]
    f(_ptr(p), _ptr(out), p.size, int(r), int(q)); return out


def extract_lwe(rlwe, q, bits=64):
    rlwe = _arr(rlwe, bits); n = rlwe.size // 2; out = np.empty(n + 1, dtype=_np(bits))
    f = getattr(lib(), f"o_extract_lwe{bits}"); f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, _ct(bits)]
    f(_ptr(rlwe), _ptr(out), n, int(q)); return out


# --------------------------------------------------------------------------- decomposition
def extract_lwe_with_index(rlwe, index, q):
    """Rlwe::extract_lwe_with_index (primus_lattice/src/rlwe/coeff.rs:194-226), numpy restatement of its three copy loops."""
    rlwe = np.asarray(rlwe); n = rlwe.size // 2; split = index + 1
    out = np.empty(n + 1, dtype=rlwe.dtype)
    out[:split] = rlwe[:split][::-1]
    tail = rlwe[split:n][::-1]
    out[split:n] = np.where(tail == 0, 0, q - tail).astype(rlwe.dtype)
    out[n] = rlwe[n + index]
    return out


def extract_first_few_lwe(rlwe, count, q):
    """Rlwe::extract_first_few_lwe (coeff.rs:229-261)."""
    rlwe = np.asarray(rlwe); n = rlwe.size // 2
    out = np.empty(n + count, dtype=rlwe.dtype)
    out[0] = rlwe[0]
    tail = rlwe[1:n][::-1]
    out[1:n] = np.where(tail == 0, 0, q - tail).astype(rlwe.dtype)
    out[n:] = rlwe[n:n + count]
    return out


class ApproxSignedBasis:
    """primus_decompose/src/primitive/basis.rs:12-407 (non-power-of-two modulus branch)."""

    def __init__(self, q, log_basis, reverse_length=None, bits=64):
        self.bits = bits; ct = _ct(bits)

        class S(C.Structure):
            _fields_ = [("q", ct), ("basis", ct), ("basis_m1", ct), ("q_minus_basis", ct), ("carry_mask", ct),
                        ("log_basis", C.c_uint), ("levels", C.c_uint), ("value_bits", C.c_uint), ("drop_bits", C.c_uint),
                        ("has_threshold", C.c_int), ("threshold", ct), ("add", ct),
                        ("has_init_mask", C.c_int), ("init_mask", ct)]
        self._s = S()
        f = getattr(lib(), f"o_basis_new{bits}"); f.restype = C.c_int
        f.argtypes = [ct, C.c_uint, C.c_uint, C.c_void_p]
        e = f(int(q), log_basis, reverse_length or 0, C.byref(self._s))
        if e:
            raise ValueError(f"ApproxSignedBasis::new failed ({e})")
        self.q = int(q)

    def decompose_length(self): return int(self._s.levels)
    def drop_bits(self): return int(self._s.drop_bits)
    def log_basis(self): return int(self._s.log_basis)
    def basis_value(self): return int(self._s.basis)
    def threshold(self): return int(self._s.threshold) if self._s.has_threshold else None
    def scalars(self): return [1 << (self.drop_bits() + l * self.log_basis()) for l in range(self.decompose_length())]

    def init_value_carry_slice_to(self, values):
        values = _arr(values, self.bits); adj = np.empty_like(values); car = np.empty(values.size, dtype=np.uint8)
        f = getattr(lib(), f"o_basis_init_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 4 + [C.c_size_t]
        f(C.byref(self._s), _ptr(values), _ptr(adj), _ptr(car), values.size); return adj, car

    def decompose_level_slice_to(self, level, adjusted, carries):
        dig = np.empty_like(adjusted)
        f = getattr(lib(), f"o_basis_decompose_level_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        f(C.byref(self._s), level, _ptr(adjusted), _ptr(dig), _ptr(carries), adjusted.size); return dig

    def decompose_slice(self, values):
        """All levels, LSB level first: returns [levels, n]."""
        values = _arr(values, self.bits); out = np.empty((self.decompose_length(), values.size), dtype=_np(self.bits))
        f = getattr(lib(), f"o_basis_decompose_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 3 + [C.c_size_t]
        f(C.byref(self._s), _ptr(values), _ptr(out), values.size); return out


class RNSBase:
    """primus_rns/src/base.rs:26-122; modulus-major batched layout (lib.rs:8-16)."""

    def __init__(self, moduli, bits=64):
        self.bits = bits; self.moduli = [int(m) for m in moduli]
        arr = _arr(self.moduli, bits) if self.moduli else np.zeros(0, dtype=_np(bits))
        err = C.c_int(0)
        f = getattr(lib(), f"o_rns_create{bits}"); f.restype = C.c_void_p
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        self._h = f(_ptr(arr), len(self.moduli), C.byref(err))
        if not self._h:
            raise OracleError(err.value)
        g = getattr(lib(), f"o_rns_value_len{bits}"); g.restype = C.c_size_t; g.argtypes = [C.c_void_p]
        self.value_len = int(g(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            f = getattr(lib(), f"o_rns_destroy{self.bits}"); f.restype = None; f.argtypes = [C.c_void_p]
            f(h); self._h = None

    def moduli_count(self): return len(self.moduli)
    def big_uint_value_len(self): return self.value_len

    def moduli_product(self):
        f = getattr(lib(), f"o_rns_product{self.bits}"); f.restype = C.POINTER(_ct(self.bits)); f.argtypes = [C.c_void_p]
        w = np.ctypeslib.as_array(f(self._h), shape=(self.value_len,))
        return sum(int(x) << (self.bits * i) for i, x in enumerate(w))

    def compose_multiple_values_to(self, multi_residues, value_count):
        r = _arr(multi_residues, self.bits); out = np.empty(value_count * self.value_len, dtype=_np(self.bits))
        f = getattr(lib(), f"o_rns_compose_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 3 + [C.c_size_t]
        f(self._h, _ptr(r), _ptr(out), value_count); return out

    def decompose_big_uint_values_to(self, big_values, value_count):
        b = _arr(big_values, self.bits); out = np.empty(value_count * len(self.moduli), dtype=_np(self.bits))
        f = getattr(lib(), f"o_rns_decompose_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 3 + [C.c_size_t]
        f(self._h, _ptr(b), _ptr(out), value_count); return out

    def wrapping_decompose_small_values_to(self, small, small_modulus):
        s = _arr(small, self.bits); out = np.empty(s.size * len(self.moduli), dtype=_np(self.bits))
        f = getattr(lib(), f"o_rns_lift_small{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, _ct(self.bits)]
        f(self._h, _ptr(s), _ptr(out), s.size, int(small_modulus)); return out

    def wrapping_decompose_small_values_scaled_add_to(self, small, acc, small_modulus, scalars):
        s = _arr(small, self.bits); sc = _arr(scalars, self.bits)
        f = getattr(lib(), f"o_rns_lift_small_scaled_acc{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 3 + [C.c_size_t, _ct(self.bits), C.c_void_p]
        f(self._h, _ptr(s), _ptr(acc), s.size, int(small_modulus), _ptr(sc)); return acc



class BaseConverter:
    """primus_rns/src/converter.rs:21-365: fast_convert_array / exact_convert_array, modulus-major layout."""

    def __init__(self, in_moduli, out_moduli, bits=64):
        self.bits, self.in_moduli, self.out_moduli = bits, [int(m) for m in in_moduli], [int(m) for m in out_moduli]
        err = C.c_int(0)
        f = getattr(lib(), f"o_baseconv_create{bits}"); f.restype = C.c_void_p
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        a, b = _arr(self.in_moduli, bits), _arr(self.out_moduli, bits)
        self._h = f(_ptr(a), len(self.in_moduli), _ptr(b), len(self.out_moduli), C.byref(err))
        if not self._h:
            raise OracleError(err.value)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            f = getattr(lib(), f"o_baseconv_destroy{self.bits}"); f.restype = None; f.argtypes = [C.c_void_p]
            f(h); self._h = None

    def matrix(self):
        f = getattr(lib(), f"o_baseconv_matrix{self.bits}"); f.restype = C.POINTER(_ct(self.bits)); f.argtypes = [C.c_void_p]
        return np.ctypeslib.as_array(f(self._h), shape=(len(self.out_moduli), len(self.in_moduli))).copy()

    def fast_convert_array(self, crt_in, poly_length):
        x = _arr(crt_in, self.bits); out = np.empty(len(self.out_moduli) * poly_length, dtype=_np(self.bits))
        f = getattr(lib(), f"o_baseconv_fast_array{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        f(self._h, _ptr(x), _ptr(out), poly_length); return out

    def exact_convert_array(self, crt_in, poly_length):
        x = _arr(crt_in, self.bits); out = np.empty(poly_length, dtype=_np(self.bits))
        f = getattr(lib(), f"o_baseconv_exact_array{self.bits}"); f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        if f(self._h, _ptr(x), _ptr(out), poly_length) != 0:
            raise ValueError("output base in exact_convert_array must be one.")
        return out


class BigUintApproxSignedBasis:
    """primus_decompose/src/big_integer/basis.rs:17-434."""

    def __init__(self, rns: RNSBase, log_basis, reverse_length=None):
        self.rns, self.bits = rns, rns.bits
        f = getattr(lib(), f"o_bigbasis_create{self.bits}"); f.restype = C.c_void_p
        f.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        self._h = f(rns._h, log_basis, reverse_length or 0)
        if not self._h:
            raise ValueError("BigUintApproxSignedBasis::new failed")
        self._log_basis = log_basis

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            f = getattr(lib(), f"o_bigbasis_destroy{self.bits}"); f.restype = None; f.argtypes = [C.c_void_p]
            f(h); self._h = None

    def _u(self, name):
        f = getattr(lib(), f"{name}{self.bits}"); f.restype = C.c_uint; f.argtypes = [C.c_void_p]
        return int(f(self._h))

    def decompose_length(self): return self._u("o_bigbasis_levels")
    def drop_bits(self): return self._u("o_bigbasis_drop_bits")
    def basis_value(self): return 1 << self._log_basis

    def init_value_carry_slice_inplace(self, big_values):
        n = big_values.size // self.rns.value_len; car = np.empty(n, dtype=np.uint8)
        f = getattr(lib(), f"o_bigbasis_init_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p] * 3 + [C.c_size_t]
        f(self._h, _ptr(big_values), _ptr(car), n); return car

    def unsigned_decompose_slice_to(self, level, big_values, carries):
        n = carries.size; dig = np.empty(n, dtype=_np(self.bits))
        f = getattr(lib(), f"o_bigbasis_unsigned_level_slice{self.bits}"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        f(self._h, level, _ptr(big_values), _ptr(dig), _ptr(carries), n); return dig



def butterfly_mul_factor(a, s, w, moduli, n):
    """DcrtPolynomial::butterfly_mul_factor_to (primus_poly/src/dcrt/mul.rs:189-222; scalar body slice_butterfly):
    (a, out) = (a + s, (a - s) * w) mod q per limb.  a, s: [rows][L][n]; w: [L][n].  Exact big-int restatement."""
    a = np.asarray(a); s_ = np.asarray(s); w = np.asarray(w)
    L = len(moduli)
    A = a.reshape(-1, L, n).astype(object); S = s_.reshape(-1, L, n).astype(object); W = w.reshape(L, n).astype(object)
    new_a = np.empty_like(A); out = np.empty_like(A)
    for li, q in enumerate(moduli):
        new_a[:, li, :] = (A[:, li, :] + S[:, li, :]) % q
        out[:, li, :] = ((A[:, li, :] - S[:, li, :]) % q) * W[li][None, :] % q
    return new_a.astype(a.dtype).reshape(a.shape), out.astype(a.dtype).reshape(a.shape)


# --------------------------------------------------------------------------- products
class DcrtTable:
    """One table per limb, looped (primus_ntt/src/dcrt/prime64.rs:11-128)."""

    def __init__(self, log_n, moduli, bits=64):
        cls = U64NttTable if bits == 64 else U32NttTable
        self.bits, self.tables = bits, [cls(log_n, m) for m in moduli]
        self.n, self.moduli = 1 << log_n, [int(m) for m in moduli]

    def poly_length(self): return self.n
    def moduli_count(self): return len(self.tables)
    def crt_poly_length(self): return self.n * len(self.tables)

    def transform_slice(self, v):
        for i, t in enumerate(self.tables):
            t.transform_slice(v[i * self.n:(i + 1) * self.n])

    def inverse_transform_slice(self, v):
        for i, t in enumerate(self.tables):
            t.inverse_transform_slice(v[i * self.n:(i + 1) * self.n])

    def _handles(self):
        return (C.c_void_p * len(self.tables))(*[t._h for t in self.tables])


def external_product(dcrt: DcrtTable, rns: RNSBase, basis: BigUintApproxSignedBasis, k, ggsw, glwe_in,
                     to_coeff=True, batch=1, threads=0):
    """CrtGlwe::mul_dcrt_ggsw_to (primus_lattice/src/glwe/crt.rs:200-227) [+ into_coeff_form]."""
    bits = dcrt.bits
    ggsw, glwe_in = _arr(ggsw, bits), _arr(glwe_in, bits); out = np.empty_like(glwe_in)
    f = getattr(lib(), f"o_external_product_batch{bits}"); f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_size_t, C.c_int]
    f(dcrt._handles(), rns._h, basis._h, k, _ptr(ggsw), _ptr(glwe_in), _ptr(out), int(to_coeff), batch,
      threads or max_threads())
    return out


def external_product_single(table: _NttTable, basis: ApproxSignedBasis, k, rgsw, glwe_in, to_coeff=True,
                            batch=1, threads=0):
    """L = 1 external product with the single-word signed basis."""
    bits = table.bits
    rgsw, glwe_in = _arr(rgsw, bits), _arr(glwe_in, bits); out = np.empty_like(glwe_in)
    f = getattr(lib(), f"o_external_product_single_batch{bits}"); f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
    f(table._h, C.byref(basis._s), k, _ptr(rgsw), _ptr(glwe_in), _ptr(out), int(to_coeff), batch,
      threads or max_threads())
    return out


def blind_rotate(table: _NttTable, basis: ApproxSignedBasis, bsk, n_lwe, lwe, test_vector, batch=1, threads=0):
    """Composed blind rotation. lwe: uint32 [batch, n_lwe+1] in Z_{2N}."""
    bits = table.bits
    bsk, tv = _arr(bsk, bits), _arr(test_vector, bits)
    lwe = np.ascontiguousarray(lwe, dtype=np.uint32)
    out = np.empty((batch, 2 * table.n), dtype=_np(bits))
    f = getattr(lib(), f"o_blind_rotate_batch{bits}"); f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    f(table._h, C.byref(basis._s), _ptr(bsk), n_lwe, _ptr(lwe), _ptr(tv), _ptr(out), batch, threads or max_threads())
    return out


def modulus_switch(lwe, q, log_2n):
    """LWE modulus switch q -> 2N = 2^log_2n, round to nearest: floor((v * 2N + floor(q/2)) / q) mod 2N.
    NOT in the reference (it has no bootstrapping); the convention is fixed here and in include/pfhe.h. Exact big-int arithmetic."""
    two_n = 1 << log_2n
    flat = [((int(v) * two_n + q // 2) // q) % two_n for v in np.asarray(lwe).reshape(-1)]
    return np.array(flat, dtype=np.uint32).reshape(np.asarray(lwe).shape)


def blind_rotate_ternary(table: _NttTable, basis: ApproxSignedBasis, bsk_plus, bsk_minus, n_lwe, lwe, test_vector, batch=1):
    """Ternary-secret blind rotation by monomial combination (NOT in the reference -- composed from its primitives):
    ACC <- (0, tv * X^(2N-b)); for every i: K = (NTT(X^a_i) - 1) .* BSK+_i + (NTT(X^-a_i) - 1) .* BSK-_i (transform_coeff_one_monomial,
    prime64/table.rs:611-651; slice ops), ACC <- ACC + into_coeff_form(mul_dcrt_ggsw_to(ACC, K)) (glwe/crt.rs:200-227)."""
    bits, n, q = table.bits, table.n, table.q
    dt = _np(bits)
    bsk_plus, bsk_minus, tv = _arr(bsk_plus, bits), _arr(bsk_minus, bits), _arr(test_vector, bits)
    lwe = np.ascontiguousarray(lwe, dtype=np.uint32).reshape(batch, n_lwe + 1)
    levels = basis.decompose_length()
    rgsw = 2 * levels * 2 * n
    out = np.empty((batch, 2 * n), dtype=dt)
    for b in range(batch):
        acc = np.zeros(2 * n, dtype=dt)
        acc[n:] = mul_monomial(tv, (2 * n - int(lwe[b, n_lwe])) % (2 * n), q, bits)
        for i in range(n_lwe):
            a = int(lwe[b, i]) % (2 * n)
            mp = (table.transform_coeff_one_monomial(a).astype(object) - 1) % q
            mm = (table.transform_coeff_one_monomial((2 * n - a) % (2 * n)).astype(object) - 1) % q
            kp = bsk_plus[i * rgsw:(i + 1) * rgsw].astype(object).reshape(-1, n)
            km = bsk_minus[i * rgsw:(i + 1) * rgsw].astype(object).reshape(-1, n)
            key = ((kp * mp + km * mm) % q).astype(dt).reshape(-1)
            prod = external_product_single(table, basis, 1, key, acc.reshape(1, -1), to_coeff=True, batch=1, threads=1)[0]
            acc = ((acc.astype(object) + prod.astype(object)) % q).astype(dt)
        out[b] = acc
    return out
