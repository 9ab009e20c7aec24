"""ctypes loader for libpfhe_cuda.so (the C-ABI declared in include/pfhe.h).

Fails loudly: if the shared library is missing it is NOT replaced by any CPU path -- importing
succeeds (so CPU-only hosts can inspect symbols) but every call site goes through `lib()`, which
raises when the library cannot be loaded.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpfhe_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pfhe.h")

STATUS_NAMES = {
    0: "Ok", 1: "NoPrimitiveRoot", 2: "DegreeConversionErr", 3: "DegreeTooLarge", 4: "NttTableErr",
    5: "ModulusTooLarge", 6: "EmptyBase", 7: "CoPrimeError", 8: "CudaError", 9: "InvalidArgument", 10: "Unsupported",
}


class PfheError(RuntimeError):
    """Raised for every non-OK pfhe_status. `.code` / `.name` mirror NttError / RNSError variants
    (primus_ntt/src/error.rs:7-49, primus_rns/src/error.rs:7-20)."""

    def __init__(self, code: int, detail: str = ""):
        self.code = int(code)
        self.name = STATUS_NAMES.get(self.code, str(code))
        super().__init__(f"{self.name}{(': ' + detail) if detail else ''}")


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m primus_fhe_b200.build` "
                "(there is no CPU fallback for the CUDA hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.pfhe_status_string.restype = C.c_char_p
        _lib.pfhe_last_cuda_error.restype = C.c_char_p
        _lib.pfhe_version.restype = C.c_char_p
        _lib.pfhe_compiled_arch.restype = C.c_char_p
        _lib.pfhe_launch_count.restype = C.c_uint64
    return _lib


def check(status: int):
    if status != 0:
        detail = ""
        if status == 8:
            detail = lib().pfhe_last_cuda_error().decode()
        raise PfheError(status, detail)


def declared_symbols() -> list[str]:
    """Every function name declared in include/pfhe.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfhe_[a-z0-9_]+)\s*\(", text)))


def launch_count() -> int:
    return int(lib().pfhe_launch_count())
