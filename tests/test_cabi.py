"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/pfhe.h declares, and reports the architecture it was compiled for.  No compute calls (no GPU here)."""
import os
import subprocess
import pytest

import primus_fhe_b200 as P


def test_library_exports_every_declared_symbol():
    lib = P._cabi.lib()
    syms = P.declared_symbols()
    assert len(syms) > 80
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.pfhe_compiled_arch() == b"sm_100a"
    assert lib.pfhe_status_string(5) == b"ModulusTooLarge" and lib.pfhe_status_string(1) == b"NoPrimitiveRoot"


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", P.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    archs = {line.split(".")[-2] for line in out.splitlines() if "sm_" in line and line.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, out[:400]


def test_no_product_code_imports_the_oracle():
    """The oracle is test infrastructure: nothing under primus_fhe_b200/ may reference it."""
    root = os.path.dirname(P.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "pfhe_oracle" not in text, f


def test_geometry_and_status_paths_without_gpu():
    """Host-only entry points work without a device; device entry points fail loudly (no CPU fallback)."""
    import ctypes as C
    lv, dr = C.c_uint32(), C.c_uint32()
    lib = P._cabi.lib()
    lib.pfhe_basis64_geometry.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    assert lib.pfhe_basis64_geometry(1125899906826241, 7, 0, C.byref(lv), C.byref(dr)) == 0
    assert (lv.value, dr.value) == (7, 1)        # l = 7, drop_bits = 1
    lib.pfhe_basis32_geometry.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    assert lib.pfhe_basis32_geometry(132120577, 7, 0, C.byref(lv), C.byref(dr)) == 0
    assert (lv.value, dr.value) == (3, 6)        # C4-A: l = 3, drop_bits = 6
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        lib.pfhe_ntt64_create.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_void_p]
        rc = lib.pfhe_ntt64_create(0, 12, 1125899906826241, C.byref(h))
        assert rc == 8 and not h.value           # PFHE_ERR_CUDA: the product path refuses to run without a GPU


def test_multiply_factor_matches_reference_tests():
    """MultiplyFactor (primus_factor/src/mul_factor/mod.rs:95-147: test_32 / test_52 / test_64): host-side setup arithmetic,
    callable without a GPU; checked against exact big-int products on seeded inputs."""
    import numpy as np
    import primus_fhe_b200 as P
    rng = np.random.default_rng(52)
    for q, shift in ((536813569, 32), (562949953392641, 52), (1152921504606830593, 64)):
        for _ in range(300):
            a, b = int(rng.integers(0, q)), int(rng.integers(0, q))
            mf = P.MultiplyFactor(a, shift, q)
            assert mf.quotient() == ((a << shift) // q) & ((1 << 64) - 1)
            assert mf.mul_modulo(b) == a * b % q
    import pytest
    with pytest.raises(P.PfheError):
        P.MultiplyFactor(5, 40, 17)          # "Unsupported BitShift"
    with pytest.raises(P.PfheError):
        P.MultiplyFactor(17, 64, 17)         # operand must be less than modulus


def test_host_only_handles_match_oracle_without_gpu():
    """RNSBase / BigUintApproxSignedBasis geometry / BaseConverter handles are host-only value types (constants travel as
    kernel parameters): constructors, error variants and geometry work -- and agree with the oracle -- on a CPU-only host."""
    import pytest
    import primus_fhe_b200 as P
    from oracle import oracle as O
    with pytest.raises(P.PfheError) as e:
        P.RNSBase([])
    assert e.value.name == "EmptyBase"                                  # primus_rns/tests/rns.rs:67-70
    with pytest.raises(P.PfheError) as e:
        P.RNSBase([21, 35])
    assert e.value.name == "CoPrimeError"                               # rns.rs:74-77
    q50, q50b, q49 = 1125899906826241, 1125899906629633, 562949953392641
    for bits, moduli in ((64, [3, 5, 7]), (64, [q50, q50b]), (64, [q50, q50b, q49]), (32, [134215681, 134176769])):
        g, o = P.RNSBase(moduli, bits), O.RNSBase(moduli, bits)
        assert g.big_uint_value_len() == o.big_uint_value_len() and g.moduli_product() == o.moduli_product()
        for beta, rev in ((7, None), (1, None), (13, 2)):
            if bits == 64 and moduli == [3, 5, 7] and beta > 6:
                continue
            gb, ob = P.BigUintApproxSignedBasis(g, beta, rev), O.BigUintApproxSignedBasis(o, beta, rev)
            assert (gb.decompose_length(), gb.drop_bits()) == (ob.decompose_length(), ob.drop_bits())
    with pytest.raises(P.PfheError):
        P.BigUintApproxSignedBasis(P.RNSBase([q50, q50b]), 7, 99)        # more levels than the modulus has
    P.BaseConverter([17, 19, 23], [29, 31])                             # rns.rs:282-284
    with pytest.raises(P.PfheError) as e:
        P.BaseConverter([17, 19, 23], [29, 58])
    assert e.value.name == "CoPrimeError"
    with pytest.raises(P.PfheError) as e:
        P.BaseConverter([], [29])
    assert e.value.name == "EmptyBase"


def test_round2_host_only_entry_points_without_gpu():
    """Byte layout (macros/mod.rs:39-97) and the UintNttTable constructor rules (primitive.rs:114-181) are host-side: they work and
    fail exactly the same on a CPU-only host."""
    import numpy as np
    import pytest
    import primus_fhe_b200 as P
    rng = np.random.default_rng(7)
    for bits, dt in ((32, np.uint32), (64, np.uint64)):
        w = rng.integers(0, 1 << 27, 257, dtype=np.uint64).astype(dt)
        b = P.to_bytes(w, bits)
        assert b == w.astype(w.dtype.newbyteorder("<")).tobytes() and len(b) == w.size * bits // 8
        assert np.array_equal(P.from_bytes(b, bits), w)
        assert np.array_equal(P.from_bytes(b"\x01" + b[1:], bits)[1:], w[1:])
        with pytest.raises(P.PfheError):
            P.from_bytes(b[:-1], bits)
    for bits, q, log_n, name in ((16, 12289, 13, "NoPrimitiveRoot"), (32, 132120577, 21, "NoPrimitiveRoot"), (16, 40961, 10, "ModulusTooLarge"),
                                 (32, 3221225473, 10, "ModulusTooLarge"), (64, 97, 7, "NoPrimitiveRoot"), (64, 1152921504606830593 * 4 + 1, 3, None)):
        with pytest.raises(P.PfheError) as e:
            P.UintNttTable(log_n, q, bits)
        assert name is None or e.value.name == name


def test_shard_rule_matches_the_c_abi_documentation():
    """pfhe_multi_* documents shard r of `total` over n parts as starting at r*floor(total/n) + min(r, total mod n): the same rule
    as primus_fhe_b200.shard.shard_range (checked here so the two cannot drift)."""
    from primus_fhe_b200.shard import shard_range
    for total in (0, 1, 7, 1250, 10000, 65537):
        for n in (1, 2, 3, 8):
            for r in range(n):
                b = r * (total // n) + min(r, total % n)
                e = b + total // n + (1 if r < total % n else 0)
                assert shard_range(total, n, r) == (b, e)


def test_rust_ffi_crate_bindings_are_complete_and_current():
    """ffi/primus_cuda/src/sys.rs is generated from include/pfhe.h (tools/gen_rust_sys.py): it must be current and declare every
    exported symbol, and the safe wrappers must only call symbols that exist."""
    import os, re, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"]).returncode == 0, \
        "run `python tools/gen_rust_sys.py` after editing include/pfhe.h"
    sys_rs = open(os.path.join(root, "ffi", "primus_cuda", "src", "sys.rs")).read()
    declared = set(re.findall(r"pub fn (pfhe_[a-z0-9_]+)\(", sys_rs))
    import primus_fhe_b200 as P
    assert declared == set(P.declared_symbols())
    for f in ("ntt.rs", "dcrt.rs", "bootstrap.rs", "multi.rs", "lib.rs"):
        used = set(re.findall(r"\b(pfhe_[a-z0-9]+_[a-z0-9_]+)\b", open(os.path.join(root, "ffi", "primus_cuda", "src", f)).read()))
        used = {u for u in used if not re.fullmatch(r"pfhe_(ntt|dcrt|bsk|rns|baseconv|uintntt)(16|32|64)", u)} - {"pfhe_status", "pfhe_cuda", "pfhe_slice_op"}
        assert used <= declared, (f, used - declared)


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    import primus_fhe_b200 as P
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(P.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_smoke.c"),
                    "-o", exe, "-L", libdir, "-lpfhe_cuda", "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    return exe


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/pfhe.h compiles as C11 with -Wall -Wextra -Werror, examples/c_abi_smoke.c links against the shared library with nothing but
    the C runtime, and -- there being no GPU in the CPU test environment -- the product path fails loudly instead of falling back."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is covered by tests/test_gpu_ext.py")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "failed" in p.stderr and "smoke ok" not in p.stdout


def test_documented_environment_hooks_exist_in_the_sources():
    """DESIGN.md section 10 lists the A/B switches; every one of them must be read somewhere under csrc/ (and vice versa for the PFHE_* getenv
    calls), so the table cannot drift from the code."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    design = open(os.path.join(root, "DESIGN.md")).read()
    section = design[design.index("## 10. Environment hooks"):design.index("## 11.")]
    documented = set(re.findall(r"`(PFHE_[A-Z0-9_]+)", section))
    src = ""
    csrc = os.path.join(root, "primus_fhe_b200", "csrc")
    for f in os.listdir(csrc):
        src += open(os.path.join(csrc, f)).read()
    read = set(re.findall(r'getenv\("(PFHE_[A-Z0-9_]+)"\)', src)) | set(re.findall(r'env_int(?:_early)?\("(PFHE_[A-Z0-9_]+)"', src))
    assert documented <= read, sorted(documented - read)
    assert read <= documented, sorted(read - documented)
