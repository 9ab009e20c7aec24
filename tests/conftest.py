import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# primes used by the reference's own tests / benches
Q27 = 132120577            # prime32/tests.rs:5, prime64/tests.rs:5
Q29 = 536813569            # tests/ntt.rs:17
Q30 = 1073692673           # benches/bench_u64.rs:8
Q28 = 268369921            # benches/bench_u32.rs:8
Q49 = 562949953392641      # tests/ntt.rs:55
Q50 = 1125899906826241     # benches/bench_u64.rs:8
Q50B = 1125899906629633    # primus_rns/benches/decompose.rs:15
Q60 = 1152921504606830593  # tests/ntt.rs:93
