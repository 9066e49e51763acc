import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_sigma")
CUDA_LIB = os.path.join(ROOT, "seqfrost_b200", "csrc", "libsigma_b200.so")

GOLDEN_CASES = sorted(f[:-3] for f in os.listdir(GOLDEN) if f.endswith(".in"))
# reference options each fixture was generated with (tests/golden/make_golden.py)
GOLDEN_OPTS = {
    "gates_bvebound": ["-bvebound"],
    "gates_phases2": ["--phases=2"],
    "gates_nofun_cm4": ["-no-function", "--boundedclausemax=4"],
    "ere_extend2": ["--redundancyextend=2"],
    "bce_gates": ["-blocked"], "bce_ksat3": ["-blocked"], "all_gates": ["-all"],
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    import __graft_entry__ as ge
    return ge.build_cuda()


@pytest.fixture(scope="session")
def emul_lib():
    import __graft_entry__ as ge
    return ge.build_emul()


@pytest.fixture(scope="session")
def gpu_sigma(cuda_lib):
    from seqfrost_b200 import sigma
    s = sigma.Sigma(0, cuda_lib)
    yield s
    s.close()
