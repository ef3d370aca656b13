import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def fixture_gt(golden_dir):
    """GT int8 [nvar, N, 2], samples, positions of the reference's example VCF (oracle reader)."""
    from oracle import ingest_ref

    return ingest_ref.read_vcf(os.path.join(golden_dir, "data", "test_genotypes.vcf.gz"))
