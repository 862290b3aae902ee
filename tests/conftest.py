import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def srb():
    import hashlib
    return hashlib.sha256(b"Mr F was here").digest()   # tests/t_batch_verifier.nim:60


@pytest.fixture(scope="session")
def br():
    from oracle import blst_ref
    return blst_ref


@pytest.fixture(scope="session")
def cache():
    import nim_blscurve_b200 as bg
    c = bg.BatchedBLSVerifierCache(max_sets=4096, device=0)
    yield c
    c.close()
