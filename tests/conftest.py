import os
import sys

import pytest

# More hardware work queues than the default 8: tests run up to 8 ranks / proofs as independent streams of ONE process
# whose kernels wait for each other (k_shard_exchange); two such streams aliased onto one queue would deadlock until the
# kernel's timeout.  Must be in the environment before the first CUDA call of the process.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle.lib()
