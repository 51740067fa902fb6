import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand from oracle/bbme_oracle.c."""
    from oracle import binding
    binding.load()
    return binding


@pytest.fixture(scope="session")
def gpu():
    """A context on cuda:0.  Fails loudly (no skip, no fallback) if the extension or the device is missing."""
    import blockbasedmotionestimation_b200 as bb
    est = bb.Estimator(64, 64, [12], [4], chunk_pairs=1)  # any plan: stage_* calls do not depend on it
    yield est
    est.close()
