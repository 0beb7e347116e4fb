import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def model_paths(tmp_path_factory):
    """{name: path} of model/*.txt files materialised from the packed fixture (no /root/reference needed)."""
    from em_model_manned_bayes_b200.model_archive import materialize
    return materialize(str(tmp_path_factory.mktemp("models")))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "vectors.npz"))
    cases = {}
    for k in z.files:
        c, f = k.rsplit("/", 1)
        cases.setdefault(c, {})[f] = z[k]
    return cases
