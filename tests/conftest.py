import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def brca():
    """The bundled brca-eu counts (config 1), committed as a fixture under tests/golden/
    because /root/reference does not exist on the GPU box."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "brca_eu_counts.npz"))
    return [(z["rowptr0"], z["term0"], z["count0"]), (z["rowptr1"], z["term1"], z["count1"])]
