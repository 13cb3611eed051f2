import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(REPO, "tests", "golden")

    def load(name):
        return np.load(os.path.join(gdir, name + ".npz"))
    return load


@pytest.fixture(autouse=True)
def _scratch_cwd(tmp_path, monkeypatch):
    """Samplers and multistarts write logs_and_csvs/ and pickles/ like the reference: keep them out of the repository."""
    monkeypatch.chdir(tmp_path)
