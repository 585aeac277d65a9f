import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the shipped checkpoint is a reference artefact: copied (git-ignored) by __graft_entry__.build()
CKPT = os.path.join(GOLDEN, "_ref", "model_15xchr19.pt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ckpt_path():
    if not os.path.exists(CKPT):
        pytest.skip("shipped checkpoint not present (tests/golden/_ref, filled by __graft_entry__.build())")
    return CKPT
