import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_DIR = os.path.join(ROOT, "models", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def model_dir():
    """models/_ref is unpacked from the reference archives by __graft_entry__.build()."""
    if not os.path.isdir(os.path.join(MODEL_DIR, "DeepFilterNet3")):
        try:
            import __graft_entry__ as g
            g._unpack_reference_models()
        except Exception:
            pass
    if not os.path.isdir(os.path.join(MODEL_DIR, "DeepFilterNet3")):
        pytest.skip("pretrained weights not unpacked (models/_ref); run __graft_entry__.build() where /root/reference exists")
    return MODEL_DIR
