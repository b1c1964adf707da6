import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nvit", "model.py"))


@pytest.fixture(scope="session")
def reference_model_module():
    """Live import of the reference's nvit.model (build container only; skipped on the GPU box)."""
    if not reference_available():
        pytest.skip("/root/reference not present on this machine")
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    try:
        import nvit.model as ref_model  # noqa: WPS433
    except Exception as e:  # flash_attn import is unconditional in the reference (model.py:7)
        pytest.skip(f"reference import failed: {e}")
    return ref_model
