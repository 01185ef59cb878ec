import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_next: experimental kernels not yet verified on a GPU (STFEM_RUN_NEXT=1, never part of -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    import dealii_stfem_b200 as st
    c = st.Context(0)
    yield c
    c.close()
