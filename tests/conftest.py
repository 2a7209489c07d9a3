import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/flappie_oracle.c); built on demand with gcc."""
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own object code (oracle/_ref); only where it has been built."""
    from oracle import pyoracle
    if not pyoracle.have_ref():
        if os.path.isdir("/root/reference/src"):
            pyoracle.build(ref=True)
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    try:
        return pyoracle.Ref()
    except OSError as e:   # e.g. bundled OpenBLAS missing on this box
        pytest.skip(f"oracle/_ref not loadable: {e}")


@pytest.fixture(scope="session")
def lib():
    from flappie_b200.api import Library
    return Library.get()


@pytest.fixture(scope="session")
def gpu_lib(lib):
    if lib.device_count() < 1:
        pytest.fail("no CUDA device visible but a gpu-marked test was selected")
    return lib
