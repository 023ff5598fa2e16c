import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def port_lib():
    import orc
    return orc.port()


@pytest.fixture(scope="session")
def ref_lib():
    import orc
    lib = orc.ref()
    if lib is None:
        pytest.skip("oracle/_ref/liborc_ref.so not available (reference tree absent and no prebuilt copy)")
    return lib
