import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def front_oracle():
    from oracle_loader import load_front_oracle
    return load_front_oracle()


@pytest.fixture(scope="session")
def ref_strict():
    from oracle_loader import load_reference
    return load_reference("strict")


@pytest.fixture(scope="session")
def b200():
    from mytinygl_b200 import load_b200
    return load_b200()
