import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref_tree(tmp_path_factory):
    """The reference's test vectors, unpacked from the committed bundle (tests/golden/)."""
    from tests.golden_util import materialize
    return materialize(tmp_path_factory.mktemp("ref"))
