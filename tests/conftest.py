import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import _lib

    _lib.load()  # raises (= test error, not skip) if the in-tree .so is missing on a GPU box
    return torch.device("cuda:0")
