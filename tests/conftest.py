import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def staged():
    from pbrt_v2_b200 import scenes
    scenes.stage()
    if not scenes.have_staged():
        pytest.skip("fixture scenes are not staged (run __graft_entry__.build() where /root/reference exists)")
    return scenes
