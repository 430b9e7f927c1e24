import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def client_key():
    from oracle import tfhe
    return tfhe.ClientKey.load(os.path.join(GOLDEN, "client_key"))


@pytest.fixture(scope="session")
def server_key(client_key):
    """ServerKey::new(&client_key) analogue (engine.rs:252), oracle keygen, seed 0."""
    from oracle import tfhe
    return tfhe.keygen_server(client_key, seed=0)
