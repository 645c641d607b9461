import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def oracle_engine(monkeypatch):
    """Route ``Engine.get()`` to the CPU oracle - TEST ONLY, to exercise host logic without a GPU."""
    from helpers import OracleEngine
    from sequencing_b200.solver import engine as engine_mod

    fake = OracleEngine()
    monkeypatch.setattr(engine_mod.Engine, "get", classmethod(lambda cls, device_index=None: fake))
    return fake


@pytest.fixture(scope="session")
def gpu_engine():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sequencing_b200.solver import Engine

    return Engine.get(0)
