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


@pytest.fixture
def plain_diag(monkeypatch):
    """SEQ_METHOD_DIAG on the plain diagonal kernel (kernel_diag.cuh) - the block-local variant (kernel_diagblk.cuh),
    which takes over whenever a small tensor factor can stay inside a thread, switched off."""
    monkeypatch.setenv("SEQ_DIAG_BLK", "0")


class _RecordingEngine:
    """The real CUDA engine with the OracleEngine's bookkeeping (how many launches, how large)."""

    def __init__(self, eng):
        self._eng = eng
        self.calls = 0
        self.batch_sizes = []

    def solve(self, bp, *a, **kw):
        self.calls += 1
        self.batch_sizes.append(bp.B)
        return self._eng.solve(bp, *a, **kw)

    def __getattr__(self, name):
        return getattr(self._eng, name)


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def kat_engine(request, monkeypatch):
    """The reference's known-answer tests run twice: through the CPU oracle (``-m "not gpu"``: pins the host
    mirror and the oracle) and through the REAL CUDA engine (``-m gpu``: the reference's own acceptance
    thresholds on the product path)."""
    from sequencing_b200.solver import engine as engine_mod

    if request.param == "oracle":
        from helpers import OracleEngine

        fake = OracleEngine()
        monkeypatch.setattr(engine_mod.Engine, "get", classmethod(lambda cls, device_index=None: fake))
        return fake
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    real = engine_mod.Engine.get(0)
    rec = _RecordingEngine(real)
    monkeypatch.setattr(engine_mod.Engine, "get", classmethod(lambda cls, device_index=None: rec))
    return rec
