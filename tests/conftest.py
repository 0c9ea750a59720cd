import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "frontend_golden.npz"))


@pytest.fixture(autouse=True)
def _eager_frontend_by_default(monkeypatch):
    """Parity tests drive the kernels launch by launch (pinned RNG streams, launch counting, engine switches): modules built
    inside a test do not capture themselves into CUDA graphs unless the test asks for it (m.graph_replay = True)."""
    try:
        import biear_b200.frontend as fe
    except Exception:  # noqa: BLE001
        return
    monkeypatch.setattr(fe, "GRAPH_REPLAY_DEFAULT", False)
    # `module.engine = "chain"` selects the per-frame cross-check engine (test infrastructure, tests/chain_engine.py)
    from tests import chain_engine
    monkeypatch.setattr(fe, "CHAIN_ENGINE", chain_engine.run)
