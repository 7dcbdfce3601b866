import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def pytest_sessionfinish(session, exitstatus):
    """GPU runs: keep the parity statistics the tests collected (per-element gradient errors, contributor-count
    mismatch rates) as gpurun_out/parity_report.json."""
    try:
        import json
        mod = sys.modules.get("tests.test_gpu_parity")
        rep = getattr(mod, "REPORT", None) if mod is not None else None
        if rep:
            out = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "parity_report.json"), "w") as f:
                json.dump(rep, f, indent=1, sort_keys=True)
    except Exception:
        pass
