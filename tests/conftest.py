"""pytest configuration: the ``gpu`` marker and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden():
    return load_golden


# ---- measured accuracy next to every FP32 bound ("stop hiding tolerances"): tests call record_measured();
# the numbers are printed in the terminal summary and written to gpurun_out/measured_tolerances.json (copied to profiles/)
_MEASURED = []


def record_measured(test: str, what: str, value: float, bound: float, note: str = "") -> float:
    _MEASURED.append({"test": test, "what": what, "measured": float(value), "bound": float(bound), "note": note})
    return value


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if not _MEASURED:
        return
    import json
    worst = {}
    for m in _MEASURED:
        key = (m["test"], m["what"])
        if key not in worst or m["measured"] > worst[key]["measured"]:
            worst[key] = m
    rows = sorted(worst.values(), key=lambda m: (m["test"], m["what"]))
    terminalreporter.write_line("measured accuracy (worst case per test / quantity):")
    for m in rows:
        terminalreporter.write_line(f"  {m['test']:58s} {m['what']:10s} measured {m['measured']:.2e}  bound {m['bound']:.1e}  {m['note']}")
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "measured_tolerances.json"), "w") as fh:
            json.dump(rows, fh, indent=1)
    except OSError:
        pass
