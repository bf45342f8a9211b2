import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name), weights_only=False)
    return load


def pytest_sessionfinish(session, exitstatus):
    """Write the measured parity errors of this session (tests/parity_util.RECORDS) where gpurun brings them back."""
    try:
        import json
        from parity_util import RECORDS
        if not RECORDS:
            return
        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        by_test = {}
        for r in RECORDS:
            key = r["where"].rsplit(":", 1)[0] + " [" + r["kind"] + "]"
            b = by_test.setdefault(key, {"comparisons": 0, "max_abs_err": 0.0, "max_err_over_ref_scale": 0.0, "ref_max_abs": 0.0,
                                         "rtol": r["rtol"], "max_atol_used": 0.0})
            b["comparisons"] += 1
            b["max_abs_err"] = max(b["max_abs_err"], r["max_abs_err"])
            b["max_err_over_ref_scale"] = max(b["max_err_over_ref_scale"], r["max_err_over_scale"])
            b["ref_max_abs"] = max(b["ref_max_abs"], r["ref_max_abs"])
            b["max_atol_used"] = max(b["max_atol_used"], r["atol"])
        with open(os.path.join(out_dir, "parity_r02.json"), "w") as f:
            json.dump({"exitstatus": int(exitstatus), "tolerance": "outputs: rtol 1e-4 / atol 1e-5 literal; gradients: atol x "
                       "max(1, max|ref|)", "tests": by_test}, f, indent=1, sort_keys=True)
    except Exception:  # noqa: BLE001
        pass
